"""GPU parity tests: the CUDA path (through the C ABI of libomnigs_b200.so) against
  (1) the committed golden fixtures = outputs of the unmodified reference rasterizer,
  (2) the CPU oracle on seeded inputs,
  (3) the reference rasterizer itself (oracle/_ref/omnigs_ref.so) when it travelled to the box,
  (4) size-independent properties at BASELINE.json's full size (C2: 1M Gaussians, 2048x1024).

Parity bar (BASELINE.json north_star): num_rendered, radii, tile ranges, sorted point list and keys
bit-exact; image <= 1e-5 max-abs (in fact bit-identical); gradients <= 1e-4 relative, "a tolerance stated to allow for
atomic-order differences", measured per tensor in the max norm (max|a-b| / max|b|) AND per element
(|a-b| <= 1e-4|b| + c * max|b|).

What that tolerance can mean for the outputs of the per-Gaussian chain through cov2D — dL_dmeans3D, dL_dcov3D, dL_dscales,
dL_drotations — is measured, not assumed (profiles/r02_grad_noise.json: all five BASELINE configs; profiles/
r02_parity_spread.json: every case of this file, worst of six runs; tools/grad_noise.py, tools/parity_spread.py):
  * the REFERENCE run twice on the same forward state differs from itself by up to 1.9e-4 on them with the lonlat camera
    (1.4e-3 with the perspective one) — its float atomics are unordered —, against < 1e-5 on the other four tensors;
  * the reference's chain, evaluated in float, is up to 6e-4 away from the same chain evaluated in DOUBLE on the
    reference's own render-backward outputs (division by det(cov2D)^2, backward.cu:395-407; differences of nearly equal
    entries of dL/dM, :544-547) — no float implementation can sit within 1e-4 of the reference on them unless it
    reproduces the reference's rounding errors;
  * our chain (gaussian_grad.cuh) stays within 1e-4 of the double evaluation of its own inputs (worst measured 6.4e-5);
  * building render_bwd with the reference's own expf / IEEE division instead of ex2.approx / rcp.approx changes none
    of these figures (the "exact_math" rows of r02_grad_noise.json).
The asserted bars (tests/_harness.py assert_gradient_parity):
  * dL_dmeans2D, dL_dcolors, dL_dopacity, dL_dsh and the intermediate dL_dconic (everything the chain consumes): <= 1e-4
    in the max norm AND per element with c = 5e-6, as stated, no noise term (measured: <= 1.5e-5 and c <= 2e-6);
  * the four chain outputs: ours within 1e-4 of the double arbiter (perspective camera, f-4: 1e-3 — the reference's own
    chain sits up to 9.3e-4 from its double evaluation there, ours measured <= 2.4e-4), and ours-vs-reference <= max(1e-4, ours-vs-double + reference-vs-double + F x the
    reference's own run-to-run difference (worst of three re-runs; eight for the perspective camera)), F = 2 (4 for the
    perspective camera, whose noise is heavy-tailed) — the triangle inequality through the double evaluation.  Together:
    every input of the chain within the stated tolerance of the reference's, and the chain within the stated tolerance of
    its exact evaluation.
No outlier allowances."""
import numpy as np
import pytest
import torch

import _harness as h
import cases

pytestmark = pytest.mark.gpu
sm = h.scene_mod

IMG_TOL = 1e-5
GRAD_REL = 1e-4
ILL_CONDITIONED = ("dL_dmeans3D", "dL_dcov3D", "dL_dscales", "dL_drotations")
GRAD_TOL = {n: 3e-4 for n in ILL_CONDITIONED}


def _np(t):
    return t.detach().cpu().numpy()


def assert_grads_close(ours, ref, names=h.GRAD_NAMES, ref_again=None, ill=ILL_CONDITIONED, cap=5e-4, floor=GRAD_REL, factor=2.0):
    """Max-norm bound per tensor.  ref_again: a second run of the reference on the same input; its distance from the
    first (the reference's own noise floor) times `factor` widens the bound for the tensors in `ill`, up to `cap`.
    Without ref_again (golden fixtures: one stored reference run) the ill-conditioned tensors get GRAD_TOL."""
    for i, (n, a, b) in enumerate(zip(names, ours, ref)):
        a = torch.as_tensor(a).double().cpu().flatten()
        b = torch.as_tensor(b).double().cpu().flatten()
        assert a.shape == b.shape, n
        if b.numel() == 0:
            continue
        scale = float(b.abs().max())
        if scale < 1e-9:
            assert float(a.abs().max()) < 1e-8, n
            continue
        diff = (a - b).abs()
        tol = GRAD_TOL.get(n, GRAD_REL) if n in ill else GRAD_REL
        if ref_again is not None and n in ill:
            noise = float((torch.as_tensor(ref_again[i]).double().cpu().flatten() - b).abs().max()) / scale
            tol = min(cap, max(floor, factor * noise))
        assert float(diff.max()) / scale <= tol, (n, float(diff.max()) / scale, tol)


def assert_own_runs_close(n, a, b, tol):
    """Two evaluations of OUR backward that differ only in float-atomic order / band decomposition."""
    scale = float(b.abs().max()) + 1e-30
    err = (a - b).abs() / scale
    # our own run-to-run noise on the ill-conditioned tensors reaches 6e-5 at C2 (profiles/r02_grad_noise.json, ours_vs_ours)
    assert float(err.max()) < (3e-4 if n in ILL_CONDITIONED else tol), (n, float(err.max()))


def bits(t):
    return t.contiguous().view(torch.int32)


# ------------------------------------------------------------------ (1) golden fixtures
@pytest.mark.parametrize("name", list(cases.CASES))
def test_against_golden_reference_outputs(name):
    gold = np.load(f"{h.ROOT}/tests/golden/{name}.npz")
    scene, view, dL_np, c = cases.build(name)
    assert bytes(gold["input_sha256"]).decode() == cases.input_hash(scene, view, dL_np)
    d = h.torch_inputs(scene, view, mode=c["mode"], bg=c["bg"], degree=c["degree"], render_depth=c.get("render_depth", False))
    fwd = h.run_forward(h.pkg, d)
    depth_only = bool(c.get("render_depth"))   # the reference's backward has no depth path
    grads = None if depth_only else h.run_backward(h.pkg, d, fwd, torch.from_numpy(dL_np).cuda())
    st = h.ours_state(d, fwd)
    torch.cuda.synchronize()

    if "present" in gold:   # checkFrustum (camera_type 1)
        assert np.array_equal(_np(h.pkg.markVisible(d["means3D"], d["viewmatrix"], d["projmatrix"], 1)), gold["present"])
    assert fwd[0] == int(gold["num_rendered"])
    assert np.array_equal(_np(fwd[2]), gold["radii"])
    assert np.array_equal(_np(st["tiles_touched"]), gold["tiles_touched"])
    vis = gold["radii"] > 0
    for k in ("means2D", "depths", "conic_opacity"):
        assert np.array_equal(_np(st[k])[vis].view(np.int32), gold[k][vis].view(np.int32)), k
    assert np.array_equal(_np(st["ranges"]), gold["ranges"])
    assert np.array_equal(_np(st["point_list"]), gold["point_list"])
    assert np.array_equal(_np(st["point_list_keys"]), gold["point_list_keys"])
    if "rgb" in gold and not depth_only:
        assert np.abs(_np(st["rgb"])[vis] - gold["rgb"][vis]).max() <= 1e-6
        assert np.array_equal(_np(st["clamped"])[vis], gold["clamped"][vis])
    # depth frames blend camera-space depths (up to 20) instead of colours in [0, 1]: same relative bound
    img_tol = IMG_TOL * (float(np.abs(gold["out_color"]).max()) if depth_only else 1.0)
    assert np.abs(_np(fwd[1]) - gold["out_color"]).max() <= max(IMG_TOL, img_tol)
    assert np.abs(_np(st["accum_alpha"]) - gold["accum_alpha"]).max() <= IMG_TOL
    assert np.array_equal(_np(st["n_contrib"]), gold["n_contrib"])
    if not depth_only:
        assert_grads_close(grads, [gold[n] for n in h.GRAD_NAMES])


# ------------------------------------------------------------------ (2) CPU oracle
@pytest.mark.parametrize("seed,W,H,degree,bg", [(1, 177, 93, 3, (0, 0, 0)), (2, 64, 48, 0, (1, 0.5, 0.25))])
def test_against_cpu_oracle(seed, W, H, degree, bg):
    from oracle import oracle
    scene = sm.make_scene(2500, W, H, 0.04, 40 + seed, pole_frac=0.1, seam_frac=0.05, near_frac=0.01)
    view = sm.random_view(50 + seed)
    dL_np = sm.make_grad_image(W, H, 60 + seed)
    d = h.torch_inputs(scene, view, bg=bg, degree=degree)
    fwd = h.run_forward(h.pkg, d)
    grads = h.run_backward(h.pkg, d, fwd, torch.from_numpy(dL_np).cuda())
    bgn = np.array(bg, np.float32)
    of = oracle.forward(scene.means3D, scene.opacities, view[0], view[1], W, H, bgn, shs=scene.shs, degree=degree,
                        scales=scene.scales, rotations=scene.rotations)
    og = oracle.backward(of, dL_np, scene.means3D, view[0], view[1], W, H, bgn, shs=scene.shs, degree=degree,
                         scales=scene.scales, rotations=scene.rotations)
    # CPU libm vs CUDA libdevice: allow isolated threshold flips, nothing systematic
    assert abs(fwd[0] - of["num_rendered"]) <= max(4, of["num_rendered"] // 500)
    assert int((_np(fwd[2]) != of["radii"]).sum()) <= 2
    diff = np.abs(_np(fwd[1]) - of["out_color"]).max(axis=0)
    assert (diff > 1e-4).mean() <= 2e-3 and diff.max() < 0.1
    for n, g in zip(h.GRAD_NAMES, grads):
        ref = og[n].reshape(tuple(g.shape))
        scale = np.abs(ref).max() + 1e-30
        assert np.abs(_np(g) - ref).max() / scale < 1e-2, n


# ------------------------------------------------------------------ (3) the reference itself
REF_CASES = {
    "C1": lambda: (sm.make_config_scene("C1"), sm.identity_view(), "sh", (0, 0, 0), 3),
    "C1_view": lambda: (sm.make_config_scene("C1"), sm.random_view(71), "sh", (1, 1, 1), 3),
    "odd_deg2": lambda: (sm.make_scene(30000, 517, 263, 0.03, 72, pole_frac=0.1, seam_frac=0.05), sm.random_view(73), "sh", (0, 0, 0), 2),
    "colors": lambda: (sm.make_scene(30000, 400, 200, 0.03, 74), sm.random_view(75), "colors", (0.1, 0.2, 0.3), 0),
    "cov_deg1": lambda: (sm.make_scene(30000, 400, 200, 0.03, 76), sm.random_view(77), "cov", (0, 0, 0), 1),
    "long_lists": lambda: (sm.make_scene(40000, 256, 128, 0.15, 78), sm.identity_view(), "sh", (0, 0, 0), 3),
    # perspective camera (camera_type 1)
    "pin_C1": lambda: (sm.make_config_scene("C1"), sm.perspective_view(81, 1024, 512, 90.0), "sh", (0, 0, 0), 3),
    "pin_odd_cov": lambda: (sm.make_scene(60000, 517, 263, 0.03, 82), sm.perspective_view(83, 517, 263, 60.0), "cov", (1, 1, 1), 2),
    "pin_colors_wide": lambda: (sm.make_scene(60000, 400, 200, 0.05, 84), sm.perspective_view(85, 400, 200, 120.0), "colors", (0.1, 0.2, 0.3), 0),
}


@pytest.mark.parametrize("case", list(REF_CASES))
def test_against_reference_rasterizer(case):
    if h.load_reference() is None:
        pytest.skip("oracle/_ref/omnigs_ref.so not present on this box")
    scene, view, mode, bg, degree = REF_CASES[case]()
    d = h.torch_inputs(scene, view, mode=mode, bg=bg, degree=degree)
    ref = h.load_reference()
    fo, fr = h.run_forward(h.pkg, d), h.run_forward(ref, d)
    so, sr = h.ours_state(d, fo), h.ref_state(ref, d, fr)
    torch.cuda.synchronize()
    assert fo[0] == fr[0]
    assert torch.equal(fo[2], fr[2])
    vis = fr[2] > 0
    assert torch.equal(so["tiles_touched"], sr["tiles_touched"])
    for k in ("means2D", "depths", "conic_opacity"):
        assert torch.equal(bits(so[k][vis]), bits(sr[k][vis])), k
    assert torch.equal(so["ranges"], sr["ranges"])
    assert torch.equal(so["point_list"], sr["point_list"])
    assert torch.equal(so["point_list_keys"], sr["point_list_keys"])
    assert float((fo[1] - fr[1]).abs().max()) <= IMG_TOL
    if d["camera_type"] == 3 and mode != "colors":
        # beyond the stated tolerance: SH colours and blend are pinned to the reference's operation order, so the
        # lonlat image is bit-identical
        assert torch.equal(bits(so["rgb"][vis]), bits(sr["rgb"][vis]))
    if d["camera_type"] == 3:
        assert torch.equal(bits(fo[1]), bits(fr[1]))
    assert float((so["accum_alpha"] - sr["accum_alpha"]).abs().max()) <= IMG_TOL
    assert torch.equal(so["n_contrib"], sr["n_contrib"])
    del fo, fr, so, sr
    # gradients: noise floor + double arbiter (module docstring; tests/_harness.py parity_report / assert_gradient_parity)
    rep = h.parity_report(scene, view, mode=mode, bg=bg, degree=degree, noise_runs=8 if case.startswith("pin") else 3)
    assert rep["integers"]["num_rendered"] and rep["integers"]["radii"] and rep["integers"]["point_list"]
    h.assert_gradient_parity(rep, tag=case)


# ------------------------------------------------------------------ (3b) every BASELINE config, noise floor + double arbiter
@pytest.mark.parametrize("name", ["C1", "C2", "C3", "C4", "C5"])
def test_baseline_config_against_reference_with_noise_floor_and_double_arbiter(name):
    """all five BASELINE configs against the live reference.  Integers and the image bit-exact; gradients
    against the reference's own run-to-run noise and a double-precision evaluation of the per-Gaussian chain (see the
    module docstring for the bars and tests/_harness.py parity_report for how they are measured).  C5 is the
    sort/bin stress (4.8e8 instances: 23 GB here, 60 GB in the reference)."""
    if h.load_reference() is None:
        pytest.skip("oracle/_ref/omnigs_ref.so not present on this box")
    rep = h.config_parity_report(name)
    ints = dict(rep["integers"])
    assert ints.pop("image_maxdiff") == 0.0
    assert all(ints.values()), rep["integers"]
    h.assert_gradient_parity(rep, tag=name)


def test_more_than_2_to_the_30_tile_instances():
    """The reference takes any `int num_rendered` (rasterizer_impl.cu:627-632); round 1 stopped at 2^30 (30-bit look-back
    words).  The C5 stress scene with every Gaussian 2.5x larger has ~1.2e9 instances (20 GB of binning state): too large
    for the reference on one GPU (it needs 36 R + CUB temporaries), so the list is checked through its defining properties."""
    scene = sm.make_config_scene("C5")
    scene.scales[:] = scene.scales * np.float32(2.5)
    d = h.torch_inputs(scene, sm.identity_view())
    fwd = h.run_forward(h.pkg, d)
    R = fwd[0]
    assert (1 << 30) < R < (1 << 31), R
    P, W, H = scene.P, scene.W, scene.H
    st = h.pkg.export_forward_state(P, W, H, R, fwd[3], fwd[4], fwd[5], want_keys=False)
    assert int(st["tiles_touched"].long().sum()) == R
    ranges = st["ranges"].long()
    lens = ranges[:, 1] - ranges[:, 0]
    assert bool((lens >= 0).all()) and int(lens.sum()) == R
    nz = lens > 0
    starts = torch.cumsum(lens, 0) - lens                       # row-major tile order, no gaps
    assert torch.equal(ranges[nz, 0], starts[nz]) and int(ranges[nz, 1].max()) == R
    # sampled tiles (the longest, the first, the last and 40 random ones): every entry is a visible Gaussian whose rect
    # covers the tile; depths ascend, equal depths keep ascending Gaussian index (the reference's stable sort order)
    gx = (W + 15) // 16
    pick = torch.nonzero(nz).flatten()
    g = torch.Generator(device="cpu").manual_seed(1)
    sample = torch.cat([lens.argmax()[None].cpu(), pick[:1].cpu(), pick[-1:].cpu(), pick.cpu()[torch.randint(0, pick.numel(), (40,), generator=g)]])
    depth_bits = bits(st["depths"]).long()
    m2d = st["means2D"]
    for t in sample.tolist():
        a, b = int(ranges[t, 0]), int(ranges[t, 1])
        ids = st["point_list"][a:b].long()
        assert bool((fwd[2][ids] > 0).all())
        key = depth_bits[ids] * (1 << 31) + ids                  # (depth, index) lexicographic
        assert bool((key[1:] > key[:-1]).all()), t
        ty, tx = divmod(t, gx)
        rad = fwd[2][ids].float()
        cx, cy = m2d[ids, 0], m2d[ids, 1]
        assert bool(((cx - rad) < (tx + 1) * 16).all() and ((cx + rad + 16) > tx * 16).all()
                    and ((cy - rad) < (ty + 1) * 16).all() and ((cy + rad + 16) > ty * 16).all()), t
    img = fwd[1]
    assert bool(torch.isfinite(img).all())
    nc = st["n_contrib"].view(H, W).long()
    tile_len = lens.view((H + 15) // 16, gx).repeat_interleave(16, 0).repeat_interleave(16, 1)[:H, :W]
    assert bool((nc <= tile_len).all())
    del st, depth_bits, tile_len, nc
    grads = h.run_backward(h.pkg, d, fwd, torch.from_numpy(sm.make_grad_image(W, H, 99)).cuda())
    for n, t in zip(h.GRAD_NAMES, grads):
        assert bool(torch.isfinite(t).all()), n
    assert float(grads[3].abs().max()) > 0


# ------------------------------------------------------------------ (4) full-size properties (C2)
@pytest.fixture(scope="module")
def c2():
    scene = sm.make_config_scene("C2")
    d = h.torch_inputs(scene, sm.random_view(21))
    fwd = h.run_forward(h.pkg, d)
    return scene, d, fwd


def test_c2_binning_invariants(c2):
    scene, d, fwd = c2
    R = fwd[0]
    st = h.ours_state(d, fwd)
    keys = st["point_list_keys"]
    assert int(st["tiles_touched"].long().sum()) == R
    assert bool((keys[1:] >= keys[:-1]).all()), "keys not sorted"
    # equal keys keep ascending Gaussian index (stability of the reference's radix sort)
    same = keys[1:] == keys[:-1]
    pl = st["point_list"].long()
    assert bool((pl[1:][same] > pl[:-1][same]).all())
    ranges = st["ranges"].long()
    lens = ranges[:, 1] - ranges[:, 0]
    assert int(lens.sum()) == R and bool((lens >= 0).all())
    nz = lens > 0
    tiles = (keys >> 32)
    assert torch.equal(torch.unique(tiles), torch.nonzero(nz).flatten())
    assert torch.equal(tiles[ranges[nz, 0]], torch.nonzero(nz).flatten())
    # every list entry is a visible Gaussian whose depth bits are the low key word
    depth_bits = bits(st["depths"]).long() & 0xFFFFFFFF
    assert torch.equal(keys & 0xFFFFFFFF, depth_bits[pl])
    assert bool((fwd[2][pl] > 0).all())
    # a checksum of checksums: per-Gaussian multiplicity in the list == tiles_touched
    mult = torch.bincount(pl, minlength=scene.P)
    assert torch.equal(mult, st["tiles_touched"].long())


def test_c2_forward_is_deterministic(c2):
    scene, d, fwd = c2
    again = h.run_forward(h.pkg, d)
    assert again[0] == fwd[0]
    assert torch.equal(bits(again[1]), bits(fwd[1]))
    assert torch.equal(again[2], fwd[2])
    img = fwd[1]
    assert bool(torch.isfinite(img).all()) and float(img.min()) >= 0.0


def test_c2_backward_is_linear_in_upstream_gradient(c2):
    scene, d, fwd = c2
    g1 = torch.from_numpy(sm.make_grad_image(scene.W, scene.H, 1)).cuda()
    g2 = torch.from_numpy(sm.make_grad_image(scene.W, scene.H, 2)).cuda()
    a = h.run_backward(h.pkg, d, fwd, g1)
    b = h.run_backward(h.pkg, d, fwd, g2)
    c = h.run_backward(h.pkg, d, fwd, 2.0 * g1 - 3.0 * g2)
    for n, x, y, z in zip(h.GRAD_NAMES, a, b, c):
        lin = 2.0 * x - 3.0 * y
        scale = float(lin.abs().max()) + 1e-30
        assert float((z - lin).abs().max()) / scale < 2e-4, n
    culled = fwd[2] == 0
    for n, x in zip(h.GRAD_NAMES, a):
        assert not bool(x[culled].any()), n     # culled Gaussians get exact zeros
    assert not bool(a[0][:, 2].any())           # dL_dmeans2D.z is never written
    assert bool(all(torch.isfinite(x).all() for x in a))


def test_c2_against_reference_if_present(c2):
    ref = h.load_reference()
    if ref is None:
        pytest.skip("oracle/_ref/omnigs_ref.so not present on this box")
    scene, d, fo = c2
    dL = torch.from_numpy(sm.make_grad_image(scene.W, scene.H, 99)).cuda()
    fr = h.run_forward(ref, d)
    assert fo[0] == fr[0] and torch.equal(fo[2], fr[2])
    so, sr = h.ours_state(d, fo), h.ref_state(ref, d, fr)
    assert torch.equal(so["point_list"], sr["point_list"])
    assert torch.equal(so["point_list_keys"], sr["point_list_keys"])
    assert torch.equal(so["ranges"], sr["ranges"])
    assert float((fo[1] - fr[1]).abs().max()) <= IMG_TOL
    assert torch.equal(so["n_contrib"], sr["n_contrib"])
    # gradients of this configuration: test_baseline_config_against_reference_with_noise_floor_and_double_arbiter[C2]


# ------------------------------------------------------------------ edge cases
def test_empty_scene_returns_zero_outputs():
    e = torch.empty((0,), device="cuda")
    m = torch.empty((0, 3), device="cuda")
    out = h.pkg.RasterizeGaussiansCUDA(torch.ones(3, device="cuda"), m, e, e, e, e, 1.0, e, torch.eye(4, device="cuda"),
                                       torch.eye(4, device="cuda"), 0.0, 0.0, 20, 30, e, 0, torch.zeros(3, device="cuda"),
                                       False, 3, False)
    assert out[0] == 0 and out[1].shape == (3, 20, 30) and not bool(out[1].any())   # rasterize_points.cu:84,97
    assert out[2].numel() == 0
    g = h.pkg.RasterizeGaussiansBackwardCUDA(torch.ones(3, device="cuda"), m, out[2], e, e, e, 1.0, e,
                                             torch.eye(4, device="cuda"), torch.eye(4, device="cuda"), 0.0, 0.0,
                                             torch.zeros(3, 20, 30, device="cuda"), e, 0, torch.zeros(3, device="cuda"),
                                             out[3], 0, out[4], out[5], 3)
    assert [tuple(x.shape) for x in g] == [(0, 3), (0, 3), (0, 1), (0, 3), (0, 6), (0, 0, 3), (0, 3), (0, 4)]


def test_all_culled_renders_background():
    scene = sm.make_scene(300, 50, 30, 0.02, 5)
    scene.means3D[:] *= 0.005                      # everything inside the r <= 0.2 cull sphere
    d = h.torch_inputs(scene, sm.identity_view(), bg=(0.25, 0.5, 0.75))
    fwd = h.run_forward(h.pkg, d)
    assert fwd[0] == 0 and not bool(fwd[2].any())
    exp = torch.tensor([0.25, 0.5, 0.75], device="cuda")[:, None, None].expand(3, 30, 50)
    assert torch.equal(fwd[1], exp)
    g = h.run_backward(h.pkg, d, fwd, torch.ones(3, 30, 50, device="cuda"))
    assert not any(bool(x.any()) for x in g)


def test_invalid_camera_type_raises_like_the_reference():
    scene = sm.make_scene(10, 32, 16, 0.02, 6)
    d = h.torch_inputs(scene, sm.identity_view())
    args = [d["background"], d["means3D"], d["colors"], d["opacity"], d["scales"], d["rotations"], 1.0,
            d["cov3D_precomp"], d["viewmatrix"], d["projmatrix"], 0.0, 0.0, 16, 32, d["sh"], 3, d["campos"], False]
    with pytest.raises(RuntimeError, match=r"\[CudaRasterizer\]Invalid camera_type"):
        h.pkg.RasterizeGaussiansCUDA(*args, 2, False)
    with pytest.raises(RuntimeError, match=r"\[CudaRasterizer\]Invalid camera_type"):
        h.pkg.markVisible(d["means3D"], d["viewmatrix"], d["projmatrix"], 0)


def test_pinhole_against_cpu_oracle():
    """camera_type 1 against the C restatement of preprocessCUDA / computeCov2DCUDA, incl. markVisible
    and a render_depth frame."""
    from oracle import oracle
    W, H = 177, 93
    scene = sm.make_scene(6000, W, H, 0.04, 91, near_frac=0.02)
    view = sm.perspective_view(92, W, H, 80.0)
    pin = dict(projmatrix=view[1], tan_fovx=view[3], tan_fovy=view[4])
    dL_np = sm.make_grad_image(W, H, 93)
    bgn = np.array((0.2, 0.1, 0.3), np.float32)
    d = h.torch_inputs(scene, view, bg=tuple(bgn), degree=3)
    fwd = h.run_forward(h.pkg, d)
    grads = h.run_backward(h.pkg, d, fwd, torch.from_numpy(dL_np).cuda())
    kw = dict(shs=scene.shs, degree=3, scales=scene.scales, rotations=scene.rotations, pinhole=pin)
    of = oracle.forward(scene.means3D, scene.opacities, view[0], view[2], W, H, bgn, **kw)
    og = oracle.backward(of, dL_np, scene.means3D, view[0], view[2], W, H, bgn, **kw)
    assert abs(fwd[0] - of["num_rendered"]) <= max(4, of["num_rendered"] // 500)
    assert int((_np(fwd[2]) != of["radii"]).sum()) <= 2
    diff = np.abs(_np(fwd[1]) - of["out_color"]).max(axis=0)
    assert (diff > 1e-4).mean() <= 2e-3 and diff.max() < 0.1
    for n, g in zip(h.GRAD_NAMES, grads):
        ref = og[n].reshape(tuple(g.shape))
        scale = np.abs(ref).max() + 1e-30
        assert np.abs(_np(g) - ref).max() / scale < 1e-2, n
    present = h.pkg.markVisible(d["means3D"], d["viewmatrix"], d["projmatrix"], 1)
    assert np.array_equal(_np(present), oracle.pinhole_mark_visible(scene.means3D, view[0]))
    assert 0 < int(present.sum()) < scene.P
    # depth frame: camera-space z blended into all three channels
    dd = h.torch_inputs(scene, view, bg=(0, 0, 0), degree=3, render_depth=True)
    fd = h.run_forward(h.pkg, dd)
    od = oracle.forward(scene.means3D, scene.opacities, view[0], view[2], W, H, np.zeros(3, np.float32),
                        **dict(kw, pinhole=dict(pin, render_depth=True)))
    img = _np(fd[1])
    assert np.array_equal(img[0], img[1]) and np.array_equal(img[1], img[2])
    ddiff = np.abs(img - od["out_color"]).max(axis=0)
    assert (ddiff > 1e-3).mean() <= 2e-3


def test_mark_visible_marks_everything():
    m = torch.randn(1000, 3, device="cuda")
    v = h.pkg.markVisible(m, torch.eye(4, device="cuda"), torch.eye(4, device="cuda"), 3)
    assert v.dtype == torch.bool and bool(v.all())


@pytest.mark.parametrize("W,H", [(1, 1), (16, 16), (17, 15), (33, 65)])
def test_tiny_and_ragged_images_match_oracle(W, H):
    from oracle import oracle
    scene = sm.make_scene(200, W, H, 0.2, 8)
    view = sm.identity_view()
    d = h.torch_inputs(scene, view)
    fwd = h.run_forward(h.pkg, d)
    of = oracle.forward(scene.means3D, scene.opacities, view[0], view[1], W, H, np.zeros(3, np.float32), shs=scene.shs,
                        degree=3, scales=scene.scales, rotations=scene.rotations)
    assert abs(fwd[0] - of["num_rendered"]) <= 2
    diff = np.abs(_np(fwd[1]) - of["out_color"])
    assert np.median(diff) < 1e-5 and (diff > 1e-3).mean() < 0.02


def test_single_huge_gaussian_fans_out_over_every_tile():
    # one Gaussian whose square rect covers the whole 2048x1024 grid: exercises the load-balanced
    # emission (one source -> 8192 slots) and a polar position
    scene = sm.make_scene(1, 2048, 1024, 0.02, 9)
    scene.means3D[0] = (0.01, -3.0, 0.02)          # almost at the pole (lat ~ -90 deg, +y is down)
    scene.scales[0] = (2.0, 2.0, 2.0)
    d = h.torch_inputs(scene, sm.identity_view())
    fwd = h.run_forward(h.pkg, d)
    st = h.ours_state(d, fwd)
    assert fwd[0] == int(st["tiles_touched"][0]) == 128 * 64
    assert bool((st["point_list"] == 0).all())
    assert torch.equal(st["ranges"][:, 0].long(), torch.arange(8192, device="cuda"))


def test_latitude_bands_partition_the_frame():
    """Stage 1 with a tile-row band: the union of the bands' lists is the full frame's list."""
    import ctypes
    lib = h.pkg.load_library()
    scene = sm.make_scene(20000, 512, 256, 0.03, 10, pole_frac=0.2)
    d = h.torch_inputs(scene, sm.random_view(11))
    full = h.run_forward(h.pkg, d)
    st_full = h.ours_state(d, full)
    P, W, H = scene.P, scene.W, scene.H
    gy = (H + 15) // 16
    total, img = 0, torch.zeros_like(full[1])
    for (b0, b1) in [(0, 5), (5, 11), (11, gy)]:
        radii = torch.empty(P, dtype=torch.int32, device="cuda")
        geom = torch.empty(lib.ogs_geom_bytes(P), dtype=torch.uint8, device="cuda")
        imgb = torch.empty(lib.ogs_img_bytes(W, H), dtype=torch.uint8, device="cuda")
        n = ctypes.c_int64(0)
        p = lambda t: ctypes.c_void_p(t.data_ptr()) if t.numel() else None
        rc = lib.ogs_lonlat_forward_stage1_band(P, 3, 16, W, H, b0, b1, p(d["means3D"]), p(d["sh"]), None, p(d["opacity"]),
                                               p(d["scales"]), 1.0, p(d["rotations"]), None, p(d["viewmatrix"]),
                                               p(d["campos"]), p(radii), p(geom), p(imgb), ctypes.byref(n), None)
        assert rc == 0
        R = n.value
        binb = torch.empty(lib.ogs_binning_bytes(R, W, H), dtype=torch.uint8, device="cuda")
        out = torch.empty(3, H, W, device="cuda")
        assert lib.ogs_lonlat_forward_stage2(P, W, H, R, p(d["background"]), p(geom), p(binb), p(imgb), p(out), None) == 0
        torch.cuda.synchronize()
        assert torch.equal(radii, full[2])          # radii are the unclipped reference radii
        st = h.pkg.export_forward_state(P, W, H, R, geom, binb, imgb)
        rows = slice(b0 * (W // 16), b1 * (W // 16))
        lens_band = (st["ranges"][:, 1] - st["ranges"][:, 0]).long()
        lens_full = (st_full["ranges"][:, 1] - st_full["ranges"][:, 0]).long()
        assert torch.equal(lens_band[rows], lens_full[rows]) and int(lens_band.sum()) == R == int(lens_full[rows].sum())
        start = int(st_full["ranges"][rows][lens_full[rows] > 0][0, 0]) if R else 0
        assert torch.equal(st["point_list"], st_full["point_list"][start:start + R])
        img[:, b0 * 16:min(H, b1 * 16)] = out[:, b0 * 16:min(H, b1 * 16)]
        total += R
    assert total == full[0]
    assert torch.equal(bits(img), bits(full[1]))


# ------------------------------------------------------------------ the C++ LibTorch drop-in
def _load_cpp_shim():
    import importlib
    import os
    import sys
    d = os.path.join(h.ROOT, "omnigs-fork_b200")
    if not os.path.exists(os.path.join(d, "omnigs_b200_torch.so")):
        return None
    if d not in sys.path:
        sys.path.insert(0, d)
    return importlib.import_module("omnigs_b200_torch")


def test_cpp_dropin_matches_python_mirror_bit_for_bit():
    """csrc/rasterize_points.cpp exports the reference's three symbols; driven through pybind it must
    give exactly what the ctypes mirror gives (same C ABI underneath) and keep the reference's errors."""
    shim = _load_cpp_shim()
    if shim is None:
        pytest.skip("omnigs_b200_torch.so not built")
    scene = sm.make_scene(30000, 517, 263, 0.03, 72, pole_frac=0.1, seam_frac=0.05)
    d = h.torch_inputs(scene, sm.random_view(73), bg=(0.3, 0.2, 0.1), degree=2)
    dL = torch.from_numpy(sm.make_grad_image(scene.W, scene.H, 5)).cuda()
    fa, fb = h.run_forward(h.pkg, d), h.run_forward(shim, d)
    assert fa[0] == fb[0] and torch.equal(fa[2], fb[2]) and torch.equal(bits(fa[1]), bits(fb[1]))
    assert fb[3].dtype == torch.uint8 and fb[4].dtype == torch.uint8 and fb[5].dtype == torch.uint8
    ga, gb = h.run_backward(h.pkg, d, fa, dL), h.run_backward(shim, d, fb, dL)
    for n, x, y in zip(h.GRAD_NAMES, ga, gb):
        assert x.shape == y.shape, n
        # only the order of the float atomics differs between two runs (amplified for the ill-conditioned tensors)
        assert_own_runs_close(n, x, y, 1e-5)
    v = shim.markVisible(d["means3D"], d["viewmatrix"], d["projmatrix"], 3)
    assert v.dtype == torch.bool and bool(v.all())
    args = [d["background"], d["means3D"], d["colors"], d["opacity"], d["scales"], d["rotations"], 1.0,
            d["cov3D_precomp"], d["viewmatrix"], d["projmatrix"], 0.0, 0.0, d["H"], d["W"], d["sh"], 2, d["campos"], False]
    with pytest.raises(RuntimeError, match=r"\[CudaRasterizer\]Invalid camera_type"):
        shim.RasterizeGaussiansCUDA(*args, 7, False)
    with pytest.raises(RuntimeError, match="means3D must have dimensions"):
        bad = list(args); bad[1] = torch.zeros(5, 2, device="cuda")
        shim.RasterizeGaussiansCUDA(*bad, 3, False)


def test_band_gradients_sum_to_full_frame_gradients():
    """render + backward per latitude band (what each GPU of a band-parallel job does);
    the bands' images tile the full image bit for bit and their gradient shares add up to the full-frame
    gradients.  Bands are balanced with parallel.band_rows on the per-row instance counts."""
    import importlib
    par = importlib.import_module("omnigs-fork_b200.parallel")
    scene = sm.make_scene(60000, 1024, 512, 0.02, 15, pole_frac=0.15, seam_frac=0.05)
    d = h.torch_inputs(scene, sm.random_view(16), bg=(0.1, 0.2, 0.3))
    dL = torch.from_numpy(sm.make_grad_image(scene.W, scene.H, 17)).cuda()
    full = h.run_forward(h.pkg, d)
    g_full = h.run_backward(h.pkg, d, full, dL)
    rows = par.tile_row_counts(h.ours_state(d, full)["ranges"], scene.W, scene.H)
    for world in (2, 4):
        bands = par.band_rows(rows, world)
        assert bands[0][0] == 0 and bands[-1][1] == len(rows)
        loads = [sum(rows[a:b]) for a, b in bands]
        assert max(loads) < 1.35 * sum(loads) / world
        img = torch.zeros_like(full[1])
        acc = None
        for band in bands:
            def rasterize(b):
                return h.pkg.RasterizeGaussiansCUDA(
                    d["background"], d["means3D"], d["colors"], d["opacity"], d["scales"], d["rotations"], 1.0,
                    d["cov3D_precomp"], d["viewmatrix"], d["projmatrix"], 0.0, 0.0, d["H"], d["W"], d["sh"], 3,
                    d["campos"], False, 3, False, band=b)
            part, fwd = par.render_band_forward(rasterize, band, scene.H)   # single process: no all-reduce
            img += part
            assert torch.equal(fwd[2], full[2])
            g = h.run_backward(h.pkg, d, fwd, dL)
            acc = [x.clone() for x in g] if acc is None else [a + x for a, x in zip(acc, g)]
        assert torch.equal(bits(img), bits(full[1]))
        for n, a, b in zip(h.GRAD_NAMES, acc, g_full):
            assert_own_runs_close(n, a, b, 2e-5)

    # cheaper exchange: sum the packed accumulators of the bands between the two backward kernels
    # (what an all-reduce of 48 B/Gaussian does) and finish once on the sums
    bands = par.band_rows(rows, 3)
    fwds = [rasterize(b) for b in bands]
    parts = []
    for fwd in fwds[:-1]:
        h.run_backward(h.pkg, d, fwd, dL, reduce_accumulators=lambda t: parts.append(t.clone()))
    def add_others(t):
        for p_ in parts:
            t += p_
    g_sum = h.run_backward(h.pkg, d, fwds[-1], dL, reduce_accumulators=add_others)
    for n, a, b in zip(h.GRAD_NAMES, g_sum, g_full):
        assert_own_runs_close(n, a, b, 2e-5)

    # pipelined exchange (ogs_lonlat_backward_finish_range): the per-Gaussian backward run range by range on caller-owned
    # accumulators gives the same bits as one launch over all Gaussians
    acc_a = torch.empty((scene.P, 12), device="cuda")
    acc_b = torch.empty((scene.P, 12), device="cuda")
    seen = []
    whole = h.run_backward(h.pkg, d, full, dL, accumulators=acc_a, reduce_accumulators=lambda t: None)
    def per_range(t, ranges):
        seen.append(ranges)
        return [None] * len(ranges)
    chunked = h.run_backward(h.pkg, d, full, dL, accumulators=acc_b, reduce_accumulators=per_range, accumulator_chunks=5)
    assert len(seen) == 1 and len(seen[0]) == 5 and sum(c for _, c in seen[0]) == scene.P
    assert all(f % 128 == 0 for f, _ in seen[0])
    for n, a, b in zip(h.GRAD_NAMES, chunked, whole):
        if n in ("dL_dmeans2D", "dL_dcolors", "dL_dopacity", "dL_dsh"):
            assert_own_runs_close(n, a, b, 2e-5)       # two render backwards: atomic order differs
    # same accumulators, whole vs ranges: bit-identical
    lib = h.pkg.load_library()
    again = h.run_backward(h.pkg, d, full, dL, accumulators=acc_b, reduce_accumulators=lambda t: t.copy_(acc_a))
    ranged = h.run_backward(h.pkg, d, full, dL, accumulators=acc_b,
                            reduce_accumulators=lambda t, r: (t.copy_(acc_a), [None] * len(r))[1], accumulator_chunks=7)
    for n, a, b in zip(h.GRAD_NAMES, ranged, again):
        assert torch.equal(bits(a), bits(b)), n


def test_caller_streams_and_side_stream_ordering():
    """The library forks tile_ranges onto its own side stream and joins it back (csrc/c_api.cu): frames issued back
    to back on the caller's current stream — default or not, alternating, with no host synchronisation in between —
    must give the same bits as a lone frame."""
    scene = sm.make_scene(60000, 640, 320, 0.02, 41)
    d = h.torch_inputs(scene, sm.random_view(42))
    dL = torch.from_numpy(sm.make_grad_image(scene.W, scene.H, 43)).cuda()
    base_f = h.run_forward(h.pkg, d)
    base_g = h.run_backward(h.pkg, d, base_f, dL)
    base_s = h.ours_state(d, base_f)
    torch.cuda.synchronize()
    streams = [torch.cuda.Stream(), torch.cuda.Stream()]
    runs = []
    for i in range(6):
        st = streams[i % 2]
        st.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(st):
            f = h.run_forward(h.pkg, d)
            g = h.run_backward(h.pkg, d, f, dL)
            runs.append((f, g, h.ours_state(d, f)))
    torch.cuda.synchronize()
    for f, g, s in runs:
        assert f[0] == base_f[0]
        assert torch.equal(f[1], base_f[1]) and torch.equal(f[2], base_f[2])
        for k in ("ranges", "point_list", "n_contrib"):
            assert np.array_equal(_np(s[k]), _np(base_s[k])), k
        assert_grads_close(g, base_g)


def _fan_out_scene(P, seed):
    """P Gaussians next to the pole, each large enough to cover all 32768 tiles of a 4096x2048 frame, at distinct depths."""
    scene = sm.make_scene(P, 4096, 2048, 0.02, seed)
    rng = np.random.Generator(np.random.PCG64(seed))
    scene.means3D[:] = np.array((0.01, -3.0, 0.02), np.float32)[None, :] * rng.uniform(0.9, 1.1, (P, 1)).astype(np.float32)
    scene.scales[:] = (4.0, 4.0, 4.0)
    return scene


def test_more_than_2_30_tile_instances_sort_like_the_reference():
    """Maximum size: the reference accepts any int num_rendered (rasterizer_impl.cu:627-632); round 1
    refused R >= 2^30.  32768 + 8 Gaussians x 32768 tiles = 2^30 + 262144 instances (17 GB of binning state here, 39 GB in
    the reference): every tile's list must be the depth order of all Gaussians; bit-exact against the reference when it
    is on the box."""
    P, T = 32768 + 8, 32768
    scene = _fan_out_scene(P, 11)
    d = h.torch_inputs(scene, sm.identity_view())
    fwd = h.run_forward(h.pkg, d)
    R = fwd[0]
    assert R == P * T and R >= (1 << 30)
    st = h.pkg.export_forward_state(P, scene.W, scene.H, R, fwd[3], fwd[4], fwd[5], want_keys=False)
    ranges = st["ranges"].long() & 0xFFFFFFFF
    assert torch.equal(ranges[:, 0], torch.arange(T, device="cuda") * P) and torch.equal(ranges[:, 1] - ranges[:, 0], torch.full((T,), P, device="cuda"))
    # depth order (ties by index) of all P Gaussians, repeated once per tile
    depth_bits = bits(st["depths"]).long() & 0xFFFFFFFF
    order = torch.argsort(depth_bits * (1 << 20) + torch.arange(P, device="cuda"))
    pl = st["point_list"].view(T, P)
    for t0 in range(0, T, 4096):
        assert bool((pl[t0:t0 + 4096] == order.to(torch.int32)[None, :]).all())
    assert bool(torch.isfinite(fwd[1]).all())
    ref = h.load_reference()
    if ref is not None:
        img, n_contrib = fwd[1].clone(), st["n_contrib"].clone()
        pl_first, pl_last = pl[:64].clone(), pl[-64:].clone()
        del fwd, st, pl
        torch.cuda.empty_cache()
        fr = h.run_forward(ref, d)
        assert fr[0] == R
        sr = h.ref_state(ref, d, fr)
        rp = sr["point_list"].view(T, P)
        assert torch.equal(rp[:64], pl_first) and torch.equal(rp[-64:], pl_last)
        for t0 in range(0, T, 4096):
            assert bool((rp[t0:t0 + 4096] == order.to(torch.int32)[None, :]).all())
        assert torch.equal(bits(fr[1]), bits(img)) and torch.equal(sr["n_contrib"], n_contrib)


def test_more_than_2_31_tile_instances_are_rejected():
    """2^31 and more overflows the reference's `int num_rendered`; here stage 1 refuses (OGS_ERR_TOO_MANY) before any
    R-sized buffer is asked for, and the library stays usable."""
    scene = _fan_out_scene(65536 + 8, 13)
    d = h.torch_inputs(scene, sm.identity_view())
    with pytest.raises(h.pkg.OgsError, match="2\\^31"):
        h.run_forward(h.pkg, d)
    small = sm.make_scene(2000, 160, 80, 0.03, 12)
    fwd = h.run_forward(h.pkg, h.torch_inputs(small, sm.identity_view()))
    assert fwd[0] > 0 and bool(torch.isfinite(fwd[1]).all())
