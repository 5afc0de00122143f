"""GPU tests of the multi-view / data-parallel extensions on ONE device (the multi-rank legs run in
tools/peer_check.py under torchrun and in tests/test_parallel_cpu.py on gloo):

  * ogs_lonlat_backward_view: views accumulate inside the per-Gaussian backward, statistics included; oracle = the sum
    of RasterizeGaussiansBackwardCUDA's per-view outputs (the "sum of the reference's per-view gradients" of §8(e-a));
  * ogs_sh_gradient_from_views: dL_dsh rebuilt from the views' dL/dRGB factors — bit-identical to adding the per-view
    rows in view order, equal to the dense dL_dsh up to the float-atomic noise of two replays;
  * the split forward (geometry + bin | colours + blend) is bit-identical to the one-call forward, and stage 2 may be
    repeated on the same buffers;
  * raw-parameter mode with SH degree 0 (M == 1, empty features_rest).
"""
import importlib

import numpy as np
import pytest
import torch

import _harness as h

pytestmark = pytest.mark.gpu
sm = h.scene_mod
par = importlib.import_module("omnigs-fork_b200.parallel")
tr = importlib.import_module("omnigs-fork_b200.trainer")


def bits(t):
    return t.contiguous().view(torch.int32)


def rel(a, b):
    return float((a - b).abs().max()) / (float(b.abs().max()) + 1e-30)


@pytest.fixture(scope="module")
def step():
    scene = sm.make_scene(60000, 800, 400, 0.02, 61, pole_frac=0.1, seam_frac=0.03)
    views = [sm.random_view(70 + v) for v in range(3)]
    d = h.torch_inputs(scene, views[0], bg=(0.1, 0.2, 0.3))
    dL = [torch.from_numpy(sm.make_grad_image(scene.W, scene.H, 80 + v)).cuda() for v in range(3)]
    return scene, views, d, dL


def _set_view(d, view):
    d["viewmatrix"] = torch.from_numpy(view[0]).cuda()
    d["campos"] = torch.from_numpy(view[1]).cuda()
    d["projmatrix"] = d["viewmatrix"]


def test_views_accumulate_in_the_backward_and_sh_is_rebuilt_from_factors(step):
    scene, views, d, dL = step
    P = scene.P
    dense, fwds = [], []
    for v, view in enumerate(views):
        _set_view(d, view)
        fwd = h.run_forward(h.pkg, d)
        fwds.append(fwd)
        dense.append([g.clone() for g in h.run_backward(h.pkg, d, fwd, dL[v])])
    bucket = par.GradientBucket(P, 16, "cuda", views_per_rank=len(views))
    bucket.flat.fill_(float("nan"))            # the first view must overwrite, not add
    bucket["dL_drgb"].zero_()
    m2d = []
    for v, view in enumerate(views):
        _set_view(d, view)
        fwd = fwds[v]
        m2d.append(h.pkg.RasterizeGaussiansBackwardView(
            d["background"], d["means3D"], fwd[2], d["scales"], d["rotations"], 1.0, d["viewmatrix"], dL[v], d["sh"], 3,
            d["campos"], fwd[3], fwd[0], fwd[4], fwd[5], bucket, v, want_means2D=True))
    campos = torch.stack([torch.from_numpy(v[1]) for v in views]).cuda()
    par.exchange_bucket(bucket, means3D=d["means3D"], campos_views=campos, degree=3)
    torch.cuda.synchronize()
    name = {n: i for i, n in enumerate(h.GRAD_NAMES)}
    # two replays of the same blend differ by the order of their float atomics only
    for n in ("dL_dmeans3D", "dL_dopacity"):
        want = sum(g[name[n]] for g in dense)
        assert rel(bucket[n], want) < 2e-5, (n, rel(bucket[n], want))
    for n in ("dL_dscales", "dL_drotations"):          # ill-conditioned: same bound as assert_own_runs_close
        want = sum(g[name[n]] for g in dense)
        assert rel(bucket[n], want) < 3e-4, (n, rel(bucket[n], want))
    want_sh = sum(g[name["dL_dsh"]] for g in dense)
    assert rel(bucket["dL_dsh"], want_sh) < 2e-5
    for v in range(len(views)):
        assert rel(m2d[v], dense[v][0]) < 2e-5
    # statistics: what ogs_densify_stats / ogs_view_stats derive per view, summed / maxed over the views
    radii = torch.stack([f[2] for f in fwds])
    assert torch.equal(bucket["denom"], (radii > 0).float().sum(dim=0))
    assert torch.equal(bucket.max_radii2D, radii.max(dim=0).values.float())
    gn = sum(torch.where(f[2] > 0, m[:, :2].norm(dim=-1), torch.zeros(P, device="cuda")) for f, m in zip(fwds, m2d))
    assert rel(bucket["xyz_gradient_accum"], gn) < 1e-5
    # culled everywhere -> exact zeros
    never = (radii > 0).sum(dim=0) == 0
    assert not bool(bucket["dL_dsh"][never].any()) and not bool(bucket["dL_dmeans3D"][never].any())

    # rebuild from factors: bit-identical to adding the per-view rows in view order
    f = [bucket["dL_drgb"][v] for v in range(len(views))]
    singles = []
    for v in range(len(views)):
        out = torch.empty((P, 16, 3), device="cuda")
        par.sh_gradient_from_views(d["means3D"], campos[v:v + 1], [f[v]], 3, out)
        singles.append(out)
    seq = singles[0] + singles[1]
    seq = seq + singles[2]
    assert torch.equal(bits(bucket["dL_dsh"]), bits(seq))
    # ... and the CUDA kernel agrees with the torch restatement of the SH weights
    chk = torch.zeros((P, 16, 3), dtype=torch.float64)
    par.sh_gradient_from_views(d["means3D"].double().cpu(), campos.double().cpu(), [x.double().cpu() for x in f], 3, chk)
    assert rel(bucket["dL_dsh"].double().cpu(), chk) < 1e-6


def test_sh_gradient_from_views_split_layout_and_degrees(step):
    """The raw-parameter trainer's layout (dL_dfeatures_dc + dL_dfeatures_rest) and lower SH degrees."""
    import ctypes
    scene, views, d, _ = step
    P = scene.P
    lib = h.pkg.load_library()
    g = torch.Generator(device="cuda").manual_seed(3)
    factors = [torch.randn((P, 3), device="cuda", generator=g) for _ in range(2)]
    campos = torch.stack([torch.from_numpy(v[1]) for v in views[:2]]).cuda()
    p = lambda t: ctypes.c_void_p(t.data_ptr())
    for deg in (0, 1, 2, 3):
        full = torch.empty((P, 16, 3), device="cuda")
        par.sh_gradient_from_views(d["means3D"], campos, factors, deg, full)
        dc, rest = torch.empty((P, 1, 3), device="cuda"), torch.empty((P, 15, 3), device="cuda")
        ptrs = (ctypes.c_void_p * 2)(*[f.data_ptr() for f in factors])
        assert lib.ogs_sh_gradient_from_views(P, deg, 16, 2, p(d["means3D"]), p(campos), ptrs, None, p(dc), p(rest), None) == 0
        assert torch.equal(bits(torch.cat([dc, rest], dim=1)), bits(full))
        assert not bool(full[:, (deg + 1) ** 2:].any())
        chk = torch.zeros((P, 16, 3), dtype=torch.float64)
        par.sh_gradient_from_views(d["means3D"].double().cpu(), campos.double().cpu(), [x.double().cpu() for x in factors], deg, chk)
        assert rel(full.double().cpu(), chk) < 1e-6


def test_split_forward_is_bit_identical_and_stage2_repeatable(step):
    scene, views, d, dL = step
    _set_view(d, views[1])
    one = h.run_forward(h.pkg, d)
    st = h.pkg.RasterizeGaussiansGeometry(d["means3D"], d["opacity"], d["scales"], d["rotations"], 1.0, d["cov3D_precomp"],
                                          d["viewmatrix"], d["campos"], scene.H, scene.W)
    two = h.pkg.RasterizeGaussiansBlend(st, d["background"], d["sh"], 3)
    assert two[0] == one[0] and torch.equal(two[2], one[2])
    assert torch.equal(bits(two[1]), bits(one[1]))
    a, b = h.ours_state(d, one), h.ours_state(d, two)
    for k in ("ranges", "point_list", "n_contrib", "clamped"):
        assert torch.equal(a[k], b[k]), k
    assert torch.equal(bits(a["rgb"]), bits(b["rgb"]))
    ga, gb = h.run_backward(h.pkg, d, one, dL[1]), h.run_backward(h.pkg, d, two, dL[1])
    for n, x, y in zip(h.GRAD_NAMES, ga, gb):
        assert rel(x, y) < (3e-4 if n in ("dL_dcov3D", "dL_dscales", "dL_drotations") else 2e-5), (n, rel(x, y))
    # stage 2 again on the same buffers (the scan's ticket / look-back words were only zeroed by stage 1)
    import ctypes
    lib = h.pkg.load_library()
    p = lambda t: ctypes.c_void_p(t.data_ptr())
    out = torch.empty_like(one[1])
    for _ in range(2):
        assert lib.ogs_lonlat_forward_stage2(scene.P, scene.W, scene.H, one[0], p(d["background"]), p(one[3]), p(one[4]),
                                            p(one[5]), p(out), None) == 0
        torch.cuda.synchronize()
        assert torch.equal(bits(out), bits(one[1]))


def test_raw_mode_with_sh_degree_zero_and_no_features_rest():
    """a degree-0 model has an empty features_rest_ (M == 1); the reference trains it through cat()."""
    scene = sm.make_scene(5000, 320, 160, 0.03, 33, sh_coeffs=1)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    op = np.clip(scene.opacities.astype(np.float64), 1e-4, 1 - 1e-4)
    pc = tr.GaussianModel(t(scene.means3D), t(scene.shs[:, :1, :]), t(scene.shs[:, 1:, :]),
                          t(np.log(op / (1 - op)).astype(np.float32)), t(np.log(scene.scales.astype(np.float64)).astype(np.float32)),
                          t(scene.rotations), sh_degree=0)
    assert pc.M == 1 and pc.features_rest_.numel() == 0
    view = sm.random_view(34)
    vm, cp = torch.from_numpy(view[0]).cuda(), torch.from_numpy(view[1]).cuda()
    bg = torch.zeros(3, device="cuda")
    img, radii, ctx = tr.render_lonlat_raw(pc, vm, cp, scene.H, scene.W, bg)
    d = h.torch_inputs(scene, view, degree=0)
    ref = h.run_forward(h.pkg, d)
    assert int((radii != ref[2]).sum()) <= 2 and float((img - ref[1]).abs().max()) < 1e-2
    dL = torch.from_numpy(sm.make_grad_image(scene.W, scene.H, 35)).cuda()
    _, grads = tr.backward_lonlat_raw(pc, ctx, dL)
    gr = h.run_backward(h.pkg, d, ref, dL)
    assert grads[2].numel() == 0
    assert rel(grads[1].view(-1, 3), gr[5].view(-1, 3)) < 1e-3      # f_dc gradient = dL_dsh[:, 0]
    assert rel(grads[0], gr[3]) < 1e-3
