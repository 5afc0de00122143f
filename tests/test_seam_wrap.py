"""Longitude-seam wrap-around (opt-in, non-parity extension).

The reference clamps tile rects at the left/right image edge, so Gaussians straddling lon = +-pi are cut
(auxiliary.h:56-66, forward.cu:678-681).  No reference output exists for the wrap mode; its oracle is the
CPU restatement extended with the same rule, and the defining property is YAW EQUIVARIANCE: yawing the
camera by 2*pi*k/W must roll the panorama by k pixels."""
import math

import numpy as np
import pytest

import _harness as h
from oracle import oracle

sm = h.scene_mod


def seam_scene(W=256, H=128, P=1200, seed=31):
    # a third of the Gaussians sit within a tile of the seam
    return sm.make_scene(P, W, H, 0.04, seed, pole_frac=0.05, seam_frac=0.35, near_frac=0.0)


def oracle_image(scene, view, wrap):
    oracle.set_seam_wrap(wrap)
    try:
        f = oracle.forward(scene.means3D, scene.opacities, view[0], view[1], scene.W, scene.H, np.zeros(3, np.float32),
                           shs=scene.shs, degree=3, scales=scene.scales, rotations=scene.rotations)
    finally:
        oracle.set_seam_wrap(False)
    return f


def rolled_mismatch(img0, imgk, k):
    """fraction of pixels whose colour differs by more than tol between view k and the rolled view 0.
    A Gaussian is blended on every pixel of the tiles its 3-sigma square touches and nowhere else
    (reference behaviour), so its faint tail (alpha <= 0.011) depends on how the tile grid falls: a yaw by
    a multiple of 16 px relabels tiles and must agree tightly, any other yaw only up to those tails."""
    tol = 2e-3 if k % 16 == 0 else 3e-2
    d = np.abs(np.roll(img0, -k, axis=2) - imgk).max(axis=0)
    return float((d > tol).mean()), float(d.max())


def test_oracle_wrap_mode_is_yaw_equivariant_and_clamp_mode_is_not():
    scene = seam_scene()
    W = scene.W
    img0_wrap = oracle_image(scene, sm.yaw_view(0.0), True)["out_color"]
    img0_clamp = oracle_image(scene, sm.yaw_view(0.0), False)["out_color"]
    for k in (16, 37, 128):
        view = sm.yaw_view(2 * math.pi * k / W)
        frac_wrap, _ = rolled_mismatch(img0_wrap, oracle_image(scene, view, True)["out_color"], k)
        frac_clamp, _ = rolled_mismatch(img0_clamp, oracle_image(scene, view, False)["out_color"], k)
        assert frac_wrap < 2e-3, (k, frac_wrap)          # only isolated threshold flips (float yaw matrix)
        assert frac_clamp > 10 * max(frac_wrap, 1e-3), (k, frac_clamp, frac_wrap)   # the reference's seam artefact


def test_wrap_mode_only_adds_instances_near_the_seam():
    # small Gaussians only (radius << W/4, no polar ones): nothing can reach the middle half across the seam
    scene = sm.make_scene(1200, 256, 128, 0.008, 34, pole_frac=0.0, seam_frac=0.35, near_frac=0.0)
    lat = np.abs(np.arcsin(np.clip(scene.means3D[:, 1] / np.linalg.norm(scene.means3D, axis=1), -1, 1)))
    keep = lat < math.radians(60)
    for f in ("means3D", "scales", "rotations", "opacities", "shs"):
        setattr(scene, f, np.ascontiguousarray(getattr(scene, f)[keep]))
    a = oracle_image(scene, sm.identity_view(), False)
    b = oracle_image(scene, sm.identity_view(), True)
    assert b["num_rendered"] > a["num_rendered"]
    assert np.array_equal(a["radii"], b["radii"])
    # columns away from the seam are unchanged
    mid = slice(scene.W // 4, 3 * scene.W // 4)
    assert np.abs(a["out_color"][:, :, mid] - b["out_color"][:, :, mid]).max() < 1e-6


@pytest.mark.gpu
def test_gpu_wrap_mode_matches_oracle_and_is_yaw_equivariant():
    import torch
    scene = seam_scene(W=512, H=256, P=6000, seed=32)
    dL_np = sm.make_grad_image(scene.W, scene.H, 33)
    prev = h.pkg.set_seam_wrap(True)
    try:
        view = sm.identity_view()
        d = h.torch_inputs(scene, view)
        fwd = h.run_forward(h.pkg, d)
        grads = h.run_backward(h.pkg, d, fwd, torch.from_numpy(dL_np).cuda())
        of = oracle_image(scene, view, True)
        oracle.set_seam_wrap(True)
        og = oracle.backward(of, dL_np, scene.means3D, view[0], view[1], scene.W, scene.H, np.zeros(3, np.float32),
                             shs=scene.shs, degree=3, scales=scene.scales, rotations=scene.rotations)
        oracle.set_seam_wrap(False)
        assert abs(fwd[0] - of["num_rendered"]) <= max(4, of["num_rendered"] // 500)
        diff = np.abs(fwd[1].cpu().numpy() - of["out_color"]).max(axis=0)
        assert (diff > 1e-4).mean() <= 2e-3 and diff.max() < 0.1
        for n, g in zip(h.GRAD_NAMES, grads):
            ref = og[n].reshape(tuple(g.shape))
            scale = np.abs(ref).max() + 1e-30
            assert np.abs(g.cpu().numpy() - ref).max() / scale < 1e-2, n
        img0 = fwd[1].cpu().numpy()
        for k in (16, 100):
            dk = h.torch_inputs(scene, sm.yaw_view(2 * math.pi * k / scene.W))
            imgk = h.run_forward(h.pkg, dk)[1].cpu().numpy()
            frac, _ = rolled_mismatch(img0, imgk, k)
            assert frac < 2e-3, (k, frac)
        # a tile-aligned yaw relabels tiles only: instance count identical
        d16 = h.torch_inputs(scene, sm.yaw_view(2 * math.pi * 16 / scene.W))
        assert abs(h.run_forward(h.pkg, d16)[0] - fwd[0]) <= max(4, fwd[0] // 500)
    finally:
        h.pkg.set_seam_wrap(prev)
    # default (parity) mode is untouched by the option having been used
    base = h.run_forward(h.pkg, h.torch_inputs(scene, sm.identity_view()))
    assert base[0] < fwd[0]


@pytest.mark.gpu
def test_wrap_mode_rejects_widths_that_are_not_tile_aligned():
    scene = sm.make_scene(50, 250, 100, 0.05, 3)
    prev = h.pkg.set_seam_wrap(True)
    try:
        with pytest.raises(RuntimeError, match="multiple of 16"):
            h.run_forward(h.pkg, h.torch_inputs(scene, sm.identity_view()))
    finally:
        h.pkg.set_seam_wrap(prev)
