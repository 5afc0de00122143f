"""CPU: host-side logic — scene generator, argument checks of the boundary mirror, band planner."""
import numpy as np
import pytest
import torch

import _harness as h

sm = h.scene_mod


def test_scene_generator_is_deterministic_and_well_formed():
    a = sm.make_scene(500, 64, 32, 0.02, 3, pole_frac=0.1, seam_frac=0.1)
    b = sm.make_scene(500, 64, 32, 0.02, 3, pole_frac=0.1, seam_frac=0.1)
    for f in ("means3D", "scales", "rotations", "opacities", "shs"):
        assert np.array_equal(getattr(a, f), getattr(b, f))
        assert getattr(a, f).dtype == np.float32
    assert np.allclose(np.linalg.norm(a.rotations, axis=1), 1.0, atol=1e-5)
    assert (a.opacities >= 0.05).all() and (a.opacities < 1.0).all()
    assert (a.scales > 0).all()
    r = np.linalg.norm(a.means3D, axis=1)
    assert (r < 0.2).sum() >= 0 and r.max() <= 20.5


def test_view_convention():
    V, c = sm.random_view(9)
    Tcw = V.T.astype(np.float64)            # viewmatrix is Tcw transposed
    R, t = Tcw[:3, :3], Tcw[:3, 3]
    assert np.allclose(R @ R.T, np.eye(3), atol=1e-6)
    assert np.allclose(-R.T @ t, c, atol=1e-6)   # campos = camera centre
    assert np.linalg.norm(c) <= 0.3 + 1e-6


def test_boundary_argument_checks_without_gpu():
    bad = torch.zeros((4, 2))
    e = torch.empty((0,))
    with pytest.raises(RuntimeError, match="means3D must have dimensions"):
        h.pkg.RasterizeGaussiansCUDA(e, bad, e, e, e, e, 1.0, e, e, e, 0.0, 0.0, 8, 8, e, 0, e, False, 3, False)
    ok = torch.zeros((4, 3))
    with pytest.raises(RuntimeError, match="CUDA tensor"):   # no CPU path, fails loudly
        h.pkg.RasterizeGaussiansCUDA(e, ok, e, e, e, e, 1.0, e, e, e, 0.0, 0.0, 8, 8, e, 0, e, False, 3, False)


def test_rasterizer_module_argument_rules():
    rs = h.pkg.GaussianRasterizationSettings(8, 8, 0.0, 0.0, torch.zeros(3), 1.0, torch.eye(4), torch.eye(4), 0,
                                             torch.zeros(3))
    r = h.pkg.GaussianRasterizer(rs)
    m = torch.zeros((2, 3))
    with pytest.raises(RuntimeError, match="excatly one of either SHs or precomputed colors"):
        r(m, m, torch.ones(2, 1), shs=None, colors_precomp=None, scales=m, rotations=torch.zeros(2, 4))
    with pytest.raises(RuntimeError, match="scale/rotation pair or precomputed 3D covariance"):
        r(m, m, torch.ones(2, 1), shs=torch.zeros(2, 1, 3), scales=m, rotations=torch.zeros(2, 4),
          cov3D_precomp=torch.zeros(2, 6))


def test_config_table_matches_baseline():
    assert sm.CONFIGS["C2"]["P"] == 1_000_000 and (sm.CONFIGS["C2"]["W"], sm.CONFIGS["C2"]["H"]) == (2048, 1024)
    assert sm.CONFIGS["C1"]["P"] == 100_000 and (sm.CONFIGS["C1"]["W"], sm.CONFIGS["C1"]["H"]) == (1024, 512)


def test_bench_algorithmic_bytes_match_the_survey_yardstick():
    """bench.py's roofline numerator is SURVEY.md section 8(d)'s fixed yardstick; its worked C2 example (V = P = 1e6, D = 3,
    M = 16, N = 2048 x 1024, T = 8192, K = 6, R = 2.3e7) gives 0.311 + (0.008 + 0.296 + 3.496 + 0.184) + 0.962 + 1.042 +
    0.559 = 6.86 GB."""
    import importlib.util
    import os
    spec = importlib.util.spec_from_file_location("bench", os.path.join(os.path.dirname(__file__), "..", "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    B = bench.alg_bytes(P=10**6, V=10**6, R=23 * 10**6, N=2048 * 1024, T=8192, D=3, M=16, K=6)
    gb = {k: v / 1e9 for k, v in B.items()}
    assert abs(gb["preprocess_fwd"] - 0.311) < 1e-3
    assert abs(gb["binning"] - (0.008 + 0.296 + 3.496 + 0.184)) < 2e-3
    assert abs(gb["render_fwd"] - 0.962) < 1e-3
    assert abs(gb["render_bwd"] - 1.042) < 1e-3
    assert abs(gb["preprocess_bwd"] - 0.559) < 1e-3
    assert abs(sum(gb.values()) - 6.86) < 0.01


def test_emit_slot_float_division_is_exact_with_its_correction():
    """binning.cu emit_slot(): q = local / w is evaluated as trunc(float_rz(local) * rcp.approx(float(w))) followed by a
    two-sided correction.  Restated in float32 numpy with the reciprocal off by -1 / 0 / +1 ulp (rcp.approx's error
    bound) on adversarial slots (remainders next to 0 and w - 1, local up to 2^31, q up to 2^16): always exact."""
    rng = np.random.default_rng(1)
    n = 1_500_000
    w = rng.integers(1, 65536, n).astype(np.int64)
    q = rng.integers(0, 65536, n).astype(np.int64)
    r = np.where(rng.random(n) < 0.5, rng.integers(0, 3, n), w - 1 - rng.integers(0, 3, n)).astype(np.int64)
    r = np.clip(r, 0, w - 1)
    local = q * w + r
    keep = local < (1 << 31)
    w, q, local = w[keep], q[keep], local[keep]

    def f32_rz(x):   # __uint2float_rz
        f = x.astype(np.float32)
        return np.where(f.astype(np.int64) > x, np.nextafter(f, np.float32(0)), f).astype(np.float32)

    lf, wf = f32_rz(local), f32_rz(w)
    rc = (np.float32(1) / wf).astype(np.float32)
    for rcp in (rc, np.nextafter(rc, np.float32(np.inf)), np.nextafter(rc, np.float32(0))):
        qq = np.floor((lf * rcp).astype(np.float32)).astype(np.int64)
        rem = local - qq * w
        qq = np.where(rem < 0, qq - 1, np.where(rem >= w, qq + 1, qq))
        rem = local - qq * w
        assert np.array_equal(qq, q) and bool(((rem >= 0) & (rem < w)).all())


def test_committed_ncu_counters_belong_to_the_committed_kernels():
    """profiles/kernel_counters.json (DRAM traffic, warp instructions, L2 reductions per launch: what bench.py's `roofline`
    block quotes as `traffic`) is captured under ncu and stamped with a hash of csrc/*.cu, *.cuh.  A kernel change without a
    fresh capture must not go unnoticed: bench.py withholds stale numbers, and this test fails."""
    import json
    import os
    import sys
    sys.path.insert(0, h.ROOT)
    import bench
    counters = json.load(open(os.path.join(h.ROOT, "profiles", "kernel_counters.json")))
    assert counters["source_stamp"] == bench.source_stamp(), "re-capture profiles/kernel_counters.json (tools/ncu_summary.py full ...)"
    for k in ("preprocess_fwd", "binning", "render_fwd", "render_bwd", "preprocess_bwd"):
        assert counters["kernels"][k]["dram_bytes"] > 0 and counters["kernels"][k]["warp_instructions"] > 0
    assert counters["kernels"]["render_bwd"]["l2_red_sectors"] > 0


def test_backward_staging_model_replays_every_reachable_position_once_in_descending_order():
    """tools/proto_bwd_batching.py mirrors the control flow of render_bwd_kernel's hit-byte path (chunked scan, queue, batches,
    remainder moves).  For random hit densities and list lengths around the chunk / batch boundaries the replayed positions
    are exactly the reachable ones, descending; batches are full except the last; the queue stays inside its capacity."""
    import os
    import sys
    sys.path.insert(0, os.path.join(h.ROOT, "tools"))
    from proto_bwd_batching import emulate
    rng = np.random.default_rng(7)
    chunk, batch = 256, 128
    lengths = [0, 1, 63, 127, 128, 129, 255, 256, 257, 383, 384, 385, 511, 512, 513, 1000, 4097]
    for n in lengths:
        for density in (0.0, 0.02, 0.26, 0.5, 0.97, 1.0):
            hits = (rng.random(n + 50) < density).astype(np.uint8) * rng.integers(1, 256, n + 50).astype(np.uint8)
            batches, high = emulate(hits, n, chunk, batch)
            want = [p for p in range(n - 1, -1, -1) if hits[p] != 0]
            got = [p for b in batches for p in b]
            assert got == want, (n, density)
            assert all(len(b) == batch for b in batches[:-1]) and all(0 < len(b) <= batch for b in batches)
            assert high < chunk + batch
