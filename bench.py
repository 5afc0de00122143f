#!/usr/bin/env python
"""bench.py — lonlat rasterizer forward+backward on a 360Roam-shaped scene (BASELINE.json configs[1]:
1M Gaussians, 2048x1024 equirect, SH degree 3), one view per rank per step.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Prints ONE JSON line (rank 0).  `value` = whole-job training views/s with every input resident in HBM
(fwd + bwd through the public RasterizeGaussiansCUDA / RasterizeGaussiansBackwardCUDA boundary, plus
the NCCL gradient all-reduce when N > 1); `ms_per_step` is the BASELINE "fwd+bwd ms/frame".
`e2e` adds, per step, the pinned host->device copy of the view (pose + target image) and the
device->host read of the loss.  `roofline_frame` carries the per-stage device times and `frame_ms`, the
per-frame distribution SURVEY.md 8(d) asks for (forward, backward, both: p10 / median / p90 over 50 frames).  Two more blocks
cover the multi-GPU configs of BASELINE.json in the same process group: `dp_views` (configs[2]: C3, 3M Gaussians,
1920x960, 8 views per step shared by the N ranks, strong scaling) and `bands` (configs[3]: C4, 5M Gaussians, 7680x3840,
latitude bands, strong scaling); `--no-extra` skips them.  `--impl reference` times the reference's own rasterizer
(oracle/_ref/omnigs_ref.so: its sources rebuilt for sm_100) through its own entry points on the same
scene; if that library is absent it falls back to the CPU oracle port.
"""
import argparse
import ctypes
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

# data-parallel step: colour kernel of step s+1 on the side stream behind the dL_dsh rebuild of step s (OGS_DP_COLOURS=main: on
# the main stream, after an event wait)
COLOURS_ON_SIDE = os.environ.get("OGS_DP_COLOURS", "side") != "main"
METRIC = "lonlat_fwd_bwd_train_views_per_s"
UNIT = "views/s"
WORKLOAD = "C2"


def alg_bytes(P, V, R, N, T, D, M, K):
    """SURVEY.md §8(d) / BASELINE.md §4: algorithmic bytes per frame and per stage."""
    return {
        "preprocess_fwd": P * (52 + 12 * (D + 1) ** 2) + V * 67,
        "binning": 8 * P + (8 * P + 12 * V + 12 * R) + R * (8 + 24 * K) + (8 * R + 8 * T),
        "render_fwd": 40 * R + 20 * N + 8 * T,
        "render_bwd": 40 * R + 20 * N + 80 * P,
        "preprocess_bwd": 4 * P + V * (107 + 12 * (D + 1) ** 2) + P * (64 + 12 * M),
    }


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed regions (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "20", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()

    def summary(self, windows):
        sm, mx, reasons = [], 0.0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for t, line in self.rows:
            if not any(a <= t <= b for a, b in windows):
                continue
            f = [x.strip() for x in line.split(",")]
            try:
                sm.append(float(f[0])); mx = max(mx, float(f[1]))
            except (ValueError, IndexError):
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx or None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_port_baseline(scene, view, dL_np, budget_s=45.0):
    """The oracle port (plain C + OpenMP) on this box's host cores: one full frame when that fits the
    budget, else BASELINE configs[0] (C1)."""
    from oracle import oracle
    import _harness as h
    bg = np.zeros(3, np.float32)

    def frame(sc, dl):
        t0 = time.time()
        f = oracle.forward(sc.means3D, sc.opacities, view[0], view[1], sc.W, sc.H, bg, shs=sc.shs, degree=3,
                           scales=sc.scales, rotations=sc.rotations)
        oracle.backward(f, dl, sc.means3D, view[0], view[1], sc.W, sc.H, bg, shs=sc.shs, degree=3,
                        scales=sc.scales, rotations=sc.rotations)
        return time.time() - t0

    c1 = h.scene_mod.make_config_scene("C1")
    c1_dl = h.scene_mod.make_grad_image(c1.W, c1.H, 99)
    t_c1 = frame(c1, c1_dl)
    # the same port on ONE host core, BASELINE configs[0]
    cores = oracle.num_threads()
    oracle.set_num_threads(1)
    t_c1_single = frame(c1, c1_dl)
    oracle.set_num_threads(cores)
    single = {"value": 1.0 / t_c1_single, "unit": UNIT, "cores": 1,
              "sample": f"1 frame of BASELINE configs[0] (C1: 100k Gaussians, 1024x512), {t_c1_single:.1f} s; the same frame on "
                        f"{cores} cores: {t_c1:.2f} s"}
    if t_c1 * 12 <= budget_s:
        t = frame(scene, dL_np)
        return {"value": 1.0 / t, "unit": UNIT, "cores": cores, "kind": "port",
                "sample": f"1 full {WORKLOAD} frame (fwd+bwd), {t:.1f} s", "single_thread": single}
    return {"value": 1.0 / t_c1, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"1 frame of BASELINE configs[0] (C1: 100k Gaussians, 1024x512), {t_c1:.1f} s; a {WORKLOAD} frame "
                      f"is ~11x larger", "single_thread": single}


def training_iteration_bench(h, scene, views, dev, K, Wm, ref_mod):
    """one whole training iteration per step (lr schedule, render from the stored tensors, L1 + SSIM
    loss, backward, densification statistics, Adam; loss read back) on the WORKLOAD scene.  ref_mod None: this
    package's fused calls (trainer.train_for_one_iteration).  ref_mod = the reference rasterizer library: the
    reference's composition (LibTorch activations / conv2d SSIM / autograd / torch.optim.Adam around its own
    rasterizer, gaussian_mapper.cpp:300-470).  Reported beside the BASELINE metric, not instead of it."""
    import torch
    from importlib import import_module
    tr = import_module("omnigs-fork_b200.trainer")
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    op = np.clip(scene.opacities.astype(np.float64), 1e-4, 1 - 1e-4)
    raw = [t(scene.means3D), t(scene.shs[:, :1, :]), t(scene.shs[:, 1:, :]), t(np.log(op / (1 - op)).astype(np.float32)),
           t(np.log(scene.scales.astype(np.float64)).astype(np.float32)), t(scene.rotations)]
    W, H = scene.W, scene.H
    gt = torch.rand((3, H, W), device=dev, generator=torch.Generator(device=dev).manual_seed(5))
    bg = torch.zeros(3, device=dev)
    view_dev = [(torch.from_numpy(v).to(dev), torch.from_numpy(c).to(dev)) for v, c in views]
    K = min(K, 50)
    if ref_mod is None:
        pc = tr.GaussianModel(*raw, sh_degree=3)

        def step(s):
            loss_out, _ = tr.train_for_one_iteration(pc, view_dev[s][0], view_dev[s][1], gt, bg, s + 1)
            return float(loss_out[0].item())
        what = "fused: 5 library calls per iteration (raw forward, loss fwd+bwd, raw backward, statistics, Adam)"
    else:
        from oracle import train_ref as ref   # reference arm only: the LibTorch composition restated
        leaves = [p.clone().requires_grad_(True) for p in raw]
        opt = tr.OptimizationParams()
        lrs = [opt.position_lr_init, opt.feature_lr, opt.feature_lr / 20.0, opt.opacity_lr, opt.scaling_lr, opt.rotation_lr]
        adam = ref.make_adam(leaves, lrs)
        P = scene.P
        stats = [torch.zeros(P, device=dev), torch.zeros((P, 1), device=dev), torch.zeros((P, 1), device=dev)]
        rasterize = h.make_autograd_rasterizer(ref_mod)

        def step(s):
            adam.param_groups[0]["lr"] = tr.expon_lr(s + 1, opt.position_lr_init, opt.position_lr_final, 0,
                                                     opt.position_lr_delay_mult, opt.position_lr_max_steps)
            return h.reference_training_iteration(rasterize, ref, leaves, adam, stats, view_dev[s][0], view_dev[s][1],
                                                  gt, bg, opt.lambda_dssim)
        what = "reference composition: LibTorch-op activations, conv2d SSIM, autograd, torch.optim.Adam around the reference rasterizer"
    for s in range(Wm):
        step(s)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for s in range(K):
        loss = step(Wm + s)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / K
    return {"ms_per_iteration": ms, "views_per_s": 1000.0 / ms, "steps": K, "final_loss": loss, "what": what}


# the kernels profiles/kernel_counters.json describes (one frame: per-Gaussian forward / backward, binning, both blends)
# live in these files and in the headers; the C ABI layer, the training-step and the collective kernels are not in it
FRAME_KERNEL_SOURCES = ("binning.cu", "preprocess_fwd.cu", "preprocess_bwd.cu", "render_fwd.cu", "render_bwd.cu")


def source_stamp():
    """Hash of the frame kernels' sources: profiles/kernel_counters.json (captured under ncu) carries it, so a stale capture
    is detectable (tests/test_host_logic.py fails, the roofline block withholds the numbers)."""
    import hashlib
    hh = hashlib.sha256()
    d = os.path.join(ROOT, "omnigs-fork_b200", "csrc")
    for f in sorted(os.listdir(d)):
        if f in FRAME_KERNEL_SOURCES or f.endswith(".cuh"):
            hh.update(f.encode())
            hh.update(open(os.path.join(d, f), "rb").read())
    return hh.hexdigest()[:16]


def exchange_check(par, bucket, means3D, campos_views, side, dev, rank, world):
    """Outside every timed region: the data-parallel exchange on seeded per-rank data, (a) overlapped (dL_dsh rebuild on the
    side stream) and (b) sequential — the two schedules must give the same bits on this rank —, (c) against NCCL's all-reduce
    of the same data (another summation order: close, not equal) and (d) across ranks (every replica must hold the same
    bits: checked with a MAX / MIN all-reduce of a checksum).  Returns a small report for the JSON line."""
    import torch
    import torch.distributed as dist
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    data = torch.randn(bucket.flat.shape, device=dev, generator=g)
    data[bucket.count_sum:bucket.count_sum + bucket.count_max].abs_()          # the radii section is non-negative
    results = []
    for overlapped in (True, False):
        bucket.flat.copy_(data)
        ev = par.exchange_bucket(bucket, means3D=means3D, campos_views=campos_views, degree=3, sh_stream=side if overlapped else None)
        if ev is not None:
            torch.cuda.current_stream().wait_event(ev)
        torch.cuda.synchronize()
        dist.barrier()
        results.append((bucket.flat[:bucket.count_sum + bucket.count_max].clone(), bucket["dL_dsh"].clone()))
    same_schedule = bool(torch.equal(results[0][0], results[1][0]) and torch.equal(results[0][1], results[1][1]))
    ref_sum = data[:bucket.count_sum].clone()
    dist.all_reduce(ref_sum, op=dist.ReduceOp.SUM)
    ref_max = data[bucket.count_sum:bucket.count_sum + bucket.count_max].clone()
    dist.all_reduce(ref_max, op=dist.ReduceOp.MAX)
    scale = float(ref_sum.abs().max())
    sum_err = float((results[0][0][:bucket.count_sum] - ref_sum).abs().max()) / scale
    max_equal = bool(torch.equal(results[0][0][bucket.count_sum:], ref_max))
    # identical bits on every replica: min and max over ranks of an integer checksum agree
    def checksum(t):
        return t.contiguous().view(torch.int32).to(torch.int64).sum().reshape(1)
    cs = torch.cat([checksum(results[0][0]), checksum(results[0][1])])
    lo, hi = cs.clone(), cs.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    return {"overlapped_equals_sequential_bitwise": same_schedule, "replicas_bitwise_equal": bool(torch.equal(lo, hi)),
            "max_rel_difference_from_nccl_sum": sum_err, "radii_max_equals_nccl": max_equal,
            "transport": "peer/multimem kernels" if bucket.peer else "NCCL"}


def dp_views_block(h, par, dev, rank, world, distributed, steps=8, warmup=3, config="C3", views=8):
    """BASELINE configs[2]: one training step = 8 views of the C3 scene shared by the ranks (rank g renders views g, g+N, ...),
    every view accumulated into the step's factored bucket inside the per-Gaussian backward, ONE exchange per step.  Strong
    scaling: the step is the same work at every N."""
    import torch
    import torch.distributed as dist
    sm = h.scene_mod
    scene = sm.make_config_scene(config)
    P = scene.P
    d = h.torch_inputs(scene, sm.random_view(0), device=dev)
    dL = torch.from_numpy(sm.make_grad_image(scene.W, scene.H, 99)).to(dev)
    mine = par.views_for_rank(views, rank, world)
    vpr = -(-views // world)
    poses = [[sm.random_view(5000 + 31 * s + v) for v in range(views)] for s in range(steps + warmup)]
    campos_all = [torch.from_numpy(np.stack([c for _, c in st] + [np.zeros(3, np.float32)] * (vpr * world - views))).to(dev) for st in poses]
    poses = [[(torch.from_numpy(a).to(dev), torch.from_numpy(c).to(dev)) for a, c in st] for st in poses]
    bucket = par.GradientBucket(P, 16, dev, views_per_rank=vpr)
    side = torch.cuda.Stream()
    pending = [None]
    cur = torch.cuda.current_stream()

    def step(s):
        if not mine:
            bucket.zero_step()
        for k, v in enumerate(mine):
            vm, cp = poses[s][v]
            if k == 0:
                # the previous step's SH gradients may still be in flight: only the colours wait for them
                st = h.pkg.RasterizeGaussiansGeometry(d["means3D"], d["opacity"], d["scales"], d["rotations"], 1.0,
                                                      d["cov3D_precomp"], vm, cp, scene.H, scene.W)
                # the colour kernel runs on the exchange's side stream, behind the dL_dsh rebuild it depends on and
                # underneath this step's sorts; the blend waits for it
                if distributed and COLOURS_ON_SIDE:
                    fwd = h.pkg.RasterizeGaussiansBlend(st, d["background"], d["sh"], 3, colors_stream=side)
                else:
                    if pending[0] is not None:
                        cur.wait_event(pending[0])
                    fwd = h.pkg.RasterizeGaussiansBlend(st, d["background"], d["sh"], 3)
            else:
                d["viewmatrix"], d["campos"], d["projmatrix"] = vm, cp, vm
                fwd = h.run_forward(h.pkg, d)
            h.pkg.RasterizeGaussiansBackwardView(d["background"], d["means3D"], fwd[2], d["scales"], d["rotations"], 1.0, vm,
                                                 dL, d["sh"], 3, cp, fwd[3], fwd[0], fwd[4], fwd[5], bucket, k)
        if not mine and pending[0] is not None:
            cur.wait_event(pending[0])
        pending[0] = par.exchange_bucket(bucket, means3D=d["means3D"], campos_views=campos_all[s], degree=3,
                                         sh_stream=side if distributed else None)

    def drain():
        if pending[0] is not None:
            cur.wait_event(pending[0])
            pending[0] = None

    for s in range(warmup):
        step(s)
    drain()
    torch.cuda.synchronize()
    if distributed:
        dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for s in range(steps):
        step(warmup + s)
    drain()
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / steps], device=dev, dtype=torch.float64)
    if distributed:
        dist.barrier()
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    out = {"metric": "lonlat_dp_train_views_per_s", "value": views / (float(ms) / 1e3), "unit": "views/s", "ms_per_step": float(ms),
           "n_gpus": world, "scaling": "strong", "steps": steps, "warmup": warmup,
           "config": {"workload": config, "gaussians": P, "image": [scene.W, scene.H], "views_per_step": views,
                      "views_per_rank": len(mine),
                      "exchange": ("factored: all-reduce of 14 floats/Gaussian + dL/dRGB factors read over NVLink by the dL_dsh "
                                   "rebuild kernel" + (" (peer/multimem kernels)" if bucket.peer else " (NCCL all-reduce + all-gather)"))
                                  if distributed else "none (views accumulate in the per-Gaussian backward)"}}
    del bucket, d, dL, poses, campos_all
    torch.cuda.empty_cache()
    return out


def bands_block(h, par, dev, rank, world, distributed, steps=6, warmup=2, config="C4"):
    """BASELINE configs[3]: one C4 panorama (5M Gaussians, 7680x3840) rendered and differentiated in latitude bands: every
    rank holds all Gaussians and bins / sorts / blends only its tile rows (balanced on the per-row instance counts of a
    full-frame forward), band rows are all-gathered, the [P,12] accumulators are summed between the two backward kernels."""
    import torch
    import torch.distributed as dist
    sm = h.scene_mod
    scene = sm.make_config_scene(config)
    P = scene.P
    d = h.torch_inputs(scene, sm.identity_view(), device=dev)
    dL = torch.from_numpy(sm.make_grad_image(scene.W, scene.H, 99)).to(dev)

    def rasterize(band):
        return h.pkg.RasterizeGaussiansCUDA(d["background"], d["means3D"], d["colors"], d["opacity"], d["scales"], d["rotations"],
                                            1.0, d["cov3D_precomp"], d["viewmatrix"], d["projmatrix"], 0.0, 0.0, d["H"], d["W"],
                                            d["sh"], 3, d["campos"], False, 3, False, band=band)
    full = rasterize(None)            # "previous frame": its per-row loads balance the bands
    R_full = full[0]
    prev = h.pkg.export_forward_state(P, scene.W, scene.H, R_full, full[3], full[4], full[5], want_keys=False)
    rows = par.tile_row_counts(prev["ranges"], scene.W, scene.H)
    # bands are balanced on estimated cost (instances for the binning, visited list entries for the blend);
    # OGS_BAND_BALANCE=instances balances on instances alone
    costs = rows if os.environ.get("OGS_BAND_BALANCE", "cost") == "instances" else \
        par.tile_row_costs(prev["ranges"], prev["n_contrib"], scene.W, scene.H)
    del full, prev
    torch.cuda.empty_cache()
    bands = par.band_rows(costs, world)
    band = bands[rank]
    loads = [sum(rows[a:b]) for a, b in bands]
    cost_share = [sum(costs[a:b]) / sum(costs) for a, b in bands]
    ex = par.BandExchange(P, scene.W, scene.H, dev) if distributed else None

    # OGS_BAND_GATHER=full all-gathers whole bands (every rank ends with the whole frame); the default exchanges the 5-row
    # halos a band-wise L1 + SSIM loss needs (11-tap window).  OGS_BAND_CHUNKS > 1 pipelines the accumulator exchange with the
    # per-Gaussian backward in Gaussian ranges (measured on four B200s: 5.61 ms with four ranges, 5.55 ms with one — the all-reduce
    # kernel fills the SMs, so the ranges do not overlap usefully; profiles/r02_bands_trace_n4.log).
    halo = None if os.environ.get("OGS_BAND_GATHER", "halo") == "full" else 5
    chunks = int(os.environ.get("OGS_BAND_CHUNKS", "1"))

    def step():
        img, fwd = par.render_band_forward(rasterize, band, scene.H, exchange=ex, halo=halo)
        if ex is not None:
            h.run_backward(h.pkg, d, fwd, dL, accumulators=ex.acc, reduce_accumulators=ex.reduce_accumulators,
                           accumulator_chunks=chunks)
        else:
            h.run_backward(h.pkg, d, fwd, dL)
        return fwd[0]

    # feedback balancing (a trainer does this as it goes): three calibration frames, each followed by a re-cut of the bands
    # from the ranks' measured band-dependent kernel times (the library's own stage events)
    lib = h.pkg.load_library()
    band_ms_history = []
    if distributed and os.environ.get("OGS_BAND_FEEDBACK", "1") != "0":
        buf8 = (ctypes.c_float * 8)()
        for _ in range(3):
            lib.ogs_profile_enable(1)
            step()
            lib.ogs_profile_read(buf8, 8)
            lib.ogs_profile_enable(0)
            mine_ms = float(buf8[3] + buf8[4] + buf8[5] + buf8[6])     # emit, tile_sort, render_fwd, render_bwd
            got = [None] * world
            dist.all_gather_object(got, mine_ms)
            band_ms_history.append(got)
            costs = par.rescale_row_costs(costs, bands, got)
            bands = par.band_rows(costs, world)
            band = bands[rank]
        loads = [sum(rows[a:b]) for a, b in bands]
        cost_share = [sum(costs[a:b]) / sum(costs) for a, b in bands]

    for _ in range(warmup):
        Rb = step()
    torch.cuda.synchronize()
    if distributed:
        dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / steps], device=dev, dtype=torch.float64)
    if distributed:
        dist.barrier()
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    # where the frame goes on rank 0: the library's per-stage events (3 frames) and the forward / backward split
    lib.ogs_profile_enable(1)
    acc8, buf8 = np.zeros(8), (ctypes.c_float * 8)()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    fwd_ms = bwd_ms = 0.0
    for _ in range(3):
        ev[0].record()
        img, fwd = par.render_band_forward(rasterize, band, scene.H, exchange=ex, halo=halo)
        ev[1].record()
        if ex is not None:
            h.run_backward(h.pkg, d, fwd, dL, accumulators=ex.acc, reduce_accumulators=ex.reduce_accumulators,
                           accumulator_chunks=chunks)
        else:
            h.run_backward(h.pkg, d, fwd, dL)
        ev[2].record()
        lib.ogs_profile_read(buf8, 8)
        acc8 += np.array(list(buf8))
        torch.cuda.synchronize()
        fwd_ms += ev[0].elapsed_time(ev[1]) / 3
        bwd_ms += ev[1].elapsed_time(ev[2]) / 3
    lib.ogs_profile_enable(0)
    # OGS_BANDS_TRACE=1: every rank's phase timeline (microseconds between successive marks of the exchange code, one frame)
    trace_all = None
    if os.environ.get("OGS_BANDS_TRACE") and ex is not None:
        frames = []
        torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
        for _ in range(5):          # back to back like the timed loop; the last frame is reported
            marks = []
            par._trace = marks
            par._mark("frame_start", torch.cuda.current_stream())
            step()
            par._mark("frame_end", torch.cuda.current_stream())
            frames.append(marks)
        par._trace = None
        torch.cuda.synchronize()
        t0 = frames[-1][0][1]
        mine_t = {lab: round(1e3 * t0.elapsed_time(evm)) for lab, evm in frames[-1][1:]}
        trace_all = [None] * world
        dist.all_gather_object(trace_all, mine_t)
    stage_names = ["preprocess_fwd", "depth_order", "tile_ranges", "emit", "tile_sort", "render_fwd", "render_bwd", "preprocess_bwd"]
    rank0 = {"forward_ms": fwd_ms, "backward_ms": bwd_ms, "band_instances": int(fwd[0]),
             "stage_ms": {n: float(v) / 3 for n, v in zip(stage_names, acc8)}}
    if trace_all is not None:
        rank0["timeline_us_since_frame_start_per_rank"] = trace_all
    out = {"metric": "lonlat_band_parallel_ms_per_frame", "value": float(ms), "unit": "ms", "n_gpus": world, "rank0": rank0,
           "higher_is_better": False, "scaling": "strong", "steps": steps, "warmup": warmup,
           "config": {"workload": config, "gaussians": P, "image": [scene.W, scene.H], "num_rendered": R_full, "bands": bands,
                      "band_instances": loads, "max_over_mean_band_load": max(loads) / (sum(loads) / world),
                      "band_cost_share": cost_share, "band_kernel_ms_during_calibration": band_ms_history,
                      "exchange": ((("5-row halos of the band" if halo else "all-gather of band rows") + " (peer stores) + all-reduce of the "
                                     f"[P,12] accumulators in {chunks} ranges pipelined with the per-Gaussian backward (peer / multimem kernel)"
                                     if ex.peer else "NCCL all_gather of band rows / halos + all_reduce of the [P,12] accumulators")
                                   if distributed else "none")}}
    del ex, d, dL
    torch.cuda.empty_cache()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the dp_views (C3) and bands (C4) blocks")
    args = ap.parse_args()
    K, Wm = args.steps, max(3, args.warmup)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference" and rank != 0:
        return 0  # the reference has no multi-GPU path: rank 0 alone runs it

    import torch
    import torch.distributed as dist
    import _harness as h
    sm = h.scene_mod
    from importlib import import_module
    par = import_module("omnigs-fork_b200.parallel")

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    distributed = world > 1 and args.impl == "ours"
    if distributed:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # NCCL prints its version banner on STDOUT at NCCL_DEBUG=VERSION; stdout carries the one JSON line only
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"
        dist.init_process_group("nccl", device_id=dev)

    scene = sm.make_config_scene(WORKLOAD)
    P, W, H, D, M = scene.P, scene.W, scene.H, 3, 16
    N, T = W * H, ((W + 15) // 16) * ((H + 15) // 16)
    views = [sm.random_view(1000 + 97 * s + rank) for s in range(K + Wm)]
    d = h.torch_inputs(scene, views[0], device=dev)
    view_dev = [(torch.from_numpy(v).to(dev), torch.from_numpy(c).to(dev)) for v, c in views]
    dL_np = sm.make_grad_image(W, H, 99)
    dL = torch.from_numpy(dL_np).to(dev)

    mod = h.pkg
    impl_note = None
    if args.impl == "reference":
        mod = h.load_reference()
        if mod is None:
            # the reference rasterizer library did not travel: time the CPU oracle port instead
            cb = cpu_port_baseline(scene, views[0], dL_np)
            line = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": 0,
                    "steps": 1, "warmup": 0, "ms_per_step": 1000.0 / cb["value"], "higher_is_better": True,
                    "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                    "config": {"workload": WORKLOAD, "gaussians": P, "image": [W, H], "sh_degree": D},
                    "cpu_baseline": cb,
                    "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
            print(json.dumps(line))
            return 0
        impl_note = "reference CUDA rasterizer rebuilt for sm_100 (oracle/_ref), its own entry points"

    names = h.GRAD_NAMES
    # data parallel (one view per rank and step): the per-Gaussian backward writes the four geometry gradients, the
    # densification statistics and the view's dL/dRGB factor into a factored bucket; ONE kernel all-reduces the 14 floats per
    # Gaussian, and the dL_dsh rebuild (which reads the peers' 3-float factors over NVLink) runs on a side stream underneath
    # the next step's geometry / depth order / tile sort (started right behind the all-reduce; later starts and a capped
    # grid measured slower, profiles/r02_dp_rebuild_sweep_n4.log) — only the colours of the next step wait for it, as they
    # would wait for Adam on the SH coefficients in a trainer.  OGS_DP_EXCHANGE=nccl forces the NCCL transport, =dense the
    # round-1 exchange (244 B/Gaussian in one all-reduce, nothing overlapped).
    dp_mode = os.environ.get("OGS_DP_EXCHANGE", "peer")
    bucket = None
    if distributed:
        bucket = par.GradientBucket(P, M, dev, peer=None if dp_mode in ("peer", "dense") else False,
                                    views_per_rank=0 if dp_mode == "dense" else 1)
    # every rank knows every rank's pose of a step (the view schedule of a trainer is deterministic)
    campos_all = [torch.from_numpy(np.stack([sm.random_view(1000 + 97 * s + r)[1] for r in range(world)])).to(dev)
                  for s in range(K + Wm)] if distributed else None
    side = torch.cuda.Stream() if distributed else None
    sh_pending = [None]

    def dp_step(s, dL_img, vm, cp):
        """forward + backward + exchange of one data-parallel step; dL_img: upstream gradient or a callable(image)."""
        cur = torch.cuda.current_stream()
        if bucket.factored:
            st = h.pkg.RasterizeGaussiansGeometry(d["means3D"], d["opacity"], d["scales"], d["rotations"], 1.0,
                                                  d["cov3D_precomp"], vm, cp, H, W)
            if COLOURS_ON_SIDE:
                # the colour kernel runs on the exchange's side stream, behind the dL_dsh rebuild it depends on and
                # underneath this step's sorts; the blend waits for it
                fwd = h.pkg.RasterizeGaussiansBlend(st, d["background"], d["sh"], 3, colors_stream=side)
            else:
                if sh_pending[0] is not None:
                    cur.wait_event(sh_pending[0])
                fwd = h.pkg.RasterizeGaussiansBlend(st, d["background"], d["sh"], 3)
            g_img = dL_img(fwd[1]) if callable(dL_img) else dL_img
            m2d = h.pkg.RasterizeGaussiansBackwardView(d["background"], d["means3D"], fwd[2], d["scales"], d["rotations"], 1.0,
                                                       vm, g_img, d["sh"], 3, cp, fwd[3], fwd[0], fwd[4], fwd[5], bucket, 0)
            sh_pending[0] = par.exchange_bucket(bucket, means3D=d["means3D"], campos_views=campos_all[s], degree=3, sh_stream=side)
            return fwd, m2d
        d["viewmatrix"], d["campos"], d["projmatrix"] = vm, cp, vm
        fwd = h.run_forward(mod, d)
        g_img = dL_img(fwd[1]) if callable(dL_img) else dL_img
        g = h.run_backward(mod, d, fwd, g_img, out=bucket)
        par.allreduce_bucket(bucket, g[0], fwd[2])
        return fwd, g

    def dp_drain():
        if sh_pending[0] is not None:
            torch.cuda.current_stream().wait_event(sh_pending[0])
            sh_pending[0] = None

    def step(s):
        if distributed:
            return dp_step(s, dL, *view_dev[s])
        d["viewmatrix"], d["campos"] = view_dev[s]
        d["projmatrix"] = d["viewmatrix"]
        fwd = h.run_forward(mod, d)
        g = h.run_backward(mod, d, fwd, dL)
        return fwd, g

    def sync_all():
        torch.cuda.synchronize()
        if distributed:
            dist.barrier()
            torch.cuda.synchronize()

    def max_over_ranks(ms):
        if not distributed:
            return ms
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    vis_env = os.environ.get("CUDA_VISIBLE_DEVICES")
    clocks = ClockSampler(vis_env.split(",")[local_rank] if vis_env else local_rank)
    clocks.start()
    windows = []

    # ---------------- device-resident timed region ----------------
    for s in range(Wm):
        fwd, _ = step(s)
    if distributed:
        dp_drain()
    R = fwd[0]
    V = int((fwd[2] > 0).sum())
    sync_all()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_a = time.time()
    e0.record()
    for s in range(K):
        step(Wm + s)
    if distributed:
        dp_drain()          # the last step's dL_dsh is part of the timed work
    e1.record()
    sync_all()
    windows.append((t_a, time.time()))
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    ms_per_step = ms_total / K
    value = world * K / (ms_total / 1000.0) if args.impl == "ours" else K / (ms_total / 1000.0)

    # ---------------- end-to-end: pinned H2D of the view + target, loss read back ----------------
    gt_host = torch.from_numpy(np.clip(dL_np * (W * H) * 0.1 + 0.5, 0, 1).astype(np.float32)).pin_memory()
    view_host = [torch.from_numpy(np.concatenate([v.reshape(-1), c.reshape(-1)])).pin_memory() for v, c in views]
    gt_dev = torch.empty_like(gt_host, device=dev)
    vbuf = torch.empty(19, device=dev)
    h2d = gt_host.numel() * 4 + 19 * 4

    copy_stream = torch.cuda.Stream()
    gt_ready = torch.cuda.Event()
    tr = import_module("omnigs-fork_b200.trainer")
    loss_host = torch.empty(3, dtype=torch.float32).pin_memory()
    loss_ready = torch.cuda.Event()
    loss_pending = [False]

    # ours: the trainer's prefetcher — the pose of a step is copied at its start; the 25 MB target image of step s+1 is
    # copied into the other of two device buffers while step s runs (issued once the loss of step s, the last reader of
    # the previous occupant's sibling, is queued), so one pose and one target cross PCIe every step inside the timed region
    gt_bufs = [gt_dev, torch.empty_like(gt_dev)]
    gt_events = [torch.cuda.Event(), torch.cuda.Event()]
    loss_queued = torch.cuda.Event()
    e2e_count = [0]

    def prefetch_target(slot):
        with torch.cuda.stream(copy_stream):
            gt_bufs[slot].copy_(gt_host, non_blocking=True)
            gt_events[slot].record()

    def e2e_step(s):
        if args.impl == "ours":
            k = e2e_count[0]
            e2e_count[0] += 1
            cur, nxt = k & 1, (k + 1) & 1
            if k == 0:   # very first step (warm-up): nothing was prefetched yet
                copy_stream.wait_stream(torch.cuda.current_stream())
                prefetch_target(cur)
            vbuf.copy_(view_host[s], non_blocking=True)
            d["viewmatrix"], d["campos"] = vbuf[:16].view(4, 4), vbuf[16:19]
            d["projmatrix"] = d["viewmatrix"]
            loss_box = [None]

            def loss_and_gradient(image):
                torch.cuda.current_stream().wait_event(gt_events[cur])
                # L1 loss and its gradient in one library pass (ogs_photometric_loss, lambda = 0)
                loss_box[0], dL_dimg = tr.photometric_loss(image, gt_bufs[cur], 0.0)
                loss_queued.record()
                copy_stream.wait_event(loss_queued)   # the other buffer's last reader (the previous step's loss) is behind this point
                prefetch_target(nxt)
                return dL_dimg
            if distributed:
                dp_step(s, loss_and_gradient, d["viewmatrix"], d["campos"])
            else:
                fwd = h.run_forward(mod, d)
                g = h.run_backward(mod, d, fwd, loss_and_gradient(fwd[1]))
            loss_out = loss_box[0]
            # device->host read of the step's result: an asynchronous copy into pinned memory every step; the host
            # consumes it one step later (a trainer's logging), so the read never drains the queue
            prev = None
            if loss_pending[0]:
                loss_ready.synchronize()
                prev = float(loss_host[0])
            loss_host.copy_(loss_out, non_blocking=True)
            loss_ready.record()
            loss_pending[0] = True
            return prev
        # reference arm: the 25 MB target image is only needed by the loss, so its copy runs on a side stream underneath
        # the forward and is waited for before the loss
        vbuf.copy_(view_host[s], non_blocking=True)
        copy_stream.wait_stream(torch.cuda.current_stream())   # the previous step's loss has consumed gt_dev
        with torch.cuda.stream(copy_stream):
            gt_dev.copy_(gt_host, non_blocking=True)
            gt_ready.record()
        d["viewmatrix"], d["campos"] = vbuf[:16].view(4, 4), vbuf[16:19]
        d["projmatrix"] = d["viewmatrix"]
        fwd = h.run_forward(mod, d)
        torch.cuda.current_stream().wait_event(gt_ready)
        diff = fwd[1] - gt_dev
        loss = diff.abs().mean()                       # L1 (the reference's main loss term)
        if distributed:
            g = h.run_backward(mod, d, fwd, torch.sign(diff) / diff.numel(), out=bucket)
            par.allreduce_bucket(bucket, g[0], fwd[2])
        else:
            g = h.run_backward(mod, d, fwd, torch.sign(diff) / diff.numel())
        return float(loss.item())                      # device->host read of the step's result

    for s in range(Wm):
        e2e_step(s)
    if distributed:
        dp_drain()
    sync_all()
    t_a = time.time()
    e0.record()
    for s in range(K):
        e2e_step(Wm + s)
    if distributed and args.impl == "ours":
        dp_drain()
    if loss_pending[0]:
        loss_ready.synchronize()   # the last step's loss
        final_e2e_loss = float(loss_host[0])
    e1.record()
    sync_all()
    windows.append((t_a, time.time()))
    ms_e2e = max_over_ranks(e0.elapsed_time(e1))
    e2e_value = (world if args.impl == "ours" else 1) * K / (ms_e2e / 1000.0)

    clocks.stop()
    # BASELINE configs[2] and [3] in the same process group (every rank takes part; rank 0 reports)
    extra = {}
    if distributed and bucket.factored:
        extra["exchange_check"] = exchange_check(par, bucket, d["means3D"], campos_all[0], side, dev, rank, world)
    if args.impl == "ours" and not args.no_extra:
        keep = (d, dL, gt_bufs, gt_dev)      # the C2 state stays alive for the per-stage section below
        extra["dp_views"] = dp_views_block(h, par, dev, rank, world, distributed)
        extra["bands"] = bands_block(h, par, dev, rank, world, distributed)
    if rank != 0:
        if distributed:
            dist.destroy_process_group()
        return 0

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world if args.impl == "ours" else 1,
        "steps": K, "warmup": Wm, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "gaussians": P, "visible": V, "image": [W, H], "sh_degree": D,
                   "num_rendered": R, "views_per_step_per_gpu": 1,
                   "parallelism": f"dp{world}" if distributed else "single",
                   "gradient_exchange": (("factored: " if bucket.factored else "dense 244 B/Gaussian: ")
                                         + ("own NVLink kernels (peer loads/stores up to 4 ranks, multimem beyond); dL_dsh rebuilt from "
                                            "the ranks' dL/dRGB factors on a side stream under the next step's geometry + sort"
                                            if bucket.peer else "NCCL all-reduce (+ all-gather of the factors)")) if distributed else "none",
                   "l2": "no flush: every step touches > 1 GB (params 236 MB, lists, accumulators), L2 is 126 MB; "
                         "a different camera pose each step"},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 12 if args.impl == "ours" else 4,
                "ms_per_step": ms_e2e / K,
                "what": "per step: pinned H2D of the pose, fwd, L1 loss + gradient (library pass), bwd, pinned H2D of the next "
                        "step's target image into the other of two device buffers (prefetch under the backward), loss "
                        "{loss, L1, SSIM} copied to pinned host memory and read by the host one step later"},
        "clocks": clocks.summary(windows),
    }
    if args.impl == "reference":
        line["impl"] = "reference"
        line["note"] = impl_note
        line["cpu_baseline"] = {"value": value, "unit": UNIT, "cores": 0, "kind": "reference", "device": "cuda",
                                "sample": f"{K} full {WORKLOAD} frames; the reference ships no CPU rasterizer, its CUDA "
                                          "kernels (rebuilt -arch=sm_100) are the reference implementation of the path"}
        line["e2e"] = {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
        line["train_step"] = training_iteration_bench(h, scene, views, dev, K, Wm, mod)
        print(json.dumps(line))
        return 0

    # ---------------- per-stage device times + roofline (ours) ----------------
    lib = h.pkg.load_library()
    stage_names = ["preprocess_fwd", "depth_order", "tile_ranges", "emit", "tile_sort", "render_fwd", "render_bwd",
                   "preprocess_bwd"]
    acc = np.zeros(8)
    reps = min(K, 10)
    lib.ogs_profile_enable(1)
    buf = (ctypes.c_float * 8)()
    for s in range(reps):
        d["viewmatrix"], d["campos"] = view_dev[Wm + s]
        fwd = h.run_forward(mod, d)
        h.run_backward(mod, d, fwd, dL)
        lib.ogs_profile_read(buf, 8)
        acc += np.array(list(buf))
    lib.ogs_profile_enable(0)
    stage_ms = dict(zip(stage_names, (acc / reps).tolist()))

    # per-frame distribution (forward, backward and both; median with p10 / p90), CUDA events around the
    # two boundary calls of every frame, one pose per frame
    n_dist = min(K, 50)
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(n_dist)]
    for s in range(n_dist):
        d["viewmatrix"], d["campos"] = view_dev[Wm + s]
        d["projmatrix"] = d["viewmatrix"]
        evs[s][0].record()
        fwd = h.run_forward(mod, d)
        evs[s][1].record()
        h.run_backward(mod, d, fwd, dL)
        evs[s][2].record()
    torch.cuda.synchronize()

    def spread(v):
        v = sorted(v)
        n = len(v)
        return {"p10": v[int(0.1 * (n - 1))], "p50": statistics.median(v), "p90": v[int(0.9 * (n - 1) + 0.5)]}

    frame_ms = {"forward": spread([e[0].elapsed_time(e[1]) for e in evs]),
                "backward": spread([e[1].elapsed_time(e[2]) for e in evs]),
                "forward_backward": spread([e[0].elapsed_time(e[2]) for e in evs]), "frames": n_dist}
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak, peak_src = (peaks["hbm_gbs"], "MEASURED_PEAKS.json") if "hbm_gbs" in peaks else (6650.0, "fallback")
    Kpass = (32 + int(np.ceil(np.log2(T + 1))) + 7) // 8
    B = alg_bytes(P, V, R, N, T, D, M, Kpass)
    groups = {"preprocess_fwd": ["preprocess_fwd"], "binning": ["depth_order", "tile_ranges", "emit", "tile_sort"],
              "render_fwd": ["render_fwd"], "render_bwd": ["render_bwd"], "preprocess_bwd": ["preprocess_bwd"]}
    per_stage = {}
    for gname, members in groups.items():
        ms = sum(stage_ms[m] for m in members)
        per_stage[gname] = {"ms": ms, "alg_bytes": B[gname], "gbs": B[gname] / ms / 1e6 if ms > 0 else None}
    dominant = max(per_stage, key=lambda k: per_stage[k]["ms"])      # every group competes, binning included
    # ncu-derived per-launch counters (DRAM traffic, warp instructions, issue-active, L2 RED sectors) are captured under a
    # profiler, so they come from a committed file — stamped with the hash of the kernel sources it was taken on; a stale
    # capture is reported as such and its numbers are withheld
    counters, stamp = {}, source_stamp()
    try:
        counters = json.load(open(os.path.join(ROOT, "profiles", "kernel_counters.json")))
    except Exception:
        pass
    fresh = counters.get("source_stamp") == stamp
    kc = counters.get("kernels", {}) if fresh else {}
    traffic = kc.get(dominant, {}).get("dram_bytes")
    ach = per_stage[dominant]["gbs"]
    line["roofline"] = {"bound": "hbm", "kernel": dominant, "achieved": ach, "peak": peak, "unit": "GB/s",
                        "frac": ach / peak, "traffic": traffic, "peak_source": peak_src,
                        "traffic_source": {"file": "profiles/kernel_counters.json", "captured_on": counters.get("source_stamp"),
                                           "current": stamp, "fresh": fresh},
                        "note": "blend kernels are FP32-issue / L2-RED bound, not HBM bound (SURVEY 8d); HBM frac reported as the "
                                "contract asks, the issue roofline is in roofline_issue"}
    # compute roofline of the two blend kernels: the frame's blending work (counted live by ogs_export_pair_counts,
    # independent of kernel organisation) x the instructions one (pixel, Gaussian) pair minimally costs / issue rate
    cnt = torch.zeros(4, dtype=torch.int64, device=dev)
    lib.ogs_export_pair_counts(P, W, H, fwd[0], ctypes.c_void_p(fwd[3].data_ptr()), ctypes.c_void_p(fwd[4].data_ptr()),
                               ctypes.c_void_p(fwd[5].data_ptr()), ctypes.c_void_p(cnt.data_ptr()), None)
    visited, blended, tile_sync, _npx = [int(x) for x in cnt.tolist()]
    sm_mhz = (line["clocks"]["sm_mhz"] or peaks.get("sm_max_mhz") or 1965.0)
    issue_rate = 148 * 4 * sm_mhz * 1e6            # warp instructions per second, all SMs
    MIN_INSTR = {"render_fwd": 30, "render_bwd": 47}   # per pair and lane, DESIGN.md section 5
    line["roofline_issue"] = {"pairs_visited": visited, "pairs_blended": blended, "pairs_tile_synchronous": tile_sync,
                              "issue_rate_warp_inst_per_s": issue_rate, "sm_mhz": sm_mhz, "kernels": {}}
    for kname in ("render_fwd", "render_bwd"):
        floor_ms = blended * MIN_INSTR[kname] / 32.0 / issue_rate * 1e3
        row = {"ms": per_stage[kname]["ms"], "min_instr_per_pair": MIN_INSTR[kname], "floor_ms": floor_ms,
               "frac_of_issue_roofline": floor_ms / per_stage[kname]["ms"]}
        if kname in kc:
            row.update({k: kc[kname].get(k) for k in ("warp_instructions", "issue_active_pct", "l2_red_sectors", "l2_red_bytes")})
            if kc[kname].get("l2_red_bytes"):
                row["l2_red_gbs"] = kc[kname]["l2_red_bytes"] / per_stage[kname]["ms"] / 1e6
            if kc[kname].get("warp_instructions"):
                row["issue_ms_at_100pct"] = kc[kname]["warp_instructions"] / issue_rate * 1e3
        line["roofline_issue"]["kernels"][kname] = row
    total_alg = sum(B.values())
    line["roofline_frame"] = {"alg_bytes": total_alg, "achieved": total_alg / ms_per_step / 1e6, "peak": peak,
                              "unit": "GB/s", "frac": total_alg / ms_per_step / 1e6 / peak, "stages": per_stage,
                              "stage_ms": stage_ms, "frame_ms": frame_ms}
    if world == 1:
        line["train_step"] = training_iteration_bench(h, scene, views, dev, K, Wm, None)
    # fwd: preprocess, histogram, 4 depth passes, tile ranges, width scan, column starts, x pass over segments, height scan,
    # y pass over instances, render = 13; bwd: 2; data parallel adds the colour kernel, the all-reduce (sum + max sections)
    # and the dL_dsh rebuild
    line["gpu_launches"] = K * (13 + 2 + (4 if (distributed and bucket.factored) else (2 if distributed else 0)))
    line.update(extra)
    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_port_baseline(scene, views[Wm], dL_np)
    print(json.dumps(line))
    if distributed:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
