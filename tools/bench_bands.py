#!/usr/bin/env python
"""Latitude-band partitioned render + backward of ONE large panorama (BASELINE configs[3]: C4, 5M Gaussians,
7680x3840) across N GPUs.  Launch:  torchrun --nproc-per-node N tools/bench_bands.py [--config C4] [--steps K]
Every rank holds all Gaussians, bins/sorts/blends only its tile rows (balanced by the previous frame's per-row
instance counts), the image is summed over ranks, each rank runs the backward for its band and the gradient
shares are all-reduced.  Prints one JSON line (rank 0): ms per frame (max over ranks, CUDA events)."""
import argparse, json, os, sys
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch, torch.distributed as dist
import _harness as h
from importlib import import_module
par = import_module("omnigs-fork_b200.parallel")
sm = h.scene_mod

ap = argparse.ArgumentParser()
ap.add_argument("--config", default="C4"); ap.add_argument("--steps", type=int, default=10); ap.add_argument("--warmup", type=int, default=3)
args = ap.parse_args()
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
scene = sm.make_config_scene(args.config)
d = h.torch_inputs(scene, sm.identity_view(), device=f"cuda:{local}")
dL = torch.from_numpy(sm.make_grad_image(scene.W, scene.H, 99)).cuda()

def rasterize(band):
    return h.pkg.RasterizeGaussiansCUDA(d["background"], d["means3D"], d["colors"], d["opacity"], d["scales"], d["rotations"], 1.0,
                                        d["cov3D_precomp"], d["viewmatrix"], d["projmatrix"], 0.0, 0.0, d["H"], d["W"], d["sh"], 3,
                                        d["campos"], False, 3, False, band=band)
# "previous frame": one full-frame forward gives the per-row loads used to balance the bands
full = rasterize(None)
rows = par.tile_row_counts(h.ours_state(d, full)["ranges"], scene.W, scene.H)
R_full = full[0]
del full
bands = par.band_rows(rows, world)
band = bands[rank]

def step():
    img, fwd = par.render_band_forward(rasterize, band, scene.H)
    # exchange the 48 B/Gaussian accumulators between the two backward kernels, not the 324 B of final gradients
    g = h.run_backward(h.pkg, d, fwd, dL, reduce_accumulators=(lambda t: dist.all_reduce(t)) if world > 1 else (lambda t: None))
    return fwd[0]

for _ in range(args.warmup): Rb = step()
torch.cuda.synchronize()
if world > 1: dist.barrier(); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(args.steps): step()
e1.record(); torch.cuda.synchronize()
if world > 1: dist.barrier()
ms = torch.tensor([e0.elapsed_time(e1) / args.steps, float(Rb)], device="cuda", dtype=torch.float64)
if world > 1:
    mx = ms.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
    sm_ = ms.clone(); dist.all_reduce(sm_, op=dist.ReduceOp.SUM)
else:
    mx, sm_ = ms, ms
if rank == 0:
    print(json.dumps({"metric": "lonlat_band_parallel_ms_per_frame", "value": float(mx[0]), "unit": "ms", "n_gpus": world,
                      "higher_is_better": False, "scaling": "strong", "steps": args.steps, "warmup": args.warmup,
                      "config": {"workload": args.config, "gaussians": scene.P, "image": [scene.W, scene.H], "num_rendered": R_full,
                                 "bands": bands, "sum_band_instances": int(sm_[1]), "max_band_instances": int(mx[1])}}))
if world > 1: dist.destroy_process_group()
