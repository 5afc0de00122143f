#!/usr/bin/env python
"""BASELINE configs[2]: C3 (3M Gaussians, 1920x960), one training step = 8 views, data parallel over N GPUs.
Launch:  [torchrun --nproc-per-node N] tools/bench_dp_views.py [--config C3] [--views 8] [--steps K]
Rank g renders views g, g+N, ... of the step's 8 poses (forward + backward each), accumulates the optimiser-facing
gradients and densification statistics locally, and ONE exchange per step sums them over ranks (peer-memory / multimem
kernel, NCCL fallback).  Strong scaling: the step is the same work at every N.  Prints one JSON line (rank 0)."""
import argparse, json, os, sys
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch, torch.distributed as dist
import _harness as h
from importlib import import_module
par = import_module("omnigs-fork_b200.parallel")
sm = h.scene_mod

ap = argparse.ArgumentParser()
ap.add_argument("--config", default="C3"); ap.add_argument("--views", type=int, default=8)
ap.add_argument("--steps", type=int, default=10); ap.add_argument("--warmup", type=int, default=3)
args = ap.parse_args()
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local); dev = torch.device("cuda", local)
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=dev)
scene = sm.make_config_scene(args.config)
d = h.torch_inputs(scene, sm.random_view(0), device=dev)
dL = torch.from_numpy(sm.make_grad_image(scene.W, scene.H, 99)).to(dev)
poses = [[sm.random_view(5000 + 31 * s + v) for v in range(args.views)] for s in range(args.steps + args.warmup)]
poses = [[(torch.from_numpy(a).to(dev), torch.from_numpy(c).to(dev)) for a, c in step] for step in poses]
mine = par.views_for_rank(args.views, rank, world)
bucket = par.GradientBucket(scene.P, 16, dev)          # exchanged once per step
work = par.GradientBucket(scene.P, 16, dev, peer=False) if len(mine) > 1 else bucket   # per-view scratch

def step(s):
    for k, v in enumerate(mine):
        d["viewmatrix"], d["campos"] = poses[s][v]; d["projmatrix"] = d["viewmatrix"]
        fwd = h.run_forward(h.pkg, d)
        tgt = bucket if k == 0 else work
        g = h.run_backward(h.pkg, d, fwd, dL, out=tgt)
        if k == 0:
            par.fill_view_stats(bucket, g[0], fwd[2])
        else:                                            # accumulate this view into the step's bucket
            par.fill_view_stats(work, g[0], fwd[2])
            bucket.flat += work.flat
            torch.maximum(bucket.max_radii2D, work.max_radii2D, out=bucket.max_radii2D)
    par.exchange_bucket(bucket)

for s in range(args.warmup): step(s)
torch.cuda.synchronize()
if world > 1: dist.barrier(); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for s in range(args.steps): step(args.warmup + s)
e1.record(); torch.cuda.synchronize()
ms = torch.tensor([e0.elapsed_time(e1) / args.steps], device=dev, dtype=torch.float64)
if world > 1: dist.all_reduce(ms, op=dist.ReduceOp.MAX)
if rank == 0:
    print(json.dumps({"metric": "lonlat_dp_train_views_per_s", "value": args.views / (float(ms) / 1e3), "unit": "views/s",
                      "ms_per_step": float(ms), "n_gpus": world, "scaling": "strong", "steps": args.steps, "warmup": args.warmup,
                      "config": {"workload": args.config, "gaussians": scene.P, "image": [scene.W, scene.H],
                                 "views_per_step": args.views, "views_per_rank": len(mine),
                                 "gradient_exchange": "peer/multimem kernel" if bucket.peer else ("NCCL" if world > 1 else "none")}}))
if world > 1: dist.destroy_process_group()
