"""One of bench.py's multi-GPU blocks on its own (A/B runs of the exchange variants):
   python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/bench_blocks.py bands|dp_views [steps]"""
import json, os, sys
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch, torch.distributed as dist
import _harness as h
import bench
from importlib import import_module
par = import_module("omnigs-fork_b200.parallel")
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local); dev = torch.device("cuda", local)
distributed = world > 1
if distributed:
    os.environ.setdefault("NCCL_DEBUG", "WARN")
    dist.init_process_group("nccl", device_id=dev)
which = sys.argv[1]
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 8
fn = {"bands": bench.bands_block, "dp_views": bench.dp_views_block}[which]
out = fn(h, par, dev, rank, world, distributed, steps=steps)
if rank == 0:
    print(json.dumps({"block": which, "value": out["value"], "unit": out["unit"], "n_gpus": world, "rank0": out.get("rank0"),
                      "env": {k: v for k, v in os.environ.items() if k.startswith("OGS_")}, "config": out["config"]}), flush=True)
if distributed:
    dist.destroy_process_group()
