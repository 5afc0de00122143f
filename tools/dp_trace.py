"""Phase timeline of the data-parallel C2 step of bench.py (one view per rank, factored exchange): CUDA events between the
phases on the main stream and on the dL_dsh side stream, host timestamps beside them.
   python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/dp_trace.py [steps]
Every rank prints one JSON line (mean microseconds per phase over the timed steps)."""
import json, os, sys, time
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch, torch.distributed as dist
import _harness as h
from importlib import import_module
par = import_module("omnigs-fork_b200.parallel")
sm = h.scene_mod
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
os.environ.setdefault("NCCL_DEBUG", "WARN")
dist.init_process_group("nccl", device_id=dev)
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 30
warm = 5
scene = sm.make_config_scene("C2")
P, W, H = scene.P, scene.W, scene.H
views = [sm.random_view(1000 + 97 * s + rank) for s in range(steps + warm)]
d = h.torch_inputs(scene, views[0], device=dev)
view_dev = [(torch.from_numpy(v).to(dev), torch.from_numpy(c).to(dev)) for v, c in views]
campos_all = [torch.from_numpy(np.stack([sm.random_view(1000 + 97 * s + r)[1] for r in range(world)])).to(dev) for s in range(steps + warm)]
dL = torch.from_numpy(sm.make_grad_image(W, H, 99)).to(dev)
bucket = par.GradientBucket(P, 16, dev, views_per_rank=1)
side = torch.cuda.Stream()
pending, deferred, prev_trace = [None], [None], [None]
DEFER = False
cur = torch.cuda.current_stream()
rows, hosts = [], []

def mark(tr, label):
    ev = torch.cuda.Event(enable_timing=True); ev.record(cur); tr.append((label, ev))

def step(s, trace):
    vm, cp = view_dev[s]
    tr = [] if trace else None
    par._trace = tr
    t = [time.perf_counter()]
    if trace: mark(tr, "start")
    def launch():
        if deferred[0] is not None:
            par._trace = prev_trace[0]
            pending[0] = deferred[0]()
            deferred[0] = None
            par._trace = tr
    st = h.pkg.RasterizeGaussiansGeometry(d["means3D"], d["opacity"], d["scales"], d["rotations"], 1.0, d["cov3D_precomp"], vm, cp, H, W)
    t.append(time.perf_counter())
    if trace: mark(tr, "geometry_sort")
    if pending[0] is not None:
        cur.wait_event(pending[0])
    if trace: mark(tr, "wait_sh")
    fwd = h.pkg.RasterizeGaussiansBlend(st, d["background"], d["sh"], 3)
    t.append(time.perf_counter())
    if trace: mark(tr, "colors_blend")
    h.pkg.RasterizeGaussiansBackwardView(d["background"], d["means3D"], fwd[2], d["scales"], d["rotations"], 1.0, vm, dL, d["sh"], 3, cp,
                                         fwd[3], fwd[0], fwd[4], fwd[5], bucket, 0)
    t.append(time.perf_counter())
    if DEFER:
        pending[0] = None
        raise RuntimeError("deferred rebuild was removed (measured slower)")
        prev_trace[0] = tr
    else:
        pending[0] = par.exchange_bucket(bucket, means3D=d["means3D"], campos_views=campos_all[s], degree=3, sh_stream=side)
    t.append(time.perf_counter())
    par._trace = None
    if trace:
        rows.append(tr); hosts.append(t)

for s in range(warm):
    step(s, False)
torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for s in range(steps):
    step(warm + s, True)
if deferred[0] is not None:
    par._trace = prev_trace[0]
    pending[0] = deferred[0]()
    par._trace = None
cur.wait_event(pending[0])
e1.record()
torch.cuda.synchronize()
out = {"rank": rank, "world": world, "ms_per_step": e0.elapsed_time(e1) / steps, "gpu_us": {}, "host_us": {}}
main_labels = ["start", "geometry_sort", "wait_sh", "colors_blend", "backward_done", "barrier0", "allreduce", "barrier1"]
for a, b in zip(main_labels[:-1], main_labels[1:]):
    out["gpu_us"][b] = 1e3 * float(np.mean([dict(tr)[a].elapsed_time(dict(tr)[b]) for tr in rows]))
out["gpu_us"]["sh_rebuild"] = 1e3 * float(np.mean([dict(tr)["sh_fork"].elapsed_time(dict(tr)["sh_rebuild"]) for tr in rows]))
out["gpu_us"]["sh_barrier2"] = 1e3 * float(np.mean([dict(tr)["sh_rebuild"].elapsed_time(dict(tr)["sh_barrier2"]) for tr in rows]))
# gap between steps on the main stream: barrier1 of step s -> start of step s+1
out["gpu_us"]["gap_to_next_start"] = 1e3 * float(np.mean([dict(rows[i])["barrier1"].elapsed_time(dict(rows[i + 1])["start"]) for i in range(len(rows) - 1)]))
for i, n in enumerate(["geometry_call (blocks on num_rendered)", "blend_call", "backward_call", "exchange_call"]):
    out["host_us"][n] = 1e6 * float(np.mean([t[i + 1] - t[i] for t in hosts]))
out["host_us"]["step_total"] = 1e6 * float(np.mean([hosts[i + 1][0] - hosts[i][0] for i in range(len(hosts) - 1)]))
out["defer_sh"] = DEFER
import ctypes
lib = h.pkg.load_library()
names = ["preprocess_fwd", "depth_order", "tile_ranges", "emit", "tile_sort", "render_fwd", "render_bwd", "preprocess_bwd"]
lib.ogs_profile_enable(1)
acc = np.zeros(8); buf = (ctypes.c_float * 8)()
for s in range(5):
    step(warm + s, False)
    lib.ogs_profile_read(buf, 8); acc += np.array(list(buf))
lib.ogs_profile_enable(0)
torch.cuda.synchronize()
out["stage_us"] = {n: round(200.0 * v, 1) for n, v in zip(names, acc)}
for r in range(world):
    if r == rank:
        print(json.dumps(out), flush=True)
    dist.barrier()
dist.destroy_process_group()
