"""2+ GPUs (torchrun): the peer-memory all-reduce kernel against NCCL on the same data; prints timings."""
import importlib, os, sys
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch, torch.distributed as dist
import _harness as h
par = importlib.import_module("omnigs-fork_b200.parallel")
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr); dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
P, M = 1_000_003, 16
b = par.GradientBucket(P, M, dev)
assert b.peer is not None, "symmetric memory unavailable"
g = torch.Generator(device=dev).manual_seed(7 + rank)
src = torch.randn(b.flat.shape, device=dev, generator=g)
ref = src.clone(); dist.all_reduce(ref)
m2d = torch.randn((P, 3), device=dev, generator=g); radii = torch.randint(0, 50, (P,), device=dev, generator=g, dtype=torch.int32)
for it in range(3):
    b.flat.copy_(src)
    acc_ref = torch.where(radii > 0, m2d[:, :2].norm(dim=-1), torch.zeros((), device=dev)); dist.all_reduce(acc_ref)
    grads, stats = par.allreduce_bucket(b, m2d, radii)
    torch.cuda.synchronize()
    # the statistics slots were overwritten by allreduce_bucket; compare the five gradient sections and the stats
    for n in par.OPTIMISED:
        o = b[n].data_ptr() - b.flat.data_ptr()
        sec = ref.view(-1)[o // 4: o // 4 + b[n].numel()].view(b[n].shape)
        err = float((b[n] - sec).abs().max())
        assert err <= 1e-6 * world, (n, err)
    assert float((stats["xyz_gradient_accum"] - acc_ref).abs().max()) <= 1e-6 * world
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
def kernel_only():
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    b.peer["handle"].barrier(channel=0)
    lib.ogs_peer_allreduce_sum(arr, world, rank, b.flat.numel(), st)
    b.peer["handle"].barrier(channel=1)
def kernel_bare():
    lib.ogs_peer_allreduce_sum(arr, world, rank, b.flat.numel(), ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
import ctypes
lib = h.pkg.load_library()
arr = (ctypes.c_void_p * world)(*b.peer["ptrs"])
def set_mm(v):
    b.peer["use_multimem"] = v
if rank == 0: print("multicast pointer:", hex(b.peer.get("multicast") or 0))
if b.peer.get("multicast"):
    set_mm(True)
    b.flat.copy_(src); par.allreduce_bucket(b, m2d, radii); torch.cuda.synchronize()
    for n in par.OPTIMISED:
        o = b[n].data_ptr() - b.flat.data_ptr()
        sec = ref.view(-1)[o // 4: o // 4 + b[n].numel()].view(b[n].shape)
        assert float((b[n] - sec).abs().max()) <= 2e-6 * world, n
    if rank == 0: print("multimem all-reduce matches NCCL")
def mm_only():
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    b.peer["handle"].barrier(channel=0)
    lib.ogs_multimem_allreduce_sum(ctypes.c_void_p(b.peer["multicast"]), world, rank, b.flat.numel(), st)
    b.peer["handle"].barrier(channel=1)
extra = (("multimem barriers + kernel", mm_only),) if b.peer.get("multicast") else ()
set_mm(False)
for name, fn in (("peer (stats + barriers + kernel + max-allreduce)", lambda: par.allreduce_bucket(b, m2d, radii)),) + extra + (
                 ("peer barriers + kernel", kernel_only),
                 ("nccl", lambda: dist.all_reduce(ref)),):
    for _ in range(3): fn()
    dist.barrier(); torch.cuda.synchronize(); e0.record()
    for _ in range(20): fn()
    e1.record(); torch.cuda.synchronize()
    if rank == 0: print(f"{name}: {e0.elapsed_time(e1) / 20:.3f} ms per {b.flat.numel() * 4 / 1e6:.0f} MB all-reduce (world {world})")
if rank == 0: print("peer all-reduce matches NCCL")
dist.destroy_process_group()
