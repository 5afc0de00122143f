"""2+ GPUs (torchrun): the library's NVLink collectives against NCCL on the same data, with timings.

  torchrun --nproc-per-node N tools/peer_check.py

Checks (asserts) and times, on every rank:
  1. dense bucket: ogs_peer_allreduce / ogs_multimem_allreduce (summed section + max section) against NCCL SUM / MAX;
  2. factored bucket: geometry all-reduce + dL_dsh rebuilt from every rank's dL/dRGB factors read over NVLink
     (ogs_sh_gradient_from_views), on the current stream and on a side stream, against "all-gather the factors with
     NCCL and rebuild locally"; every replica must end with identical bits;
  3. latitude-band exchange: ogs_band_rows_allgather and the [P,12] accumulator all-reduce against NCCL.
"""
import ctypes, importlib, os, sys
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch, torch.distributed as dist
import _harness as h
par = importlib.import_module("omnigs-fork_b200.parallel")
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr); dev = torch.device("cuda", lr)
os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
dist.init_process_group("nccl", device_id=dev)
lib = h.pkg.load_library()
P, M = 1_000_003, 16
say = lambda *a: print(*a, flush=True) if rank == 0 else None
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)


def timed(name, fn, nbytes, reps=20):
    for _ in range(3): fn()
    dist.barrier(); torch.cuda.synchronize(); e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    say(f"  {name}: {e0.elapsed_time(e1) / reps:.3f} ms ({nbytes / 1e6:.0f} MB, world {world})")


# ---------------------------------------------------------------- 1. dense bucket, sum + max sections
b = par.GradientBucket(P, M, dev)
assert b.peer is not None, f"symmetric memory unavailable: {b.peer_error}"
say("multicast pointer:", hex(b.peer.get("multicast") or 0))
g = torch.Generator(device=dev).manual_seed(7 + rank)
src = torch.randn(b.flat.shape, device=dev, generator=g)
src[b.count_sum:] = torch.randint(0, 300, (b.flat.numel() - b.count_sum,), device=dev, generator=g).float()   # radii
ref_sum = src[:b.count_sum].clone(); dist.all_reduce(ref_sum)
ref_max = src[b.count_sum:b.count_sum + b.count_max].clone(); dist.all_reduce(ref_max, op=dist.ReduceOp.MAX)
for mm in ([False, True] if b.peer.get("multicast") else [False]):
    b.peer["use_multimem"] = mm
    b.flat.copy_(src)
    par.exchange_bucket(b)
    torch.cuda.synchronize()
    assert float((b.flat[:b.count_sum] - ref_sum).abs().max()) <= 2e-6 * world, ("sum", mm)
    assert torch.equal(b.flat[b.count_sum:b.count_sum + b.count_max], ref_max), ("max", mm)
    say(f"dense bucket all-reduce (sum + max sections) matches NCCL, multimem={mm}")
    timed(f"dense exchange, multimem={mm} (2 barriers + kernels)", lambda: par.exchange_bucket(b), b.flat.numel() * 4)
tmp = src[:b.count_sum].clone()
timed("NCCL all-reduce of the same summed section", lambda: dist.all_reduce(tmp), b.count_sum * 4)
b.peer.pop("use_multimem", None)
del b, src, ref_sum, ref_max, tmp
torch.cuda.empty_cache()

# ---------------------------------------------------------------- 2. factored bucket
for vpr in (1, 2):
    fb = par.GradientBucket(P, M, dev, views_per_rank=vpr)
    assert fb.peer is not None
    g = torch.Generator(device=dev).manual_seed(100 + rank)
    src = torch.randn(fb.flat.shape, device=dev, generator=g)
    src[fb.count_sum:fb.count_sum + fb.count_max] = torch.randint(0, 300, (fb.count_max,), device=dev, generator=g).float()
    means = torch.randn((P, 3), device=dev, generator=torch.Generator(device=dev).manual_seed(5)) * 5
    campos = torch.randn((vpr * world, 3), device=dev, generator=torch.Generator(device=dev).manual_seed(6)) * 0.3
    # checker: NCCL all-reduce / all-gather + local rebuild
    ref_sum = src[:fb.count_sum].clone(); dist.all_reduce(ref_sum)
    mine = src[fb.offsets["dL_drgb"]:fb.offsets["dL_drgb"] + vpr * P * 3].view(vpr, P, 3).contiguous()
    gathered = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(gathered, mine)
    want = torch.empty((P, M, 3), device=dev)
    par.sh_gradient_from_views(means, campos, [gathered[r][s] for s in range(vpr) for r in range(world)], 3, want)
    side = torch.cuda.Stream()
    for use_side in (False, True):
        fb.flat.copy_(src)
        fb["dL_dsh"].fill_(float("nan"))
        ev = par.exchange_bucket(fb, means3D=means, campos_views=campos, degree=3, sh_stream=side if use_side else None)
        if ev is not None:
            torch.cuda.current_stream().wait_event(ev)
        torch.cuda.synchronize()
        assert float((fb.flat[:fb.count_sum] - ref_sum).abs().max()) <= 2e-6 * world
        assert torch.equal(fb["dL_dsh"].view(torch.int32), want.view(torch.int32)), ("dL_dsh bits", vpr, use_side)
    # replicas identical: compare a checksum of the bits across ranks
    chk = fb["dL_dsh"].view(torch.int32).long().sum().reshape(1)
    lo, hi = chk.clone(), chk.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    assert int(lo) == int(hi)
    say(f"factored bucket, {vpr} view(s) per rank: sums match NCCL, dL_dsh bit-identical to all-gather + local rebuild on every rank")

    def seq():
        par.exchange_bucket(fb, means3D=means, campos_views=campos, degree=3)

    def overlapped():
        ev = par.exchange_bucket(fb, means3D=means, campos_views=campos, degree=3, sh_stream=side)
        torch.cuda.current_stream().wait_event(ev)
    moved = (fb.count_sum + fb.count_max) * 4 + (world - 1) * vpr * P * 12
    timed(f"factored exchange, {vpr} view(s)/rank, rebuild on the same stream", seq, moved)
    timed(f"factored exchange, {vpr} view(s)/rank, rebuild on a side stream (waited for)", overlapped, moved)

    def critical_path_only():
        ev = par.exchange_bucket(fb, means3D=means, campos_views=campos, degree=3, sh_stream=side)
        critical_path_only.ev = ev
    timed(f"factored exchange, {vpr} view(s)/rank, what stays on the main stream", critical_path_only, (fb.count_sum + fb.count_max) * 4)
    torch.cuda.current_stream().wait_event(critical_path_only.ev)
    torch.cuda.synchronize()
    del fb, src, gathered, want
    torch.cuda.empty_cache()

# ---------------------------------------------------------------- 3. latitude-band exchange
W, H, Pb = 3840, 1920, 2_000_003
ex = par.BandExchange(Pb, W, H, dev)
assert ex.peer is not None
gy = (H + 15) // 16
bands = [(gy * r // world, gy * (r + 1) // world) for r in range(world)]
frame = torch.randn((3, H, W), device=dev, generator=torch.Generator(device=dev).manual_seed(9))
y0, y1 = min(H, bands[rank][0] * 16), min(H, bands[rank][1] * 16)
mine = torch.full((3, H, W), float("nan"), device=dev)
mine[:, y0:y1] = frame[:, y0:y1]
full = ex.gather_image(mine, bands[rank])
torch.cuda.synchronize()
assert torch.equal(full, frame)
acc = torch.randn((Pb, 12), device=dev, generator=torch.Generator(device=dev).manual_seed(20 + rank))
ref = acc.clone(); dist.all_reduce(ref)
ex.acc.copy_(acc); ex.reduce_accumulators(); torch.cuda.synchronize()
assert float((ex.acc - ref).abs().max()) <= 2e-6 * world
say("band exchange: all-gather of band rows is exact, accumulator all-reduce matches NCCL")
timed("band rows all-gather (peer stores + barrier)", lambda: ex.gather_image(mine, bands[rank]), 3 * H * W * 4)
pad = torch.zeros_like(frame)
timed("NCCL all-reduce of zero-padded full frames (round 1)", lambda: dist.all_reduce(pad), 3 * H * W * 4)
timed("accumulator all-reduce (2 barriers + kernel)", ex.reduce_accumulators, Pb * 48)
timed("NCCL all-reduce of the accumulators", lambda: dist.all_reduce(ref), Pb * 48)
say("peer_check done")
dist.destroy_process_group()
