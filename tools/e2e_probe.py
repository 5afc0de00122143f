"""Where does the end-to-end step's time beyond the device-resident frame go?  Adds the pieces of bench.py's e2e step one
at a time (C2, one GPU) and prints ms/step for each variant."""
import importlib, os, sys
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import _harness as h
sm = h.scene_mod
tr = importlib.import_module("omnigs-fork_b200.trainer")
scene = sm.make_config_scene("C2")
W, H = scene.W, scene.H
views = [sm.random_view(1000 + 97 * s) for s in range(60)]
d = h.torch_inputs(scene, views[0])
view_dev = [(torch.from_numpy(v).cuda(), torch.from_numpy(c).cuda()) for v, c in views]
dL_np = sm.make_grad_image(W, H, 99)
dL = torch.from_numpy(dL_np).cuda()
gt_host = torch.from_numpy(np.clip(dL_np * (W * H) * 0.1 + 0.5, 0, 1).astype(np.float32)).pin_memory()
view_host = [torch.from_numpy(np.concatenate([v.reshape(-1), c.reshape(-1)])).pin_memory() for v, c in views]
gt_dev = torch.empty_like(gt_host, device="cuda"); gt_dev.copy_(gt_host)
vbuf = torch.empty(19, device="cuda")
copy_stream, gt_ready = torch.cuda.Stream(), torch.cuda.Event()
loss_host, loss_ready = torch.empty(3).pin_memory(), torch.cuda.Event()

gt_bufs = [gt_dev, torch.empty_like(gt_dev)]
gt_events = [torch.cuda.Event(), torch.cuda.Event()]
loss_done = torch.cuda.Event()

def step_prefetch(s):
    """Target image of step s+1 copied while step s runs (double-buffered), pose copy, L1 pass, deferred loss read."""
    cur, nxt = s & 1, (s + 1) & 1
    vbuf.copy_(view_host[s], non_blocking=True)
    d["viewmatrix"], d["campos"] = vbuf[:16].view(4, 4), vbuf[16:19]
    d["projmatrix"] = d["viewmatrix"]
    fwd = h.run_forward(h.pkg, d)
    torch.cuda.current_stream().wait_event(gt_events[cur])
    loss_out, g_in = tr.photometric_loss(fwd[1], gt_bufs[cur], 0.0)
    loss_done.record()
    # the other buffer was last read by the previous step's loss, which is behind us on this stream
    copy_stream.wait_event(loss_done)
    with torch.cuda.stream(copy_stream):
        gt_bufs[nxt].copy_(gt_host, non_blocking=True); gt_events[nxt].record()
    h.run_backward(h.pkg, d, fwd, g_in)
    loss_ready.synchronize(); _ = float(loss_host[0])
    loss_host.copy_(loss_out, non_blocking=True); loss_ready.record()

def step(s, pose_copy, gt_copy, loss, readback):
    if pose_copy:
        vbuf.copy_(view_host[s], non_blocking=True)
        d["viewmatrix"], d["campos"] = vbuf[:16].view(4, 4), vbuf[16:19]
    else:
        d["viewmatrix"], d["campos"] = view_dev[s]
    d["projmatrix"] = d["viewmatrix"]
    if gt_copy:
        copy_stream.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(copy_stream):
            gt_dev.copy_(gt_host, non_blocking=True); gt_ready.record()
    fwd = h.run_forward(h.pkg, d)
    if gt_copy:
        torch.cuda.current_stream().wait_event(gt_ready)
    g_in = dL
    if loss:
        loss_out, g_in = tr.photometric_loss(fwd[1], gt_dev, 0.0)
    h.run_backward(h.pkg, d, fwd, g_in)
    if loss and readback == "deferred":
        loss_ready.synchronize(); _ = float(loss_host[0])
        loss_host.copy_(loss_out, non_blocking=True); loss_ready.record()
    elif loss and readback == "item":
        _ = float(loss_out[0].item())

for name, args in [("frame only", (False, False, False, None)), ("+ pose H2D", (True, False, False, None)),
                   ("+ target H2D (side stream)", (True, True, False, None)), ("+ L1 loss pass", (True, True, True, None)),
                   ("+ deferred loss read", (True, True, True, "deferred")), ("+ blocking loss read", (True, True, True, "item")),
                   ("frame only again", (False, False, False, None))]:
    loss_ready.record()
    for s in range(5): step(s, *args)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for s in range(5, 55): step(s, *args)
    e1.record(); torch.cuda.synchronize()
    print(f"{name:32s} {e0.elapsed_time(e1) / 50:.3f} ms/step")
gt_events[0].record(); gt_events[1].record(); loss_ready.record()
for s in range(5): step_prefetch(s)
torch.cuda.synchronize()
e0.record()
for s in range(5, 55): step_prefetch(s)
e1.record(); torch.cuda.synchronize()
print(f"{'prefetched target (under bwd)':32s} {e0.elapsed_time(e1) / 50:.3f} ms/step")
