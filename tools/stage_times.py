"""Print per-stage device times (ms) of ours at C2 (profiling API)."""
import ctypes, os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import numpy as np, torch
import _harness as h
sm = h.scene_mod
cfg = sys.argv[1] if len(sys.argv) > 1 else "C2"
scene = sm.make_config_scene(cfg)
d = h.torch_inputs(scene, sm.random_view(21) if cfg in ("C1", "C2", "C3") else sm.identity_view())
dL = torch.from_numpy(sm.make_grad_image(scene.W, scene.H, 99)).cuda()
lib = h.pkg.load_library()
names = ["preprocess_fwd", "depth_order", "tile_ranges", "emit", "tile_sort", "render_fwd", "render_bwd", "preprocess_bwd"]
for _ in range(3):
    f = h.run_forward(h.pkg, d); h.run_backward(h.pkg, d, f, dL)
lib.ogs_profile_enable(1)
acc = np.zeros(8); buf = (ctypes.c_float * 8)()
for _ in range(10):
    f = h.run_forward(h.pkg, d); h.run_backward(h.pkg, d, f, dL); lib.ogs_profile_read(buf, 8); acc += np.array(list(buf))
lib.ogs_profile_enable(0)
# whole frames without the per-stage events (stages on the side stream overlap, so their sum is not the frame)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize(); e0.record()
for _ in range(20):
    f = h.run_forward(h.pkg, d); h.run_backward(h.pkg, d, f, dL)
e1.record(); torch.cuda.synchronize()
frame_ms = e0.elapsed_time(e1) / 20
print(cfg, "R", f[0], " ".join(f"{n}={v/10:.3f}" for n, v in zip(names, acc)), f"sum={acc.sum()/10:.3f} frame={frame_ms:.3f}")
