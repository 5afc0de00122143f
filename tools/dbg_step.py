"""25 timed forward+backward frames at C2 after 5 warm-up frames (ncu target for the frame kernels)."""
import sys, os, time
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import _harness as h
sm = h.scene_mod
scene = sm.make_config_scene("C2")
views = [sm.random_view(1000 + 97 * s) for s in range(30)]
d = h.torch_inputs(scene, views[0])
view_dev = [(torch.from_numpy(v).cuda(), torch.from_numpy(c).cuda()) for v, c in views]
dL = torch.from_numpy(sm.make_grad_image(scene.W, scene.H, 99)).cuda()
def step(s):
    d["viewmatrix"], d["campos"] = view_dev[s]; d["projmatrix"] = d["viewmatrix"]
    fwd = h.run_forward(h.pkg, d); g = h.run_backward(h.pkg, d, fwd, dL); return fwd, g
for s in range(5): step(s)
torch.cuda.synchronize()
for s in range(5, 30):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.time(); e0.record(); fwd, g = step(s); e1.record(); t1 = time.time(); torch.cuda.synchronize(); t2 = time.time()
    print(f"step {s}: gpu {e0.elapsed_time(e1):.3f} ms  cpu-launch {1e3*(t1-t0):.3f} ms  total {1e3*(t2-t0):.3f}  R {fwd[0]}  mem {torch.cuda.memory_reserved()/1e9:.2f} GB")
