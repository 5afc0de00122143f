"""A few training iterations at C2 through trainer.train_for_one_iteration (ncu target for the loss / Adam /
raw-parameter kernels)."""
import importlib, os, sys
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import _harness as h
sm = h.scene_mod
tr = importlib.import_module("omnigs-fork_b200.trainer")
scene = sm.make_config_scene(sys.argv[1] if len(sys.argv) > 1 else "C2")
t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
op = np.clip(scene.opacities.astype(np.float64), 1e-4, 1 - 1e-4)
pc = tr.GaussianModel(t(scene.means3D), t(scene.shs[:, :1, :]), t(scene.shs[:, 1:, :]), t(np.log(op / (1 - op)).astype(np.float32)),
                      t(np.log(scene.scales.astype(np.float64)).astype(np.float32)), t(scene.rotations))
gt = torch.rand((3, scene.H, scene.W), device="cuda")
bg = torch.zeros(3, device="cuda")
N = int(sys.argv[2]) if len(sys.argv) > 2 else 8
views = [sm.random_view(1000 + 97 * s) for s in range(N)]
for s, (v, c) in enumerate(views):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    loss, _ = tr.train_for_one_iteration(pc, t(v), t(c), gt, bg, s + 1)
    e1.record(); torch.cuda.synchronize()
    print(f"iteration {s}: {e0.elapsed_time(e1):.3f} ms loss {float(loss[0]):.5f}")
