"""Small forward+backward (plain, band and seam-wrap modes) for compute-sanitizer runs."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import numpy as np, torch
import _harness as h
sm = h.scene_mod
scene = sm.make_scene(6000, 333, 171, 0.04, 5, pole_frac=0.1, seam_frac=0.1, near_frac=0.01)
d = h.torch_inputs(scene, sm.random_view(3), bg=(0.2, 0.3, 0.4))
dL = torch.from_numpy(sm.make_grad_image(scene.W, scene.H, 9)).cuda()
f = h.run_forward(h.pkg, d); g = h.run_backward(h.pkg, d, f, dL)
st = h.ours_state(d, f)
fb = h.pkg.RasterizeGaussiansCUDA(d["background"], d["means3D"], d["colors"], d["opacity"], d["scales"], d["rotations"], 1.0,
                                  d["cov3D_precomp"], d["viewmatrix"], d["projmatrix"], 0.0, 0.0, d["H"], d["W"], d["sh"], 3,
                                  d["campos"], False, 3, False, band=(2, 7))
h.run_backward(h.pkg, d, fb, dL, reduce_accumulators=lambda t: None)
scene2 = sm.make_scene(3000, 256, 128, 0.04, 6, seam_frac=0.3)
d2 = h.torch_inputs(scene2, sm.identity_view())
h.pkg.set_seam_wrap(True)
f2 = h.run_forward(h.pkg, d2); h.run_backward(h.pkg, d2, f2, torch.ones(3, 128, 256, device="cuda") * 1e-4)
h.pkg.set_seam_wrap(False)
torch.cuda.synchronize()
print("sanitize_case ok", f[0], fb[0], f2[0])
