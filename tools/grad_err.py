"""Gradient errors of ours against the live reference on two small cases (A/B tool)."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import torch, _harness as h
sm = h.scene_mod
ref = h.load_reference()
for name, mk in {"colors": lambda: (sm.make_scene(30000, 400, 200, 0.03, 74), sm.random_view(75), "colors", (0.1, 0.2, 0.3), 0),
                 "cov_deg1": lambda: (sm.make_scene(30000, 400, 200, 0.03, 76), sm.random_view(77), "cov", (0, 0, 0), 1),
                 "C1": lambda: (sm.make_config_scene("C1"), sm.identity_view(), "sh", (0, 0, 0), 3)}.items():
    scene, view, mode, bg, degree = mk()
    d = h.torch_inputs(scene, view, mode=mode, bg=bg, degree=degree)
    dL = torch.from_numpy(sm.make_grad_image(scene.W, scene.H, 99)).cuda()
    fo, fr = h.run_forward(h.pkg, d), h.run_forward(ref, d)
    go, gr, gr2 = h.run_backward(h.pkg, d, fo, dL), h.run_backward(ref, d, fr, dL), h.run_backward(ref, d, fr, dL)
    print(name, " ".join(f"{n[3:]}={h.grad_error(a, b)[0]:.1e}({h.grad_error(c, b)[0]:.0e})" for n, a, b, c in zip(h.GRAD_NAMES, go, gr, gr2)))
