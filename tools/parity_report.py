"""Verbose parity + timing report: ours vs the unmodified reference rasterizer (oracle/_ref) on a GPU.
Usage: python tools/parity_report.py [case ...]   cases: tiny odd C1 C2 C2views colors cov bgwhite
"""
import os
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import numpy as np
import torch

import _harness as h

ref = h.load_reference()
sm = h.scene_mod


def timeit(fn, iters=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ts = []
    for _ in range(iters):
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))


def report(name, scene, view, mode="sh", bg=(0, 0, 0), degree=3, time_it=True):
    print(f"\n=== {name}: P={scene.P} {scene.W}x{scene.H} mode={mode} bg={bg} deg={degree}", flush=True)
    d = h.torch_inputs(scene, view, mode=mode, bg=bg, degree=degree)
    dL = torch.from_numpy(sm.make_grad_image(scene.W, scene.H, 99)).cuda()
    fo = h.run_forward(h.pkg, d)
    torch.cuda.synchronize()
    print("ours  num_rendered", fo[0], flush=True)
    go = h.run_backward(h.pkg, d, fo, dL)
    torch.cuda.synchronize()
    so = h.ours_state(d, fo)
    if ref is not None:
        fr = h.run_forward(ref, d)
        gr = h.run_backward(ref, d, fr, dL)
        torch.cuda.synchronize()
        sr = h.ref_state(ref, d, fr)
        print("ref   num_rendered", fr[0])
        vis = fr[2] > 0
        print("radii mismatches", int((fo[2] != fr[2]).sum()), "/", scene.P, " visible", int(vis.sum()))
        print("tiles_touched mismatches", int((so["tiles_touched"] != sr["tiles_touched"]).sum()))
        keys = ["means2D", "depths", "conic_opacity"] + ([] if mode == "colors" else ["rgb"]) + ([] if mode == "cov" else ["cov3D"])
        for k in keys:
            a, b = so[k][vis], sr[k][vis]
            bits = int((a.view(torch.int32) != b.view(torch.int32)).sum())
            print(f"  {k}: bit mismatches {bits} / {a.numel()}  max abs diff {float((a - b).abs().max()) if a.numel() else 0:.3e}")
        if mode != "colors":
            print("  clamped mismatches", int((so["clamped"][vis] != sr["clamped"][vis].to(torch.uint8)).sum()))
        if fo[0] == fr[0]:
            print("ranges mismatches", int((so["ranges"] != sr["ranges"]).sum()),
                  " point_list mismatches", int((so["point_list"] != sr["point_list"]).sum()),
                  " keys mismatches", int((so["point_list_keys"] != sr["point_list_keys"]).sum()))
        print(f"out_color max abs diff {float((fo[1] - fr[1]).abs().max()):.3e}  final_T max diff "
              f"{float((so['accum_alpha'] - sr['accum_alpha']).abs().max()):.3e}  n_contrib mismatches "
              f"{int((so['n_contrib'] != sr['n_contrib']).sum())} / {scene.W * scene.H}")
        for nm, a, b in zip(h.GRAD_NAMES, go, gr):
            rel, exc = h.grad_error(a, b)
            print(f"  {nm:14s} rel(max-norm) {rel:.3e}  elem excess {exc:.3e}  max|ref| {float(b.abs().max()) if b.numel() else 0:.3e}")
        # reference vs itself (atomic-order noise floor)
        gr2 = h.run_backward(ref, d, fr, dL)
        print("  ref-vs-ref noise:", " ".join(f"{h.grad_error(a, b)[0]:.1e}" for a, b in zip(gr2, gr)))
    if time_it:
        t_f = timeit(lambda: h.run_forward(h.pkg, d))
        t_b = timeit(lambda: h.run_backward(h.pkg, d, fo, dL))
        line = f"time ours fwd {t_f:.3f} ms  bwd {t_b:.3f} ms  total {t_f + t_b:.3f}"
        if ref is not None:
            r_f = timeit(lambda: h.run_forward(ref, d))
            r_b = timeit(lambda: h.run_backward(ref, d, fr, dL))
            line += f" | ref fwd {r_f:.3f}  bwd {r_b:.3f}  total {r_f + r_b:.3f} | speedup {(r_f + r_b) / (t_f + t_b):.2f}x"
        print(line, flush=True)


cases = sys.argv[1:] or ["tiny", "odd", "C1"]
for c in cases:
    if c == "tiny":
        report("tiny", sm.make_scene(3000, 256, 128, 0.02, 11), sm.identity_view())
    elif c == "odd":
        report("odd", sm.make_scene(20000, 333, 171, 0.03, 12), sm.random_view(5), bg=(1, 1, 1))
    elif c == "colors":
        report("colors", sm.make_scene(20000, 320, 160, 0.03, 13), sm.random_view(6), mode="colors")
    elif c == "cov":
        report("cov", sm.make_scene(20000, 320, 160, 0.03, 14), sm.random_view(7), mode="cov", degree=1)
    elif c == "simple":
        s = sm.simple_cloud(); s.W, s.H = 400, 200
        report("simple_cloud", s, sm.identity_view())
    elif c in ("C1", "C2", "C3", "C5"):
        report(c, sm.make_config_scene(c), sm.identity_view())
    elif c == "C2view":
        report(c, sm.make_config_scene("C2"), sm.random_view(21))
    else:
        print("unknown case", c)
print("\nparity_report done", flush=True)
