"""All BASELINE configs on one GPU: ours vs the reference (if present): parity + time."""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import numpy as np, torch
import _harness as h
sm = h.scene_mod
ref = h.load_reference()

def timeit(fn, iters=5, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))

for name in (sys.argv[1:] or ["C1", "C2", "C3", "C4", "C5"]):
    t0 = time.time()
    scene = sm.make_config_scene(name)
    view = sm.random_view(300 + sm.CONFIG_INDEX[name]) if name in ("C3",) else sm.identity_view()
    d = h.torch_inputs(scene, view)
    dL = torch.from_numpy(sm.make_grad_image(scene.W, scene.H, 99)).cuda()
    print(f"=== {name}: P={scene.P} {scene.W}x{scene.H} (scene gen {time.time()-t0:.1f}s)", flush=True)
    fo = h.run_forward(h.pkg, d); go = h.run_backward(h.pkg, d, fo, dL); torch.cuda.synchronize()
    print(f"ours R={fo[0]} visible={int((fo[2]>0).sum())} finite={bool(torch.isfinite(fo[1]).all())} mem={torch.cuda.max_memory_allocated()/1e9:.1f} GB", flush=True)
    tf = timeit(lambda: h.run_forward(h.pkg, d)); tb = timeit(lambda: h.run_backward(h.pkg, d, fo, dL))
    line = f"time ours fwd {tf:.3f} bwd {tb:.3f} total {tf+tb:.3f} ms"
    if ref is not None:
        torch.cuda.reset_peak_memory_stats()
        fr = h.run_forward(ref, d); gr = h.run_backward(ref, d, fr, dL); torch.cuda.synchronize()
        so = h.ours_state(d, fo); sr = h.ref_state(ref, d, fr)
        print("ref R", fr[0], "radii mism", int((fo[2] != fr[2]).sum()), "ranges mism", int((so["ranges"] != sr["ranges"]).sum()),
              "point_list mism", int((so["point_list"] != sr["point_list"]).sum()) if fo[0] == fr[0] else "n/a",
              "keys mism", int((so["point_list_keys"] != sr["point_list_keys"]).sum()) if fo[0] == fr[0] else "n/a",
              f"img maxdiff {float((fo[1]-fr[1]).abs().max()):.2e}", "n_contrib mism", int((so["n_contrib"] != sr["n_contrib"]).sum()),
              f"ref mem={torch.cuda.max_memory_allocated()/1e9:.1f} GB", flush=True)
        print("  grads rel:", " ".join(f"{h.grad_error(a,b)[0]:.1e}" for a, b in zip(go, gr)))
        del so, sr
        rf = timeit(lambda: h.run_forward(ref, d), iters=3, warm=1); rb = timeit(lambda: h.run_backward(ref, d, fr, dL), iters=3, warm=1)
        line += f" | ref fwd {rf:.3f} bwd {rb:.3f} total {rf+rb:.3f} | speedup {(rf+rb)/(tf+tb):.2f}x"
        del fr, gr
    print(line, flush=True)
    del fo, go, d, dL, scene
    torch.cuda.empty_cache()
