"""Generate the golden fixtures tests/golden/*.npz by running the UNMODIFIED reference rasterizer
(oracle/_ref/omnigs_ref.so, built from /root/reference by oracle/build_ref.sh) on a GPU.

Run on the GPU box:   python tests/golden/make_golden.py gpurun_out/golden
then copy gpurun_out/golden/*.npz into tests/golden/ and commit them.
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import cases  # noqa: E402

h = cases.h
out_dir = sys.argv[1] if len(sys.argv) > 1 else os.path.dirname(os.path.abspath(__file__))
os.makedirs(out_dir, exist_ok=True)
ref = h.load_reference()
assert ref is not None, "oracle/_ref/omnigs_ref.so missing: run oracle/build_ref.sh where /root/reference exists"

only = sys.argv[2].split(",") if len(sys.argv) > 2 else None
for name in cases.CASES:
    if only and name not in only:
        continue
    scene, view, dL_np, c = cases.build(name)
    d = h.torch_inputs(scene, view, mode=c["mode"], bg=c["bg"], degree=c["degree"], render_depth=c.get("render_depth", False))
    dL = torch.from_numpy(dL_np).cuda()
    fwd = h.run_forward(ref, d)
    # the reference's backward has no depth path: render_depth fixtures are forward-only
    grads = h.run_backward(ref, d, fwd, dL) if not c.get("render_depth") else []
    torch.cuda.synchronize()
    st = h.ref_state(ref, d, fwd)
    vis = (fwd[2] > 0)
    z = lambda t, m=None: (t if m is None else torch.where(m, t, torch.zeros_like(t))).cpu().numpy()
    m1 = vis
    out = dict(
        input_sha256=np.frombuffer(cases.input_hash(scene, view, dL_np).encode(), dtype=np.uint8),
        num_rendered=np.int64(fwd[0]), out_color=z(fwd[1]), radii=z(fwd[2]),
        means2D=z(st["means2D"], m1[:, None]), depths=z(st["depths"], m1), conic_opacity=z(st["conic_opacity"], m1[:, None]),
        tiles_touched=z(st["tiles_touched"]), ranges=z(st["ranges"]), point_list=z(st["point_list"]),
        point_list_keys=z(st["point_list_keys"]), accum_alpha=z(st["accum_alpha"]), n_contrib=z(st["n_contrib"]))
    if d["camera_type"] == 1:
        out["present"] = ref.markVisible(d["means3D"], d["viewmatrix"], d["projmatrix"], 1).cpu().numpy()
    if c["mode"] != "colors":   # uninitialised in the reference when colours are precomputed
        out["rgb"] = z(st["rgb"], m1[:, None])
        out["clamped"] = z(st["clamped"].to(torch.uint8), m1[:, None])
    if c["mode"] != "cov":
        out["cov3D"] = z(st["cov3D"], m1[:, None])
    for n, g in zip(h.GRAD_NAMES, grads):
        out[n] = z(g)
    path = os.path.join(out_dir, name + ".npz")
    np.savez_compressed(path, **out)
    print(f"{name}: P={scene.P} {scene.W}x{scene.H} R={fwd[0]} visible={int(vis.sum())} -> {path} "
          f"({os.path.getsize(path) / 1e3:.0f} kB)", flush=True)
