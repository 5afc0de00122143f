"""Bit-level comparison of the per-Gaussian colours and the image against the live reference (C1)."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import torch, _harness as h
sm = h.scene_mod
ref = h.load_reference()
scene = sm.make_config_scene("C1")
d = h.torch_inputs(scene, sm.identity_view())
fo, fr = h.run_forward(h.pkg, d), h.run_forward(ref, d)
so, sr = h.ours_state(d, fo), h.ref_state(ref, d, fr)
vis = fr[2] > 0
b = lambda t: t.contiguous().view(torch.int32)
print("rgb bit mismatches", int((b(so["rgb"][vis]) != b(sr["rgb"][vis])).sum()), "of", int(vis.sum()) * 3,
      "| image bit mismatches", int((b(fo[1]) != b(fr[1])).sum()), "maxdiff", float((fo[1] - fr[1]).abs().max()))
