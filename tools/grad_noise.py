#!/usr/bin/env python
"""Gradient parity against the live reference with the reference's own run-to-run noise beside it.

For each BASELINE config (C1..C5) and each of the 8 gradient tensors: ours-vs-reference and reference-vs-reference
(second backward on the same forward state; its float atomics are unordered) in the max norm
(max|a-b| / max|b|), the SURVEY 8(c) per-element excess (max of |a-b| - (1e-4|b| + 1e-6 max|b|); <= 0 passes),
the fraction of elements over the per-element bound, ours-vs-ours, and for the outputs of the per-Gaussian chain the
distance of each implementation from a DOUBLE evaluation of that chain on its own inputs (tests/_harness.py,
config_parity_report).  Run once per library build
(OMNIGS_B200_LIB selects an A/B build, --tag names it); results are merged into --out.

  python tools/grad_noise.py [--configs C1 C2 ...] [--tag default] [--out gpurun_out/r02_grad_noise.json]
"""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import torch
import _harness as h

sm = h.scene_mod


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--configs", nargs="*", default=["C1", "C2", "C3", "C4", "C5"])
    ap.add_argument("--tag", default="default")
    ap.add_argument("--out", default="gpurun_out/r02_grad_noise.json")
    args = ap.parse_args()
    res = {}
    if os.path.exists(args.out):
        res = json.load(open(args.out))
    mine = res.setdefault(args.tag, {"library": os.environ.get("OMNIGS_B200_LIB", "libomnigs_b200.so")})
    for name in args.configs:
        t0 = time.time()
        print(f"{args.tag} {name}", flush=True)
        mine[name] = h.config_parity_report(name, verbose=True)
        print(f"  integers {mine[name]['integers']}  R={mine[name]['num_rendered']} ({time.time() - t0:.0f} s)", flush=True)
        torch.cuda.empty_cache()
        os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
        json.dump(res, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
