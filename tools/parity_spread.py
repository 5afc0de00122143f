"""Spread of the gradient-parity statistics over repeated runs (float atomics are unordered, so every run differs): for each
reference case of tests/test_parity_gpu.py, `reps` parity reports with different upstream gradients; prints and stores the
worst value of every statistic the tests bound.   python tools/parity_spread.py out.json [reps] [case ...]"""
import json, os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden"))
import torch
import _harness as h
import test_parity_gpu as T

out_path = sys.argv[1]
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
cases = sys.argv[3:] or list(T.REF_CASES)
res = {}
for case in cases:
    scene, view, mode, bg, degree = T.REF_CASES[case]()
    worst, runs = {}, []
    for r in range(reps):
        rep = h.parity_report(scene, view, mode=mode, bg=bg, degree=degree, dL_seed=99 + r)
        runs.append({n: {k: {f: st[f] for f in ("rel", "excess")} for k, st in row.items()} for n, row in rep["tensors"].items()})
        for n, row in rep["tensors"].items():
            w = worst.setdefault(n, {})
            for k, st in row.items():
                for f in ("rel", "excess"):
                    key = f"{k}.{f}"
                    w[key] = max(w.get(key, -1.0), st[f])
        torch.cuda.empty_cache()
    res[case] = {"worst": worst, "runs": runs}
    print(f"== {case} ({reps} runs)")
    for n, w in worst.items():
        print(f"  {n:14s} " + "  ".join(f"{k}={v:+.2e}" for k, v in w.items() if k.endswith("rel") or k in ("ours_vs_ref.excess", "ref_vs_ref.excess")), flush=True)
json.dump(res, open(out_path, "w"), indent=1)
