#!/usr/bin/env python
"""What the tail of the two blend kernels costs ("no tail measurement at C4/C5 is committed").

Runs with the OGS_TILE_TIMELINE build of the library (tools/build_variant.sh timeline -DOGS_TILE_TIMELINE), whose blend
kernels stamp %globaltimer at the start and end of every tile.  For each config and tile order (OGS_TILE_ORDER: 0 =
row-major, 1 = tile rows alternately from both poles) it reports, per kernel:
  span_us        first tile start .. last tile end
  work_us        sum of the tile durations / peak concurrency   (the span a perfectly packed schedule would need)
  tail_us        span - time at which the last tile was DISPATCHED: from then on SMs only drain
  drain_loss     idle CTA-slot time inside the tail / (span x peak concurrency): the fraction of the kernel's capacity
                 lost to the tail
  longest_tile_us, median_tile_us, per-row mean tile time at the poles and at the equator.

  OMNIGS_B200_LIB=gpurun_variants/libomnigs_b200_timeline.so python tools/tile_timeline.py C2 C4 C5 > profiles/r02_tile_timeline.json
"""
import ctypes, json, os, subprocess, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import numpy as np, torch
import _harness as h
sm = h.scene_mod


def analyse(clk, gx, gy):
    st, en = clk[:, 0].astype(np.int64), clk[:, 1].astype(np.int64)
    ok = en > 0
    st, en = st[ok], en[ok]
    t0 = st.min()
    st, en = st - t0, en - t0
    span = float(en.max())
    dur = (en - st).astype(np.float64)
    # concurrency over time
    ev = np.concatenate([np.stack([st, np.ones_like(st)], 1), np.stack([en, -np.ones_like(en)], 1)])
    ev = ev[np.argsort(ev[:, 0], kind="stable")]
    conc = np.cumsum(ev[:, 1])
    peak = int(conc.max())
    last_dispatch = float(st.max())
    # idle slot-time after the last dispatch
    t = ev[:, 0].astype(np.float64)
    idle = 0.0
    for i in range(len(t) - 1):
        if t[i + 1] > last_dispatch:
            a = max(t[i], last_dispatch)
            idle += (peak - conc[i]) * (t[i + 1] - a)
    rows = dur.reshape(gy, gx).mean(axis=1) if ok.all() else None
    out = {"span_us": span / 1e3, "work_us": dur.sum() / peak / 1e3, "peak_concurrent_tiles": peak,
           "tail_us": (span - last_dispatch) / 1e3, "drain_loss": idle / (span * peak),
           "longest_tile_us": dur.max() / 1e3, "median_tile_us": float(np.median(dur)) / 1e3}
    if rows is not None:
        out["row_mean_tile_us"] = {"top_pole": rows[0] / 1e3, "equator": rows[gy // 2] / 1e3, "bottom_pole": rows[-1] / 1e3}
    return out


def run(cfg):
    lib = h.pkg.load_library()
    raw = ctypes.CDLL(h.pkg.library_path())
    scene = sm.make_config_scene(cfg)
    d = h.torch_inputs(scene, sm.random_view(21) if cfg in ("C2", "C3") else sm.identity_view())
    dL = torch.from_numpy(sm.make_grad_image(scene.W, scene.H, 99)).cuda()
    gx, gy = (scene.W + 15) // 16, (scene.H + 15) // 16
    T = gx * gy
    cf = torch.zeros((T, 2), dtype=torch.int64, device="cuda")
    cb = torch.zeros((T, 2), dtype=torch.int64, device="cuda")
    assert raw.ogs_debug_set_fwd_tile_clock(ctypes.c_void_p(cf.data_ptr())) == 0
    assert raw.ogs_debug_set_bwd_tile_clock(ctypes.c_void_p(cb.data_ptr())) == 0
    for _ in range(3):
        f = h.run_forward(h.pkg, d); h.run_backward(h.pkg, d, f, dL)
    torch.cuda.synchronize()
    cf.zero_(); cb.zero_()
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    e[0].record(); f = h.run_forward(h.pkg, d); e[1].record(); h.run_backward(h.pkg, d, f, dL); e[2].record()
    torch.cuda.synchronize()
    return {"num_rendered": f[0], "tiles": [gx, gy], "forward_ms": e[0].elapsed_time(e[1]), "backward_ms": e[1].elapsed_time(e[2]),
            "render_fwd": analyse(cf.cpu().numpy(), gx, gy), "render_bwd": analyse(cb.cpu().numpy(), gx, gy)}


if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "--one":
        print(json.dumps(run(sys.argv[2])))
        sys.exit(0)
    res = {}
    for cfg in (sys.argv[1:] or ["C2", "C4", "C5"]):
        res[cfg] = {}
        for order in ("0", "1"):
            env = dict(os.environ, OGS_TILE_ORDER=order)
            out = subprocess.run([sys.executable, __file__, "--one", cfg], env=env, capture_output=True, text=True)
            res[cfg]["row_major" if order == "0" else "poles_first"] = json.loads(out.stdout.strip().splitlines()[-1]) if out.returncode == 0 else {"error": out.stderr[-400:]}
    print(json.dumps(res, indent=1))
