"""Turn ncu reports into the markdown tables of profiles/*_ncu_summary.md (run where ncu is installed; no GPU needed).

  python tools/ncu_summary.py launches gpurun_out/ev_launches.csv
  python tools/ncu_summary.py full gpurun_out/ev_frame.ncu-rep [profiles/kernel_counters.json to write] [source stamp]
"""
import csv, io, json, re, subprocess, sys
from collections import OrderedDict


def short(name):
    name = re.sub(r"^void ", "", name)
    name = re.sub(r"^ogs::", "", name)
    m = re.match(r"([A-Za-z0-9_:]+)(<[^(]*>)?", name)
    base = m.group(1) + (m.group(2) or "") if m else name
    return base.replace("(int)", "").replace("(bool)", "")[:70]


def launches(path):
    rows = [r for r in csv.reader(l for l in open(path) if not l.startswith("==")) if r]
    hdr = rows[0]
    iname, ival, imet = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Name")
    agg = OrderedDict()
    for r in rows[1:]:
        if len(r) <= ival or r[imet] != "gpu__time_duration.sum":
            continue
        k = short(r[iname])
        ns = float(r[ival].replace(",", ""))
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += ns
    total = sum(v[1] for v in agg.values())
    n = sum(v[0] for v in agg.values())
    print(f"## Launch list ({total / 1e6:.1f} ms of device time in {n} launches)\n")
    print("| kernel | launches | avg us | share |\n|---|---|---|---|")
    for k, (c, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{k}` | {c} | {ns / c / 1e3:.1f} | {100 * ns / total:.1f} % |")


def full(path, traffic_path=None, stamp=None):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = rows[0]
    col = {n: i for i, n in enumerate(hdr)}
    def g(r, name, scale=1.0):
        v = r[col[name]].replace(",", "")
        return float(v) * scale if v not in ("", "n/a") else float("nan")
    units = rows[1]
    print("| kernel | us | DRAM read MB | DRAM write MB | DRAM % of peak | warps active % | regs | issue active % | warp inst (M) | L1/shared % |")
    print("|---|---|---|---|---|---|---|---|---|---|")
    traffic = {}
    counters = {}
    for r in rows[2:]:
        name = short(r[col["Kernel Name"]])
        def bytes_of(metric):
            v = g(r, metric)
            u = units[col[metric]]
            return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
        rd, wr = bytes_of("dram__bytes_read.sum"), bytes_of("dram__bytes_write.sum")
        dur = g(r, "gpu__time_duration.sum")
        du = units[col["gpu__time_duration.sum"]]
        dur_us = dur * {"ns": 1e-3, "us": 1, "ms": 1e3, "usecond": 1, "nsecond": 1e-3, "msecond": 1e3}.get(du, 1)
        print(f"| `{name}` | {dur_us:.1f} | {rd / 1e6:.1f} | {wr / 1e6:.1f} | "
              f"{g(r, 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'):.1f} | "
              f"{g(r, 'sm__warps_active.avg.pct_of_peak_sustained_active'):.0f} | {g(r, 'launch__registers_per_thread'):.0f} | "
              f"{g(r, 'sm__inst_issued.avg.pct_of_peak_sustained_active'):.0f} | {g(r, 'inst_executed') / 1e6:.1f} | "
              f"{g(r, 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed'):.0f} |")
        traffic[name] = int(rd + wr)
        red = g(r, "lts__t_sectors_srcunit_tex_op_red.sum") if "lts__t_sectors_srcunit_tex_op_red.sum" in col else float("nan")
        redq = g(r, "lts__t_requests_srcunit_tex_op_red.sum") if "lts__t_requests_srcunit_tex_op_red.sum" in col else float("nan")
        counters.setdefault(name, []).append({
            "duration_us": dur_us, "dram_bytes": int(rd + wr), "warp_instructions": int(g(r, "inst_executed")),
            "issue_active_pct": g(r, "sm__inst_issued.avg.pct_of_peak_sustained_active"),
            "registers": int(g(r, "launch__registers_per_thread")),
            "l2_red_sectors": None if red != red else int(red), "l2_red_requests": None if redq != redq else int(redq),
            "l2_red_bytes": None if red != red else int(red) * 32})
    if traffic_path:
        # profiles/kernel_counters.json: per stage group of bench.py, summed over the group's launches of ONE frame
        groups = {"preprocess_fwd": ["preprocess_lonlat_fwd_kernel"], "render_fwd": ["render_fwd_kernel"],
                  "render_bwd": ["render_bwd_kernel"], "preprocess_bwd": ["preprocess_lonlat_bwd_kernel"],
                  "binning": ["depth_histogram_kernel", "onesweep_pass_kernel", "gather_scan_kernel", "tile_ranges_kernel"]}
        out = {"source_stamp": stamp, "workload": "C2, one frame (tools/dbg_step.py)", "capture": path,
               "how": "ncu --set full --clock-control none; dram_bytes = dram__bytes_read.sum + dram__bytes_write.sum, "
                      "l2_red_sectors = lts__t_sectors_srcunit_tex_op_red.sum (32 B each), per launch, summed per group", "kernels": {}}
        for key, pats in groups.items():
            launches = [c for n, cs in counters.items() if any(n.startswith(p) for p in pats) for c in cs]
            if not launches:
                continue
            agg = {"launches": len(launches)}
            for f in ("duration_us", "dram_bytes", "warp_instructions", "l2_red_sectors", "l2_red_requests", "l2_red_bytes"):
                vals = [c[f] for c in launches if c[f] is not None]
                agg[f] = sum(vals) if vals else None
            tot = sum(c["duration_us"] for c in launches)
            agg["issue_active_pct"] = sum(c["issue_active_pct"] * c["duration_us"] for c in launches) / tot if tot else None
            agg["registers"] = max(c["registers"] for c in launches)
            out["kernels"][key] = agg
        json.dump(out, open(traffic_path, "w"), indent=1)


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2])
    else:
        full(sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else None, sys.argv[4] if len(sys.argv) > 4 else None)
