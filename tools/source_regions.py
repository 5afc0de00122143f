"""Per-region instruction and stall-sample shares of one kernel from an ncu report's source page (SASS view): consecutive
instructions with the same execution count form a region (a loop level / branch arm).
  python tools/source_regions.py report.ncu-rep kernel_regex [min_share_pct]"""
import csv, io, subprocess, sys

rep, pat = sys.argv[1], sys.argv[2]
min_share = float(sys.argv[3]) if len(sys.argv) > 3 else 0.3
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "-k", f"regex:{pat}"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[1]
ia, isrc, iex, ipr, ismp = (hdr.index(n) for n in ("Address", "Source", "Instructions Executed",
                                                    "Predicated-On Thread Instructions Executed", "# Samples"))
data = [(int(r[ia], 16), r[isrc].strip(), int(r[iex]), int(r[ipr]), int(r[ismp])) for r in rows[2:] if len(r) > ismp]
base = data[0][0]
tot, tots = sum(d[2] for d in data), sum(d[4] for d in data)
print(f"kernel {rows[0][1][:90]}\n{tot / 1e6:.1f} M warp instructions, {tots} stall samples\n")
print("| SASS offsets | instructions | executions (M) | warp inst (M) | share | stall samples | active lanes | first instruction |")
print("|---|---|---|---|---|---|---|---|")
seg = []
def flush():
    if not seg:
        return
    n, inst, smp, thr = len(seg), sum(d[2] for d in seg), sum(d[4] for d in seg), sum(d[3] for d in seg)
    if 100.0 * inst / tot >= min_share:
        print(f"| {seg[0][0] - base:#06x}-{seg[-1][0] - base:#06x} | {n} | {seg[0][2] / 1e6:.3f} | {inst / 1e6:.1f} | {100 * inst / tot:.1f} % | "
              f"{100 * smp / tots:.1f} % | {thr / max(inst, 1):.1f} | `{seg[0][1][:40]}` |")
for d in data:
    if seg and abs(d[2] - seg[-1][2]) > 0.02 * max(d[2], seg[-1][2], 1):
        flush(); seg = []
    seg.append(d)
flush()
