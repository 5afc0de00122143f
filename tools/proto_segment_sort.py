"""CPU prototype (numpy) of the tile sort on column segments proposed in DESIGN.md section 10 — not product code.

Reference order (rasterizer_impl.cu:94-167 + the stable 64-bit sort): instances of the depth-ordered Gaussians, each
rect walked row-major, stably sorted by tile id y * gx + x.
Proposed: (1) emit one COLUMN SEGMENT (x, g) per rect column in depth order and stable-sort the segments by x;
(2) expand every sorted segment along its rows y0..y1-1 and stable-sort the instances by y.  The result is ordered by
(y, x, depth) = the reference's list, with one R-sized pass instead of two.  `python tools/proto_segment_sort.py` checks
the equality on random rect sets (small and pole-sized rects)."""
import numpy as np


def reference_list(rects, gx):
    keys, vals = [], []
    for g, (x0, x1, y0, y1) in enumerate(rects):
        for y in range(y0, y1):
            for x in range(x0, x1):
                keys.append(y * gx + x)
                vals.append(g)
    keys, vals = np.asarray(keys, np.int64), np.asarray(vals, np.int64)
    return vals[np.argsort(keys, kind="stable")]


def segment_list(rects, gx):
    seg_x, seg_g = [], []
    for g, (x0, x1, y0, y1) in enumerate(rects):
        if y1 > y0:
            for x in range(x0, x1):
                seg_x.append(x)
                seg_g.append(g)
    seg_x, seg_g = np.asarray(seg_x, np.int64), np.asarray(seg_g, np.int64)
    seg_g = seg_g[np.argsort(seg_x, kind="stable")]          # pass 1: S = sum of widths items
    inst_y, inst_g = [], []
    for g in seg_g:                                          # pass 2: expansion along y, R items
        _, _, y0, y1 = rects[g]
        inst_y.extend(range(y0, y1))
        inst_g.extend([g] * (y1 - y0))
    inst_y, inst_g = np.asarray(inst_y, np.int64), np.asarray(inst_g, np.int64)
    return inst_g[np.argsort(inst_y, kind="stable")]


def random_rects(rng, n, gx, gy, big_frac=0.05):
    rects = []
    for _ in range(n):
        big = rng.random() < big_frac
        w = int(rng.integers(1, (gx if big else min(gx, 5)) + 1))
        h = int(rng.integers(0, (gy if big else min(gy, 5)) + 1))   # h = 0: culled rows appear in the depth order too
        x0 = int(rng.integers(0, gx - w + 1))
        y0 = int(rng.integers(0, gy - h + 1))
        rects.append((x0, x0 + w, y0, y0 + h))
    return rects


if __name__ == "__main__":
    rng = np.random.default_rng(20240403)
    for gx, gy, n in ((8, 4, 50), (32, 16, 400), (128, 64, 300)):
        for trial in range(5):
            rects = random_rects(rng, n, gx, gy)
            a, b = reference_list(rects, gx), segment_list(rects, gx)
            assert a.shape == b.shape and np.array_equal(a, b), (gx, gy, trial)
            S = sum((x1 - x0) for x0, x1, y0, y1 in rects if y1 > y0)
        print(f"grid {gx}x{gy}: {n} rects, R = {a.size}, column segments = {S} ({a.size / max(S, 1):.1f} instances each): identical lists")
