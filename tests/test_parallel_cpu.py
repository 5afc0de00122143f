"""CPU: the multi-GPU plumbing on the gloo backend (world_size 2) and the latitude-band planner."""
import importlib
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import _harness as h  # noqa: F401  (registers the package)

par = importlib.import_module("omnigs-fork_b200.parallel")
rp = importlib.import_module("omnigs-fork_b200.rasterize_points")
NAMES = h.GRAD_NAMES


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _make(rank, P=257, M=16):
    g = torch.Generator().manual_seed(100 + rank)
    shapes = {"dL_dmeans2D": (P, 3), "dL_dcolors": (P, 3), "dL_dopacity": (P, 1), "dL_dmeans3D": (P, 3),
              "dL_dcov3D": (P, 6), "dL_dsh": (P, M, 3), "dL_dscales": (P, 3), "dL_drotations": (P, 4)}
    grads = {n: torch.randn(s, generator=g) for n, s in shapes.items()}
    radii = torch.randint(0, 40, (P,), generator=g, dtype=torch.int32)
    return grads, radii


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    grads, radii = _make(rank)
    grads, stats = par.allreduce_gradients(grads, radii)
    if rank == 0:
        torch.save({"grads": grads, "stats": stats}, out)
    dist.barrier()
    dist.destroy_process_group()


def test_allreduce_gradients_world2_gloo(tmp_path):
    out = str(tmp_path / "r0.pt")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    got = torch.load(out)
    (g0, r0), (g1, r1) = _make(0), _make(1)
    for n in NAMES:
        if n in par.OPTIMISED:      # summed: what the optimiser consumes
            assert torch.allclose(got["grads"][n], g0[n] + g1[n], atol=1e-6), n
        else:                        # per-view quantities stay local
            assert torch.equal(got["grads"][n], g0[n]), n
    acc = sum(torch.where(r > 0, g["dL_dmeans2D"][:, :2].norm(dim=-1), torch.zeros(r.shape)) for g, r in ((g0, r0), (g1, r1)))
    assert torch.allclose(got["stats"]["xyz_gradient_accum"], acc, atol=1e-6)
    assert torch.equal(got["stats"]["denom"], (r0 > 0).float() + (r1 > 0).float())
    assert torch.equal(got["stats"]["max_radii2D"], torch.maximum(r0, r1).float())


def _bucket_worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    grads, radii = _make(rank)
    bucket = par.GradientBucket(257, 16, "cpu")
    for n in par.OPTIMISED:          # what the library does through RasterizeGaussiansBackwardCUDA(..., out=bucket)
        bucket[n].copy_(grads[n])
    summed, stats = par.allreduce_bucket(bucket, grads["dL_dmeans2D"], radii)
    if rank == 0:
        torch.save({"grads": {k: v.clone() for k, v in summed.items()}, "stats": {k: v.clone() for k, v in stats.items()}}, out)
    dist.barrier()
    dist.destroy_process_group()


def test_bucket_allreduce_matches_per_tensor_allreduce_world2_gloo(tmp_path):
    out = str(tmp_path / "b0.pt")
    mp.spawn(_bucket_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    got = torch.load(out)
    (g0, r0), (g1, r1) = _make(0), _make(1)
    for n in par.OPTIMISED:
        assert torch.allclose(got["grads"][n], g0[n] + g1[n], atol=1e-6), n
    acc = sum(torch.where(r > 0, g["dL_dmeans2D"][:, :2].norm(dim=-1), torch.zeros(r.shape)) for g, r in ((g0, r0), (g1, r1)))
    assert torch.allclose(got["stats"]["xyz_gradient_accum"], acc, atol=1e-6)
    assert torch.equal(got["stats"]["denom"], (r0 > 0).float() + (r1 > 0).float())
    assert torch.equal(got["stats"]["max_radii2D"], torch.maximum(r0, r1).float())


def test_bucket_sections_are_aligned_views_of_one_buffer():
    b = par.GradientBucket(1001, 16, "cpu")
    base = b.flat.data_ptr()
    for n, t in b.tensors.items():
        assert t.is_contiguous() and (t.data_ptr() - base) % 16 == 0, n
        assert base <= t.data_ptr() < base + b.flat.numel() * 4
    assert b["dL_dsh"].shape == (1001, 16, 3) and b["dL_dopacity"].shape == (1001, 1)


def test_allreduce_is_identity_on_one_rank():
    grads, radii = _make(0)
    ref = {k: v.clone() for k, v in grads.items()}
    out, stats = par.allreduce_gradients(grads, radii)
    for n in NAMES:
        assert torch.equal(out[n], ref[n])
    assert stats["denom"].sum() == (radii > 0).sum()


def test_tile_row_costs_weigh_instances_and_visited_pairs():
    """tile_row_costs: per tile row, 15 ps per instance (ranges) + 1.3 ps per list entry a pixel walks (n_contrib), with the
    last tile row of an image whose height is not a multiple of 16 counted over its real pixel rows only."""
    W, H = 48, 40                               # 3 x 3 tiles, the last tile row has 8 pixel rows
    ranges = torch.tensor([[0, 10], [10, 10], [10, 30], [30, 31], [31, 31], [31, 40], [40, 40], [40, 40], [40, 100]], dtype=torch.int32)
    n_contrib = torch.zeros((H, W), dtype=torch.int32)
    n_contrib[0:16] = 2
    n_contrib[16:32] = 1
    n_contrib[32:40] = 5
    costs = par.tile_row_costs(ranges, n_contrib.flatten(), W, H)
    inst = [30, 10, 60]
    visited = [2 * 16 * W, 1 * 16 * W, 5 * 8 * W]
    assert costs == pytest.approx([15.0 * i + 1.3 * v for i, v in zip(inst, visited)])
    assert par.tile_row_counts(ranges, W, H) == inst


def test_feedback_rebalancing_moves_boundaries_towards_equal_times():
    """rescale_row_costs + band_rows: a frame whose true cost per instance doubles towards the equator, bands first cut on
    instances alone; three feedback rounds with the bands' true times bring max/mean from 1.16 to below 1.03."""
    gy, world = 240, 4
    inst = np.full(gy, 1000.0)
    true = inst * (1.0 + np.sin(np.linspace(0, np.pi, gy)))           # what a row really costs
    costs = inst.tolist()
    bands = par.band_rows(costs, world)
    first = [true[a:b].sum() for a, b in bands]
    for _ in range(3):
        measured = [true[a:b].sum() for a, b in bands]
        costs = par.rescale_row_costs(costs, bands, measured)
        for (a, b), t in zip(bands, measured):
            assert abs(sum(costs[a:b]) - t) < 1e-6 * t                 # every band's estimate now equals its measurement
        bands = par.band_rows(costs, world)
    last = [true[a:b].sum() for a, b in bands]
    assert max(first) / np.mean(first) > 1.15 and max(last) / np.mean(last) < 1.03
    assert bands[0][0] == 0 and bands[-1][1] == gy and all(x[1] == y[0] for x, y in zip(bands, bands[1:]))


def test_views_round_robin():
    assert par.views_for_rank(8, 1, 4) == [1, 5]
    assert sorted(sum((par.views_for_rank(8, r, 3) for r in range(3)), [])) == list(range(8))


def test_band_rows_balances_instances_and_covers_all_rows():
    rng = np.random.default_rng(0)
    counts = rng.integers(100, 1000, 64)
    counts[:6] *= 8      # heavy polar rows
    counts[-6:] *= 8
    for world in (1, 2, 3, 4, 8):
        bands = par.band_rows(counts, world)
        assert len(bands) == world and bands[0][0] == 0 and bands[-1][1] == 64
        assert all(a[1] == b[0] for a, b in zip(bands, bands[1:]))
        loads = [int(counts[a:b].sum()) for a, b in bands]
        assert max(loads) <= 1.6 * sum(loads) / world + counts.max()
    assert par.band_rows([5, 5], 4)[-1][1] == 2   # more ranks than rows: trailing bands may be empty


def _view_data(v, P=257):
    """Fake per-view backward outputs: geometry gradients, radii, the clamp-masked dL/dRGB factor and a camera centre."""
    grads, radii = _make(v, P)
    g = torch.Generator().manual_seed(900 + v)
    drgb = torch.randn((P, 3), generator=g) * (radii > 0)[:, None]
    campos = torch.randn(3, generator=g) * 0.3
    return grads, radii, drgb, campos


def _means(P=257):
    return torch.randn((P, 3), generator=torch.Generator().manual_seed(7)) * 5


def _write_view(bucket, slot, grads, radii, drgb):
    """What RasterizeGaussiansBackwardView does on the GPU (ogs_lonlat_backward_view): the step's first view writes the
    geometry gradients and statistics, later views are added; the view's dL/dRGB factor goes to its slot."""
    vis = radii > 0
    gn = torch.where(vis, grads["dL_dmeans2D"][:, :2].norm(dim=-1), torch.zeros(radii.shape))
    if slot == 0:
        for n in ("dL_dmeans3D", "dL_dopacity", "dL_dscales", "dL_drotations"):
            bucket[n].copy_(grads[n])
        bucket["xyz_gradient_accum"].copy_(gn)
        bucket["denom"].copy_(vis)
        bucket.max_radii2D.copy_(radii)
    else:
        for n in ("dL_dmeans3D", "dL_dopacity", "dL_dscales", "dL_drotations"):
            bucket[n].add_(grads[n])
        bucket["xyz_gradient_accum"].add_(gn)
        bucket["denom"].add_(vis)
        torch.maximum(bucket.max_radii2D, radii.float(), out=bucket.max_radii2D)
    bucket["dL_drgb"][slot].copy_(drgb)


def _multiview_worker(rank, world, port, out, views):
    """One step = `views` views (BASELINE configs[2]): every rank accumulates its own views into the step's factored
    bucket inside the backward, ONE exchange sums the ranks and rebuilds dL_dsh from all views' dL/dRGB factors."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = par.views_for_rank(views, rank, world)
    vpr = -(-views // world)
    bucket = par.GradientBucket(257, 16, "cpu", views_per_rank=vpr)
    if not mine:
        bucket.zero_step()
    for k, v in enumerate(mine):
        grads, radii, drgb, _ = _view_data(v)
        _write_view(bucket, k, grads, radii, drgb)
    campos = torch.stack([_view_data(v)[3] if v < views else torch.zeros(3) for v in range(vpr * world)])
    par.exchange_bucket(bucket, means3D=_means(), campos_views=campos, degree=3)
    torch.save({n: bucket[n].clone() for n in ("dL_dmeans3D", "dL_dopacity", "dL_dscales", "dL_drotations", "dL_dsh",
                                               "xyz_gradient_accum", "denom", "max_radii2D")}, out + f".{rank}")
    dist.barrier()
    dist.destroy_process_group()


def _check_multiview(out, world, views):
    per_view = [_view_data(v) for v in range(views)]
    means = _means()
    got = [torch.load(out + f".{r}") for r in range(world)]
    for r in range(1, world):                       # replicas end bit-identical
        for n, t in got[0].items():
            assert torch.equal(t, got[r][n]), (n, r)
    g = got[0]
    for n in ("dL_dmeans3D", "dL_dopacity", "dL_dscales", "dL_drotations"):
        assert torch.allclose(g[n], sum(pv[0][n] for pv in per_view), atol=1e-5), n
    # dL_dsh = sum over views of b(direction) (x) dL/dRGB: the per-view rows the reference's backward writes
    # (backward.cu:60-150: dL_dsh[k] = basis_k * dL_dRGB)
    sh = torch.zeros(257, 16, 3)
    for _, _, drgb, campos in per_view:
        d = means - campos[None, :]
        sh += par.sh_weights(d / d.norm(dim=-1, keepdim=True), 3)[:, :, None] * drgb[:, None, :]
    assert torch.allclose(g["dL_dsh"], sh, atol=1e-5)
    acc = sum(torch.where(r > 0, gr["dL_dmeans2D"][:, :2].norm(dim=-1), torch.zeros(r.shape)) for gr, r, _, _ in per_view)
    assert torch.allclose(g["xyz_gradient_accum"], acc, atol=1e-5)
    assert torch.equal(g["denom"], sum((r > 0).float() for _, r, _, _ in per_view))
    assert torch.equal(g["max_radii2D"], torch.stack([r for _, r, _, _ in per_view]).max(dim=0).values.float())


def test_multi_view_step_accumulates_locally_and_exchanges_once_world2_gloo(tmp_path):
    out, views = str(tmp_path / "mv.pt"), 4
    mp.spawn(_multiview_worker, args=(2, _free_port(), out, views), nprocs=2, join=True)
    _check_multiview(out, 2, views)


def test_multi_view_step_with_fewer_views_than_ranks_world2_gloo(tmp_path):
    """A rank without a view in the step contributes zeros (it used to exchange an uninitialised bucket)."""
    out = str(tmp_path / "mv1.pt")
    mp.spawn(_multiview_worker, args=(2, _free_port(), out, 1), nprocs=2, join=True)
    _check_multiview(out, 2, 1)


def _subgroup_worker(rank, world, port, out):
    """a bucket built for a sub-group must be exchanged over exactly that group."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    groups = [dist.new_group([0, 1]), dist.new_group([2, 3])]
    grp = groups[rank // 2]
    grads, radii = _make(rank)
    bucket = par.GradientBucket(257, 16, "cpu", group=grp)
    for n in par.OPTIMISED:
        bucket[n].copy_(grads[n])
    par.allreduce_bucket(bucket, grads["dL_dmeans2D"], radii)
    try:
        par.exchange_bucket(bucket, group=groups[1 - rank // 2])
        wrong_group_refused = False
    except RuntimeError:
        wrong_group_refused = True
    torch.save({"m3": bucket["dL_dmeans3D"].clone(), "radii": bucket.max_radii2D.clone(), "refused": wrong_group_refused}, out + f".{rank}")
    dist.barrier()
    dist.destroy_process_group()


def test_bucket_honours_its_process_group_world4_gloo(tmp_path):
    out = str(tmp_path / "sg.pt")
    mp.spawn(_subgroup_worker, args=(4, _free_port(), out), nprocs=4, join=True)
    data = [_make(r) for r in range(4)]
    for r in range(4):
        got = torch.load(out + f".{r}")
        a, b = (0, 1) if r < 2 else (2, 3)
        assert torch.allclose(got["m3"], data[a][0]["dL_dmeans3D"] + data[b][0]["dL_dmeans3D"], atol=1e-6)
        assert torch.equal(got["radii"], torch.maximum(data[a][1], data[b][1]).float())
        assert got["refused"]


def _band_worker(rank, world, port, out):
    """SURVEY 8(e-b) on gloo: each rank owns a band of pixel rows of the frame and a partial [P,12] accumulator; the
    band exchange all-gathers the rows and sums the accumulators."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    P, W, H = 300, 70, 100
    bands = [(0, 3), (3, 7)]                      # tile rows; the last band ends beyond H (7 * 16 = 112 > 100)
    frame = torch.randn((3, H, W), generator=torch.Generator().manual_seed(5))
    ex = par.BandExchange(P, W, H, "cpu")
    y0, y1 = min(H, bands[rank][0] * 16), min(H, bands[rank][1] * 16)
    mine = torch.full((3, H, W), float("nan"))    # rows outside the band are garbage as far as the exchange knows
    mine[:, y0:y1] = frame[:, y0:y1]
    full, _ = par.render_band_forward(lambda b: (0, mine, None, None, None, None), bands[rank], H, exchange=ex)
    ex.acc.copy_(torch.randn((P, 12), generator=torch.Generator().manual_seed(50 + rank)))
    ex.reduce_accumulators()
    summed = ex.acc.clone()
    # pipelined variant: the same sums range by range (ranges as RasterizeGaussiansBackwardCUDA(accumulator_chunks=3) cuts them)
    ex.acc.copy_(torch.randn((P, 12), generator=torch.Generator().manual_seed(50 + rank)))
    ranges = rp.gaussian_ranges(P, 3)
    events = ex.reduce_accumulators(ex.acc, ranges)
    assert len(events) == len(ranges)
    # band-wise loss: only the 5-row halos cross ranks
    halo_in = torch.full((3, H, W), float("nan"))
    halo_in[:, y0:y1] = frame[:, y0:y1]
    halo_img, _ = par.render_band_forward(lambda b: (0, halo_in, None, None, None, None), bands[rank], H, exchange=ex, halo=5)
    torch.save({"image": full.clone(), "acc": summed, "acc_ranges": ex.acc.clone(), "halo": halo_img.clone(), "rows": (y0, y1),
                "ranges": ranges}, out + f".{rank}")
    dist.barrier()
    dist.destroy_process_group()


def test_band_exchange_world2_gloo(tmp_path):
    out = str(tmp_path / "band.pt")
    mp.spawn(_band_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    frame = torch.randn((3, 100, 70), generator=torch.Generator().manual_seed(5))
    acc = sum(torch.randn((300, 12), generator=torch.Generator().manual_seed(50 + r)) for r in range(2))
    for r in range(2):
        got = torch.load(out + f".{r}")
        assert torch.equal(got["image"], frame)
        assert torch.allclose(got["acc"], acc, atol=1e-6)
        assert torch.equal(got["acc_ranges"], got["acc"])
        assert got["ranges"] == [(0, 128), (128, 128), (256, 44)]
        y0, y1 = got["rows"]
        lo, hi = max(0, y0 - 5), min(100, y1 + 5)
        assert torch.equal(got["halo"][:, lo:hi], frame[:, lo:hi])            # own rows + both neighbours' halos
        rest = torch.cat([got["halo"][:, :lo], got["halo"][:, hi:]], dim=1)
        assert bool(torch.isnan(rest).all())                                   # nothing else was exchanged


def test_band_rows_properties_hold_for_arbitrary_row_loads():
    from hypothesis import given, settings, strategies as st

    @settings(max_examples=200, deadline=None)
    @given(st.lists(st.integers(min_value=0, max_value=10**6), min_size=1, max_size=300), st.integers(min_value=1, max_value=8))
    def check(counts, world):
        bands = par.band_rows(counts, world)
        gy = len(counts)
        assert len(bands) == world and bands[0][0] == 0 and bands[-1][1] == gy
        assert all(0 <= a <= b <= gy for a, b in bands)
        assert all(x[1] == y[0] for x, y in zip(bands, bands[1:]))          # contiguous, disjoint, covering
        if gy >= world:
            assert all(b > a for a, b in bands)                              # nobody idles while rows are left

    check()
