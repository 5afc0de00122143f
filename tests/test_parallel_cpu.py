"""CPU: the multi-GPU plumbing on the gloo backend (world_size 2) and the latitude-band planner."""
import importlib
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import _harness as h  # noqa: F401  (registers the package)

par = importlib.import_module("omnigs-fork_b200.parallel")
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


def _multiview_worker(rank, world, port, out, views):
    """One step = `views` views: every rank accumulates its own views into the step's bucket and ONE exchange sums the
    ranks (the flow of tools/bench_dp_views.py / BASELINE configs[2])."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    bucket = par.GradientBucket(257, 16, "cpu")
    work = par.GradientBucket(257, 16, "cpu", peer=False)
    for k, v in enumerate(par.views_for_rank(views, rank, world)):
        grads, radii = _make(v)
        tgt = bucket if k == 0 else work
        for n in par.OPTIMISED:
            tgt[n].copy_(grads[n])
        par.fill_view_stats(tgt, grads["dL_dmeans2D"], radii)
        if k > 0:
            bucket.flat += work.flat
            torch.maximum(bucket.max_radii2D, work.max_radii2D, out=bucket.max_radii2D)
    par.exchange_bucket(bucket)
    if rank == 0:
        torch.save({"flat": bucket.flat.clone(), "max_radii2D": bucket.max_radii2D.clone(),
                    "sh": bucket["dL_dsh"].clone(), "acc": bucket["xyz_gradient_accum"].clone(), "denom": bucket["denom"].clone()}, out)
    dist.barrier()
    dist.destroy_process_group()


def test_multi_view_step_accumulates_locally_and_exchanges_once_world2_gloo(tmp_path):
    out, views = str(tmp_path / "mv.pt"), 4
    mp.spawn(_multiview_worker, args=(2, _free_port(), out, views), nprocs=2, join=True)
    got = torch.load(out)
    per_view = [_make(v) for v in range(views)]
    assert torch.allclose(got["sh"], sum(g["dL_dsh"] for g, _ in per_view), atol=1e-5)
    acc = sum(torch.where(r > 0, g["dL_dmeans2D"][:, :2].norm(dim=-1), torch.zeros(r.shape)) for g, r in per_view)
    assert torch.allclose(got["acc"], acc, atol=1e-5)
    assert torch.equal(got["denom"], sum((r > 0).float() for _, r in per_view))
    assert torch.equal(got["max_radii2D"], torch.stack([r for _, r in per_view]).max(dim=0).values.float())


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
