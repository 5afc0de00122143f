"""Multi-GPU plumbing for the two ways the lonlat path shards.

(a) Data parallel over camera views: every rank holds all Gaussians, renders its own view(s) and the
    per-Gaussian gradients are summed with one all-reduce per tensor (NCCL over NVLink on GPUs, gloo in
    the CPU tests).  The reference itself has no multi-GPU code; the single-GPU equivalent is "sum of the
    per-view gradients computed serially", which is what the tests compare against.
(b) Latitude bands of one panorama: `band_rows` splits the tile rows so that every rank gets about
    the same number of tile instances (pole rows are heavier); each rank then runs
    ogs_lonlat_forward_stage1_band on its rows.
"""
import torch
import torch.distributed as dist

# tensors of the backward 8-tuple the optimiser consumes (236 B per Gaussian at SH degree 3)
OPTIMISED = ("dL_dmeans3D", "dL_dsh", "dL_dopacity", "dL_dscales", "dL_drotations")


def is_distributed():
    return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1


def allreduce_gradients(grads, radii=None, group=None):
    """Sum the optimiser-facing gradients over ranks, in place.  `grads` maps the names of
    RasterizeGaussiansBackwardCUDA's outputs to tensors.  With `radii`, also reduces the densification
    statistics the trainer derives per view (reference src/gaussian_mapper.cpp:427-434,
    src/gaussian_model.cpp:839-853): sum of |dL_dmeans2D.xy|, visibility count, max radius.
    Returns (grads, stats-or-None).  A no-op on a single rank."""
    stats = None
    if radii is not None:
        vis = radii > 0
        stats = {
            "xyz_gradient_accum": torch.where(vis, grads["dL_dmeans2D"][:, :2].norm(dim=-1), torch.zeros_like(radii, dtype=torch.float32)),
            "denom": vis.to(torch.float32),
            "max_radii2D": radii.to(torch.float32),
        }
    if not is_distributed():
        return grads, stats
    work = [dist.all_reduce(grads[n], op=dist.ReduceOp.SUM, group=group, async_op=True) for n in OPTIMISED if grads[n].numel()]
    if stats is not None:
        work.append(dist.all_reduce(stats["xyz_gradient_accum"], op=dist.ReduceOp.SUM, group=group, async_op=True))
        work.append(dist.all_reduce(stats["denom"], op=dist.ReduceOp.SUM, group=group, async_op=True))
        work.append(dist.all_reduce(stats["max_radii2D"], op=dist.ReduceOp.MAX, group=group, async_op=True))
    for w in work:
        w.wait()
    return grads, stats


class GradientBucket:
    """One flat float32 buffer holding, back to back, the five gradient tensors the optimiser consumes
    (dL_dmeans3D [P,3], dL_dsh [P,M,3], dL_dopacity [P,1], dL_dscales [P,3], dL_drotations [P,4]) and the two
    summed densification statistics ([P] each).  RasterizeGaussiansBackwardCUDA(..., out=bucket) lets the library
    write the gradients straight into it, so the data-parallel exchange is ONE all-reduce(SUM) over
    (59 + 2) floats per Gaussian (plus one small all-reduce(MAX) for the radii) instead of eight collectives:
    collective launches are latency-bound, and one large message uses the NVLink/NVSwitch bandwidth better."""

    def __init__(self, P, M, device, peer=None):
        """peer: None = use NVLink peer memory when available (NCCL otherwise), False = always NCCL/gloo."""
        self.P, self.M = int(P), int(M)
        self.peer = None
        self.peer_error = None
        sizes = [("dL_dmeans3D", (P, 3)), ("dL_dsh", (P, M, 3)), ("dL_dopacity", (P, 1)), ("dL_dscales", (P, 3)),
                 ("dL_drotations", (P, 4)), ("xyz_gradient_accum", (P,)), ("denom", (P,))]
        # every section starts 16-byte aligned (the SH rows go through 16-byte bulk copies)
        offs, total = [], 0
        for _, shape in sizes:
            offs.append(total)
            n = 1
            for d in shape:
                n *= int(d)
            total += -(-n // 4) * 4
        self.flat = None
        if peer is not False and is_distributed() and torch.device(device).type == "cuda":
            flat = self._symmetric(total, device)
            # every rank must take the same path (the peer kernel is bracketed by cross-rank barriers)
            ok = torch.tensor([1 if flat is not None else 0], dtype=torch.int32, device=device)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)
            if int(ok.item()) == 1:
                self.flat = flat
            else:
                self.peer = None
        if self.flat is None:
            self.flat = torch.empty((total,), dtype=torch.float32, device=device)
        self.tensors = {}
        for (name, shape), o in zip(sizes, offs):
            n = 1
            for d in shape:
                n *= int(d)
            self.tensors[name] = self.flat[o:o + n].view(shape)
        self.max_radii2D = torch.empty((P,), dtype=torch.float32, device=device)

    def __getitem__(self, name):
        return self.tensors[name]

    def _symmetric(self, total, device):
        """Allocate the bucket in torch.distributed symmetric memory and map every peer's copy into this process.
        Returns None (-> NCCL path) when that is not possible on this system."""
        try:
            import torch.distributed._symmetric_memory as symm_mem
            world = dist.get_world_size()
            if world > 8:
                return None
            flat = symm_mem.empty(total, dtype=torch.float32, device=device)
            hdl = symm_mem.rendezvous(flat, dist.group.WORLD)
            ptrs = [int(p) for p in hdl.buffer_ptrs]
            if len(ptrs) != world or any(p == 0 or p % 16 for p in ptrs):
                return None
            mc = 0
            try:
                mc = int(hdl.multicast_ptr) if hdl.has_multicast_support(torch.device(device).type, torch.device(device).index or 0) else 0
            except Exception:
                try:
                    mc = int(hdl.multicast_ptr)
                except Exception:
                    mc = 0
            self.peer = dict(handle=hdl, ptrs=ptrs, rank=dist.get_rank(), world=world, multicast=mc)
            return flat
        except Exception as e:   # no symmetric memory on this build / topology: remember why, use NCCL
            self.peer = None
            self.peer_error = repr(e)
            return None


def fill_view_stats(bucket, dL_dmeans2D, radii):
    """Write one view's densification-statistics increments into the bucket's slots (one kernel on the GPU):
    ||dL_dmeans2D.xy|| and 1 where radii > 0, and the radius as float."""
    if bucket.flat.is_cuda:
        import ctypes
        from ._lib import load_library, check
        ptr = lambda t: ctypes.c_void_p(t.data_ptr())
        check(load_library().ogs_view_stats(
            bucket.P, ptr(radii.contiguous()), ptr(dL_dmeans2D.contiguous()), ptr(bucket["xyz_gradient_accum"]),
            ptr(bucket["denom"]), ptr(bucket.max_radii2D),
            ctypes.c_void_p(torch.cuda.current_stream(bucket.flat.device).cuda_stream)))
    else:   # gloo tests on CPU tensors
        vis = radii > 0
        torch.where(vis, dL_dmeans2D[:, :2].norm(dim=-1), dL_dmeans2D.new_zeros(()), out=bucket["xyz_gradient_accum"])
        bucket["denom"].copy_(vis)
        bucket.max_radii2D.copy_(radii)


def exchange_bucket(bucket, group=None):
    """Sum the bucket (gradients + summed statistics) and take the maximum of the radii over ranks, in place.
    One kernel over NVLink peer memory / the multicast address when the bucket lives in symmetric memory
    (csrc/peer_collective.cu), NCCL / gloo all-reduce otherwise.  A no-op on a single rank."""
    if is_distributed() and bucket.peer is not None:
        # reduce-scatter + all-gather in one kernel, bracketed by the symmetric-memory barrier on this stream:
        # all buckets written before, all slices stored after
        import ctypes
        from ._lib import load_library, check
        pr = bucket.peer
        w2 = dist.all_reduce(bucket.max_radii2D, op=dist.ReduceOp.MAX, group=group, async_op=True)
        arr = (ctypes.c_void_p * pr["world"])(*pr["ptrs"])
        stream = ctypes.c_void_p(torch.cuda.current_stream(bucket.flat.device).cuda_stream)
        pr["handle"].barrier(channel=0)
        if pr.get("multicast") and pr.get("use_multimem", pr["world"] > 4):
            # NVSwitch in-switch reduction: one inbound copy per element instead of world - 1.  Measured on the 244 MB
            # bucket: peer loads/stores win up to four ranks (2: 0.36 ms; 4: 0.575 against 0.598), the multicast path
            # beyond (8: 0.63 against 0.74)
            check(load_library().ogs_multimem_allreduce_sum(ctypes.c_void_p(pr["multicast"]), pr["world"], pr["rank"],
                                                            bucket.flat.numel(), stream))
        else:
            check(load_library().ogs_peer_allreduce_sum(arr, pr["world"], pr["rank"], bucket.flat.numel(), stream))
        pr["handle"].barrier(channel=1)
        w2.wait()
    elif is_distributed():
        w1 = dist.all_reduce(bucket.flat, op=dist.ReduceOp.SUM, group=group, async_op=True)
        w2 = dist.all_reduce(bucket.max_radii2D, op=dist.ReduceOp.MAX, group=group, async_op=True)
        w1.wait()
        w2.wait()


def allreduce_bucket(bucket, dL_dmeans2D, radii, group=None):
    """One view per rank and step: fill the statistics slots of `bucket` from this rank's view and sum the whole
    bucket over ranks in one collective.  Returns (dict of the five summed gradients, stats dict) like
    allreduce_gradients.  (Several views per rank: fill_view_stats + local accumulation + exchange_bucket,
    see tools/bench_dp_views.py.)"""
    fill_view_stats(bucket, dL_dmeans2D, radii)
    exchange_bucket(bucket, group)
    grads = {n: bucket[n] for n in OPTIMISED}
    stats = {"xyz_gradient_accum": bucket["xyz_gradient_accum"], "denom": bucket["denom"], "max_radii2D": bucket.max_radii2D}
    return grads, stats


def tile_row_counts(ranges, W, H):
    """Instances per tile row from a frame's `ranges` [T,2] (e.g. of the previous frame of a sequence)."""
    gx, gy = (W + 15) // 16, (H + 15) // 16
    lens = (ranges[:, 1] - ranges[:, 0]).to(torch.int64).view(gy, gx)
    return lens.sum(dim=1).tolist()


def render_band_forward(rasterize, band, H, group=None):
    """Latitude-band forward for this rank.  `rasterize(band)` must call RasterizeGaussiansCUDA(...,
    band=band) and return its 6-tuple.  Pixel rows outside the band are zeroed and the per-rank images
    are summed, so every rank ends with the full frame (the loss of the trainer needs it whole; a band-
    wise loss would only need an 11x11-SSIM halo of 5 rows).  Returns (image, forward_tuple)."""
    fwd = rasterize(band)
    img = fwd[1]
    y0, y1 = min(H, band[0] * 16), min(H, band[1] * 16)
    full = torch.zeros_like(img)
    full[:, y0:y1] = img[:, y0:y1]
    if is_distributed():
        dist.all_reduce(full, op=dist.ReduceOp.SUM, group=group)
    return full, fwd


def reduce_band_gradients(grads, group=None):
    """Sum every rank's band share of the backward 8-tuple (all of them are per-band partial sums,
    including dL_dmeans2D / dL_dcolors that the data-parallel path keeps local)."""
    if is_distributed():
        work = [dist.all_reduce(g, op=dist.ReduceOp.SUM, group=group, async_op=True) for g in grads if g.numel()]
        for w in work:
            w.wait()
    return grads


def views_for_rank(num_views, rank, world):
    """Round-robin assignment of a step's views: rank g renders views g, g+G, ..."""
    return list(range(rank, num_views, world))


def band_rows(row_instance_counts, world):
    """Split tile rows [0, gy) into `world` contiguous bands with balanced instance counts.
    row_instance_counts[y] = number of tile instances in tile row y.  Returns [(y0, y1)] * world;
    bands may be empty when world > gy."""
    counts = [int(c) for c in row_instance_counts]
    gy, total = len(counts), sum(counts)
    bands, y, acc = [], 0, 0
    for g in range(world):
        target = total * (g + 1) / world
        y0 = y
        while y < gy and (acc + counts[y] <= target or y == y0) and (gy - y) > (world - g - 1):
            acc += counts[y]
            y += 1
        if g == world - 1:
            y = gy
        bands.append((y0, y))
    return bands
