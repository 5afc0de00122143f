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


def _numel(shape):
    n = 1
    for d in shape:
        n *= int(d)
    return n


class GradientBucket:
    """The per-step exchange buffer of data-parallel training: ONE flat float32 allocation (symmetric memory when the
    ranks share NVLink, so that our own kernels can read and write every rank's copy) laid out as

        [ summed section | max section | factor section ]

    summed : the gradients the optimiser consumes and the two summed densification statistics ([P] each)
    max    : max_radii2D [P] (radii are non-negative, so the same kernel max-reduces them on their bit patterns)
    factors: (factored mode) this rank's views' clamp-masked dL/dRGB [views_per_rank, P, 3]

    Dense mode (views_per_rank == 0): dL_dsh [P,M,3] sits in the summed section, 61 + 1 floats per Gaussian are exchanged.
    Factored mode (views_per_rank >= 1): row k of dL/dsh is b_k(view direction) * dL/dRGB per view (reference
    backward.cu:60-150), so ranks only publish the 3-float factor per view; every rank rebuilds the step's dL_dsh from ALL
    views' factors, reading the peers' factors straight over NVLink inside the rebuild kernel
    (ogs_sh_gradient_from_views).  13 + 1 floats per Gaussian are all-reduced and 12 bytes per Gaussian and remote view
    are read, instead of 244 bytes; dL_dsh lives in a plain local tensor.  RasterizeGaussiansBackwardView writes a view
    straight into the bucket (accumulating views after the first), statistics included."""

    def __init__(self, P, M, device, peer=None, group=None, views_per_rank=0):
        """peer: None = use NVLink peer memory when available (NCCL otherwise), False = always NCCL/gloo.
        group: the process group the bucket is exchanged over (None = WORLD); the symmetric-memory rendezvous, the
        barriers, the slices and the NCCL fallback all use it."""
        self.P, self.M = int(P), int(M)
        self.group = group
        self.views_per_rank = int(views_per_rank)
        self.peer = None
        self.peer_error = None
        summed = [("dL_dmeans3D", (P, 3))]
        if self.views_per_rank == 0:
            summed.append(("dL_dsh", (P, M, 3)))
        summed += [("dL_dopacity", (P, 1)), ("dL_dscales", (P, 3)), ("dL_drotations", (P, 4)),
                   ("xyz_gradient_accum", (P,)), ("denom", (P,))]
        sizes = summed + [("max_radii2D", (P,))]
        if self.views_per_rank:
            sizes.append(("dL_drgb", (self.views_per_rank, P, 3)))
        # every section starts 16-byte aligned (128-bit accesses, bulk copies of the SH rows)
        offs, total = {}, 0
        for name, shape in sizes:
            offs[name] = total
            total += -(-_numel(shape) // 4) * 4
        self.count_sum = offs["max_radii2D"]
        self.count_max = -(-int(P) // 4) * 4
        self.offsets = offs
        self.flat = None
        distributed = dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
        if peer is not False and distributed and torch.device(device).type == "cuda":
            flat = self._symmetric(total, device)
            # every rank must take the same path (the peer kernel is bracketed by cross-rank barriers)
            ok = torch.tensor([1 if flat is not None else 0], dtype=torch.int32, device=device)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
            if int(ok.item()) == 1:
                self.flat = flat
            else:
                self.peer = None
        if self.flat is None:
            self.flat = torch.empty((total,), dtype=torch.float32, device=device)
        self.tensors = {}
        for name, shape in sizes:
            self.tensors[name] = self.flat[offs[name]:offs[name] + _numel(shape)].view(shape)
        self.max_radii2D = self.tensors["max_radii2D"]
        if self.views_per_rank:
            self.tensors["dL_drgb"].zero_()                      # unused view slots contribute exact zeros
            self.tensors["dL_dsh"] = torch.empty((P, M, 3), dtype=torch.float32, device=device)

    def __getitem__(self, name):
        return self.tensors[name]

    @property
    def factored(self):
        return self.views_per_rank > 0

    def world(self):
        return dist.get_world_size(self.group) if (dist.is_available() and dist.is_initialized()) else 1

    def zero_step(self):
        """A rank that renders no view in a step contributes zeros (its bucket must not hold a previous step)."""
        self.flat.zero_()

    def _symmetric(self, total, device):
        """Allocate the bucket in torch.distributed symmetric memory and map every peer's copy into this process.
        Returns None (-> NCCL path) when that is not possible on this system."""
        try:
            import torch.distributed._symmetric_memory as symm_mem
            grp = self.group if self.group is not None else dist.group.WORLD
            world = dist.get_world_size(grp)
            if world > 8:
                return None
            flat = symm_mem.empty(total, dtype=torch.float32, device=device)
            hdl = symm_mem.rendezvous(flat, grp)
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
            self.peer = dict(handle=hdl, ptrs=ptrs, rank=dist.get_rank(grp), world=world, multicast=mc)
            return flat
        except Exception as e:   # no symmetric memory on this build / topology: remember why, use NCCL
            self.peer = None
            self.peer_error = repr(e)
            return None


def fill_view_stats(bucket, dL_dmeans2D, radii):
    """Write one view's densification-statistics increments into the bucket's slots (one kernel on the GPU):
    ||dL_dmeans2D.xy|| and 1 where radii > 0, and the radius as float.  (RasterizeGaussiansBackwardView does this inside
    the per-Gaussian backward; this is for the dense bucket fed by RasterizeGaussiansBackwardCUDA.)"""
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


SH_C0 = 0.28209479177387814
SH_C1 = 0.4886025119029199
SH_C2 = (1.0925484305920792, -1.0925484305920792, 0.31539156525252005, -1.0925484305920792, 0.5462742152960396)
SH_C3 = (-0.5900435899266435, 2.890611442640554, -0.4570457994644658, 0.3731763325901154, -0.4570457994644658,
         1.445305721320277, -0.5900435899266435)


def sh_weights(dirs, degree):
    """The 16 real SH weights b_k(dir) of the reference's colour evaluation (forward.cu:30-83) for unit vectors
    dirs [...,3] -> [...,16] (zeros beyond (degree+1)^2).  Torch restatement used where no CUDA kernel can run (gloo
    tests on CPU tensors) and as the checker of ogs_sh_gradient_from_views."""
    x, y, z = dirs[..., 0], dirs[..., 1], dirs[..., 2]
    b = [torch.full_like(x, SH_C0)]
    if degree > 0:
        b += [-SH_C1 * y, SH_C1 * z, -SH_C1 * x]
    if degree > 1:
        xx, yy, zz, xy, yz, xz = x * x, y * y, z * z, x * y, y * z, x * z
        b += [SH_C2[0] * xy, SH_C2[1] * yz, SH_C2[2] * (2 * zz - xx - yy), SH_C2[3] * xz, SH_C2[4] * (xx - yy)]
    if degree > 2:
        b += [SH_C3[0] * y * (3 * xx - yy), SH_C3[1] * xy * z, SH_C3[2] * y * (4 * zz - xx - yy),
              SH_C3[3] * z * (2 * zz - 3 * xx - 3 * yy), SH_C3[4] * x * (4 * zz - xx - yy), SH_C3[5] * z * (xx - yy),
              SH_C3[6] * x * (xx - 3 * yy)]
    b += [torch.zeros_like(x)] * (16 - len(b))
    return torch.stack(b, dim=-1)


def sh_gradient_from_views(means3D, campos_views, factors, degree, out):
    """out[P,16,3] = sum over views v (in order) of b(dir_v) (x) factors[v]; factors: list of [P,3] tensors (CUDA tensors
    may live on peer GPUs: pass raw pointers with `factor_ptrs` through rebuild_sh_gradient instead)."""
    if out.is_cuda:
        import ctypes
        from ._lib import load_library, check
        P, M = int(out.size(0)), int(out.size(1))
        ptrs = (ctypes.c_void_p * len(factors))(*[int(f) if isinstance(f, int) else f.data_ptr() for f in factors])
        campos_views = campos_views.contiguous()
        check(load_library().ogs_sh_gradient_from_views(
            P, int(degree), M, len(factors), ctypes.c_void_p(means3D.data_ptr()), ctypes.c_void_p(campos_views.data_ptr()),
            ptrs, ctypes.c_void_p(out.data_ptr()), None, None,
            ctypes.c_void_p(torch.cuda.current_stream(out.device).cuda_stream)))
        return out
    out.zero_()
    for v, f in enumerate(factors):
        d = means3D - campos_views[v][None, :]
        b = sh_weights(d / d.norm(dim=-1, keepdim=True), degree)          # [P,16]
        out += b[:, :out.size(1), None] * f[:, None, :]
    return out


_trace = None   # tools/dp_trace.py sets this to a list: exchange_bucket appends (label, CUDA event) pairs of its phases


def _mark(label, stream):
    if _trace is not None:
        ev = torch.cuda.Event(enable_timing=True)
        ev.record(stream)
        _trace.append((label, ev))


def exchange_bucket(bucket, group=None, means3D=None, campos_views=None, degree=3, sh_stream=None):
    """Sum the bucket's summed section and max-reduce its radii over ranks, in place: one kernel over NVLink peer
    memory / the multicast address when the bucket lives in symmetric memory (csrc/peer_collective.cu), NCCL / gloo
    all-reduce otherwise.  Factored buckets then rebuild bucket["dL_dsh"] from every rank's view factors (means3D and
    campos_views [views_per_rank * world, 3], view v = slot * world + rank, are needed for the directions).  With
    `sh_stream` the rebuild runs on that stream underneath whatever the caller queues next on the current one; the
    returned event marks dL_dsh ready AND every rank done reading this rank's factors — wait for it before reading
    dL_dsh and before the next step's backward.  (Measured on four B200s, profiles/r02_dp_rebuild_sweep_n4.log: the rebuild is
    a bandwidth-bound kernel and costs about its own duration wherever it overlaps the next step's per-Gaussian forward or
    sorts; queueing it later — after that step's stage 1 — or capping its grid so that it trickles along both measured
    slower than starting it right behind the all-reduce.)
    A no-op on a single rank (apart from the rebuild)."""
    group = group if group is not None else bucket.group
    if group is not bucket.group:
        raise RuntimeError("exchange_bucket: the bucket was built for another process group")
    distributed = dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
    world = dist.get_world_size(group) if distributed else 1
    rank = dist.get_rank(group) if distributed else 0
    cs, cm = bucket.count_sum, bucket.count_max
    factor_srcs = None
    pr = bucket.peer if distributed else None
    cur = torch.cuda.current_stream(bucket.flat.device) if bucket.flat.is_cuda else None
    if pr is not None:
        import ctypes
        from ._lib import load_library, check
        stream = ctypes.c_void_p(cur.cuda_stream)
        _mark("backward_done", cur)
        pr["handle"].barrier(channel=0)      # every rank's bucket (and factors) of this step are written
        _mark("barrier0", cur)
        if pr.get("multicast") and pr.get("use_multimem", pr["world"] > 4):
            # NVSwitch in-switch reduction: one inbound copy per element instead of world - 1.  Measured on a 244 MB
            # bucket: peer loads/stores win up to four ranks, the multicast path beyond (profiles/r01_peer_allreduce_*.log)
            check(load_library().ogs_multimem_allreduce(ctypes.c_void_p(pr["multicast"]), pr["world"], pr["rank"], cs, cm, stream))
        else:
            arr = (ctypes.c_void_p * pr["world"])(*pr["ptrs"])
            check(load_library().ogs_peer_allreduce(arr, pr["world"], pr["rank"], cs, cm, stream))
        _mark("allreduce", cur)
        fork = None
        if bucket.factored and sh_stream is not None:
            # the dL_dsh rebuild starts BEHIND the all-reduce: run side by side the two share NVLink and the all-reduce — the
            # one the optimiser waits for — takes 0.36 ms instead of 0.15 (profiles/r02_dp_trace_n8_before.log); behind it the
            # rebuild overlaps the next step's geometry / depth order / tile sort instead
            fork = torch.cuda.Event()
            fork.record(cur)
        if bucket.factored:
            off = bucket.offsets["dL_drgb"] * 4
            step = bucket.P * 3 * 4
            factor_srcs = [pr["ptrs"][r] + off + s * step for s in range(bucket.views_per_rank) for r in range(world)]
            if sh_stream is None:
                sh_gradient_from_views(means3D, campos_views, factor_srcs, degree, bucket["dL_dsh"])
                pr["handle"].barrier(channel=1)   # all sums stored, all factors read
                return None
            sh_stream.wait_event(fork)
            with torch.cuda.stream(sh_stream):
                _mark("sh_fork", sh_stream)
                sh_gradient_from_views(means3D, campos_views, factor_srcs, degree, bucket["dL_dsh"])
                _mark("sh_rebuild", sh_stream)
                pr["handle"].barrier(channel=2)   # every rank has read this rank's factors
                _mark("sh_barrier2", sh_stream)
                done = torch.cuda.Event()
                done.record(sh_stream)
            pr["handle"].barrier(channel=1)       # all sums stored
            _mark("barrier1", cur)
            return done
        pr["handle"].barrier(channel=1)
        return None
    if distributed:
        w1 = dist.all_reduce(bucket.flat[:cs], op=dist.ReduceOp.SUM, group=group, async_op=True)
        w2 = dist.all_reduce(bucket.flat[cs:cs + cm], op=dist.ReduceOp.MAX, group=group, async_op=True)
        gathered = None
        if bucket.factored:
            gathered = [torch.empty_like(bucket["dL_drgb"]) for _ in range(world)]
            dist.all_gather(gathered, bucket["dL_drgb"].contiguous(), group=group)
        w1.wait()
        w2.wait()
        if bucket.factored:
            factor_srcs = [gathered[r][s] for s in range(bucket.views_per_rank) for r in range(world)]
    elif bucket.factored:
        factor_srcs = [bucket["dL_drgb"][s] for s in range(bucket.views_per_rank)]
    if bucket.factored:
        sh_gradient_from_views(means3D, campos_views, factor_srcs, degree, bucket["dL_dsh"])
    return None


def allreduce_bucket(bucket, dL_dmeans2D, radii, group=None):
    """One view per rank and step with a DENSE bucket: fill the statistics slots of `bucket` from this rank's view and
    sum the whole bucket over ranks in one collective.  Returns (dict of the five summed gradients, stats dict) like
    allreduce_gradients."""
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


def tile_row_costs(ranges, n_contrib, W, H, ps_per_instance=15.0, ps_per_visited_pair=1.3):
    """Estimated cost of every tile row of a frame from the previous frame's state: binning work goes with the row's
    instances, blending work with the list entries its pixels walk (sum of n_contrib).  An equatorial instance costs about
    twice a polar one in the blend kernels (profiles/r02_tile_timeline.json), so bands balanced on instances alone leave the
    polar ranks waiting.  The two weights are the C2 frame's measured picoseconds per instance (emission + tile sort) and per
    visited pair (both blend kernels); only their ratio matters.  Returns a list of gy floats for band_rows."""
    gx, gy = (W + 15) // 16, (H + 15) // 16
    inst = (ranges[:, 1] - ranges[:, 0]).to(torch.float64).view(gy, gx).sum(dim=1)
    nc = n_contrib.view(H, W).to(torch.float64).sum(dim=1)
    pad = gy * 16 - H
    if pad:
        nc = torch.cat([nc, nc.new_zeros(pad)])
    visited = nc.view(gy, 16).sum(dim=1)
    return (ps_per_instance * inst + ps_per_visited_pair * visited).tolist()


def rescale_row_costs(costs, bands, measured):
    """Feedback for band_rows: `measured[r]` is what rank r's band-dependent kernels (emission, tile sort, both blend
    kernels) took on the previous frame with the bands `bands` cut from `costs`.  Every row of band r is rescaled so that
    the band's estimate equals its measurement; cutting the result with band_rows again moves the boundaries towards equal
    times.  A static model (tile_row_costs) misjudges scenes whose cost per instance varies with latitude — C4's equatorial
    instances cost 1.6 x the polar ones, where C2's cost twice as much at the poles."""
    out = [float(c) for c in costs]
    for (a, b), t in zip(bands, measured):
        est = sum(out[a:b])
        if b > a and est > 0 and t > 0:
            f = float(t) / est
            for y in range(a, b):
                out[y] *= f
    return out


class BandExchange:
    """Buffers and exchanges of latitude-band rendering: every rank bins / sorts / blends only its
    tile rows of one large panorama and holds all Gaussians.  Per frame the ranks exchange
      forward : their pixel rows — an all-gather of band rows into every rank's image (the trainer's loss wants the
                frame whole); peer stores over NVLink (ogs_band_rows_allgather), NCCL/gloo all_gather otherwise;
      backward: the packed [P,12] render-backward accumulators (48 B/Gaussian, summed) between the two backward kernels,
                through the same NVLink all-reduce kernel as the data-parallel bucket (peer loads/stores up to four
                ranks, multimem beyond), NCCL/gloo otherwise.
    `image` and `acc` live in one symmetric allocation."""

    def __init__(self, P, W, H, device, group=None, peer=None):
        self.P, self.W, self.H, self.group = int(P), int(W), int(H), group
        self.peer = None
        n_acc = -(-12 * self.P // 4) * 4
        total = n_acc + 3 * self.W * self.H
        self.flat = None
        self.distributed = dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
        if peer is not False and self.distributed and torch.device(device).type == "cuda":
            helper = GradientBucket.__new__(GradientBucket)      # reuse the rendezvous
            helper.group, helper.peer, helper.peer_error = group, None, None
            flat = helper._symmetric(total, device)
            ok = torch.tensor([1 if flat is not None else 0], dtype=torch.int32, device=device)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
            if int(ok.item()) == 1:
                self.flat, self.peer = flat, helper.peer
        if self.flat is None:
            self.flat = torch.empty((total,), dtype=torch.float32, device=device)
        self.acc = self.flat[:12 * self.P].view(self.P, 12)
        self.image = self.flat[n_acc:].view(3, self.H, self.W)
        self._img_off = n_acc * 4

    def gather_image(self, band_image, band, halo=None):
        """band_image: this rank's render [3,H,W] (valid in its rows); band = (ty0, ty1) tile rows.  Returns the full
        frame (every rank's rows in place).  halo = h: a band-wise loss needs only the h pixel rows either side of the
        band (5 for the 11-tap SSIM window, loss_utils.h:73-131) — every rank then publishes just the first and last h
        rows of its band and band_image is returned with its neighbours' rows [y0 - h, y0) and [y1, y1 + h) filled in."""
        y0, y1 = min(self.H, band[0] * 16), min(self.H, band[1] * 16)
        if not self.distributed:
            return band_image
        if halo is not None:
            return self._exchange_halo(band_image, y0, y1, int(halo))
        if self.peer is not None:
            import ctypes
            from ._lib import load_library, check
            pr = self.peer
            imgs = (ctypes.c_void_p * pr["world"])(*[p + self._img_off for p in pr["ptrs"]])
            stream = ctypes.c_void_p(torch.cuda.current_stream(self.flat.device).cuda_stream)
            check(load_library().ogs_band_rows_allgather(imgs, pr["world"], pr["rank"], ctypes.c_void_p(band_image.data_ptr()),
                                                         self.W, self.H, y0, y1, stream))
            pr["handle"].barrier(channel=3)      # every band has landed in every image
            return self.image
        world = dist.get_world_size(self.group)
        bands = [None] * world
        dist.all_gather_object(bands, (y0, y1), group=self.group)
        rows = max(b - a for a, b in bands)
        mine = band_image.new_zeros((3, rows, self.W))
        mine[:, :y1 - y0] = band_image[:, y0:y1]
        parts = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(parts, mine, group=self.group)
        for (a, b), part in zip(bands, parts):
            self.image[:, a:b] = part[:, :b - a]
        return self.image

    def _exchange_halo(self, band_image, y0, y1, h):
        top = (y0, min(y0 + h, y1))                # my first rows: the halo of the band above
        bottom = (max(y1 - h, top[1]), y1)         # my last rows (not overlapping `top` on bands shorter than 2h)
        if self.peer is not None:
            import ctypes
            from ._lib import load_library, check
            pr = self.peer
            imgs = (ctypes.c_void_p * pr["world"])(*[p + self._img_off for p in pr["ptrs"]])
            stream = ctypes.c_void_p(torch.cuda.current_stream(self.flat.device).cuda_stream)
            _mark("band_rendered", torch.cuda.current_stream(self.flat.device))
            for a, b in (top, bottom):
                if b > a:
                    check(load_library().ogs_band_rows_allgather(imgs, pr["world"], pr["rank"], ctypes.c_void_p(band_image.data_ptr()),
                                                                 self.W, self.H, a, b, stream))
            _mark("halo_sent", torch.cuda.current_stream(self.flat.device))
            pr["handle"].barrier(channel=3)
            _mark("halo_barrier", torch.cuda.current_stream(self.flat.device))
        else:
            world = dist.get_world_size(self.group)
            mine = band_image.new_zeros((3, 2 * h, self.W))
            mine[:, :top[1] - top[0]] = band_image[:, top[0]:top[1]]
            mine[:, h:h + bottom[1] - bottom[0]] = band_image[:, bottom[0]:bottom[1]]
            spans = [None] * world
            dist.all_gather_object(spans, (top, bottom), group=self.group)
            parts = [torch.empty_like(mine) for _ in range(world)]
            dist.all_gather(parts, mine, group=self.group)
            for (t, b), part in zip(spans, parts):
                self.image[:, t[0]:t[1]] = part[:, :t[1] - t[0]]
                self.image[:, b[0]:b[1]] = part[:, h:h + b[1] - b[0]]
        lo, hi = max(0, y0 - h), min(self.H, y1 + h)
        band_image[:, lo:y0] = self.image[:, lo:y0]
        band_image[:, y1:hi] = self.image[:, y1:hi]
        return band_image

    def reduce_accumulators(self, acc=None, ranges=None):
        """Sum self.acc over the ranks, in place (call between the two backward kernels:
        RasterizeGaussiansBackwardCUDA(..., accumulators=ex.acc, reduce_accumulators=ex.reduce_accumulators)).
        With `ranges` ([(first Gaussian, count)], accumulator_chunks > 1 of that call) the ranges are summed one after the
        other on a stream of their own and one event per range is returned: the caller differentiates range k while
        range k+1 crosses NVLink."""
        if ranges is not None:
            return self._reduce_ranges(ranges)
        if not self.distributed:
            return
        if self.peer is not None:
            import ctypes
            from ._lib import load_library, check
            pr = self.peer
            n = -(-12 * self.P // 4) * 4
            stream = ctypes.c_void_p(torch.cuda.current_stream(self.flat.device).cuda_stream)
            pr["handle"].barrier(channel=0)
            if pr.get("multicast") and pr.get("use_multimem", pr["world"] > 4):
                check(load_library().ogs_multimem_allreduce(ctypes.c_void_p(pr["multicast"]), pr["world"], pr["rank"], n, 0, stream))
            else:
                arr = (ctypes.c_void_p * pr["world"])(*pr["ptrs"])
                check(load_library().ogs_peer_allreduce(arr, pr["world"], pr["rank"], n, 0, stream))
            pr["handle"].barrier(channel=1)
        else:
            dist.all_reduce(self.acc, op=dist.ReduceOp.SUM, group=self.group)


    def _reduce_ranges(self, ranges):
        if not self.distributed:
            return [None] * len(ranges)
        if self.peer is None:
            # NCCL / gloo: asynchronous all-reduces of the row slices; the "event" is the work handle's completion
            events = []
            for first, count in ranges:
                dist.all_reduce(self.acc[first:first + count], op=dist.ReduceOp.SUM, group=self.group)
                events.append(None)
            return events
        import ctypes
        from ._lib import load_library, check
        pr = self.peer
        cur = torch.cuda.current_stream(self.flat.device)
        if getattr(self, "_comm", None) is None:
            # high priority: the one-block barrier kernels of the ranges must not queue behind the per-Gaussian backward's
            # CTAs (measured: 130 us per barrier instead of 20 on a default-priority stream)
            self._comm = torch.cuda.Stream(self.flat.device, priority=-1)
        _mark("render_bwd_done", cur)
        pr["handle"].barrier(channel=0)                      # every rank's render backward has written its partial sums
        _mark("acc_barrier0", cur)
        ready = torch.cuda.Event()
        ready.record(cur)
        self._comm.wait_event(ready)
        events = []
        multimem = pr.get("multicast") and pr.get("use_multimem", pr["world"] > 4)
        with torch.cuda.stream(self._comm):
            stream = ctypes.c_void_p(self._comm.cuda_stream)
            for first, count in ranges:
                off, n = first * 48, -(-12 * count // 4) * 4
                if multimem:
                    check(load_library().ogs_multimem_allreduce(ctypes.c_void_p(pr["multicast"] + off), pr["world"], pr["rank"], n, 0, stream))
                else:
                    arr = (ctypes.c_void_p * pr["world"])(*[p + off for p in pr["ptrs"]])
                    check(load_library().ogs_peer_allreduce(arr, pr["world"], pr["rank"], n, 0, stream))
                # a range's sums are complete on THIS rank once every rank has stored its slice of it
                _mark(f"acc_range{len(events)}_reduced", self._comm)
                pr["handle"].barrier(channel=1)
                _mark(f"acc_range{len(events)}_barrier", self._comm)
                ev = torch.cuda.Event()
                ev.record(self._comm)
                events.append(ev)
        return events


def render_band_forward(rasterize, band, H, group=None, exchange=None, halo=None):
    """Latitude-band forward for this rank.  `rasterize(band)` must call RasterizeGaussiansCUDA(...,
    band=band) and return its 6-tuple.  With a BandExchange the ranks all-gather their pixel rows (each pixel crosses
    NVLink once per peer); without one, rows outside the band are zeroed and the per-rank images are summed
    (all-reduce of full frames: twice the bytes, kept for callers without symmetric memory).  Every rank ends with
    the full frame.  Returns (image, forward_tuple)."""
    fwd = rasterize(band)
    img = fwd[1]
    if exchange is not None:
        return exchange.gather_image(img, band, halo=halo), fwd
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
    counts = [float(c) for c in row_instance_counts]
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
