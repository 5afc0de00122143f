"""Host-side mirror of the reference's LibTorch boundary for the lonlat path.

Same names, argument order/meaning, return tuples and error behaviour as
``RasterizeGaussiansCUDA`` / ``RasterizeGaussiansBackwardCUDA`` / ``markVisible`` in the reference's
src/rasterize_points.cu:49-319 (declared in include/rasterize_points.h:29-80), implemented on the
C ABI of libomnigs_b200.so.  ``camera_type == 3`` (LONLAT) is the hot path this package covers;
``camera_type == 1`` (pinhole, with ``render_depth``) runs through the same kernels;
any other value raises the reference's RuntimeError.  The C++ twin of this file is
csrc/rasterize_points.cpp.
"""
import ctypes

import torch

from ._lib import load_library, check

PINHOLE = 1
LONLAT = 3
NUM_CHANNELS = 3  # reference cuda_rasterizer/config.h:25
_BINNING_GRANULE = 64 << 20


def set_seam_wrap(on):
    """Opt-in longitude-seam wrap-around (non-parity extension, see include/omnigs_b200.h).  Process-global;
    returns the previous setting."""
    lib = load_library()
    prev = bool(lib.ogs_get_seam_wrap())
    check(lib.ogs_set_seam_wrap(1 if on else 0))
    return prev


def _ptr(t):
    """Device pointer of a tensor; an empty tensor is the reference's "None" (nullptr)."""
    if t is None or t.numel() == 0:
        return None
    return ctypes.c_void_p(t.data_ptr())


def _f32c(t):
    if t is None:
        return None
    if t.numel() and t.dtype != torch.float32:
        raise RuntimeError("expected a float32 tensor")
    return t.contiguous()


def _stream(device):
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def _require_cuda(t, name):
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor (libomnigs_b200 has no CPU path)")


def RasterizeGaussiansCUDA(background, means3D, colors, opacity, scales, rotations, scale_modifier,
                           cov3D_precomp, viewmatrix, projmatrix, tan_fovx, tan_fovy,
                           image_height, image_width, sh, degree, campos, prefiltered,
                           camera_type=PINHOLE, render_depth=False, band=None):
    """reference src/rasterize_points.cu:49-164.

    Returns (num_rendered, out_color[3,H,W], radii[P] int32, geomBuffer, binningBuffer, imgBuffer);
    the three uint8 buffers are opaque and only meaningful to RasterizeGaussiansBackwardCUDA.
    projmatrix, tan_fovx, tan_fovy, prefiltered and render_depth are ignored in lonlat mode, as in
    the reference; with camera_type == 1 they drive the perspective path (rasterize_points.cu:105-132).

    ``band=(ty0, ty1)`` (extension) restricts binning and blending to tile rows
    [ty0, ty1): pixels outside the band come back as background, ``radii`` stay the full-frame radii, and
    the returned buffers feed RasterizeGaussiansBackwardCUDA unchanged (it then yields this band's
    share of every gradient; shares of disjoint bands add up to the full-frame gradient).
    """
    if means3D.dim() != 2 or means3D.size(1) != 3:
        raise RuntimeError("means3D must have dimensions (num_points, 3)")
    _require_cuda(means3D, "means3D")
    lib = load_library()
    device = means3D.device
    P, H, W = int(means3D.size(0)), int(image_height), int(image_width)

    byte_opts = dict(dtype=torch.uint8, device=device)
    radii = torch.empty((P,), dtype=torch.int32, device=device)
    geomBuffer = torch.empty((0,), **byte_opts)
    binningBuffer = torch.empty((0,), **byte_opts)
    imgBuffer = torch.empty((0,), **byte_opts)

    rendered = 0
    if P != 0:
        if camera_type not in (PINHOLE, LONLAT):
            raise RuntimeError("[CudaRasterizer]Invalid camera_type")
        if camera_type == PINHOLE and band is not None:
            raise RuntimeError("latitude bands are a lonlat extension")
        M = int(sh.size(1)) if sh.size(0) != 0 else 0
        with torch.cuda.device(device):
            out_color = torch.empty((NUM_CHANNELS, H, W), dtype=torch.float32, device=device)
            background, means3D_c, colors, opacity, scales, rotations, cov3D_precomp, viewmatrix, sh, campos = (
                _f32c(t) for t in (background, means3D, colors, opacity, scales, rotations, cov3D_precomp,
                                   viewmatrix, sh, campos))
            geomBuffer = torch.empty((lib.ogs_geom_bytes(P),), **byte_opts)
            imgBuffer = torch.empty((lib.ogs_img_bytes(W, H),), **byte_opts)
            n = ctypes.c_int64(0)
            st = _stream(device)
            common = (_ptr(means3D_c), _ptr(sh), _ptr(colors), _ptr(opacity),
                      _ptr(scales), float(scale_modifier), _ptr(rotations), _ptr(cov3D_precomp),
                      _ptr(viewmatrix), _ptr(campos),
                      _ptr(radii), _ptr(geomBuffer), _ptr(imgBuffer), ctypes.byref(n), st)
            if camera_type == PINHOLE:
                projmatrix = _f32c(projmatrix)
                check(lib.ogs_pinhole_forward_stage1(
                    P, int(degree), M, W, H, _ptr(means3D_c), _ptr(sh), _ptr(colors), _ptr(opacity),
                    _ptr(scales), float(scale_modifier), _ptr(rotations), _ptr(cov3D_precomp),
                    _ptr(viewmatrix), _ptr(projmatrix), _ptr(campos), float(tan_fovx), float(tan_fovy),
                    1 if render_depth else 0, _ptr(radii), _ptr(geomBuffer), _ptr(imgBuffer), ctypes.byref(n), st))
            elif band is None:
                check(lib.ogs_lonlat_forward_stage1(P, int(degree), M, W, H, *common))
            else:
                check(lib.ogs_lonlat_forward_stage1_band(P, int(degree), M, W, H, int(band[0]), int(band[1]), *common))
            rendered = int(n.value)
            # num_rendered changes from view to view: round the request to a 64 MiB size class so the
            # caching allocator can hand back last frame's block instead of calling cudaMalloc
            need = lib.ogs_binning_bytes(rendered, W, H)
            binningBuffer = torch.empty((-(-need // _BINNING_GRANULE) * _BINNING_GRANULE,), **byte_opts)
            check(lib.ogs_lonlat_forward_stage2(
                P, W, H, rendered, _ptr(background),
                _ptr(geomBuffer), _ptr(binningBuffer), _ptr(imgBuffer), _ptr(out_color), st))
    else:
        # reference: outputs are the zero-filled tensors, nothing is launched (:84-85,97)
        out_color = torch.zeros((NUM_CHANNELS, H, W), dtype=torch.float32, device=device)
    return rendered, out_color, radii, geomBuffer, binningBuffer, imgBuffer


def RasterizeGaussiansBackwardCUDA(background, means3D, radii, colors, scales, rotations, scale_modifier,
                                   cov3D_precomp, viewmatrix, projmatrix, tan_fovx, tan_fovy,
                                   dL_dout_color, sh, degree, campos, geomBuffer, R, binningBuffer,
                                   imageBuffer, camera_type=PINHOLE, reduce_accumulators=None, out=None,
                                   accumulators=None, conic_out=None, accumulator_chunks=0):
    """reference src/rasterize_points.cu:166-285.

    Returns (dL_dmeans2D[P,3], dL_dcolors[P,3], dL_dopacity[P,1], dL_dmeans3D[P,3], dL_dcov3D[P,6],
    dL_dsh[P,M,3], dL_dscales[P,3], dL_drotations[P,4]).

    ``reduce_accumulators`` (extension for latitude bands): a callable that receives the packed
    [P,12] float32 render-backward accumulators of this rank (a view into geomBuffer) between the two
    backward kernels, e.g. ``lambda t: dist.all_reduce(t)``; the per-Gaussian backward then runs on the
    sums, so every rank ends with the full-frame gradients after exchanging 48 B/Gaussian.

    ``accumulators`` (with ``reduce_accumulators``): a caller-owned [P,12] float32 tensor to use instead of the
    geometry buffer's (e.g. ``parallel.BandExchange`` keeps it in symmetric memory and sums it with the library's own
    NVLink kernel).

    ``accumulator_chunks`` (with ``accumulators``): n > 1 pipelines exchange and differentiation — ``reduce_accumulators``
    is then called as ``reduce_accumulators(accumulators, ranges)`` with n Gaussian ranges [(first, count)] (multiples of
    128) and returns one CUDA event (or None) per range; the per-Gaussian backward of a range is queued behind its event
    (ogs_lonlat_backward_finish_range), so range k+1 is summed over NVLink while range k is differentiated.

    ``conic_out`` (tests): a [P,4] float32 tensor that receives the reference's intermediate dL_dconic (.x, .y, .w used).

    ``out`` (extension for data-parallel training): a dense ``parallel.GradientBucket``; the five optimiser-facing
    gradients are then written into its flat buffer (and returned as views of it) instead of fresh tensors.
    """
    _require_cuda(means3D, "means3D")
    lib = load_library()
    device = means3D.device
    P = int(means3D.size(0))
    H, W = int(dL_dout_color.size(1)), int(dL_dout_color.size(2))
    M = int(sh.size(1)) if sh.size(0) != 0 else 0
    opts = dict(dtype=torch.float32, device=device)
    # the library writes every element (zeros for culled rows): no zero-fill needed
    # (the reference zero-fills 81 floats per Gaussian here, :200-208,246-247)
    alloc = torch.zeros if P == 0 else torch.empty
    dL_dmeans3D = alloc((P, 3), **opts) if out is None else out["dL_dmeans3D"]
    dL_dmeans2D = alloc((P, 3), **opts)
    dL_dcolors = alloc((P, NUM_CHANNELS), **opts)
    dL_dopacity = alloc((P, 1), **opts) if out is None else out["dL_dopacity"]
    dL_dcov3D = alloc((P, 6), **opts)
    dL_dsh = alloc((P, M, 3), **opts) if out is None else out["dL_dsh"]
    dL_dscales = alloc((P, 3), **opts) if out is None else out["dL_dscales"]
    dL_drotations = alloc((P, 4), **opts) if out is None else out["dL_drotations"]
    if out is not None and (out.P != P or out.M != M):
        raise RuntimeError("gradient bucket was built for a different P / M")
    if P != 0:
        if camera_type not in (PINHOLE, LONLAT):
            raise RuntimeError("[CudaRasterizer]Invalid camera_type")
        with torch.cuda.device(device):
            background, means3D_c, colors, scales, rotations, cov3D_precomp, viewmatrix, sh, campos, dL = (
                _f32c(t) for t in (background, means3D, colors, scales, rotations, cov3D_precomp, viewmatrix,
                                   sh, campos, dL_dout_color))
            radii_c = radii.contiguous()
            outs = (_ptr(dL_dmeans2D), _ptr(conic_out), _ptr(dL_dopacity), _ptr(dL_dcolors),
                    _ptr(dL_dmeans3D), _ptr(dL_dcov3D), _ptr(dL_dsh), _ptr(dL_dscales), _ptr(dL_drotations))
            if camera_type == PINHOLE:
                projmatrix = _f32c(projmatrix)
                check(lib.ogs_pinhole_backward(
                    P, int(degree), M, int(R), W, H,
                    _ptr(background), _ptr(means3D_c), _ptr(sh), _ptr(colors),
                    _ptr(scales), float(scale_modifier), _ptr(rotations), _ptr(cov3D_precomp),
                    _ptr(viewmatrix), _ptr(projmatrix), _ptr(campos), float(tan_fovx), float(tan_fovy), _ptr(radii_c),
                    _ptr(geomBuffer), _ptr(binningBuffer), _ptr(imageBuffer), _ptr(dL), *outs, _stream(device)))
            elif reduce_accumulators is None:
                check(lib.ogs_lonlat_backward(
                    P, int(degree), M, int(R), W, H,
                    _ptr(background), _ptr(means3D_c), _ptr(sh), _ptr(colors),
                    _ptr(scales), float(scale_modifier), _ptr(rotations), _ptr(cov3D_precomp),
                    _ptr(viewmatrix), _ptr(campos), _ptr(radii_c),
                    _ptr(geomBuffer), _ptr(binningBuffer), _ptr(imageBuffer), _ptr(dL), *outs, _stream(device)))
            elif accumulators is not None:
                if accumulators.dtype != torch.float32 or accumulators.numel() != 12 * P or not accumulators.is_contiguous():
                    raise RuntimeError("accumulators must be a contiguous float32 [P,12] tensor")
                check(lib.ogs_lonlat_backward_render_into(
                    P, int(R), W, H, _ptr(background), _ptr(geomBuffer), _ptr(binningBuffer), _ptr(imageBuffer),
                    _ptr(dL), _ptr(accumulators), _stream(device)))
                if accumulator_chunks and accumulator_chunks > 1:
                    ranges = gaussian_ranges(P, int(accumulator_chunks))
                    events = reduce_accumulators(accumulators, ranges)
                    cur = torch.cuda.current_stream(device)
                    for (first, count), ev in zip(ranges, events):
                        if ev is not None:
                            cur.wait_event(ev)
                        check(lib.ogs_lonlat_backward_finish_range(
                            P, int(degree), M, W, H, _ptr(means3D_c), _ptr(sh), _ptr(scales), float(scale_modifier),
                            _ptr(rotations), _ptr(cov3D_precomp), _ptr(viewmatrix), _ptr(campos), _ptr(radii_c),
                            _ptr(geomBuffer), _ptr(accumulators), first, count, *outs, _stream(device)))
                else:
                    reduce_accumulators(accumulators)
                    check(lib.ogs_lonlat_backward_finish_from(
                        P, int(degree), M, W, H, _ptr(means3D_c), _ptr(sh), _ptr(scales), float(scale_modifier),
                        _ptr(rotations), _ptr(cov3D_precomp), _ptr(viewmatrix), _ptr(campos), _ptr(radii_c),
                        _ptr(geomBuffer), _ptr(accumulators), *outs, _stream(device)))
            else:
                check(lib.ogs_lonlat_backward_render(
                    P, int(R), W, H, _ptr(background), _ptr(geomBuffer), _ptr(binningBuffer), _ptr(imageBuffer),
                    _ptr(dL), _stream(device)))
                base = (-geomBuffer.data_ptr()) % 256 + lib.ogs_grad_acc_offset(P)   # carving starts 256-B aligned
                reduce_accumulators(geomBuffer[base:base + 48 * P].view(torch.float32).view(P, 12))
                check(lib.ogs_lonlat_backward_finish(
                    P, int(degree), M, W, H, _ptr(means3D_c), _ptr(sh), _ptr(scales), float(scale_modifier),
                    _ptr(rotations), _ptr(cov3D_precomp), _ptr(viewmatrix), _ptr(campos), _ptr(radii_c),
                    _ptr(geomBuffer), *outs, _stream(device)))
            if M == 0:
                dL_dsh = torch.zeros((P, 0, 3), **opts)
    return dL_dmeans2D, dL_dcolors, dL_dopacity, dL_dmeans3D, dL_dcov3D, dL_dsh, dL_dscales, dL_drotations


def gaussian_ranges(P, n):
    """Split [0, P) into at most n contiguous ranges whose starts are multiples of 128 (the per-Gaussian kernels' CTA size):
    [(first, count)], the last one runs to P."""
    blocks = -(-int(P) // 128)
    n = max(1, min(int(n), blocks))
    per = -(-blocks // n)
    out = []
    for k in range(n):
        first = k * per * 128
        if first >= P:
            break
        out.append((first, min(per * 128, P - first)))
    return out


def RasterizeGaussiansBackwardView(background, means3D, radii, scales, rotations, scale_modifier, viewmatrix,
                                   dL_dout_color, sh, degree, campos, geomBuffer, R, binningBuffer, imageBuffer,
                                   bucket, view_slot, want_means2D=False):
    """One view of a multi-view / data-parallel step into a FACTORED ``parallel.GradientBucket`` (extension,
    SURVEY.md §8(e-a); C ABI ogs_lonlat_backward_view).  ``view_slot`` 0 starts the step (the bucket's geometry
    gradients and statistics are written), later slots are added inside the per-Gaussian backward; the view's
    clamp-masked dL/dRGB goes to ``bucket["dL_drgb"][view_slot]`` — dL_dsh of the step is rebuilt from all views'
    factors by ``parallel.exchange_bucket``.  lonlat camera, SH colours, scales + rotations (the live combination).
    Returns dL_dmeans2D [P,3] when asked for (else None)."""
    _require_cuda(means3D, "means3D")
    lib = load_library()
    device = means3D.device
    P = int(means3D.size(0))
    H, W = int(dL_dout_color.size(1)), int(dL_dout_color.size(2))
    M = int(sh.size(1)) if sh.size(0) != 0 else 0
    if not bucket.factored or bucket.P != P or bucket.M != M or not (0 <= view_slot < bucket.views_per_rank):
        raise RuntimeError("bucket does not fit this call (factored, same P / M, view_slot < views_per_rank)")
    m2d = torch.empty((P, 3), dtype=torch.float32, device=device) if want_means2D else None
    if P == 0:
        return m2d
    with torch.cuda.device(device):
        background, means3D_c, scales, rotations, viewmatrix, sh, campos, dL = (
            _f32c(t) for t in (background, means3D, scales, rotations, viewmatrix, sh, campos, dL_dout_color))
        check(lib.ogs_lonlat_backward_view(
            P, int(degree), M, int(R), W, H, _ptr(background), _ptr(means3D_c), _ptr(sh), _ptr(scales),
            float(scale_modifier), _ptr(rotations), _ptr(viewmatrix), _ptr(campos), _ptr(radii.contiguous()),
            _ptr(geomBuffer), _ptr(binningBuffer), _ptr(imageBuffer), _ptr(dL),
            0 if view_slot == 0 else 1, _ptr(bucket["dL_dmeans3D"]), _ptr(bucket["dL_dopacity"]), _ptr(bucket["dL_dscales"]),
            _ptr(bucket["dL_drotations"]), _ptr(bucket["dL_drgb"][view_slot]), _ptr(bucket["xyz_gradient_accum"]),
            _ptr(bucket["denom"]), _ptr(bucket.max_radii2D), _ptr(m2d), _stream(device)))
    return m2d


def RasterizeGaussiansGeometry(means3D, opacity, scales, rotations, scale_modifier, cov3D_precomp, viewmatrix, campos,
                               image_height, image_width):
    """First half of a pipelined lonlat forward (extension for data-parallel training): everything that does not read
    the SH coefficients — per-Gaussian geometry, depth order, emission and tile sort (ogs_lonlat_forward_stage1_geometry
    + ogs_lonlat_forward_bin).  Returns the state RasterizeGaussiansBlend completes."""
    _require_cuda(means3D, "means3D")
    lib = load_library()
    device = means3D.device
    P, H, W = int(means3D.size(0)), int(image_height), int(image_width)
    byte_opts = dict(dtype=torch.uint8, device=device)
    with torch.cuda.device(device):
        radii = torch.empty((P,), dtype=torch.int32, device=device)
        means3D_c, opacity, scales, rotations, cov3D_precomp, viewmatrix, campos = (
            _f32c(t) for t in (means3D, opacity, scales, rotations, cov3D_precomp, viewmatrix, campos))
        geomBuffer = torch.empty((lib.ogs_geom_bytes(P),), **byte_opts)
        imgBuffer = torch.empty((lib.ogs_img_bytes(W, H),), **byte_opts)
        n = ctypes.c_int64(0)
        st = _stream(device)
        check(lib.ogs_lonlat_forward_stage1_geometry(
            P, W, H, _ptr(means3D_c), _ptr(opacity), _ptr(scales), float(scale_modifier), _ptr(rotations),
            _ptr(cov3D_precomp), _ptr(viewmatrix), _ptr(campos), _ptr(radii), _ptr(geomBuffer), _ptr(imgBuffer),
            ctypes.byref(n), st))
        rendered = int(n.value)
        need = lib.ogs_binning_bytes(rendered, W, H)
        binningBuffer = torch.empty((-(-need // _BINNING_GRANULE) * _BINNING_GRANULE,), **byte_opts)
        check(lib.ogs_lonlat_forward_bin(P, W, H, rendered, _ptr(geomBuffer), _ptr(binningBuffer), _ptr(imgBuffer), st))
    return dict(R=rendered, radii=radii, geom=geomBuffer, binning=binningBuffer, img=imgBuffer, P=P, H=H, W=W,
                means3D=means3D_c, campos=campos)


def RasterizeGaussiansBlend(state, background, sh, degree, colors_stream=None, colors_after=None):
    """Second half: colours from the SH coefficients (ogs_lonlat_forward_colors) and the blend (ogs_lonlat_forward_blend).
    Returns RasterizeGaussiansCUDA's 6-tuple; results are bit-identical to the one-call forward.
    ``colors_stream``: run the colour kernel on that stream (after ``colors_after``, an event, if given) instead of the
    current one — it only needs the per-Gaussian records, which are complete when RasterizeGaussiansGeometry returns, so it
    can run underneath the depth and tile sorts still queued on the current stream; the blend waits for it."""
    lib = load_library()
    device = state["geom"].device
    P, H, W, R = state["P"], state["H"], state["W"], state["R"]
    with torch.cuda.device(device):
        out_color = torch.empty((NUM_CHANNELS, H, W), dtype=torch.float32, device=device)
        background, sh = _f32c(background), _f32c(sh)
        st = _stream(device)
        M = int(sh.size(1)) if sh.size(0) != 0 else 0
        if colors_stream is not None:
            cur = torch.cuda.current_stream(device)
            if colors_after is not None:
                colors_stream.wait_event(colors_after)
            check(lib.ogs_lonlat_forward_colors(P, int(degree), M, _ptr(state["means3D"]), _ptr(sh), _ptr(state["campos"]),
                                                _ptr(state["radii"]), _ptr(state["geom"]),
                                                ctypes.c_void_p(colors_stream.cuda_stream)))
            done = torch.cuda.Event()
            done.record(colors_stream)
            cur.wait_event(done)
            sh.record_stream(colors_stream)
        else:
            check(lib.ogs_lonlat_forward_colors(P, int(degree), M, _ptr(state["means3D"]), _ptr(sh), _ptr(state["campos"]),
                                                _ptr(state["radii"]), _ptr(state["geom"]), st))
        check(lib.ogs_lonlat_forward_blend(P, W, H, R, _ptr(background), _ptr(state["geom"]), _ptr(state["binning"]),
                                           _ptr(state["img"]), _ptr(out_color), st))
    return R, out_color, state["radii"], state["geom"], state["binning"], state["img"]


def markVisible(means3D, viewmatrix, projmatrix, camera_type=PINHOLE):
    """reference src/rasterize_points.cu:287-319 (lonlat: every Gaussian is marked visible; pinhole: near-plane test)."""
    _require_cuda(means3D, "means3D")
    P = int(means3D.size(0))
    present = torch.zeros((P,), dtype=torch.bool, device=means3D.device)
    if P != 0:
        if camera_type not in (PINHOLE, LONLAT):
            raise RuntimeError("[CudaRasterizer]Invalid camera_type")
        with torch.cuda.device(means3D.device):
            if camera_type == PINHOLE:
                check(load_library().ogs_mark_visible_pinhole(
                    P, _ptr(_f32c(means3D)), _ptr(_f32c(viewmatrix)), _ptr(_f32c(projmatrix)), _ptr(present),
                    _stream(means3D.device)))
            else:
                check(load_library().ogs_mark_all_visible(P, _ptr(present), _stream(means3D.device)))
    return present


def export_forward_state(P, W, H, R, geomBuffer, binningBuffer, imgBuffer, want_keys=True):
    """Test helper: our private buffers re-expressed as the reference's GeometryState /
    BinningState / ImageState arrays (rasterizer_impl.cu:198-245).  Returns a dict of tensors."""
    lib = load_library()
    device = geomBuffer.device
    T = ((W + 15) // 16) * ((H + 15) // 16)
    f = dict(dtype=torch.float32, device=device)
    out = {
        "means2D": torch.zeros((P, 2), **f), "depths": torch.zeros((P,), **f),
        "conic_opacity": torch.zeros((P, 4), **f), "rgb": torch.zeros((P, 3), **f),
        "tiles_touched": torch.zeros((P,), dtype=torch.int32, device=device),
        "clamped": torch.zeros((P, 3), dtype=torch.uint8, device=device),
        "cov3D": torch.zeros((P, 6), **f),
        "point_list": torch.zeros((R,), dtype=torch.int32, device=device),
        "point_list_keys": torch.zeros((R if want_keys else 0,), dtype=torch.int64, device=device),
        "ranges": torch.zeros((T, 2), dtype=torch.int32, device=device),
        "accum_alpha": torch.zeros((H * W,), **f),
        "n_contrib": torch.zeros((H * W,), dtype=torch.int32, device=device),
    }
    with torch.cuda.device(device):
        st = _stream(device)
        check(lib.ogs_export_geometry(P, _ptr(geomBuffer), _ptr(out["means2D"]), _ptr(out["depths"]),
                                      _ptr(out["conic_opacity"]), _ptr(out["rgb"]), _ptr(out["tiles_touched"]),
                                      _ptr(out["clamped"]), _ptr(out["cov3D"]), st))
        check(lib.ogs_export_binning(P, W, H, R, _ptr(geomBuffer), _ptr(binningBuffer), _ptr(imgBuffer),
                                     _ptr(out["point_list"]), _ptr(out["point_list_keys"]) if want_keys else None,
                                     _ptr(out["ranges"]), _ptr(out["accum_alpha"]), _ptr(out["n_contrib"]), st))
    return out
