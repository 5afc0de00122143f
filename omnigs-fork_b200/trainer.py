"""Host-side mirror of the reference's training iteration around the lonlat rasterizer.

Reference call chain (raikuma/OmniGS-fork): ``GaussianMapper::trainForOneIteration``
(src/gaussian_mapper.cpp:300-470) -> ``GaussianRenderer::renderLonlat`` (src/gaussian_renderer.cpp:172-290:
activations of the stored tensors, then the rasterizer) -> L1 + SSIM loss (include/loss_utils.h) ->
``loss.backward()`` -> densification statistics -> ``optimizer_->step()`` (torch::optim::Adam over six
parameter groups, src/gaussian_model.cpp:485-518).  The reference runs each of these as separate LibTorch
ops; here each is one call into libomnigs_b200.so (hand-written CUDA):

    render_lonlat_raw        ogs_lonlat_forward_raw_stage1 + ogs_lonlat_forward_stage2
    photometric_loss         ogs_photometric_loss          (forward and backward)
    backward_lonlat_raw      ogs_lonlat_backward_raw
    densify_stats            ogs_densify_stats
    adam_step                ogs_adam_step

Names follow the reference (GaussianModel's xyz_, features_dc_, ...; exponLrFunc; param group order).
There is no PyTorch fallback for any of them.
"""
import ctypes
import math

import torch

from ._lib import load_library, check
from .rasterize_points import _ptr, _stream, _f32c, _BINNING_GRANULE

# param group order of the reference's optimiser (gaussian_model.cpp:534-541)
PARAM_GROUPS = ("xyz", "f_dc", "f_rest", "opacity", "scaling", "rotation")


class OptimizationParams:
    """Defaults of GaussianOptimizationParams (include/gaussian_parameters.h:66-85)."""

    def __init__(self, position_lr_init=0.00016, position_lr_final=0.0000016, position_lr_delay_mult=0.01,
                 position_lr_max_steps=30000, feature_lr=0.0025, opacity_lr=0.05, scaling_lr=0.005,
                 rotation_lr=0.001, lambda_dssim=0.2, densify_until_iter=15000):
        self.position_lr_init = position_lr_init
        self.position_lr_final = position_lr_final
        self.position_lr_delay_mult = position_lr_delay_mult
        self.position_lr_max_steps = position_lr_max_steps
        self.feature_lr = feature_lr
        self.opacity_lr = opacity_lr
        self.scaling_lr = scaling_lr
        self.rotation_lr = rotation_lr
        self.lambda_dssim = lambda_dssim
        self.densify_until_iter = densify_until_iter


def expon_lr(step, lr_init, lr_final, lr_delay_steps=0, lr_delay_mult=1.0, max_steps=1000000):
    """GaussianModel::exponLrFunc (src/gaussian_model.cpp:1141-1155)."""
    if step < 0 or (lr_init == 0.0 and lr_final == 0.0):
        return 0.0
    if lr_delay_steps > 0:
        delay_rate = lr_delay_mult + (1.0 - lr_delay_mult) * math.sin(
            0.5 * math.pi * min(max(step / lr_delay_steps, 0.0), 1.0))
    else:
        delay_rate = 1.0
    t = min(max(step / max_steps, 0.0), 1.0)
    return delay_rate * math.exp(math.log(lr_init) * (1 - t) + math.log(lr_final) * t)


class GaussianModel:
    """The stored (pre-activation) tensors of the reference's GaussianModel, its Adam state and its
    densification statistics (src/gaussian_model.cpp:485-518, include/gaussian_model.h)."""

    def __init__(self, xyz, features_dc, features_rest, opacity, scaling, rotation, sh_degree=3,
                 spatial_lr_scale=1.0, opt=None):
        self.opt = opt or OptimizationParams()
        self.xyz_ = _f32c(xyz)
        self.features_dc_ = _f32c(features_dc)        # [P,1,3]
        self.features_rest_ = _f32c(features_rest)    # [P,M-1,3]
        self.opacity_ = _f32c(opacity)                # [P,1] logit
        self.scaling_ = _f32c(scaling)                # [P,3] log
        self.rotation_ = _f32c(rotation)              # [P,4] unnormalised (w,x,y,z)
        self.active_sh_degree_ = sh_degree
        self.spatial_lr_scale_ = spatial_lr_scale
        P, dev = self.xyz_.size(0), self.xyz_.device
        z = lambda *s: torch.zeros(s, dtype=torch.float32, device=dev)
        # trainingSetup (gaussian_model.cpp:485-518)
        self.max_radii2D_ = z(P)
        self.xyz_gradient_accum_ = z(P, 1)
        self.denom_ = z(P, 1)
        self.exp_avg = [torch.zeros_like(p) for p in self.params()]
        self.exp_avg_sq = [torch.zeros_like(p) for p in self.params()]
        self.step_count = 0
        o = self.opt
        self.lr = [o.position_lr_init * spatial_lr_scale, o.feature_lr, o.feature_lr / 20.0, o.opacity_lr,
                   o.scaling_lr, o.rotation_lr]

    def params(self):
        return [self.xyz_, self.features_dc_, self.features_rest_, self.opacity_, self.scaling_, self.rotation_]

    @property
    def M(self):
        return 1 + int(self.features_rest_.size(1))

    def updateLearningRate(self, step):
        """GaussianModel::updateLearningRate (gaussian_model.cpp:520-532): xyz group only."""
        o = self.opt
        self.lr[0] = expon_lr(step, o.position_lr_init * self.spatial_lr_scale_,
                              o.position_lr_final * self.spatial_lr_scale_, 0, o.position_lr_delay_mult,
                              o.position_lr_max_steps)
        return self.lr[0]


def render_lonlat_raw(pc, viewmatrix, campos, H, W, bg_color, scaling_modifier=1.0):
    """GaussianRenderer::renderLonlat on the stored tensors (activations inside the kernel).
    Returns (rendered_image[3,H,W], radii[P] int32, ctx) with ctx the opaque forward state."""
    lib = load_library()
    dev = pc.xyz_.device
    P = int(pc.xyz_.size(0))
    byte = dict(dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        radii = torch.empty((P,), dtype=torch.int32, device=dev)
        out = torch.empty((3, H, W), dtype=torch.float32, device=dev)
        geom = torch.empty((lib.ogs_geom_bytes(P),), **byte)
        img = torch.empty((lib.ogs_img_bytes(W, H),), **byte)
        viewmatrix, campos, bg_color = _f32c(viewmatrix), _f32c(campos), _f32c(bg_color)
        n = ctypes.c_int64(0)
        st = _stream(dev)
        check(lib.ogs_lonlat_forward_raw_stage1(
            P, int(pc.active_sh_degree_), pc.M, W, H, _ptr(pc.xyz_), _ptr(pc.features_dc_), _ptr(pc.features_rest_),
            _ptr(pc.opacity_), _ptr(pc.scaling_), float(scaling_modifier), _ptr(pc.rotation_),
            _ptr(viewmatrix), _ptr(campos), _ptr(radii), _ptr(geom), _ptr(img), ctypes.byref(n), st))
        R = int(n.value)
        need = lib.ogs_binning_bytes(R, W, H)
        binning = torch.empty((-(-need // _BINNING_GRANULE) * _BINNING_GRANULE,), **byte)
        check(lib.ogs_lonlat_forward_stage2(P, W, H, R, _ptr(bg_color), _ptr(geom), _ptr(binning), _ptr(img),
                                            _ptr(out), st))
    ctx = dict(R=R, geom=geom, binning=binning, img=img, viewmatrix=viewmatrix, campos=campos, bg=bg_color,
               H=H, W=W, scaling_modifier=float(scaling_modifier), radii=radii)
    return out, radii, ctx


def backward_lonlat_raw(pc, ctx, dL_dout_color, want_means2D=True):
    """Gradients of the stored tensors (what LibTorch autograd hands the optimiser) for one rendered view.
    Returns (dL_dmeans2D[P,3] or None, [dL_dxyz, dL_df_dc, dL_df_rest, dL_dopacity, dL_dscaling, dL_drotation])."""
    lib = load_library()
    dev = pc.xyz_.device
    P = int(pc.xyz_.size(0))
    with torch.cuda.device(dev):
        grads = [torch.empty_like(p) for p in pc.params()]
        m2d = torch.empty((P, 3), dtype=torch.float32, device=dev) if want_means2D else None
        dL = _f32c(dL_dout_color)
        check(lib.ogs_lonlat_backward_raw(
            P, int(pc.active_sh_degree_), pc.M, ctx["R"], ctx["W"], ctx["H"], _ptr(ctx["bg"]),
            _ptr(pc.xyz_), _ptr(pc.features_dc_), _ptr(pc.features_rest_), _ptr(pc.scaling_),
            ctx["scaling_modifier"], _ptr(pc.rotation_), _ptr(ctx["viewmatrix"]), _ptr(ctx["campos"]),
            _ptr(ctx["radii"]), _ptr(ctx["geom"]), _ptr(ctx["binning"]), _ptr(ctx["img"]), _ptr(dL),
            _ptr(m2d), _ptr(grads[0]), _ptr(grads[1]), _ptr(grads[2]), _ptr(grads[3]), _ptr(grads[4]), _ptr(grads[5]),
            _stream(dev)))
    return m2d, grads


_loss_ws = {}


def photometric_loss(rendered, gt, lambda_dssim=0.2, mask=None, rows_used=None):
    """(1 - lambda) * l1_loss + lambda * (1 - ssim) of rendered * mask vs gt and its gradient
    (gaussian_mapper.cpp:391-413, loss_utils.h:31-34,58-131).  Returns (loss_out[3] = {loss, L1, SSIM}
    device tensor, dL_drendered[3,H,W])."""
    lib = load_library()
    dev = rendered.device
    H, W = int(rendered.size(1)), int(rendered.size(2))
    rows = H if rows_used is None else int(rows_used)
    with torch.cuda.device(dev):
        rendered, gt, mask = _f32c(rendered), _f32c(gt), _f32c(mask)
        key = (dev.index, W, H)
        ws = _loss_ws.get(key)
        if ws is None:
            ws = _loss_ws[key] = torch.empty((lib.ogs_photometric_loss_workspace_bytes(W, H),), dtype=torch.uint8, device=dev)
        loss_out = torch.empty((3,), dtype=torch.float32, device=dev)
        dL = torch.empty_like(rendered)
        mch = 0 if mask is None else (3 if mask.dim() == 3 and mask.size(0) == 3 else 1)
        check(lib.ogs_photometric_loss(W, H, rows, float(lambda_dssim), _ptr(rendered), _ptr(gt), _ptr(mask), mch,
                                       _ptr(ws), _ptr(loss_out), _ptr(dL), _stream(dev)))
    return loss_out, dL


def adam_step(params, grads, exp_avg, exp_avg_sq, lrs, step, beta1=0.9, beta2=0.999, eps=1e-15):
    """torch::optim::Adam::step over all parameter groups in one launch (eps as set at gaussian_model.cpp:493)."""
    lib = load_library()
    n = len(params)
    dev = params[0].device
    for t in list(params) + list(grads) + list(exp_avg) + list(exp_avg_sq):
        if t.dtype != torch.float32 or not t.is_contiguous():
            raise RuntimeError("adam_step needs contiguous float32 tensors")
    arr = lambda ts: (ctypes.c_void_p * n)(*[t.data_ptr() for t in ts])
    counts = (ctypes.c_size_t * n)(*[t.numel() for t in params])
    lr_arr = (ctypes.c_float * n)(*[float(x) for x in lrs])
    with torch.cuda.device(dev):
        check(lib.ogs_adam_step(n, arr(params), arr(grads), arr(exp_avg), arr(exp_avg_sq), counts, lr_arr, int(step),
                                float(beta1), float(beta2), float(eps), _stream(dev)))


def densify_stats(pc, radii, dL_dmeans2D):
    """max_radii2D / addDensificationStats (gaussian_mapper.cpp:427-434, gaussian_model.cpp:839-853)."""
    lib = load_library()
    dev = pc.xyz_.device
    with torch.cuda.device(dev):
        check(lib.ogs_densify_stats(int(pc.xyz_.size(0)), _ptr(radii), _ptr(dL_dmeans2D), _ptr(pc.max_radii2D_),
                                    _ptr(pc.xyz_gradient_accum_), _ptr(pc.denom_), _stream(dev)))


def train_for_one_iteration(pc, viewmatrix, campos, gt_image, bg_color, iteration, mask=None, rows_used=None,
                            group=None):
    """One training iteration on one view (GaussianMapper::trainForOneIteration, gaussian_mapper.cpp:300-470,
    without densification / pruning): lr schedule, render, loss, backward, densification statistics (while
    iteration < densify_until_iter, as :427), Adam.  Under torch.distributed with more than one rank in `group` every
    rank renders its own view and the six gradients are summed, the two summed statistics added and the radii maxed over
    the ranks BEFORE they are applied, so parameters AND statistics stay identical on all replicas (the fast path for
    that is parallel.GradientBucket; this is the plain collective form).  Returns (loss_out[3] device tensor, rendered)."""
    import torch.distributed as dist
    H, W = int(gt_image.size(1)), int(gt_image.size(2))
    pc.updateLearningRate(iteration)
    rendered, radii, ctx = render_lonlat_raw(pc, viewmatrix, campos, H, W, bg_color)
    loss_out, dL = photometric_loss(rendered, gt_image, pc.opt.lambda_dssim, mask, rows_used)
    m2d, grads = backward_lonlat_raw(pc, ctx, dL)
    distributed = dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
    track = iteration < pc.opt.densify_until_iter
    if not distributed:
        if track:
            densify_stats(pc, radii, m2d)
    else:
        work = [dist.all_reduce(g, op=dist.ReduceOp.SUM, group=group, async_op=True) for g in grads if g.numel()]
        if track:
            lib = load_library()
            P, dev = int(pc.xyz_.size(0)), pc.xyz_.device
            gn, vis, rad = (torch.empty((P,), dtype=torch.float32, device=dev) for _ in range(3))
            with torch.cuda.device(dev):
                check(lib.ogs_view_stats(P, _ptr(radii), _ptr(m2d), _ptr(gn), _ptr(vis), _ptr(rad), _stream(dev)))
            work += [dist.all_reduce(gn, op=dist.ReduceOp.SUM, group=group, async_op=True),
                     dist.all_reduce(vis, op=dist.ReduceOp.SUM, group=group, async_op=True),
                     dist.all_reduce(rad, op=dist.ReduceOp.MAX, group=group, async_op=True)]
        for w in work:
            w.wait()
        if track:
            apply_view_stats(pc, gn, vis, rad)
    pc.step_count += 1
    adam_step(pc.params(), grads, pc.exp_avg, pc.exp_avg_sq, pc.lr, pc.step_count)
    return loss_out, rendered


def apply_view_stats(pc, grad_norm, visible, radius):
    """Apply statistics summed / maxed over several views (ranks) of one step: what ogs_densify_stats does per view,
    gaussian_mapper.cpp:427-434 and gaussian_model.cpp:839-853, for all of them at once."""
    seen = visible > 0
    pc.max_radii2D_.copy_(torch.where(seen, torch.maximum(pc.max_radii2D_, radius), pc.max_radii2D_))
    pc.xyz_gradient_accum_.add_(grad_norm.view_as(pc.xyz_gradient_accum_))
    pc.denom_.add_(visible.view_as(pc.denom_))
