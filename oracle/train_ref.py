"""TEST INFRASTRUCTURE — a PyTorch float32 restatement of the LibTorch code the reference runs either side of
its rasterizer in a training iteration.  Only tests/, __graft_entry__.smoke() and
bench.py's reference arm may import this; the product path (omnigs-fork_b200/trainer.py) never does.

The reference expresses these steps with stock LibTorch ops and autograd, so the restatement is the same ops
in Python (same operator sequence, same dtypes):

    activations          src/gaussian_model.cpp:54-77 (exp, normalize, cat, sigmoid)
    l1_loss, ssim        include/loss_utils.h:31-34, 58-131
    iteration            src/gaussian_mapper.cpp:391-434 (mask, skip-bottom crop, loss, backward, statistics)
    optimiser            src/gaussian_model.cpp:485-518 (torch::optim::Adam, eps 1e-15, six groups)

Parity is pinned by construction (identical operators); tests/test_trainer_host.py additionally checks ssim()
against a direct float64 evaluation of the SSIM definition on CPU.
"""
import math

import torch
import torch.nn.functional as F


def activations(xyz, features_dc, features_rest, opacity, scaling, rotation):
    """GaussianModel::get*Activation / getFeatures (gaussian_model.cpp:54-77)."""
    return dict(means3D=xyz, opacity=torch.sigmoid(opacity), scales=torch.exp(scaling),
                rotations=F.normalize(rotation), shs=torch.cat([features_dc, features_rest], dim=1))


def gaussian(window_size, sigma, device):
    # loss_utils.h:58-68: float32 values, normalised by their float32 sum
    g = torch.tensor([math.exp(-((x - window_size // 2) ** 2) / (2.0 * sigma * sigma)) for x in range(window_size)],
                     dtype=torch.float32, device=device)
    return g / g.sum()


def create_window(window_size, channel, device):
    # loss_utils.h:70-79
    w1 = gaussian(window_size, 1.5, device).unsqueeze(1)
    w2 = w1.mm(w1.t()).float().unsqueeze(0).unsqueeze(0)
    return w2.expand(channel, 1, window_size, window_size).contiguous()


def ssim(img1, img2, window_size=11):
    """loss_utils.h:81-131 (size_average = true).  img: [3,H,W]."""
    channel = img1.size(-3)
    window = create_window(window_size, channel, img1.device).type_as(img1)
    pad = window_size // 2
    conv = lambda x: F.conv2d(x.unsqueeze(0), window, padding=pad, groups=channel).squeeze(0)
    mu1, mu2 = conv(img1), conv(img2)
    mu1_sq, mu2_sq, mu1_mu2 = mu1.pow(2), mu2.pow(2), mu1 * mu2
    sigma1_sq = conv(img1 * img1) - mu1_sq
    sigma2_sq = conv(img2 * img2) - mu2_sq
    sigma12 = conv(img1 * img2) - mu1_mu2
    C1, C2 = 0.01 * 0.01, 0.03 * 0.03
    ssim_map = ((2 * mu1_mu2 + C1) * (2 * sigma12 + C2)) / ((mu1_sq + mu2_sq + C1) * (sigma1_sq + sigma2_sq + C2))
    return ssim_map.mean()


def l1_loss(a, b):
    return torch.abs(a - b).mean()   # loss_utils.h:31-34


def photometric_loss(rendered, gt, lambda_dssim, mask=None, rows_used=None):
    """gaussian_mapper.cpp:391-413.  Returns (loss, Ll1, ssim)."""
    img = rendered * mask if mask is not None else rendered
    if rows_used is not None and rows_used < rendered.size(1):
        img, gt = img[:, :rows_used, :], gt[:, :rows_used, :]
    Ll1 = l1_loss(img, gt)
    s = ssim(img, gt)
    return (1.0 - lambda_dssim) * Ll1 + lambda_dssim * (1.0 - s), Ll1, s


def make_adam(params, lrs):
    """trainingSetup (gaussian_model.cpp:485-511): one Adam, six groups, eps 1e-15.  foreach/fused off so the
    update is the element-wise sequence torch::optim::Adam::step performs."""
    groups = [{"params": [p], "lr": lr} for p, lr in zip(params, lrs)]
    return torch.optim.Adam(groups, lr=0.0, eps=1e-15, foreach=False, fused=False)


def densify_stats(max_radii2D, xyz_gradient_accum, denom, radii, viewspace_grad):
    """gaussian_mapper.cpp:427-434 + GaussianModel::addDensificationStats (gaussian_model.cpp:839-853)."""
    vis = radii > 0
    max_radii2D[vis] = torch.max(max_radii2D[vis], radii[vis].to(max_radii2D.dtype))
    xyz_gradient_accum[vis] += torch.norm(viewspace_grad[vis, :2], dim=-1, keepdim=True)
    denom[vis] += 1
