"""CPU: host logic of the training-iteration mirror (trainer.py) and pinning of its oracle (oracle/train_ref.py).
The reference ships no golden vectors for these steps; the oracle is its LibTorch operator sequence restated in
PyTorch, checked here against a direct float64 evaluation of the definitions."""
import importlib
import math

import numpy as np
import torch

import _harness as h
from oracle import train_ref as ref

tr = importlib.import_module("omnigs-fork_b200.trainer")


def test_expon_lr_endpoints_and_log_linear_midpoint():
    # GaussianModel::exponLrFunc (gaussian_model.cpp:1141-1155) with the defaults of gaussian_parameters.h:69-72
    a, b, n = 0.00016, 0.0000016, 30000
    assert tr.expon_lr(0, a, b, 0, 0.01, n) == a
    assert abs(tr.expon_lr(n, a, b, 0, 0.01, n) - b) < 1e-12
    assert abs(tr.expon_lr(n // 2, a, b, 0, 0.01, n) - math.sqrt(a * b)) < 1e-12
    assert tr.expon_lr(2 * n, a, b, 0, 0.01, n) == tr.expon_lr(n, a, b, 0, 0.01, n)      # clamped
    assert tr.expon_lr(-1, a, b) == 0.0 and tr.expon_lr(5, 0.0, 0.0) == 0.0
    # delay: starts at lr_init * delay_mult, reaches the undelayed rate after lr_delay_steps
    assert abs(tr.expon_lr(0, a, b, 100, 0.01, n) - 0.01 * a) < 1e-15
    assert abs(tr.expon_lr(100, a, b, 100, 0.01, n) - tr.expon_lr(100, a, b, 0, 0.01, n)) < 1e-15


def ssim_float64(x, y):
    """SSIM as defined in loss_utils.h:81-131, evaluated directly (no conv2d) in float64."""
    C, H, W = x.shape
    g = np.array([math.exp(-((k - 5) ** 2) / (2 * 1.5 * 1.5)) for k in range(11)])
    g /= g.sum()
    w2 = np.outer(g, g)
    xp, yp = np.pad(x, ((0, 0), (5, 5), (5, 5))), np.pad(y, ((0, 0), (5, 5), (5, 5)))
    tot = 0.0
    for c in range(C):
        for i in range(H):
            for j in range(W):
                a, b = xp[c, i:i + 11, j:j + 11], yp[c, i:i + 11, j:j + 11]
                mu1, mu2 = (w2 * a).sum(), (w2 * b).sum()
                s1, s2, s12 = (w2 * a * a).sum() - mu1 * mu1, (w2 * b * b).sum() - mu2 * mu2, (w2 * a * b).sum() - mu1 * mu2
                tot += ((2 * mu1 * mu2 + 1e-4) * (2 * s12 + 9e-4)) / ((mu1 * mu1 + mu2 * mu2 + 1e-4) * (s1 + s2 + 9e-4))
    return tot / (C * H * W)


def test_oracle_ssim_against_direct_definition():
    rng = np.random.Generator(np.random.PCG64(3))
    x = rng.uniform(0, 1, (3, 19, 23))
    y = np.clip(x + rng.normal(0, 0.1, x.shape), 0, 1)
    got = float(ref.ssim(torch.from_numpy(x).float(), torch.from_numpy(y).float()))
    assert abs(got - ssim_float64(x, y)) < 2e-6
    assert abs(float(ref.ssim(torch.from_numpy(x).float(), torch.from_numpy(x).float())) - 1.0) < 1e-6


def test_oracle_loss_mask_and_crop():
    g = torch.Generator().manual_seed(0)
    img, gt = torch.rand((3, 24, 32), generator=g), torch.rand((3, 24, 32), generator=g)
    mask = (torch.rand((1, 24, 32), generator=g) > 0.3).float()
    loss, l1, s = ref.photometric_loss(img, gt, 0.2, mask, 20)
    assert abs(float(l1) - float((img * mask - gt)[:, :20].abs().mean())) < 1e-7
    assert abs(float(loss) - (0.8 * float(l1) + 0.2 * (1 - float(s)))) < 1e-6


def test_model_state_follows_training_setup():
    P = 10
    z = lambda *s: torch.zeros(s)
    pc = tr.GaussianModel(z(P, 3), z(P, 1, 3), z(P, 15, 3), z(P, 1), z(P, 3), z(P, 4), spatial_lr_scale=2.0)
    o = pc.opt
    # gaussian_model.cpp:495-511: per-group learning rates, f_rest at feature_lr / 20
    assert pc.lr == [o.position_lr_init * 2.0, o.feature_lr, o.feature_lr / 20.0, o.opacity_lr, o.scaling_lr, o.rotation_lr]
    assert pc.M == 16 and tr.PARAM_GROUPS == ("xyz", "f_dc", "f_rest", "opacity", "scaling", "rotation")
    assert pc.xyz_gradient_accum_.shape == (P, 1) and pc.denom_.shape == (P, 1) and pc.max_radii2D_.shape == (P,)
    lr = pc.updateLearningRate(15000)
    assert abs(lr - 2.0 * math.sqrt(o.position_lr_init * o.position_lr_final)) < 1e-12 and pc.lr[0] == lr


def test_trainer_calls_need_the_cuda_library():
    import pytest
    x = torch.rand((3, 8, 8))
    with pytest.raises(Exception):          # CPU tensors: no fallback path exists
        tr.photometric_loss(x, x)
