"""TEST INFRASTRUCTURE ONLY — CPU restatement (plain PyTorch ops) of the photometric terms of the LGDWT-GS base loss:
`l1_loss` (LG/utils/loss_utils.py:40-41) and `ssim` / `_ssim` (LG/utils/loss_utils.py:46-86).

PINNED: tests/golden/photometric_reference.npz holds the outputs and autograd gradients of the reference's own
LG/utils/loss_utils.py functions executed in the build container (script: tests/golden/make_photometric_golden.py);
tests/test_oracle_cpu.py checks this restatement against them.
"""
from math import exp

import torch
import torch.nn.functional as F

C1 = 0.01 ** 2
C2 = 0.03 ** 2


def gaussian_window(window_size=11, sigma=1.5):
    """loss_utils.gaussian (:46-48) + create_window (:50-54): normalised fp32 taps and their outer product"""
    g = torch.tensor([exp(-(x - window_size // 2) ** 2 / float(2 * sigma ** 2)) for x in range(window_size)],
                     dtype=torch.float32)
    g = g / g.sum()
    return g, g.unsqueeze(1).mm(g.unsqueeze(0))


def l1_loss(pred, gt):
    return torch.abs(pred - gt).mean()


def ssim(img1, img2, window_size=11):
    """loss_utils._ssim (:67-86) with size_average=True; img: (C,H,W) or (N,C,H,W)"""
    C = img1.size(-3)
    _, w2 = gaussian_window(window_size)
    window = w2.to(img1.dtype).to(img1.device).expand(C, 1, window_size, window_size).contiguous()
    pad = window_size // 2
    x = img1 if img1.dim() == 4 else img1.unsqueeze(0)
    y = img2 if img2.dim() == 4 else img2.unsqueeze(0)
    mu1 = F.conv2d(x, window, padding=pad, groups=C)
    mu2 = F.conv2d(y, window, padding=pad, groups=C)
    mu1_sq, mu2_sq, mu1_mu2 = mu1.pow(2), mu2.pow(2), mu1 * mu2
    sigma1_sq = F.conv2d(x * x, window, padding=pad, groups=C) - mu1_sq
    sigma2_sq = F.conv2d(y * y, window, padding=pad, groups=C) - mu2_sq
    sigma12 = F.conv2d(x * y, window, padding=pad, groups=C) - mu1_mu2
    ssim_map = ((2 * mu1_mu2 + C1) * (2 * sigma12 + C2)) / ((mu1_sq + mu2_sq + C1) * (sigma1_sq + sigma2_sq + C2))
    return ssim_map.mean()


def photometric_terms(pred, gt):
    """(l1, ssim) as fused_photometric_loss returns them"""
    return l1_loss(pred, gt), ssim(pred, gt)
