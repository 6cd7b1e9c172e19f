"""Drop-in `fused_ssim` package backed by the B200-native photometric kernels.

The reference's training loops use it when importable (`from fused_ssim import fused_ssim`, LG/train.py:36-40,182-185;
the bundled CUDA submodule is gaussian-splatting/submodules/fused-ssim) and fall back to the PyTorch `ssim` otherwise.
Same call, same value: mean of the SSIM map with an 11x11 Gaussian window (sigma 1.5), zero "same" padding,
C1 = 0.01^2, C2 = 0.03^2 (fused_ssim/__init__.py:38-46 of the submodule, ssim.cu), differentiable w.r.t. img1.
`padding="valid"` (never used by the reference's callers) is not provided.
"""
import torch

from lgdwt_b200.photometric import fused_photometric_loss

allowed_padding = ["same", "valid"]


def fused_ssim(img1, img2, padding="same", train=True):
    assert padding in allowed_padding
    if padding != "same":
        raise NotImplementedError("fused_ssim (B200-native): only padding='same' is implemented")
    if img1.dim() != 4 or img1.shape != img2.shape:
        raise RuntimeError("fused_ssim: img1 and img2 must both be (N, C, H, W)")
    if not train:
        img1 = img1.detach()
    if img1.size(0) == 1:
        return fused_photometric_loss(img1, img2)[1]
    # the mean over a batch of equally sized images is the mean of the per-image means
    return torch.stack([fused_photometric_loss(img1[n:n + 1], img2[n:n + 1])[1] for n in range(img1.size(0))]).mean()
