"""The whole image-space loss of one LGDWT-GS iteration as ONE differentiable op (LG/train.py:128-202):

    base = (1 - lambda) * L1 + lambda * (1 - SSIM)                       (:128,182-188)
    rm   = 0.95 * rm + 0.05 * base / (dwt + 1e-8);  scale = clamp(rm, 0.1, 10)      (:190-196)
    loss = base + scale * dwt + patch_weight * patch                     (:196-202)

with dwt = the weighted 2-level global sub-band L1 (:131-164) and patch = the ELF-selected 128-px patch loss (:166-180).
One autograd node and ONE library call per direction (lg_image_loss_forward / lg_image_loss_backward): the photometric
and wavelet forward kernels, a one-thread assembly kernel that also advances the running mean on the device (the
reference does it on the host after an `.item()`), and in the backward two gradient kernels, the second adding in place.  The reference runs ~100 PyTorch launches and two autograd graphs for the same value.
"""
import ctypes

import torch

from . import _lib
from .dwt_loss import DWTLossConfig


class _FusedImageLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pred, gt, running_mean, cfg, lambda_dssim, patch_weight, update_running_mean):
        if not pred.is_cuda:
            raise RuntimeError("fused_image_loss (B200-native): CUDA tensors only; there is no CPU path")
        pred_c, gt_c = pred.contiguous().float(), gt.contiguous().float()
        if pred_c.dim() == 4:
            if pred_c.size(0) != 1:
                raise RuntimeError("fused_image_loss: batch size must be 1 (LG/train.py renders one view per step)")
            pred_c, gt_c = pred_c[0], gt_c[0]
        if pred_c.shape != gt_c.shape or pred_c.dim() != 3:
            raise RuntimeError("fused_image_loss: pred and gt must both be (C,H,W) or (1,C,H,W)")
        if running_mean.dtype != torch.float32 or running_mean.numel() != 1 or running_mean.device != pred_c.device:
            raise RuntimeError("fused_image_loss: running_mean must be a float32 scalar tensor on the image's device")
        C, H, W = pred_c.shape
        dev = pred_c.device
        lib = _lib.lib
        ps = int(cfg.patch_size) if cfg.patch_enable else 0
        L = (H // ps) * (W // ps) if ps > 0 else 0
        # ONE library call (lg_image_loss_forward) and two allocations: the small device results
        # [photometric 2 | dwt 12 | loss 2 | coefficients 4 | spare 4] followed by the ELF patch mask (zeroed by the
        # library), and one block for both workspaces (the photometric part holds the derivative maps for the backward)
        small_mask = torch.empty(96 + max(L, 1), dtype=torch.uint8, device=dev)
        small, mask = small_mask[:96].view(torch.float32), small_mask[96:]
        ph_bytes = lib.lg_photometric_workspace_bytes(C, H, W)
        dwt_bytes = lib.lg_dwt_workspace_bytes(C, H, W, ps)
        ph_pad = (ph_bytes + 255) // 256 * 256
        ws = torch.empty(ph_pad + dwt_bytes, dtype=torch.uint8, device=dev)
        weights = (ctypes.c_float * 8)(*[float(w) for w in cfg.band_weights])
        need_grad = bool(ctx.needs_input_grad[0])
        with torch.cuda.device(dev):
            _lib.check(lib.lg_image_loss_forward(
                pred_c.data_ptr(), gt_c.data_ptr(), C, H, W, weights, ps, float(cfg.patch_percentile),
                float(cfg.patch_lh1_weight), float(cfg.patch_hl1_weight), running_mean.data_ptr(), float(lambda_dssim),
                float(patch_weight), int(update_running_mean), small.data_ptr(), mask.data_ptr(), mask.numel(),
                ws.data_ptr(), ph_bytes, ws.data_ptr() + ph_pad, dwt_bytes, int(need_grad), _lib.stream_ptr(dev)),
                RuntimeError)
        ctx.save_for_backward(pred_c, gt_c, small, mask, ws)
        ctx.cfg, ctx.ps, ctx.in_shape = cfg, ps, pred.shape
        ctx.mark_non_differentiable(small)
        return small[14], small

    @staticmethod
    def backward(ctx, g_loss, _g_small):
        pred_c, gt_c, small, mask, ws = ctx.saved_tensors
        cfg = ctx.cfg
        C, H, W = pred_c.shape
        dev = pred_c.device
        g = g_loss.reshape(()).float().contiguous()
        grad = torch.empty_like(pred_c)
        weights = (ctypes.c_float * 8)(*[float(w) for w in cfg.band_weights])
        with torch.cuda.device(dev):
            # two launches: the photometric gradient, then the wavelet gradient added in place
            _lib.check(_lib.lib.lg_image_loss_backward(
                pred_c.data_ptr(), gt_c.data_ptr(), C, H, W, weights, ctx.ps, float(cfg.patch_lh1_weight),
                float(cfg.patch_hl1_weight), small.data_ptr(), g.data_ptr(), mask.data_ptr(), ws.data_ptr(),
                grad.data_ptr(), _lib.stream_ptr(dev)), RuntimeError)
        return grad.reshape(ctx.in_shape), None, None, None, None, None, None


def fused_image_loss(pred, gt, running_mean, cfg=DWTLossConfig(), lambda_dssim=0.2, patch_weight=0.1,
                     update_running_mean=True):
    """Returns (loss, terms).  `running_mean`: float32 device scalar, advanced IN PLACE when update_running_mean (the
    reference's `dwt_running_mean`, initial value 1.0).  `terms` (24 floats, no grad): [0] L1, [1] SSIM, [2] dwt,
    [3] patch, [4:12] unweighted band L1s, [12] selected patches, [13] ELF threshold, [14] loss, [15] base,
    [16:20] d(loss)/d(L1, SSIM, dwt, patch).  Differentiable w.r.t. `pred` only."""
    return _FusedImageLoss.apply(pred, gt, running_mean, cfg, lambda_dssim, patch_weight, update_running_mean)
