"""Fused Haar-DWT loss op (one forward pass + one backward pass on the GPU) equal to the PyTorch chain in
LG/train.py:131-180: the weighted 2-level global sub-band L1 plus the ELF-selected 128-px patch loss."""
import ctypes
from typing import NamedTuple, Tuple

import torch

from . import _lib


class DWTLossConfig(NamedTuple):
    """Defaults = LG/arguments/__init__.py:103-121."""
    band_weights: Tuple[float, ...] = (1.0, 1.0, 1.0, 0.0, 0.0, 0.0, 0.0, 0.0)  # LL1 LH1 HL1 HH1 LL2 LH2 HL2 HH2
    patch_size: int = 128
    patch_percentile: float = 0.2
    patch_lh1_weight: float = 1.0
    patch_hl1_weight: float = 1.0
    patch_enable: bool = True


class _FusedDWTLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pred, gt, cfg):
        if not pred.is_cuda:
            raise RuntimeError("fused_dwt_loss (B200-native): CUDA tensors only; there is no CPU path")
        pred_c = pred.contiguous().float()
        gt_c = gt.contiguous().float()
        if pred_c.dim() == 4:
            if pred_c.size(0) != 1:
                raise RuntimeError("fused_dwt_loss: batch size must be 1 (LG/train.py renders one view per step)")
            pred_c, gt_c = pred_c[0], gt_c[0]
        C, H, W = pred_c.shape
        ps = int(cfg.patch_size) if cfg.patch_enable else 0
        L = (H // ps) * (W // ps) if ps > 0 else 0
        dev = pred_c.device
        out = torch.zeros(12, dtype=torch.float32, device=dev)
        mask = torch.zeros(max(L, 1), dtype=torch.uint8, device=dev)
        nbytes = _lib.lib.lg_dwt_workspace_bytes(C, H, W, ps)
        ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        weights = (ctypes.c_float * 8)(*[float(w) for w in cfg.band_weights])
        with torch.cuda.device(dev):
            rc = _lib.lib.lg_dwt_loss_forward(pred_c.data_ptr(), gt_c.data_ptr(), C, H, W, weights, ps,
                                              float(cfg.patch_percentile), float(cfg.patch_lh1_weight),
                                              float(cfg.patch_hl1_weight), out.data_ptr(), mask.data_ptr(),
                                              ws.data_ptr(), nbytes, _lib.stream_ptr(dev))
        _lib.check(rc, RuntimeError)
        ctx.save_for_backward(pred_c, gt_c, out, mask)
        ctx.cfg, ctx.ps, ctx.in_shape = cfg, ps, pred.shape
        ctx.set_materialize_grads(False)
        details = out.clone()
        ctx.mark_non_differentiable(details)
        return out[0].clone(), out[1].clone(), details

    @staticmethod
    def backward(ctx, g_dwt, g_patch, _g_details):
        pred_c, gt_c, out, mask = ctx.saved_tensors
        cfg = ctx.cfg
        C, H, W = pred_c.shape
        dev = pred_c.device
        zero = torch.zeros((), dtype=torch.float32, device=dev)
        g_dwt = zero if g_dwt is None else g_dwt.reshape(()).float().contiguous()
        g_patch = zero if g_patch is None else g_patch.reshape(()).float().contiguous()
        grad = torch.empty_like(pred_c)
        weights = (ctypes.c_float * 8)(*[float(w) for w in cfg.band_weights])
        with torch.cuda.device(dev):
            rc = _lib.lib.lg_dwt_loss_backward(pred_c.data_ptr(), gt_c.data_ptr(), C, H, W, weights, ctx.ps,
                                               float(cfg.patch_lh1_weight), float(cfg.patch_hl1_weight),
                                               g_dwt.data_ptr(), g_patch.data_ptr(), mask.data_ptr(), out.data_ptr(),
                                               grad.data_ptr(), _lib.stream_ptr(dev))
        _lib.check(rc, RuntimeError)
        return grad.reshape(ctx.in_shape), None, None


def fused_dwt_loss(pred, gt, cfg=DWTLossConfig()):
    """Returns (dwt_loss, patch_loss, details): `dwt_loss` = sum_b w_b * L1(band_b(pred), band_b(gt))
    (LG/train.py:138-164), `patch_loss` = compute_patch_dwt_loss(pred, gt, compute_elf_map(gt), ...)
    (LG/train.py:166-180); `details` (12 floats, no grad): [2:10] the unweighted band L1s, [10] selected patches,
    [11] ELF threshold.  Differentiable w.r.t. `pred` only (the reference's GT carries no gradient)."""
    return _FusedDWTLoss.apply(pred, gt, cfg)
