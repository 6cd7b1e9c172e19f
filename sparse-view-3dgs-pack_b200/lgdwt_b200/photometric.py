"""Fused photometric loss terms (mean |pred - gt| and mean SSIM) — one forward and one backward launch on the GPU,
equal to `l1_loss` + `ssim` of LG/utils/loss_utils.py:40-41,58-86 as used by LG/train.py:128,182-188."""
import torch

from . import _lib


class _FusedPhotometric(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pred, gt):
        if not pred.is_cuda:
            raise RuntimeError("fused_photometric_loss (B200-native): CUDA tensors only; there is no CPU path")
        pred_c, gt_c = pred.contiguous().float(), gt.contiguous().float()
        if pred_c.dim() == 4:
            if pred_c.size(0) != 1:
                raise RuntimeError("fused_photometric_loss: batch size must be 1 (LG/train.py renders one view per step)")
            pred_c, gt_c = pred_c[0], gt_c[0]
        if pred_c.shape != gt_c.shape or pred_c.dim() != 3:
            raise RuntimeError("fused_photometric_loss: pred and gt must both be (C,H,W) or (1,C,H,W)")
        C, H, W = pred_c.shape
        dev = pred_c.device
        out = torch.empty(2, dtype=torch.float32, device=dev)
        nbytes = _lib.lib.lg_photometric_workspace_bytes(C, H, W)
        ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        need_grad = bool(ctx.needs_input_grad[0])
        with torch.cuda.device(dev):
            rc = _lib.lib.lg_photometric_loss_forward(pred_c.data_ptr(), gt_c.data_ptr(), C, H, W, out.data_ptr(),
                                                      ws.data_ptr(), nbytes, int(need_grad), _lib.stream_ptr(dev))
        _lib.check(rc, RuntimeError)
        ctx.save_for_backward(pred_c, gt_c, ws)
        ctx.in_shape = pred.shape
        ctx.set_materialize_grads(False)
        return out[0], out[1]

    @staticmethod
    def backward(ctx, g_l1, g_ssim):
        pred_c, gt_c, ws = ctx.saved_tensors
        C, H, W = pred_c.shape
        dev = pred_c.device
        g_l1 = None if g_l1 is None else g_l1.reshape(()).float().contiguous()
        g_ssim = None if g_ssim is None else g_ssim.reshape(()).float().contiguous()
        grad = torch.empty_like(pred_c)
        with torch.cuda.device(dev):
            rc = _lib.lib.lg_photometric_loss_backward(pred_c.data_ptr(), gt_c.data_ptr(), C, H, W, ws.data_ptr(),
                                                       _lib.ptr(g_l1), _lib.ptr(g_ssim), grad.data_ptr(),
                                                       _lib.stream_ptr(dev))
        _lib.check(rc, RuntimeError)
        return grad.reshape(ctx.in_shape), None


def fused_photometric_loss(pred, gt):
    """Returns (l1, ssim) scalars: l1 = mean |pred - gt|, ssim = mean of the 11x11-Gaussian SSIM map (zero padded).
    The reference's base loss is `(1 - lambda_dssim) * l1 + lambda_dssim * (1 - ssim)` (LG/train.py:188).
    Differentiable w.r.t. `pred` only."""
    return _FusedPhotometric.apply(pred, gt)


class _FusedL1(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pred, gt):
        if not pred.is_cuda:
            raise RuntimeError("fused_l1_loss (B200-native): CUDA tensors only; there is no CPU path")
        if pred.shape != gt.shape:
            raise RuntimeError("fused_l1_loss: pred and gt must have the same shape")
        pred_c, gt_c = pred.contiguous().float(), gt.contiguous().float()
        dev = pred_c.device
        out = torch.empty(1, dtype=torch.float32, device=dev)
        scratch = torch.empty(2, dtype=torch.float64, device=dev)
        with torch.cuda.device(dev):
            rc = _lib.lib.lg_l1_loss_forward(pred_c.data_ptr(), gt_c.data_ptr(), pred_c.numel(), out.data_ptr(),
                                             scratch.data_ptr(), _lib.stream_ptr(dev))
        _lib.check(rc, RuntimeError)
        ctx.save_for_backward(pred_c, gt_c)
        ctx.in_shape = pred.shape
        return out[0]

    @staticmethod
    def backward(ctx, g):
        pred_c, gt_c = ctx.saved_tensors
        dev = pred_c.device
        g = g.reshape(()).float().contiguous()
        grad = torch.empty_like(pred_c)
        with torch.cuda.device(dev):
            rc = _lib.lib.lg_l1_loss_backward(pred_c.data_ptr(), gt_c.data_ptr(), pred_c.numel(), g.data_ptr(),
                                              grad.data_ptr(), _lib.stream_ptr(dev))
        _lib.check(rc, RuntimeError)
        return grad.reshape(ctx.in_shape), None


def fused_l1_loss(pred, gt):
    """mean |pred - gt| (l1_loss, LG/utils/loss_utils.py:40-41) in one launch forward and one backward;
    differentiable w.r.t. `pred` only."""
    return _FusedL1.apply(pred, gt)
