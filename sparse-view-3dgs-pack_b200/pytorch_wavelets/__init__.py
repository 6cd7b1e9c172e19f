"""Compatibility stand-in for the slice of `pytorch_wavelets` that LGDWT-GS uses:
`DWTForward(J, mode='symmetric', wave='db1')` (LG/utils/loss_utils.py:104,121,140-148).

The real package is an unpinned, un-vendored dependency of the reference; this module restates
its documented Haar behaviour on top of the B200-native single-level transform kernels, so that
LG/utils/loss_utils.py imports and runs unchanged.  Only Haar/db1 with symmetric extension is provided; anything
else raises instead of silently computing something different.
"""
import torch
import torch.nn as nn

from lgdwt_b200 import _lib

__all__ = ["DWTForward"]


class _HaarLevel(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        if not x.is_cuda:
            raise RuntimeError("pytorch_wavelets compat (B200-native): CUDA tensors only; there is no CPU path")
        x = x.contiguous().float()
        N, C, H, W = x.shape
        H2, W2 = (H + 1) // 2, (W + 1) // 2
        ll = torch.empty((N, C, H2, W2), dtype=torch.float32, device=x.device)
        yh = torch.empty((N, C, 3, H2, W2), dtype=torch.float32, device=x.device)
        if x.numel() > 0:
            with torch.cuda.device(x.device):
                rc = _lib.lib.lg_haar_dwt2_forward(x.data_ptr(), N * C, H, W, ll.data_ptr(), yh.data_ptr(),
                                                   _lib.stream_ptr(x.device))
            _lib.check(rc, RuntimeError)
        ctx.shape = (N, C, H, W)
        ctx.set_materialize_grads(False)
        return ll, yh

    @staticmethod
    def backward(ctx, g_ll, g_yh):
        N, C, H, W = ctx.shape
        ref = g_ll if g_ll is not None else g_yh
        if ref is None:
            return None
        gx = torch.empty((N, C, H, W), dtype=torch.float32, device=ref.device)
        g_ll = g_ll.contiguous().float() if g_ll is not None else None
        g_yh = g_yh.contiguous().float() if g_yh is not None else None
        if gx.numel() > 0:
            with torch.cuda.device(ref.device):
                rc = _lib.lib.lg_haar_dwt2_backward(_lib.ptr(g_ll), _lib.ptr(g_yh), N * C, H, W, gx.data_ptr(),
                                                    _lib.stream_ptr(ref.device))
            _lib.check(rc, RuntimeError)
        return gx


class DWTForward(nn.Module):
    """2-D forward DWT, J levels.  Returns (yl, [yh_1 .. yh_J]) with yh_j of shape (N, C, 3, H_j, W_j) holding
    (LH, HL, HH), finest level first — the pytorch_wavelets convention."""

    def __init__(self, J=1, wave="db1", mode="zero"):
        super().__init__()
        if wave not in ("db1", "haar"):
            raise NotImplementedError("pytorch_wavelets compat: only wave='db1'/'haar' is provided, got %r" % (wave,))
        if mode != "symmetric":
            raise NotImplementedError("pytorch_wavelets compat: only mode='symmetric' is provided, got %r" % (mode,))
        self.J = int(J)

    def forward(self, x):
        yh = []
        ll = x
        for _ in range(self.J):
            ll, high = _HaarLevel.apply(ll)
            yh.append(high)
        return ll, yh
