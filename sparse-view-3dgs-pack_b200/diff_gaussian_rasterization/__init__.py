"""Drop-in `diff_gaussian_rasterization` operator surface backed by the B200-native C-ABI library.

Mirrors the reference package DGR/dgr_3dgs/__init__.py (GaussianRasterizationSettings :143-156,
GaussianRasterizer :158-207, _RasterizeGaussians :44-141): same names, argument order, return values
(color, radii, invdepth), gradient contract and error behaviour, so LG/gaussian_renderer/__init__.py,
MS/gaussian_renderer/__init__.py, train.py, train_nir.py and render.py run unchanged.

Extensions (not in the reference): colors_precomp may carry 1..4 channels (RGB+NIR in one pass); kernels run on
the current torch stream instead of the legacy default stream; gradient buffers are not pre-zeroed by the caller;
`GaussianRasterizer.forward(..., grad_sinks=GradSinks(...))` makes the backward write (or accumulate) the parameter
gradients straight into caller-owned tensors — e.g. slices of the flat bucket a data-parallel step all-reduces —
instead of returning fresh tensors for autograd to add up.
`SparseGaussianAdam` is intentionally NOT exported: its presence would switch the callers to the `separate_sh`
calling convention the bundled reference surface does not have.
"""
import ctypes
from typing import NamedTuple

import torch
import torch.nn as nn

from lgdwt_b200 import _lib

__all__ = ["GaussianRasterizationSettings", "GaussianRasterizer", "GradSinks", "rasterize_gaussians"]


class GradSinks:
    """Caller-owned destinations for the parameter gradients of one rasterizer call (extension, see module docstring).

    means3D (P,3), shs (P,M,3), opacities (P,1), scales (P,3), rotations (P,4): contiguous fp32 CUDA tensors (typically
    views of one flat bucket).  With accumulate=False the backward overwrites them, with accumulate=True it adds to
    them.  The autograd gradients of those five inputs are then None (nothing for autograd to add a second time);
    means2D still receives its gradient the usual way (densification statistics read means2D.grad).

    raw_rot_norm (P,): when given, the rasterizer's opacities / scales / rotations inputs are declared to be
    sigmoid / exp / normalize of a trainer's raw parameters (`lgdwt_b200.activate`), and the opacities / scales /
    rotations sinks receive the gradients w.r.t. those RAW parameters (chain rule applied inside the
    preprocess-backward kernel, lg_rasterize_backward_raw)."""
    __slots__ = ("means3D", "shs", "opacities", "scales", "rotations", "accumulate", "raw_rot_norm")

    def __init__(self, means3D, shs, opacities, scales, rotations, accumulate=False, raw_rot_norm=None):
        self.means3D, self.shs, self.opacities, self.scales, self.rotations = means3D, shs, opacities, scales, rotations
        self.accumulate = bool(accumulate)
        self.raw_rot_norm = raw_rot_norm


class GaussianRasterizationSettings(NamedTuple):
    image_height: int
    image_width: int
    tanfovx: float
    tanfovy: float
    bg: torch.Tensor
    scale_modifier: float
    viewmatrix: torch.Tensor
    projmatrix: torch.Tensor
    sh_degree: int
    campos: torch.Tensor
    prefiltered: bool
    debug: bool
    antialiasing: bool


def _f32c(t):
    """contiguous fp32 view of an optional tensor (the reference applies .contiguous() at the boundary,
    DGR/rasterize_points.cu:101-120)."""
    if t is None or t.numel() == 0:
        return None
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


def rasterize_gaussians(means3D, means2D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp,
                        raster_settings, grad_sinks=None):
    return _RasterizeGaussians.apply(means3D, means2D, sh, colors_precomp, opacities, scales, rotations,
                                     cov3Ds_precomp, raster_settings, grad_sinks)


# Binning-capacity hints (performance only): the library sizes the binning buffer for `hint` entries and queues the whole
# forward before it waits for the exact count (lg_rasterize_forward_hinted), instead of stalling the GPU in the middle of
# every forward as the reference does (rasterizer_impl.cu:283-288).  The hint is 1.25 x the largest entry count of the
# last few calls on that device; a hint that turns out too small only costs a re-queue of the tail of that one call.
_CAPACITY_HISTORY = {}
_HISTORY_LEN = 8
_inspection_hook = None


def set_inspection_hook(fn):
    """Opt-in hook for benchmarks / tests: `fn(info)` is called after every forward with the sizes and the three state
    buffers of that call (keeping them alive for as long as the callee holds on to `info`).  None disables it."""
    global _inspection_hook
    _inspection_hook = fn


def _capacity_hint(device, P, W, H):
    hist = _CAPACITY_HISTORY.get((device, P, W, H))
    return 0 if not hist else int(1.25 * max(hist)) + 1024


def _record_num_rendered(device, P, W, H, R):
    hist = _CAPACITY_HISTORY.setdefault((device, P, W, H), [])
    hist.append(R)
    if len(hist) > _HISTORY_LEN:
        del hist[0]
    if len(_CAPACITY_HISTORY) > 64:  # scenes come and go (densification changes P): forget the oldest shapes
        for key in list(_CAPACITY_HISTORY)[:32]:
            if key != (device, P, W, H):
                del _CAPACITY_HISTORY[key]


class _RasterizeGaussians(torch.autograd.Function):

    @staticmethod
    def forward(ctx, means3D, means2D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp,
                raster_settings, grad_sinks=None):
        rs = raster_settings
        ctx.sinks = grad_sinks
        if means3D.dim() != 2 or means3D.size(1) != 3:
            # same check and exception type as AT_ERROR in DGR/rasterize_points.cu:58-60
            raise RuntimeError("means3D must have dimensions (num_points, 3)")
        if not means3D.is_cuda:
            raise RuntimeError("diff_gaussian_rasterization (B200-native): tensors must live on a CUDA device; "
                               "there is no CPU path")
        device = means3D.device
        P = means3D.size(0)
        H, W = int(rs.image_height), int(rs.image_width)

        means3D_c = _f32c(means3D)
        sh_c, colors_c = _f32c(sh), _f32c(colors_precomp)
        opac_c, scales_c, rots_c, cov_c = _f32c(opacities), _f32c(scales), _f32c(rotations), _f32c(cov3Ds_precomp)
        bg_c, view_c, proj_c, campos_c = _f32c(rs.bg), _f32c(rs.viewmatrix), _f32c(rs.projmatrix), _f32c(rs.campos)
        channels = 3 if colors_c is None else int(colors_c.size(-1))
        M = 0 if sh_c is None else int(sh_c.size(1))

        # the library writes every element of the three outputs when P > 0 (the reference fills them first,
        # DGR/rasterize_points.cu:62-64); P == 0 returns its fill values
        alloc = torch.zeros if P == 0 else torch.empty
        color = alloc((channels, H, W), dtype=torch.float32, device=device)
        invdepth = alloc((1, H, W), dtype=torch.float32, device=device)
        radii = alloc((P,), dtype=torch.int32, device=device)
        geom, binning, img = (_lib.ResizableBuffer(device) for _ in range(3))
        num_rendered, capacity = ctypes.c_int(0), ctypes.c_int(0)
        # debug keeps the reference's synchronous protocol (it syncs after every stage anyway)
        hint = 0 if rs.debug else _capacity_hint(device, P, W, H)
        with torch.cuda.device(device):
            rc = _lib.lib.lg_rasterize_forward_hinted(
                geom.callback, None, binning.callback, None, img.callback, None,
                P, int(rs.sh_degree), M, channels,
                _lib.ptr(bg_c), W, H,
                _lib.ptr(means3D_c), _lib.ptr(sh_c), _lib.ptr(colors_c), _lib.ptr(opac_c), _lib.ptr(scales_c),
                float(rs.scale_modifier), _lib.ptr(rots_c), _lib.ptr(cov_c),
                _lib.ptr(view_c), _lib.ptr(proj_c), _lib.ptr(campos_c),
                float(rs.tanfovx), float(rs.tanfovy), int(bool(rs.prefiltered)),
                _lib.ptr(color), _lib.ptr(invdepth), int(bool(rs.antialiasing)), _lib.ptr(radii),
                int(bool(rs.debug)), _lib.stream_ptr(device), hint, ctypes.byref(num_rendered),
                ctypes.byref(capacity))
        _lib.check(rc, RuntimeError)
        _record_num_rendered(device, P, W, H, num_rendered.value)
        if _inspection_hook is not None:
            _inspection_hook(dict(num_rendered=num_rendered.value, binning_capacity=capacity.value, P=P,
                                  channels=channels, W=W, H=H, geom=geom.tensor, binning=binning.tensor,
                                  img=img.tensor))
        ctx.raster_settings = rs
        ctx.num_rendered = capacity.value  # the entry count the binning state was laid out for (>= num_rendered)
        ctx.channels = channels
        ctx.M = M
        ctx.cam = (bg_c, view_c, proj_c, campos_c)
        ctx.had = (sh is not None and sh.numel() > 0, colors_precomp is not None and colors_precomp.numel() > 0,
                   scales is not None and scales.numel() > 0, cov3Ds_precomp is not None and cov3Ds_precomp.numel() > 0)
        empty = torch.empty(0, device=device)
        ctx.save_for_backward(*(t if t is not None else empty for t in
                                (colors_c, means3D_c, scales_c, rots_c, cov_c, radii, sh_c, opac_c)),
                              geom.tensor, binning.tensor, img.tensor)
        ctx.mark_non_differentiable(radii)
        ctx.set_materialize_grads(False)
        return color, radii, invdepth

    @staticmethod
    def backward(ctx, grad_out_color, _grad_radii, grad_out_depth):
        rs = ctx.raster_settings
        (colors_c, means3D_c, scales_c, rots_c, cov_c, radii, sh_c, opac_c, geom, binning, img) = ctx.saved_tensors
        bg_c, view_c, proj_c, campos_c = ctx.cam
        device = means3D_c.device
        P = means3D_c.size(0)
        H, W = int(rs.image_height), int(rs.image_width)
        C, M = ctx.channels, ctx.M
        has_sh, has_colors, has_scales, has_cov = ctx.had

        if grad_out_color is None:
            grad_out_color = torch.zeros((C, H, W), dtype=torch.float32, device=device)
        grad_out_color = _f32c(grad_out_color)
        grad_out_depth = _f32c(grad_out_depth)  # None => the inverse-depth branch is skipped entirely

        new = lambda *shape: torch.empty(shape, dtype=torch.float32, device=device)
        sinks = ctx.sinks
        if sinks is not None:
            if not (has_sh and has_scales):
                raise RuntimeError("grad_sinks needs the SH + scale/rotation input path")
            for name, shape in (("means3D", (P, 3)), ("shs", (P, M, 3)), ("opacities", (P, 1)), ("scales", (P, 3)),
                                ("rotations", (P, 4))):
                t = getattr(sinks, name)
                if (t is None or t.dtype != torch.float32 or not t.is_contiguous() or t.device != device
                        or t.numel() != P * shape[-1] * (shape[1] if len(shape) == 3 else 1)):
                    raise RuntimeError("grad_sinks.%s must be a contiguous fp32 CUDA tensor of shape %s" % (name, shape))
            dL_dmeans3D, dL_dsh, dL_dopacity = sinks.means3D, sinks.shs, sinks.opacities
            dL_dscales, dL_drotations = sinks.scales, sinks.rotations
            dL_dmeans2D, dL_dcolors, dL_dcov3D = new(P, 3), None, None
        else:
            dL_dmeans3D, dL_dmeans2D = new(P, 3), new(P, 3)
            dL_dopacity = new(P, 1)
            # per-Gaussian colour / covariance gradients are intermediates unless the caller supplied those inputs
            dL_dcolors = new(P, C) if has_colors else None
            dL_dcov3D = new(P, 6) if has_cov else None
            dL_dsh = new(P, M, 3) if has_sh else None
            dL_dscales = new(P, 3) if has_scales else None
            dL_drotations = new(P, 4) if has_scales else None
        if P > 0:
            with torch.cuda.device(device):
                rot_norm = None if sinks is None else sinks.raw_rot_norm
                if rot_norm is not None and (rot_norm.dtype != torch.float32 or rot_norm.numel() != P
                                             or not rot_norm.is_contiguous() or rot_norm.device != device):
                    raise RuntimeError("grad_sinks.raw_rot_norm must be a contiguous fp32 CUDA tensor of P elements")
                rc = _lib.lib.lg_rasterize_backward_raw(
                    P, int(rs.sh_degree), M, ctx.num_rendered, C,
                    _lib.ptr(bg_c), W, H,
                    _lib.ptr(means3D_c), _lib.ptr(sh_c), _lib.ptr(colors_c), _lib.ptr(opac_c), _lib.ptr(scales_c),
                    float(rs.scale_modifier), _lib.ptr(rots_c), _lib.ptr(cov_c),
                    _lib.ptr(view_c), _lib.ptr(proj_c), _lib.ptr(campos_c),
                    float(rs.tanfovx), float(rs.tanfovy),
                    _lib.ptr(radii), _lib.ptr(geom), _lib.ptr(binning), _lib.ptr(img),
                    _lib.ptr(grad_out_color), _lib.ptr(grad_out_depth),
                    _lib.ptr(dL_dmeans2D), None, _lib.ptr(dL_dopacity), _lib.ptr(dL_dcolors), None,
                    _lib.ptr(dL_dmeans3D), _lib.ptr(dL_dcov3D), _lib.ptr(dL_dsh), _lib.ptr(dL_dscales),
                    _lib.ptr(dL_drotations), int(bool(rs.antialiasing)), int(bool(rs.debug)),
                    _lib.stream_ptr(device), int(sinks is not None and sinks.accumulate), _lib.ptr(rot_norm))
            _lib.check(rc, RuntimeError)
        elif sinks is not None and not sinks.accumulate:
            for t in (dL_dmeans3D, dL_dsh, dL_dopacity, dL_dscales, dL_drotations):
                t.zero_()
        if sinks is not None:  # the gradients live in the caller's tensors; autograd has nothing to add
            return (None, dL_dmeans2D, None, None, None, None, None, None, None, None)
        # (means3D, means2D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp, raster_settings, sinks)
        return (dL_dmeans3D, dL_dmeans2D, dL_dsh, dL_dcolors if has_colors else None, dL_dopacity, dL_dscales,
                dL_drotations, dL_dcov3D if has_cov else None, None, None)


class GaussianRasterizer(nn.Module):
    def __init__(self, raster_settings):
        super().__init__()
        self.raster_settings = raster_settings

    def markVisible(self, positions):
        """bool (P,) mask of Gaussians in front of the near plane (DGR/dgr_3dgs/__init__.py:163-172)."""
        with torch.no_grad():
            rs = self.raster_settings
            pos = _f32c(positions)
            P = positions.size(0)
            visible = torch.zeros((P,), dtype=torch.bool, device=positions.device)
            if P > 0:
                with torch.cuda.device(positions.device):
                    rc = _lib.lib.lg_mark_visible(P, _lib.ptr(pos), _lib.ptr(_f32c(rs.viewmatrix)),
                                                  _lib.ptr(_f32c(rs.projmatrix)), visible.data_ptr(),
                                                  _lib.stream_ptr(positions.device))
                _lib.check(rc, RuntimeError)
        return visible

    def forward(self, means3D, means2D, opacities, shs=None, colors_precomp=None, scales=None, rotations=None,
                cov3D_precomp=None, grad_sinks=None):
        rs = self.raster_settings
        # same validation, exception type and messages as DGR/dgr_3dgs/__init__.py:178-182
        if (shs is None and colors_precomp is None) or (shs is not None and colors_precomp is not None):
            raise Exception('Please provide excatly one of either SHs or precomputed colors!')
        if ((scales is None or rotations is None) and cov3D_precomp is None) or \
                ((scales is not None or rotations is not None) and cov3D_precomp is not None):
            raise Exception('Please provide exactly one of either scale/rotation pair or precomputed 3D covariance!')
        return rasterize_gaussians(means3D, means2D, shs, colors_precomp, opacities, scales, rotations,
                                   cov3D_precomp, rs, grad_sinks)
