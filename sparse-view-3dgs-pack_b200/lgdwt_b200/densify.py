"""Adaptive density control on the flat Gaussian buffer (host side of csrc/densify.cu).

Mirrors the reference's GaussianModel methods for the `dp.FlatGaussians` storage:
  add_densification_stats(...)   LG/scene/gaussian_model.py:478-480 + the max_radii2D update of LG/train.py:268
  densify_and_prune(...)         LG/scene/gaussian_model.py:456-476 (clone :437-454, split :411-435, prune :347-363,
                                 Adam-state edits :331-409)
  reset_opacity(...)             LG/scene/gaussian_model.py:258-261
One decision pass + one scan + one gather pass per field instead of four rounds of boolean indexing and torch.cat
over six parameter tensors and twelve moment tensors.  CUDA only: there is no CPU fallback.
"""
import torch

from . import _lib


class DensifyStats:
    """xyz_gradient_accum, denom (LG/scene/gaussian_model.py:170-171) and max_radii2D (:167), flat (P,) fp32"""

    def __init__(self, P, device):
        self.xyz_gradient_accum = torch.zeros(P, dtype=torch.float32, device=device)
        self.denom = torch.zeros(P, dtype=torch.float32, device=device)
        self.max_radii2D = torch.zeros(P, dtype=torch.float32, device=device)


def _need_cuda(t, what):
    if not t.is_cuda:
        raise RuntimeError("%s: tensors must live on a CUDA device; there is no CPU path" % what)


def add_densification_stats(stats, viewspace_grad, radii):
    """visible = radii > 0: accum += |grad.xy|, denom += 1, max_radii2D = max(max_radii2D, radii) — one launch"""
    _need_cuda(viewspace_grad, "add_densification_stats")
    P = stats.denom.numel()
    grad = viewspace_grad.contiguous()
    radii = radii.contiguous()
    if grad.dtype != torch.float32 or grad.numel() != 3 * P or radii.dtype != torch.int32 or radii.numel() != P:
        raise RuntimeError("add_densification_stats: expected grad (P,3) fp32 and radii (P,) int32")
    with torch.cuda.device(grad.device):
        rc = _lib.lib.lg_densify_stats(P, _lib.ptr(grad), _lib.ptr(radii), _lib.ptr(stats.xyz_gradient_accum),
                                       _lib.ptr(stats.denom), _lib.ptr(stats.max_radii2D), _lib.stream_ptr(grad.device))
    _lib.check(rc, RuntimeError)


def reset_opacity(g):
    _need_cuda(g.data, "reset_opacity")
    with torch.cuda.device(g.device):
        rc = _lib.lib.lg_reset_opacity(g.P, _lib.ptr(g.slab("opacity")), _lib.ptr(g.slab("opacity", g.exp_avg)),
                                       _lib.ptr(g.slab("opacity", g.exp_avg_sq)), _lib.stream_ptr(g.device))
    _lib.check(rc, RuntimeError)


def densify_and_prune(g, stats, max_grad, min_opacity, extent, max_screen_size, generator=None, percent_dense=0.01,
                      eps=None):
    """Rebuilds g (parameters + Adam moments) in place of GaussianModel.densify_and_prune.  `eps`: optional (2S, 3)
    unit normals for the S split parents (row k*S + j = k-th child of the j-th parent); by default they are drawn
    from `generator` — in data-parallel runs a generator every rank seeded alike.  Returns the four counts
    {kept, cloned, split_parents, split_kept} and the new P.  One host read-back (the counts size the new buffers);
    the reference syncs several times per call through boolean indexing."""
    _need_cuda(g.data, "densify_and_prune")
    dev, P = g.device, g.P
    totals = torch.zeros(4, dtype=torch.int32, device=dev)
    src_index = torch.empty(2 * max(P, 1), dtype=torch.int32, device=dev)
    eps_row = torch.empty(2 * max(P, 1), dtype=torch.int32, device=dev)
    scratch = torch.empty(_lib.lib.lg_densify_scratch_bytes(P), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        rc = _lib.lib.lg_densify_plan(P, _lib.ptr(g.slab("scaling")), _lib.ptr(g.slab("opacity")),
                                      _lib.ptr(stats.xyz_gradient_accum), _lib.ptr(stats.denom), float(max_grad),
                                      float(min_opacity), float(extent), float(percent_dense),
                                      -1.0 if max_screen_size is None else float(max_screen_size),
                                      _lib.ptr(src_index), _lib.ptr(eps_row), totals.data_ptr(), _lib.ptr(scratch),
                                      scratch.numel(), _lib.stream_ptr(dev))
    _lib.check(rc, RuntimeError)
    kept, cloned, split_parents, split_kept = (int(v) for v in totals.tolist())
    P_new = kept + cloned + 2 * split_kept
    if eps is None:
        eps = torch.randn((2 * split_parents, 3), dtype=torch.float32, device=dev, generator=generator)
    eps = eps.contiguous()
    if eps.dtype != torch.float32 or eps.numel() != 6 * split_parents or eps.device != g.data.device:
        raise RuntimeError("densify_and_prune: eps must be a (2*%d, 3) fp32 tensor on %s" % (split_parents, dev))
    stride_new = g.padded(P_new)
    new = [torch.zeros(g.floats * stride_new, dtype=torch.float32, device=dev) for _ in range(3)]
    with torch.cuda.device(dev):
        rc = _lib.lib.lg_densify_apply(P, g.stride, P_new, stride_new, 3 * g.M, _lib.ptr(g.data), _lib.ptr(g.exp_avg), _lib.ptr(g.exp_avg_sq),
                                       _lib.ptr(new[0]), _lib.ptr(new[1]), _lib.ptr(new[2]), _lib.ptr(src_index),
                                       _lib.ptr(eps_row), _lib.ptr(eps), 2 * split_parents, _lib.stream_ptr(dev))
    _lib.check(rc, RuntimeError)
    g.replace(P_new, *new)
    # densification_postfix (:404-409): statistics of the new set start from zero
    stats.__init__(P_new, dev)
    return dict(kept=kept, cloned=cloned, split_parents=split_parents, split_kept=split_kept, P=P_new)
