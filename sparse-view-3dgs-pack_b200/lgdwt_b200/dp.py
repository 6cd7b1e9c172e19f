"""View-parallel data-parallel training of LGDWT-GS (one process per GPU, NCCL over NVLink / NVSwitch).

The reference is single-process, single-GPU and steps Adam after every single view (LG/train.py:105-119,278-288);
it has no distributed code at all.  The path shards naturally by camera view: every rank holds a
replica of the Gaussian parameters, renders its slice of the step's view batch (views r, r+N, ...), accumulates the
parameter gradients in ONE flat fp32 bucket (59 floats = 236 B per Gaussian: xyz 3, f_dc 3, f_rest 45, opacity 1,
scaling 3, rotation 4 — the six Adam groups of LG/scene/gaussian_model.py:178-211), and the only exchange step is a
single in-place all-reduce of that bucket before an identical Adam update on every rank.  N = 1 with one view per
step reproduces the reference iteration.

`render_fn` / `loss_fn` are injectable so the sharding / bucketing / optimiser logic can be tested on CPU with the
gloo backend (tests/test_dp_cpu.py); the defaults use the B200-native rasterizer and fused DWT loss.
"""
import math
from typing import Callable, List, NamedTuple, Optional, Sequence

import torch
import torch.distributed as dist

FIELDS = (("xyz", 3), ("f_dc", 3), ("f_rest", 45), ("opacity", 1), ("scaling", 3), ("rotation", 4))
FLOATS_PER_GAUSSIAN = sum(w for _, w in FIELDS)  # 59


class AdamConfig(NamedTuple):
    """Per-group learning rates of LG/arguments/__init__.py:84-92 (position LR is scheduled by the caller)."""
    lr_xyz: float = 0.00016
    lr_f_dc: float = 0.0025
    lr_f_rest: float = 0.0025 / 20.0
    lr_opacity: float = 0.025
    lr_scaling: float = 0.005
    lr_rotation: float = 0.001
    beta1: float = 0.9
    beta2: float = 0.999
    eps: float = 1e-15


class FlatGaussians:
    """Raw (pre-activation) Gaussian parameters in one flat (59, P) fp32 buffer, field-major so that every field is
    a contiguous slab (coalesced for the activation / Adam passes and a single NCCL message for the gradients)."""

    def __init__(self, P, device, sh_degree=3):
        self.P, self.device, self.sh_degree = int(P), device, sh_degree
        self.data = torch.zeros(FLOATS_PER_GAUSSIAN * self.P, dtype=torch.float32, device=device)
        self.grad = torch.zeros_like(self.data)
        self.exp_avg = torch.zeros_like(self.data)
        self.exp_avg_sq = torch.zeros_like(self.data)
        self.step_count = 0
        self._slices = {}
        off = 0
        for name, w in FIELDS:
            self._slices[name] = (off * self.P, (off + w) * self.P, w)
            off += w

    def field(self, name, buf=None):
        a, b, w = self._slices[name]
        return (self.data if buf is None else buf)[a:b].view(self.P, w)

    @classmethod
    def from_scene(cls, scene, device):
        """Initialise from an lgdwt_b200.scenes.Scene (activated values -> raw parameters, the inverse of the
        activations in LG/scene/gaussian_model.py:36-50)."""
        g = cls(scene.means3D.shape[0], device, scene.sh_degree)
        t = lambda a: torch.as_tensor(a, dtype=torch.float32, device=device)
        g.field("xyz").copy_(t(scene.means3D))
        g.field("f_dc").copy_(t(scene.shs[:, 0, :]))
        g.field("f_rest").copy_(t(scene.shs[:, 1:, :]).reshape(g.P, 45))
        op = t(scene.opacities).clamp(1e-6, 1 - 1e-6)
        g.field("opacity").copy_(torch.log(op / (1 - op)))
        g.field("scaling").copy_(torch.log(t(scene.scales)))
        g.field("rotation").copy_(t(scene.rotations))
        return g

    def leaves(self):
        """Fresh autograd leaves viewing the flat buffer; their .grad is routed into the flat gradient bucket."""
        out = {}
        for name, _ in FIELDS:
            leaf = self.field(name).detach().requires_grad_(True)
            out[name] = leaf
        return out

    def accumulate(self, leaves):
        for name, _ in FIELDS:
            g = leaves[name].grad
            if g is not None:
                self.field(name, self.grad).add_(g)

    def activated(self, leaves):
        """get_xyz / get_features / get_opacity / get_scaling / get_rotation of LG/scene/gaussian_model.py:102-130"""
        shs = torch.cat([leaves["f_dc"].view(self.P, 1, 3), leaves["f_rest"].view(self.P, 15, 3)], dim=1)
        return dict(means3D=leaves["xyz"], shs=shs, opacities=torch.sigmoid(leaves["opacity"]),
                    scales=torch.exp(leaves["scaling"]), rotations=torch.nn.functional.normalize(leaves["rotation"]))

    def adam_step(self, cfg: AdamConfig, grad_scale: float = 1.0):
        """torch.optim.Adam semantics (bias-corrected, eps outside the sqrt), one learning rate per field slab;
        identical on every rank because the reduced gradient bucket is identical.  CUDA buffers take ONE fused
        kernel launch (lg_adam_step, 28 B of traffic per element); the slab-by-slab torch expression below is the
        host-side statement of the same update, used by the CPU (gloo) tests of the trainer logic."""
        self.step_count += 1
        b1, b2 = cfg.beta1, cfg.beta2
        lrs = dict(xyz=cfg.lr_xyz, f_dc=cfg.lr_f_dc, f_rest=cfg.lr_f_rest, opacity=cfg.lr_opacity,
                   scaling=cfg.lr_scaling, rotation=cfg.lr_rotation)
        if self.data.is_cuda:
            import ctypes
            from . import _lib
            n_seg = len(FIELDS)
            ends = (ctypes.c_longlong * n_seg)(*[self._slices[name][1] for name, _ in FIELDS])
            lr_arr = (ctypes.c_float * n_seg)(*[float(lrs[name]) for name, _ in FIELDS])
            with torch.cuda.device(self.data.device):
                rc = _lib.lib.lg_adam_step(self.data.data_ptr(), self.grad.data_ptr(), self.exp_avg.data_ptr(),
                                           self.exp_avg_sq.data_ptr(), self.data.numel(), n_seg, ends, lr_arr, b1, b2,
                                           cfg.eps, self.step_count, float(grad_scale),
                                           _lib.stream_ptr(self.data.device))
            _lib.check(rc, RuntimeError)
            return
        bc1, bc2 = 1 - b1 ** self.step_count, 1 - b2 ** self.step_count
        if grad_scale != 1.0:
            self.grad.mul_(grad_scale)
        self.exp_avg.mul_(b1).add_(self.grad, alpha=1 - b1)
        self.exp_avg_sq.mul_(b2).addcmul_(self.grad, self.grad, value=1 - b2)
        for name, _ in FIELDS:
            a, b, _w = self._slices[name]
            denom = (self.exp_avg_sq[a:b].sqrt() / math.sqrt(bc2)).add_(cfg.eps)
            self.data[a:b].addcdiv_(self.exp_avg[a:b], denom, value=-lrs[name] / bc1)

    def checksum(self):
        return self.data.double().sum()


def views_of_rank(num_views, rank, world):
    """round-robin view sharding: rank r renders views r, r+N, r+2N, ..."""
    return list(range(rank, num_views, world))


def default_render(act, cam, bg, sh_degree=3, antialiasing=False):
    """One view through the B200-native operator, mirroring LG/gaussian_renderer/__init__.py:18-128."""
    from diff_gaussian_rasterization import GaussianRasterizationSettings, GaussianRasterizer
    rs = GaussianRasterizationSettings(
        image_height=cam["H"], image_width=cam["W"], tanfovx=cam["tanfovx"], tanfovy=cam["tanfovy"], bg=bg,
        scale_modifier=1.0, viewmatrix=cam["viewmatrix"], projmatrix=cam["projmatrix"], sh_degree=sh_degree,
        campos=cam["campos"], prefiltered=False, debug=False, antialiasing=antialiasing)
    means2D = torch.zeros_like(act["means3D"], requires_grad=True)
    color, radii, invdepth = GaussianRasterizer(rs)(means3D=act["means3D"], means2D=means2D, shs=act["shs"],
                                                     colors_precomp=None, opacities=act["opacities"],
                                                     scales=act["scales"], rotations=act["rotations"],
                                                     cov3D_precomp=None)
    return color.clamp(0, 1), radii


def default_loss(image, gt, lambda_dssim=0.2, dwt_scale=1.0, patch_weight=0.1, cfg=None):
    """The iteration loss of LG/train.py:128-202 with fused kernels: base = (1-lambda)*L1 + lambda*(1-SSIM)
    (`fused_photometric_loss`), plus `dwt_scale * dwt + patch_weight * patch` (`fused_dwt_loss`).  `dwt_scale` is the
    caller's running-mean ratio (train.py:190-196); three launches forward, two backward."""
    from .dwt_loss import DWTLossConfig, fused_dwt_loss
    from .photometric import fused_photometric_loss
    l1, ssim = fused_photometric_loss(image, gt)
    dwt, patch, _ = fused_dwt_loss(image, gt, cfg or DWTLossConfig())
    return (1.0 - lambda_dssim) * l1 + lambda_dssim * (1.0 - ssim) + dwt_scale * dwt + patch_weight * patch


class ViewParallelTrainer:
    def __init__(self, gaussians: FlatGaussians, adam: AdamConfig = AdamConfig(),
                 render_fn: Callable = default_render, loss_fn: Callable = default_loss,
                 group: Optional[dist.ProcessGroup] = None):
        self.g, self.adam, self.render_fn, self.loss_fn, self.group = gaussians, adam, render_fn, loss_fn, group
        self.distributed = dist.is_available() and dist.is_initialized()
        self.rank = dist.get_rank(group) if self.distributed else 0
        self.world = dist.get_world_size(group) if self.distributed else 1

    def accumulate_views(self, cams: Sequence[dict], gts: Sequence[torch.Tensor], bg: torch.Tensor):
        """forward + backward of this rank's views, gradients summed into the flat bucket; returns the local loss sum"""
        self.g.grad.zero_()
        total = torch.zeros((), dtype=torch.float32, device=self.g.device)
        for v in views_of_rank(len(cams), self.rank, self.world):
            leaves = self.g.leaves()
            image, _radii = self.render_fn(self.g.activated(leaves), cams[v], bg)
            loss = self.loss_fn(image, gts[v])
            loss.backward()
            self.g.accumulate(leaves)
            total += loss.detach()
        return total

    def reduce_gradients(self):
        """THE exchange step: one in-place sum all-reduce of the 236 B/Gaussian bucket (NCCL picks NVLS in-switch
        reduction on NVSwitch systems)."""
        if self.distributed and self.world > 1:
            dist.all_reduce(self.g.grad, op=dist.ReduceOp.SUM, group=self.group)

    def step(self, cams, gts, bg, num_views_scale=True):
        loss = self.accumulate_views(cams, gts, bg)
        self.reduce_gradients()
        scale = 1.0 / len(cams) if (num_views_scale and len(cams) > 1) else 1.0  # mean over the view batch
        self.g.adam_step(self.adam, grad_scale=scale)
        return loss

    def replicas_in_sync(self):
        """parameter checksum min == max over ranks"""
        c = self.g.checksum().reshape(1)
        if not (self.distributed and self.world > 1):
            return True
        lo, hi = c.clone(), c.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN, group=self.group)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX, group=self.group)
        return bool((lo == hi).all())


def camera_to_device(cam, device):
    """lgdwt_b200.scenes.Camera -> dict of device tensors / scalars used by default_render"""
    t = lambda a: torch.as_tensor(a, dtype=torch.float32, device=device)
    return dict(W=cam.image_width, H=cam.image_height, tanfovx=cam.tanfovx, tanfovy=cam.tanfovy,
                viewmatrix=t(cam.viewmatrix), projmatrix=t(cam.projmatrix), campos=t(cam.campos))
