"""View-parallel data-parallel training of LGDWT-GS (one process per GPU, NCCL over NVLink / NVSwitch).

The reference is single-process, single-GPU and steps Adam after every single view (LG/train.py:105-119,278-288);
it has no distributed code at all.  The path shards naturally by camera view: every rank holds a
replica of the Gaussian parameters, renders its slice of the step's view batch (views r, r+N, ...), accumulates the
parameter gradients in ONE flat fp32 bucket (59 floats = 236 B per Gaussian at SH degree 3), and the only exchange
step is a single in-place all-reduce of that bucket before an identical Adam update on every rank.  N = 1 with one
view per step and a `TrainSchedule` reproduces the reference iteration: the running-mean DWT scale (LG/train.py:190-196,
kept on the device), `oneupSHdegree` every 1000 iterations (:101-103), the exponential position learning rate
(LG/scene/gaussian_model.py:203-223) and the inverse-depth L1 term (LG/train.py:204-216) — checked parameter for
parameter against the stock op chain in tests/test_trainer_gpu.py.  Every `densification_interval` steps the densification statistics
are all-reduced too (sum, sum, max) and `densify_and_prune` runs identically on every rank with unit normals drawn
from a generator all ranks seed alike.

Storage: raw (pre-activation) parameters in one flat, field-major buffer — xyz 3 | SH 3M | opacity 1 | scaling 3 |
rotation 4, every field a contiguous (P, w) slab.  A Gaussian's SH row is f_dc followed by f_rest, i.e. exactly the
(P, M, 3) `shs` tensor the rasterizer takes (the reference concatenates its two parameters per call,
LG/scene/gaussian_model.py:121-124); the six Adam groups of the reference (:178-211) remain addressable through
`field("f_dc")` / `field("f_rest")` and keep their own learning rates (lg_adam_step_split).

On CUDA the whole step runs in fused kernels: `lg_activate_forward` (sigmoid / exp / normalize), the rasterizer, the
fused image loss, a backward whose preprocess kernel applies the activation chain rule and adds straight into the
flat bucket (`GradSinks(raw_rot_norm=...)`), `lg_densify_stats`, `lg_adam_step_split`, `lg_densify_plan/apply`.
`render_fn` / `loss_fn` / `densify_fn` are injectable so the sharding / bucketing / optimiser / densification
protocol can be tested on CPU with the gloo backend (tests/test_host_cpu.py).
"""
import ctypes
import contextlib
import math
from typing import Callable, NamedTuple, Optional, Sequence

import torch
import torch.distributed as dist

GROUPS = ("xyz", "f_dc", "f_rest", "opacity", "scaling", "rotation")  # the reference's Adam groups


def fields_for(sh_degree):
    """storage slabs of the flat buffer: (name, floats per Gaussian)"""
    return (("xyz", 3), ("shs", 3 * (sh_degree + 1) ** 2), ("opacity", 1), ("scaling", 3), ("rotation", 4))


FIELDS = fields_for(3)
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


class TrainSchedule(NamedTuple):
    """What LG/train.py changes from one iteration to the next besides the parameters themselves
    (defaults: LG/arguments/__init__.py:77-100)."""
    iterations: int = 30_000
    position_lr_init: float = 0.00016
    position_lr_final: float = 0.0000016
    position_lr_delay_mult: float = 0.01      # unused by the reference too: it never passes lr_delay_steps
    position_lr_max_steps: int = 30_000
    spatial_lr_scale: float = 1.0             # = scene.cameras_extent (create_from_pcd, gaussian_model.py:149-150)
    oneup_sh_every: int = 1000                # LG/train.py:101-103
    start_sh_degree: int = 0                  # GaussianModel.active_sh_degree starts at 0 (gaussian_model.py:53)
    depth_l1_weight_init: float = 1.0
    depth_l1_weight_final: float = 0.01

    def position_lr(self, iteration):
        """get_expon_lr_func(lr_init * scale, lr_final * scale, max_steps) — LG/utils/general_utils.py:29-63 with
        lr_delay_steps = 0, as GaussianModel.training_setup calls it (gaussian_model.py:203-206)"""
        a, b = self.position_lr_init * self.spatial_lr_scale, self.position_lr_final * self.spatial_lr_scale
        if iteration < 0 or (a == 0.0 and b == 0.0):
            return 0.0
        t = min(max(iteration / self.position_lr_max_steps, 0.0), 1.0)
        return math.exp(math.log(a) * (1 - t) + math.log(b) * t)

    def depth_l1_weight(self, iteration):
        """get_expon_lr_func(depth_l1_weight_init, depth_l1_weight_final, max_steps=iterations) — LG/train.py:69"""
        a, b = self.depth_l1_weight_init, self.depth_l1_weight_final
        if iteration < 0 or (a == 0.0 and b == 0.0):
            return 0.0
        t = min(max(iteration / self.iterations, 0.0), 1.0)
        return math.exp(math.log(a) * (1 - t) + math.log(b) * t)


class DensifyConfig(NamedTuple):
    """LG/arguments/__init__.py:91-97 and the constants of LG/train.py:265-276"""
    densify_from_iter: int = 500
    densify_until_iter: int = 15_000
    densification_interval: int = 100
    opacity_reset_interval: int = 3000
    densify_grad_threshold: float = 0.0002
    min_opacity: float = 0.005
    percent_dense: float = 0.01
    size_threshold: float = 20.0  # used once iteration > opacity_reset_interval (train.py:272)
    cameras_extent: float = 1.0
    white_background: bool = False  # extra opacity reset at densify_from_iter (train.py:275)


class FlatGaussians:
    """Raw Gaussian parameters, their gradient bucket and Adam moments in flat field-major fp32 buffers."""

    def __init__(self, P, device, sh_degree=3):
        self.device, self.sh_degree = device, sh_degree
        self.active_sh_degree = sh_degree  # a TrainSchedule starts it at 0 and raises it (oneupSHdegree)
        self.fields = fields_for(sh_degree)
        self.floats = sum(w for _, w in self.fields)
        self.M = (sh_degree + 1) ** 2
        self.step_count = 0
        n = self.floats * self.padded(int(P))
        self._install(int(P), *(torch.zeros(n, dtype=torch.float32, device=device) for _ in range(3)))

    @staticmethod
    def padded(P):
        """row stride of the slabs: P rounded up to a multiple of 4, so that every slab starts 16-byte aligned (the
        rasterizer reads rotations and writes their gradients as float4).  Padding rows stay zero."""
        return (int(P) + 3) // 4 * 4

    def _install(self, P, data, exp_avg, exp_avg_sq):
        self.P, self.stride = P, self.padded(P)
        self.data, self.exp_avg, self.exp_avg_sq = data, exp_avg, exp_avg_sq
        self.grad = torch.zeros_like(data)
        self._slices, off = {}, 0
        for name, w in self.fields:
            self._slices[name] = (off * self.stride, (off + w) * self.stride, w)
            off += w
        self._act = None

    def replace(self, P_new, data, exp_avg, exp_avg_sq):
        """install the buffers a densification produced (every per-Gaussian side array is re-created by the caller)"""
        assert data.numel() == exp_avg.numel() == exp_avg_sq.numel() == self.floats * self.padded(P_new)
        self._install(int(P_new), data, exp_avg, exp_avg_sq)

    def slab(self, name, buf=None):
        a, _b, w = self._slices[name]
        return (self.data if buf is None else buf)[a:a + self.P * w].view(self.P, w)

    def field(self, name, buf=None):
        """(P, w) view of one of the reference's six parameter groups (f_dc / f_rest are column ranges of the SH slab)"""
        if name == "f_dc":
            return self.slab("shs", buf)[:, :3]
        if name == "f_rest":
            return self.slab("shs", buf)[:, 3:]
        return self.slab(name, buf)

    @classmethod
    def from_scene(cls, scene, device):
        """Initialise from an lgdwt_b200.scenes.Scene (activated values -> raw parameters, the inverse of the
        activations in LG/scene/gaussian_model.py:36-50)."""
        g = cls(scene.means3D.shape[0], device, scene.sh_degree)
        t = lambda a: torch.as_tensor(a, dtype=torch.float32, device=device)
        g.slab("xyz").copy_(t(scene.means3D))
        g.slab("shs").copy_(t(scene.shs).reshape(g.P, -1))
        op = t(scene.opacities).clamp(1e-6, 1 - 1e-6)
        g.slab("opacity").copy_(torch.log(op / (1 - op)))
        g.slab("scaling").copy_(torch.log(t(scene.scales)))
        g.slab("rotation").copy_(t(scene.rotations))
        return g

    # ---- generic autograd path (CPU tests, custom render functions)
    def leaves(self):
        """Fresh autograd leaves viewing the flat buffer; their .grad is added into the flat gradient bucket."""
        return {name: self.slab(name).detach().requires_grad_(True) for name, _ in self.fields}

    def accumulate(self, leaves):
        for name, _ in self.fields:
            g = leaves[name].grad
            if g is not None:
                self.slab(name, self.grad).add_(g)

    def activated(self, leaves):
        """get_xyz / get_features / get_opacity / get_scaling / get_rotation of LG/scene/gaussian_model.py:102-130"""
        return dict(means3D=leaves["xyz"], shs=leaves["shs"].view(self.P, self.M, 3),
                    opacities=torch.sigmoid(leaves["opacity"]), scales=torch.exp(leaves["scaling"]),
                    rotations=torch.nn.functional.normalize(leaves["rotation"]))

    # ---- fused CUDA path
    def activate(self):
        """one launch: activated opacity / scale / rotation (+ |raw rotation| for the backward chain rule); xyz and
        the SH slab are used in place.  No autograd graph: gradients flow through GradSinks."""
        from . import _lib
        if not self.data.is_cuda:
            raise RuntimeError("FlatGaussians.activate: the fused path needs CUDA buffers (no CPU fallback)")
        if self._act is None or self._act["opacities"].shape[0] != self.P:
            new = lambda w: torch.empty((self.P, w), dtype=torch.float32, device=self.device)
            self._act = dict(opacities=new(1), scales=new(3), rotations=new(4),
                             rot_norm=torch.empty(self.P, dtype=torch.float32, device=self.device))
        a = self._act
        with torch.cuda.device(self.device):
            rc = _lib.lib.lg_activate_forward(self.P, _lib.ptr(self.slab("opacity")), _lib.ptr(self.slab("scaling")),
                                              _lib.ptr(self.slab("rotation")), _lib.ptr(a["opacities"]),
                                              _lib.ptr(a["scales"]), _lib.ptr(a["rotations"]), _lib.ptr(a["rot_norm"]),
                                              _lib.stream_ptr(self.device))
        _lib.check(rc, RuntimeError)
        return dict(means3D=self.slab("xyz"), shs=self.slab("shs").view(self.P, self.M, 3), opacities=a["opacities"],
                    scales=a["scales"], rotations=a["rotations"], rot_norm=a["rot_norm"])

    def grad_sinks(self, rot_norm, accumulate):
        from diff_gaussian_rasterization import GradSinks
        s = lambda n: self.slab(n, self.grad)
        return GradSinks(s("xyz"), s("shs").view(self.P, self.M, 3), s("opacity"), s("scaling"), s("rotation"),
                         accumulate=accumulate, raw_rot_norm=rot_norm)

    def adopt(self, data_storage, grad_storage):
        """move the parameters and the gradient bucket into caller-provided storage of the same size (the IPC-shared
        block of a PeerExchange); views handed out earlier become stale"""
        assert data_storage.numel() == grad_storage.numel() == self.data.numel()
        data_storage.copy_(self.data)
        grad_storage.copy_(self.grad)
        self.data, self.grad = data_storage, grad_storage
        self._act = None

    def adam_segments(self, cfg: AdamConfig):
        """(ends, lr_a, lr_b, row_width, row_split) ctypes arrays for lg_adam_step_split / lg_peer_reduce_adam: one
        segment per slab, the SH slab split at column 3 into the f_dc and f_rest learning rates"""
        lrs = dict(xyz=cfg.lr_xyz, f_dc=cfg.lr_f_dc, f_rest=cfg.lr_f_rest, opacity=cfg.lr_opacity,
                   scaling=cfg.lr_scaling, rotation=cfg.lr_rotation)
        n_seg = len(self.fields)
        names = [name for name, _ in self.fields]
        ends = (ctypes.c_longlong * n_seg)(*[self._slices[n][1] for n in names])
        lr_a = (ctypes.c_float * n_seg)(*[float(lrs["f_dc" if n == "shs" else n]) for n in names])
        lr_b = (ctypes.c_float * n_seg)(*[float(lrs["f_rest" if n == "shs" else n]) for n in names])
        width = (ctypes.c_int * n_seg)(*[self._slices[n][2] if n == "shs" else 1 for n in names])
        split = (ctypes.c_int * n_seg)(*[3 if n == "shs" else 0 for n in names])
        return ends, lr_a, lr_b, width, split

    def adam_step(self, cfg: AdamConfig, grad_scale: float = 1.0):
        """torch.optim.Adam semantics (bias-corrected, eps outside the sqrt), the reference's six learning rates;
        identical on every rank because the reduced gradient bucket is identical.  CUDA buffers take ONE fused
        kernel launch (lg_adam_step_split, 28 B of traffic per element); the group-by-group torch expression below is
        the host-side statement of the same update, used by the CPU (gloo) tests of the trainer logic."""
        self.step_count += 1
        b1, b2 = cfg.beta1, cfg.beta2
        lrs = dict(xyz=cfg.lr_xyz, f_dc=cfg.lr_f_dc, f_rest=cfg.lr_f_rest, opacity=cfg.lr_opacity,
                   scaling=cfg.lr_scaling, rotation=cfg.lr_rotation)
        if self.P == 0:
            return
        if self.data.is_cuda:
            from . import _lib
            ends, lr_a, lr_b, width, split = self.adam_segments(cfg)
            with torch.cuda.device(self.data.device):
                rc = _lib.lib.lg_adam_step_split(self.data.data_ptr(), self.grad.data_ptr(), self.exp_avg.data_ptr(),
                                                 self.exp_avg_sq.data_ptr(), self.data.numel(), len(ends), ends, lr_a,
                                                 lr_b, width, split, b1, b2, cfg.eps, self.step_count,
                                                 float(grad_scale), _lib.stream_ptr(self.data.device))
            _lib.check(rc, RuntimeError)
            return
        bc1, bc2 = 1 - b1 ** self.step_count, 1 - b2 ** self.step_count
        if grad_scale != 1.0:
            self.grad.mul_(grad_scale)
        self.exp_avg.mul_(b1).add_(self.grad, alpha=1 - b1)
        self.exp_avg_sq.mul_(b2).addcmul_(self.grad, self.grad, value=1 - b2)
        for name in GROUPS:
            denom = (self.field(name, self.exp_avg_sq).sqrt() / math.sqrt(bc2)).add_(cfg.eps)
            self.field(name).addcdiv_(self.field(name, self.exp_avg), denom, value=-lrs[name] / bc1)

    def checksum(self):
        return self.data.double().sum()


def views_of_rank(num_views, rank, world):
    """round-robin view sharding: rank r renders views r, r+N, r+2N, ..."""
    return list(range(rank, num_views, world))


def _raster_settings(cam, bg, sh_degree, antialiasing):
    from diff_gaussian_rasterization import GaussianRasterizationSettings
    return GaussianRasterizationSettings(
        image_height=cam["H"], image_width=cam["W"], tanfovx=cam["tanfovx"], tanfovy=cam["tanfovy"], bg=bg,
        scale_modifier=1.0, viewmatrix=cam["viewmatrix"], projmatrix=cam["projmatrix"], sh_degree=sh_degree,
        campos=cam["campos"], prefiltered=False, debug=False, antialiasing=antialiasing)


def default_render(act, cam, bg, sh_degree=3, antialiasing=False):
    """One view through the B200-native operator, mirroring LG/gaussian_renderer/__init__.py:18-128 (autograd path:
    `act` holds differentiable activated parameters)."""
    from diff_gaussian_rasterization import GaussianRasterizer
    means2D = torch.zeros_like(act["means3D"], requires_grad=True)
    color, radii, invdepth = GaussianRasterizer(_raster_settings(cam, bg, sh_degree, antialiasing))(
        means3D=act["means3D"], means2D=means2D, shs=act["shs"], colors_precomp=None, opacities=act["opacities"],
        scales=act["scales"], rotations=act["rotations"], cov3D_precomp=None)
    return color.clamp(0, 1), radii, means2D


def fused_render(g: FlatGaussians, cam, bg, accumulate, antialiasing=False, with_depth=False, act=None):
    """One view of the fused path: raw parameters -> lg_activate_forward -> rasterizer at the active SH degree; the
    backward adds the RAW-parameter gradients into g.grad.  Returns (image, radii, means2D[, invdepth]) like the
    reference's render() (LG/gaussian_renderer/__init__.py:119-128).  `act`: the result of an earlier g.activate() of
    the same step (the parameters do not change between the views of a step)."""
    from diff_gaussian_rasterization import GaussianRasterizer
    if act is None:
        act = g.activate()
    means2D = torch.zeros((g.P, 3), dtype=torch.float32, device=g.device, requires_grad=True)
    color, radii, invdepth = GaussianRasterizer(_raster_settings(cam, bg, g.active_sh_degree, antialiasing))(
        means3D=act["means3D"], means2D=means2D, shs=act["shs"], colors_precomp=None, opacities=act["opacities"],
        scales=act["scales"], rotations=act["rotations"], cov3D_precomp=None,
        grad_sinks=g.grad_sinks(act["rot_norm"], accumulate))
    if with_depth:
        return color.clamp(0, 1), radii, means2D, invdepth
    return color.clamp(0, 1), radii, means2D


def default_loss(image, gt, lambda_dssim=0.2, dwt_scale=1.0, patch_weight=0.1, cfg=None):
    """The iteration loss of LG/train.py:128-202 with fused kernels: base = (1-lambda)*L1 + lambda*(1-SSIM)
    (`fused_photometric_loss`), plus `dwt_scale * dwt + patch_weight * patch` (`fused_dwt_loss`).  `dwt_scale` is the
    caller's running-mean ratio (train.py:190-196); three launches forward, two backward."""
    from .dwt_loss import DWTLossConfig, fused_dwt_loss
    from .photometric import fused_photometric_loss
    l1, ssim = fused_photometric_loss(image, gt)
    dwt, patch, _ = fused_dwt_loss(image, gt, cfg or DWTLossConfig())
    return (1.0 - lambda_dssim) * l1 + lambda_dssim * (1.0 - ssim) + dwt_scale * dwt + patch_weight * patch


class RunningMeanLoss:
    """The whole iteration loss of LG/train.py:128-216 as one callable with the reference's state: the running-mean
    ratio that rescales the DWT term every iteration (:190-196 — `rm = 0.95 rm + 0.05 base/(dwt + 1e-8)`, used clamped
    to [0.1, 10] in the SAME iteration) lives in a device scalar, so there is no `.item()` in the loop.  With a depth
    target the inverse-depth L1 term of :204-216 is added."""

    def __init__(self, device, lambda_dssim=0.2, patch_weight=0.1, cfg=None):
        self.running_mean = torch.ones((), dtype=torch.float32, device=device)
        self.lambda_dssim, self.patch_weight, self.cfg = lambda_dssim, patch_weight, cfg
        self.terms = None  # the last call's term values (fused_image_loss), for logging

    def __call__(self, image, gt, invdepth=None, depth_target=None, depth_weight=0.0):
        from .dwt_loss import DWTLossConfig
        from .image_loss import fused_image_loss
        loss, self.terms = fused_image_loss(image, gt, self.running_mean, self.cfg or DWTLossConfig(), self.lambda_dssim,
                                            self.patch_weight, update_running_mean=True)
        if depth_target is not None and depth_weight > 0.0 and invdepth is not None:
            mono, mask = depth_target
            loss = loss + depth_weight * torch.abs((invdepth - mono) * mask).mean()
        return loss

    def sync(self, group=None):
        """data-parallel runs: the ranks saw different views; average the running mean so that the replicas weigh the
        next step alike (one 4-byte all-reduce, issued before the exchange step)"""
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            dist.all_reduce(self.running_mean, op=dist.ReduceOp.SUM, group=group)
            self.running_mean.div_(dist.get_world_size(group))


def _default_densify(g, stats, cfg, max_screen_size, generator):
    from . import densify
    return densify.densify_and_prune(g, stats, cfg.densify_grad_threshold, cfg.min_opacity, cfg.cameras_extent,
                                     max_screen_size, generator=generator, percent_dense=cfg.percent_dense)


class ViewParallelTrainer:
    def __init__(self, gaussians: FlatGaussians, adam: AdamConfig = AdamConfig(),
                 render_fn: Optional[Callable] = None, loss_fn: Callable = default_loss,
                 group: Optional[dist.ProcessGroup] = None, densify: Optional[DensifyConfig] = None,
                 densify_fn: Optional[Callable] = None, stats_fn: Optional[Callable] = None,
                 reset_opacity_fn: Optional[Callable] = None, seed: int = 0, exchange: str = "nccl",
                 schedule: Optional[TrainSchedule] = None, view_streams: int = 1):
        """render_fn=None selects the fused CUDA path (`fused_render`); a custom render_fn(act, cam, bg) ->
        (image, radii[, viewspace_points]) runs through autograd leaves.  densify=None disables density control.
        exchange="nccl": all-reduce of the bucket + the same full Adam pass on every rank; exchange="peer": the
        fused reduce-scatter + Adam + all-gather kernel over NVLink peer memory (lgdwt_b200.peer), falling back to
        "nccl" (with `peer_unavailable` saying why) when the ranks cannot map each other's memory.
        schedule: a TrainSchedule turns on what LG/train.py varies per iteration (active SH degree from 0, position
        learning rate decay, depth-L1 weight); loss_fn may be a RunningMeanLoss (stateful DWT scale).
        view_streams=2 (fused path): consecutive views of a step alternate between two CUDA streams, so that the
        latency-bound head of view k+1 (activation is done once per step; preprocess, tile scan, scatter, tile sort)
        runs under the issue-bound blend backward of view k.  Everything with state stays in view order — the loss
        (running mean) waits for the previous view's loss, the backward (bucket, densification statistics) for the
        previous view's backward — so a step computes exactly what the single-stream step computes."""
        if view_streams not in (1, 2):
            raise ValueError("view_streams must be 1 or 2")
        self.view_streams = ([torch.cuda.Stream(device=gaussians.device) for _ in range(2)]
                             if view_streams == 2 and render_fn is None and gaussians.data.is_cuda else None)
        self.g, self.adam, self.render_fn, self.loss_fn, self.group = gaussians, adam, render_fn, loss_fn, group
        self.distributed = dist.is_available() and dist.is_initialized()
        self.rank = dist.get_rank(group) if self.distributed else 0
        self.world = dist.get_world_size(group) if self.distributed else 1
        self.densify_cfg = densify
        self.densify_fn = densify_fn or _default_densify
        self.stats_fn, self.reset_opacity_fn = stats_fn, reset_opacity_fn
        self.iteration = 0
        self.schedule = schedule
        if schedule is not None:
            gaussians.active_sh_degree = min(schedule.start_sh_degree, gaussians.sh_degree)
        self.stats = None
        # every rank seeds this generator alike: the split samples are identical on all replicas
        self.generator = torch.Generator(device=gaussians.device).manual_seed(seed)
        if densify is not None:
            self._new_stats()
        self.peer, self.peer_unavailable, self._want_peer = None, "", exchange == "peer"
        if exchange not in ("nccl", "peer"):
            raise ValueError("exchange must be 'nccl' or 'peer'")
        self._make_peer()

    def _make_peer(self):
        """(re)create the IPC-shared parameter / gradient block for the current buffer size — collective"""
        self.peer = None
        if not (self._want_peer and self.distributed and self.world > 1 and self.g.data.is_cuda and self.g.P > 0):
            return
        if (3 * self.g.M) % 4 != 0:
            # the 16-byte Adam path of the peer kernels needs the SH row (3, 12, 27 or 48 floats) to be a multiple of 4
            # floats; SH degree 0 and 2 take the NCCL exchange (same decision on every rank: it only depends on M)
            self.peer_unavailable = "SH row of %d floats is not a multiple of 4 (peer kernels use 16-byte accesses)" % (3 * self.g.M)
            return
        from .peer import PeerExchange
        self.peer, self.peer_unavailable = PeerExchange.create(self.g.data.numel(), self.g.device, self.group)
        if self.peer is not None:
            self.g.adopt(self.peer.param, self.peer.grad)

    def _collect_moments(self):
        """peer exchange keeps every Adam moment only on its owner rank; a densification needs them everywhere"""
        if self.peer is None:
            return
        self.peer.check()  # a sync point anyway: surface a timed-out flag barrier before the set is rebuilt
        n4 = self.g.data.numel() // 4
        lo, hi = 4 * (n4 * self.rank // self.world), 4 * (n4 * (self.rank + 1) // self.world)
        for buf in (self.g.exp_avg, self.g.exp_avg_sq):
            buf[:lo].zero_()
            buf[hi:].zero_()
            dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=self.group)

    def _new_stats(self):
        from .densify import DensifyStats
        self.stats = DensifyStats(self.g.P, self.g.device)

    def _add_stats(self, viewspace, radii):
        if self.stats is None or viewspace is None or viewspace.grad is None:
            return
        if self.stats_fn is not None:
            self.stats_fn(self.stats, viewspace.grad, radii)
        else:
            from . import densify
            densify.add_densification_stats(self.stats, viewspace.grad, radii)

    def accumulate_views(self, cams: Sequence[dict], gts: Sequence[torch.Tensor], bg: torch.Tensor,
                         depths: Optional[Sequence] = None):
        """forward + backward of this rank's views, gradients summed into the flat bucket; returns the local loss sum.
        depths[v] = (mono_invdepth (1,H,W), depth_mask (1,H,W)) or None: the view's inverse-depth target
        (`viewpoint_cam.invdepthmap` / `depth_mask` of a depth-reliable camera, LG/train.py:206-213)."""
        total = torch.zeros((), dtype=torch.float32, device=self.g.device)
        mine = views_of_rank(len(cams), self.rank, self.world)
        fused = self.render_fn is None
        if not fused or not mine:
            self.g.grad.zero_()
        depth_w = self.schedule.depth_l1_weight(self.iteration) if self.schedule is not None else 0.0
        overlap = fused and self.view_streams is not None and len(mine) > 1
        act, prev, prev_loss_done = None, None, None
        if overlap:
            main = torch.cuda.current_stream(self.g.device)
            act = self.g.activate()
            for s in self.view_streams:
                s.wait_stream(main)
        for k, v in enumerate(mine):
            stream = self.view_streams[k % 2] if overlap else None
            with (torch.cuda.stream(stream) if overlap else contextlib.nullcontext()):
                target = depths[v] if depths is not None else None
                invdepth = None
                if fused:  # the first view of the step overwrites the bucket: no zero-fill pass
                    if target is not None and depth_w > 0.0:
                        image, radii, viewspace, invdepth = fused_render(self.g, cams[v], bg, accumulate=k > 0,
                                                                         with_depth=True, act=act)
                    else:
                        image, radii, viewspace = fused_render(self.g, cams[v], bg, accumulate=k > 0, act=act)
                    leaves = None
                else:
                    leaves = self.g.leaves()
                    out = self.render_fn(self.g.activated(leaves), cams[v], bg)
                    image, radii = out[0], out[1]
                    viewspace = out[2] if len(out) > 2 else None
                    invdepth = out[3] if len(out) > 3 else None
                if prev_loss_done is not None:  # a stateful loss (running mean) advances view by view
                    stream.wait_event(prev_loss_done)
                if target is not None and depth_w > 0.0 and invdepth is not None:
                    loss = self.loss_fn(image, gts[v], invdepth=invdepth, depth_target=target, depth_weight=depth_w)
                else:
                    loss = self.loss_fn(image, gts[v])
                if overlap:
                    prev_loss_done = torch.cuda.Event()
                    prev_loss_done.record(stream)
                    if prev is not None:  # bucket / statistics / `total` order
                        stream.wait_stream(prev)
                loss.backward()
                if leaves is not None:
                    self.g.accumulate(leaves)
                if self.densify_cfg is not None and self.iteration < self.densify_cfg.densify_until_iter:
                    self._add_stats(viewspace, radii)
                total += loss.detach()
                prev = stream
        if overlap:
            for s in self.view_streams:
                main.wait_stream(s)
        return total

    def reduce_gradients(self):
        """THE exchange step: one in-place sum all-reduce of the 236 B/Gaussian bucket (NCCL picks NVLS in-switch
        reduction on NVSwitch systems)."""
        if self.distributed and self.world > 1:
            dist.all_reduce(self.g.grad, op=dist.ReduceOp.SUM, group=self.group)

    def reduce_stats(self):
        """densification statistics of the ranks' view slices -> the statistics of the whole batch"""
        if self.distributed and self.world > 1 and self.stats is not None:
            dist.all_reduce(self.stats.xyz_gradient_accum, op=dist.ReduceOp.SUM, group=self.group)
            dist.all_reduce(self.stats.denom, op=dist.ReduceOp.SUM, group=self.group)
            dist.all_reduce(self.stats.max_radii2D, op=dist.ReduceOp.MAX, group=self.group)

    def maybe_densify(self):
        """the density-control schedule of LG/train.py:265-276, applied identically on every rank.  Returns True when
        the Gaussian set was rebuilt (the reference's optimizer then has nothing to step on: the new Parameters have
        no .grad, LG/scene/gaussian_model.py:331-409 — the Adam update of this iteration is skipped)."""
        c, it = self.densify_cfg, self.iteration
        if c is None or it >= c.densify_until_iter:
            return False
        rebuilt = False
        if it > c.densify_from_iter and it % c.densification_interval == 0:
            self.reduce_stats()
            self._collect_moments()
            mss = c.size_threshold if it > c.opacity_reset_interval else None
            self.densify_fn(self.g, self.stats, c, mss, self.generator)
            self._new_stats()
            self._make_peer()
            rebuilt = True
        if it % c.opacity_reset_interval == 0 or (c.white_background and it == c.densify_from_iter):
            if self.reset_opacity_fn is not None:
                self.reset_opacity_fn(self.g)
            else:
                from . import densify
                densify.reset_opacity(self.g)
        return rebuilt

    def begin_iteration(self):
        """what LG/train.py does at the top of an iteration (:99-103): position learning rate of this iteration, and
        one more SH degree every 1000 iterations"""
        self.iteration += 1
        sch = self.schedule
        if sch is not None:
            self.adam = self.adam._replace(lr_xyz=sch.position_lr(self.iteration))
            if sch.oneup_sh_every > 0 and self.iteration % sch.oneup_sh_every == 0:
                self.g.active_sh_degree = min(self.g.active_sh_degree + 1, self.g.sh_degree)

    def step(self, cams, gts, bg, num_views_scale=True, depths=None):
        self.begin_iteration()
        loss = self.accumulate_views(cams, gts, bg, depths)
        if hasattr(self.loss_fn, "sync"):
            self.loss_fn.sync(self.group)
        if not self.maybe_densify():
            scale = 1.0 / len(cams) if (num_views_scale and len(cams) > 1) else 1.0  # mean over the view batch
            self.exchange_and_update(scale)
        return loss

    def exchange_and_update(self, grad_scale=1.0):
        """the exchange step + optimizer: NCCL all-reduce then Adam everywhere, or the fused peer-memory kernel"""
        if self.peer is not None:
            self.g.step_count += 1
            self.peer.reduce_adam(self.g.exp_avg, self.g.exp_avg_sq, self.g.adam_segments(self.adam), self.adam,
                                  self.g.step_count, grad_scale)
            if self.g.step_count % 64 == 0:
                self.peer.check()  # a timed-out flag barrier must not go unnoticed once densification is over
        else:
            self.reduce_gradients()
            self.g.adam_step(self.adam, grad_scale=grad_scale)

    def replicas_in_sync(self):
        """parameter count and checksum min == max over ranks; also raises if a peer-exchange barrier
        timed out since the last check"""
        if self.peer is not None:
            self.peer.check()
        c = torch.stack([self.g.checksum(), torch.tensor(float(self.g.P), dtype=torch.float64, device=self.g.device)])
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
