#!/usr/bin/env python
"""bench.py — headline benchmark of the B200-native LGDWT-GS hot path.

Metric (BASELINE.json): rasterize forward+backward per view at 1 M Gaussians, 800x800, SH degree 3, reported as
whole-job views/s.  A step is a batch of VIEWS_PER_RANK = 4 views per rank (8 ranks x 4 = the 32-view batch of
BASELINE config 5; per-rank work is the same at every N: weak scaling by camera view): every view is rendered and
back-propagated, the parameter gradients (59 floats = 236 B per Gaussian) of the rank's views are accumulated in one
flat bucket by the backward kernel itself, and at N > 1 the bucket is summed over ranks with ONE all-reduce inside
the timed step — the path's only exchange step (this repo's peer-memory kernel over NVLink, csrc/peer.cu; NCCL when the
ranks cannot map each other's memory or with LGDWT_EXCHANGE=nccl).  ms/view = ms_per_step / 4 (also printed as ms_per_view).

    python bench.py --gpus 1 --steps 20 --warmup 5
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \\
        bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...     # the UNMODIFIED reference extension (baseline/_ref/site/dgr_3dgs), same metric

One JSON line on rank 0.  `value`: operator fwd+bwd with all inputs resident in HBM.  `e2e`: the same step through
the public torch operator with the step's inputs (camera matrices + ground-truth image) copied from pinned host
memory, an L1 loss against that image (`lgdwt_b200.fused_l1_loss` here, the stock torch expression in the
reference arm), backward, and a device->host read of the loss.  `dropin` (N = 1): both numbers again through the
operator exactly as the unchanged LG/train.py calls it (no GradSinks: fresh gradient tensors, autograd accumulation).
`roofline`: the dominant kernel against its bound — HBM peak from MEASURED_PEAKS.json, FP32 / MUFU / LDS / SHFL peaks
measured on the box at the start of the run (csrc/microbench.cu) — timed live with CUDA events on the launching stream;
`stages`: every stage likewise.  `cpu_baseline`: the CPU oracle (a port: the reference has no CPU rasterizer) timed on
this box's cores.  `exchange` (N > 1): NCCL vs the peer-memory kernels, plus the correctness of the peer all-reduce
against NCCL and the bit-identity of the replicas.  `image_loss` / `train_iteration` (N = 1): BASELINE configs 1 and 3.
`sustained` (N = 1): the resident step back to back for ~3 s with its own clock samples.
`view_overlap` (N = 1): the step with two views in flight on two CUDA streams.
`cfg5` (every N): BASELINE config 5 — 6 M Gaussians, 1920x1080, the 32-view batch split over the ranks (strong scaling).
The reference arm imports nothing of this repo's package: its process maps only the reference's own libraries.
"""
import argparse
import gc
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "sparse-view-3dgs-pack_b200")
REF_SITE = os.path.join(ROOT, "baseline", "_ref", "site")   # the stock reference extensions (baseline/install_reference.sh)
sys.path.insert(0, ROOT)


def _load_scenes():
    """the numpy-only scene generators, loaded by file path: the reference arm must not put this repo's package
    (whose import loads liblgdwt_b200.so) on its path"""
    import importlib.util
    spec = importlib.util.spec_from_file_location("lgdwt_scenes", os.path.join(PKG, "lgdwt_scenes.py"))
    mod = importlib.util.module_from_spec(spec)
    sys.modules["lgdwt_scenes"] = mod
    spec.loader.exec_module(mod)
    return mod


scenes = _load_scenes()

P_GAUSSIANS = 1_000_000
WIDTH = HEIGHT = 800
SH_DEGREE = 3
N_CAMERAS = 8
CFG5_VIEW_STREAMS = 2     # measured 130.0 -> 122.0 ms/step at N = 1; view streams of the trainer in the cfg5 section (ViewParallelTrainer(view_streams=...))
VIEW_STREAMS_DEFAULT = 1  # 2: consecutive views of a step alternate between two CUDA streams (Stepper.overlap)
VIEWS_PER_RANK = 4  # views per rank per step (gradient accumulation); x 8 ranks = the 32-view batch of config 5
FP32_SIMT_NOMINAL_TFLOPS = 148 * 128 * 2 * 1.965e9 / 1e12  # only a fallback: the FFMA peak is measured on the box


def env_int(name, default):
    return int(os.environ.get(name, default))


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc, self.t_mark = index, [], None, None

    def mark(self):
        """start of the window whose samples are reported (the sampler itself is started earlier: nvidia-smi needs
        ~0.2 s before its first line, longer than a short timed region)"""
        self.t_mark = time.monotonic()

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.monotonic(), [c.strip() for c in line.split(",")]))

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        t_end = time.monotonic()
        time.sleep(0.12)
        self.proc.terminate()
        window = "timed regions (value + e2e)"
        rows = [r for t, r in self.rows if self.t_mark is None or self.t_mark <= t <= t_end + 0.05]
        if not rows:  # region shorter than one sampling period: fall back to the samples of the whole loaded run
            rows, window = [r for _, r in self.rows], "whole run incl. warm-up (timed regions shorter than one sample)"
        sm = [float(r[0]) for r in rows if len(r) >= 6 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in rows if len(r) >= 6 and r[1].replace(".", "").isdigit()]
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        reasons = [n for k, n in enumerate(names) if any(len(r) >= 6 and r[2 + k].lower().startswith("active") for r in rows)]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm), "window": window}


def make_workload(device):
    sc = scenes.trained_like_scene(P_GAUSSIANS, seed=1)
    cams = [scenes.metric_camera(WIDTH, HEIGHT)] + scenes.orbit_cameras(N_CAMERAS - 1, WIDTH, HEIGHT, phase=0.4)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(device)
    params = {k: t(getattr(sc, k)).requires_grad_(True) for k in ("means3D", "shs", "opacities", "scales", "rotations")}
    return sc, cams, params


def cam_dict(cam, device):
    t = lambda a: torch.as_tensor(a, dtype=torch.float32, device=device)
    return dict(W=cam.image_width, H=cam.image_height, tanfovx=cam.tanfovx, tanfovy=cam.tanfovy,
                viewmatrix=t(cam.viewmatrix), projmatrix=t(cam.projmatrix), campos=t(cam.campos))


class Stepper:
    """One benchmark step for either implementation; identical work outside the operator.
    mode "sinks"  : this repo's operator with GradSinks (gradients accumulated by the backward kernel itself in one flat
                    bucket: the data-parallel design, what `value` / `e2e` report);
    mode "dropin" : this repo's operator exactly as the unchanged LG/train.py calls it (fresh gradient tensors per view,
                    summed by autograd) — reported next to the headline as `dropin`;
    mode "stock"  : the UNMODIFIED reference extension dgr_3dgs (built by its own setup.py into baseline/_ref/site)
                    through its own GaussianRasterizationSettings / GaussianRasterizer;
    mode "mirror" : fallback when baseline/_ref is absent: the reference's CUDA files behind oracle/ref_cuda.py."""

    def __init__(self, mode, params, device, world, view_streams=0):
        self.mode, self.p, self.device, self.world = mode, params, device, world
        self.bg = torch.zeros(3, device=device)
        P = params["means3D"].shape[0]
        self.P = P
        gen = torch.Generator(device=device).manual_seed(1234)
        self.dL = torch.randn((3, HEIGHT, WIDTH), device=device, generator=gen)
        self.means2D = torch.zeros((P, 3), device=device, requires_grad=True)
        self.copy_stream = torch.cuda.Stream(device=device)
        self.copy_done = [torch.cuda.Event() for _ in range(VIEWS_PER_RANK)]
        self.dev_gt = [[torch.empty((3, HEIGHT, WIDTH), device=device) for _ in range(VIEWS_PER_RANK)] for _ in range(2)]
        self.host_loss = [torch.zeros((), dtype=torch.float32).pin_memory() for _ in range(2)]
        self.loss_ready = [torch.cuda.Event() for _ in range(2)]
        self.e2e_steps, self.last_loss = 0, float("nan")
        # view overlap (this repo's "sinks" arm only): consecutive views of a step alternate between two streams, so that
        # the latency-bound head of view v+1 (preprocess, tile scan, scatter, tile sort) runs under the issue-bound
        # blend backward of view v.  The backward kernels that add into the shared bucket stay in view order (each
        # backward waits for the previous view's), so the step's result is the single-stream one bit for bit.
        self.overlap = mode == "sinks" and (view_streams or env_int("LGDWT_VIEW_STREAMS", VIEW_STREAMS_DEFAULT)) > 1
        self.view_streams = [torch.cuda.Stream(device=device) for _ in range(2)] if self.overlap else None
        # the flat gradient bucket (what a data-parallel step all-reduces): means3D 3 | shs 48 | opacity 1 | scales 3 | rot 4
        self.peer, self.peer_unavailable = None, ""
        if mode == "sinks" and world > 1:
            # the exchange step over NVLink peer memory (csrc/peer.cu); NCCL all-reduce when the ranks cannot map
            # each other's memory (LGDWT_EXCHANGE=nccl forces it)
            if os.environ.get("LGDWT_EXCHANGE", "peer") == "peer":
                from lgdwt_b200.peer import PeerExchange
                self.peer, self.peer_unavailable = PeerExchange.create(59 * P, device)
            else:
                self.peer_unavailable = "LGDWT_EXCHANGE=nccl"
        self.bucket = self.peer.grad if self.peer is not None else torch.zeros(59 * P, device=device)
        # which all-reduce the timed step uses: measured here, on this box and this bucket (the NVLink paths of the
        # leased GPUs differ from box to box: at N = 2 the peer kernel took 0.37 ms on one box and 0.74 ms on another, NCCL
        # 0.47 / 0.65 ms); every rank takes the same decision from the max-over-ranks times
        self.use_peer, self.exchange_pick = self.peer is not None, None
        if self.peer is not None and os.environ.get("LGDWT_EXCHANGE_PICK", "auto") == "auto":
            def t(fn):
                for _ in range(3):
                    fn()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                dist.barrier()
                torch.cuda.synchronize(device)
                a.record()
                for _ in range(5):
                    fn()
                b.record()
                torch.cuda.synchronize(device)
                ms = torch.tensor([a.elapsed_time(b) / 5], device=device)
                dist.all_reduce(ms, op=dist.ReduceOp.MAX)
                return float(ms.item())
            t_peer, t_nccl = t(lambda: self.peer.allreduce(1.0)), t(lambda: dist.all_reduce(self.bucket))
            self.use_peer = t_peer <= t_nccl
            self.exchange_pick = {"peer_ms": round(t_peer, 4), "nccl_ms": round(t_nccl, 4),
                                  "picked": "peer" if self.use_peer else "nccl"}
            self.bucket.zero_()
        self.fields = (("means3D", 3), ("shs", 48), ("opacities", 1), ("scales", 3), ("rotations", 4))
        views, off = {}, 0
        for k, w in self.fields:
            views[k] = self.bucket[off * P:(off + w) * P].view(params[k].shape)
            off += w
        self.views = views
        if mode in ("sinks", "dropin"):
            from diff_gaussian_rasterization import GaussianRasterizationSettings, GaussianRasterizer, GradSinks
            self.RS, self.R = GaussianRasterizationSettings, GaussianRasterizer
            from lgdwt_b200 import fused_l1_loss
            self.l1 = fused_l1_loss
            self.sinks = GradSinks(views["means3D"], views["shs"], views["opacities"], views["scales"],
                                   views["rotations"]) if mode == "sinks" else None
        elif mode == "stock":
            if REF_SITE not in sys.path:
                sys.path.insert(0, REF_SITE)
            from dgr_3dgs import GaussianRasterizationSettings, GaussianRasterizer
            self.RS, self.R = GaussianRasterizationSettings, GaussianRasterizer
        else:
            from oracle import ref_cuda
            self.ref = ref_cuda

    def render(self, cam, first):
        p = self.p
        if self.mode == "mirror":
            return self.ref.RefRasterize.apply(p["means3D"], self.means2D, p["shs"], p["opacities"], p["scales"],
                                               p["rotations"], cam, self.bg, SH_DEGREE)
        rs = self.RS(cam["H"], cam["W"], cam["tanfovx"], cam["tanfovy"], self.bg, 1.0, cam["viewmatrix"],
                     cam["projmatrix"], SH_DEGREE, cam["campos"], False, False, False)
        kw = dict(means3D=p["means3D"], means2D=self.means2D, shs=p["shs"], opacities=p["opacities"],
                  scales=p["scales"], rotations=p["rotations"])
        if self.mode == "sinks":
            self.sinks.accumulate = not first  # the first view of a step overwrites the bucket: no zero-fill pass
            kw["grad_sinks"] = self.sinks
        return self.R(rs)(**kw)

    def begin_step(self):
        self.means2D.grad = None
        if self.mode != "sinks":  # autograd sums the views into .grad
            for t in self.p.values():
                t.grad = None

    def exchange(self):
        """the path's only collective: sum the 236 B/Gaussian gradient bucket over ranks"""
        if self.world <= 1:
            return
        if self.mode != "sinks":
            for k, _ in self.fields:
                self.views[k].copy_(self.p[k].grad)
        if self.use_peer:
            self.peer.allreduce(1.0)
        else:
            dist.all_reduce(self.bucket)

    def step_resident(self, cams):
        self.begin_step()
        if self.overlap:
            main = torch.cuda.current_stream(self.device)
            for s in self.view_streams:
                s.wait_stream(main)
            for v, cam in enumerate(cams):
                s = self.view_streams[v % 2]
                with torch.cuda.stream(s):
                    color, radii, invd = self.render(cam, v == 0)
                    if v > 0:  # bucket order: this view's backward after the previous view's
                        s.wait_stream(self.view_streams[(v - 1) % 2])
                    color.backward(self.dL)
            for s in self.view_streams:
                main.wait_stream(s)
        else:
            for v, cam in enumerate(cams):
                color, radii, invd = self.render(cam, v == 0)
                color.backward(self.dL)
        self.exchange()

    def step_e2e(self, host_cams, host_gts):
        """every view's inputs come from pinned host memory; the step's loss goes back to the host (read one step
        late, see below).  The camera (needed first) is copied on the compute stream; the 7.7 MB ground-truth image is
        copied on a side stream while the rasterizer runs and joined right before the loss (same harness for both
        implementations)."""
        self.begin_step()
        main = torch.cuda.current_stream(self.device)
        total = torch.zeros((), device=self.device)
        if self.overlap:
            for s in self.view_streams:
                s.wait_stream(main)
        for v, (host_cam, host_gt) in enumerate(zip(host_cams, host_gts)):
            s = self.view_streams[v % 2] if self.overlap else main
            with torch.cuda.stream(s):
                cam = dict(host_cam)
                pk = host_cam["packed"].to(self.device, non_blocking=True)  # one 140-byte upload per view
                cam["viewmatrix"], cam["projmatrix"], cam["campos"] = pk[:16].view(4, 4), pk[16:32].view(4, 4), pk[32:35]
                gt_buf = self.dev_gt[self.e2e_steps % 2][v]
                with torch.cuda.stream(self.copy_stream):
                    # double-buffered: the readers of this buffer belong to the step before the previous one, whose loss
                    # has been read back (so it has completed) before this step was started
                    gt_buf.copy_(host_gt, non_blocking=True)
                    self.copy_done[v].record(self.copy_stream)
                color, radii, invd = self.render(cam, v == 0)
                s.wait_event(self.copy_done[v])
                if self.mode in ("sinks", "dropin"):   # this repo's public loss op; the reference arm keeps the stock torch expression
                    loss = self.l1(color, gt_buf)
                else:
                    loss = (color - gt_buf).abs().mean()      # l1_loss, LG/utils/loss_utils.py:40-41
                if self.overlap and v > 0:  # bucket (and `total`) order: after the previous view's backward
                    s.wait_stream(self.view_streams[(v - 1) % 2])
                loss.backward()
                total += loss.detach()
        if self.overlap:
            for s in self.view_streams:
                main.wait_stream(s)
        self.exchange()
        # device -> host read of the step's loss: an asynchronous copy into pinned memory, waited for only after the NEXT
        # step has been queued (a trainer logs the loss one step late rather than draining the GPU every step; the
        # reference's train.py blocks on .item() every iteration).  Every step's loss is read inside the timed region.
        slot = self.e2e_steps % 2
        self.e2e_steps += 1
        self.host_loss[slot].copy_(total, non_blocking=True)
        self.loss_ready[slot].record(main)
        if self.e2e_steps >= 2:
            self.loss_ready[1 - slot].synchronize()
            self.last_loss = float(self.host_loss[1 - slot])
        return self.last_loss


STEP_MS = {}   # label -> per-step device times of the last timed_loop with that label (diagnostic: median vs mean)


def timed_loop(fn, steps, world, device, label=None):
    """K steps between two CUDA events, barrier + synchronize on both sides, max over ranks.  The Python garbage
    collector is parked for the duration (a generation-2 collection in the middle of an 80 ms region costs milliseconds;
    a trainer does the same with gc.freeze()); one extra event per step records the per-step times for `label`."""
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(device)
    gc.collect()
    gc.disable()
    try:
        marks = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)] if label else None
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for i in range(steps):
            if marks:
                marks[i].record()
            fn(i)
        if marks:
            marks[steps].record()
        b.record()
        torch.cuda.synchronize(device)
    finally:
        gc.enable()
    if world > 1:
        dist.barrier()
    if marks:
        STEP_MS[label] = [marks[i].elapsed_time(marks[i + 1]) for i in range(steps)]
    ms = torch.tensor([a.elapsed_time(b)], device=device)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return float(ms.item())


def exchange_section(stepper, device, world):
    """The exchange step of one data-parallel training step on the benchmark's bucket (236 B per Gaussian), timed
    alone: NCCL all-reduce vs the peer-memory all-reduce the timed step uses, and — with the optimizer — NCCL
    all-reduce + the fused Adam pass on every rank vs ONE fused reduce-scatter + Adam + all-gather kernel."""
    from lgdwt_b200 import dp
    out = {"step_uses": ("peer all-reduce (csrc/peer.cu, %s)" % stepper.peer.backend) if stepper.use_peer
           else "nccl all-reduce",
           "picked_at_setup": stepper.exchange_pick,
           "bucket_bytes": int(stepper.bucket.numel() * 4)}
    if stepper.peer is None:
        out["peer_unavailable"] = stepper.peer_unavailable

    def t(fn, reps=10):
        for _ in range(3):
            fn()
        return timed_loop(lambda i: fn(), reps, world, device) / reps

    scratch = torch.zeros_like(stepper.bucket)
    out["nccl_allreduce_ms"] = round(t(lambda: dist.all_reduce(scratch)), 4)
    del scratch
    if stepper.peer is not None:
        ex = stepper.peer
        out["peer_allreduce_ms"] = round(t(lambda: ex.allreduce(1.0)), 4)
        g = dp.FlatGaussians(stepper.P, device)
        cfg = dp.AdamConfig()
        moments = (g.exp_avg, g.exp_avg_sq)
        g.data, g.grad = ex.param, ex.grad   # time on the shared block itself

        def nccl_adam():
            dist.all_reduce(g.grad)
            g.adam_step(cfg)

        def fused():
            g.step_count += 1
            ex.reduce_adam(moments[0], moments[1], g.adam_segments(cfg), cfg, g.step_count, 1.0)

        out["nccl_allreduce_plus_adam_ms"] = round(t(nccl_adam), 4)
        out["peer_fused_reduce_adam_gather_ms"] = round(t(fused), 4)
        try:
            ex.check()
        except RuntimeError as e:   # a flag barrier timed out somewhere in this run: say so instead of dying
            out["peer_barrier_timeout"] = str(e)
        # correctness of the exchange the timed step uses (the 1-GPU test box cannot run tests/peer_worker.py): the
        # peer-memory all-reduce of rank-dependent data against NCCL's, and bit-identity of the result across ranks
        rank = dist.get_rank()
        gen = torch.Generator(device=device).manual_seed(100 + rank)
        data = torch.randn(stepper.bucket.numel(), device=device, generator=gen)
        expect = data.clone()
        dist.all_reduce(expect)
        ex.grad.copy_(data)
        ex.allreduce(1.0)
        got = ex.grad.clone()
        err = (got - expect).abs().max()
        dist.all_reduce(err, op=dist.ReduceOp.MAX)
        lo, hi = got.clone(), got.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        out["max_abs_err_vs_nccl"] = float(err)
        out["max_abs_value"] = float(expect.abs().max())
        out["replicas_bit_identical"] = bool(torch.equal(lo.view(torch.int32), hi.view(torch.int32)))
        del data, expect, got, lo, hi
        n = stepper.bucket.numel() * 4
        out["nvlink_frac_of_900GBps_each_way"] = round(n * (world - 1) / world / (out["peer_allreduce_ms"] * 1e-3) / 900e9, 3)
        out["nvlink_bytes_per_rank_each_way"] = int(n * (world - 1) / world)
        out["peer_fused_GBps_each_way"] = round(n * (world - 1) / world / (out["peer_fused_reduce_adam_gather_ms"] * 1e-3) / 1e9, 1)
    return out


def cpu_baseline(sc, cam):
    """CPU oracle (port of the reference algorithm) on ONE view of the same workload, all host threads for the
    parallel stages.  Bounded sample: forward + backward of a single view."""
    from oracle import oracle
    kw = oracle.scene_kwargs(sc, cam, np.zeros(3, np.float32))
    rng = np.random.default_rng(0)
    dL = rng.standard_normal((3, HEIGHT, WIDTH)).astype(np.float32)
    t0 = time.perf_counter()
    f = oracle.rasterize_forward(**kw)
    t1 = time.perf_counter()
    oracle.rasterize_backward(f, dL_dpix=dL, **kw)
    t2 = time.perf_counter()
    return {"value": 1.0 / (t2 - t0), "unit": "views/s", "cores": oracle.num_threads(), "kind": "port",
            "sample": "1 view fwd+bwd of the same 1M-Gaussian 800x800 scene (fwd %.2f s, bwd %.2f s; OpenMP over "
                      "Gaussians / tiles, the radix sort is single-threaded)" % (t1 - t0, t2 - t1)}


def image_loss_section(device, hbm_peak):
    """BASELINE config 1 (LGDWT-GS Haar DWT loss on a 3x800x800 render/GT pair) plus the photometric terms: the
    fused kernels forward+backward on the GPU against their HBM roofline, and the CPU port of the reference's
    PyTorch op chain on this box's cores as a reported baseline."""
    from lgdwt_b200 import fused_image_loss
    from oracle import dwt_oracle, photometric_oracle
    pred_np, gt_np = scenes.dwt_pair(3, HEIGHT, WIDTH, seed=0)
    pred = torch.from_numpy(pred_np).to(device).requires_grad_(True)
    gt = torch.from_numpy(gt_np).to(device)
    rm = torch.ones((), device=device)

    def gpu_step():
        pred.grad = None
        loss, _ = fused_image_loss(pred, gt, rm)   # L1 + SSIM + DWT + patch + running-mean scale: one autograd node
        loss.backward()

    for _ in range(5):
        gpu_step()
    n = 20
    from lgdwt_b200 import _lib
    c0 = _lib.lib.lg_launch_count()
    gpu_step()
    loss_launches = int(_lib.lib.lg_launch_count() - c0)
    ms = timed_loop(lambda i: gpu_step(), n, 1, device) / n

    def device_ms(step):
        """device time of `step` with the host running ahead (as inside a training iteration, where the rasterizer's
        kernels keep the GPU busy while the host queues the loss): n steps queued behind a ~10 ms spin kernel"""
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(device)
        torch.cuda._sleep(int(2.0e7))
        a.record()
        for _ in range(n):
            step()
        b.record()
        torch.cuda.synchronize(device)
        return a.elapsed_time(b) / n

    ms_dev = device_ms(gpu_step)
    # the L1 + SSIM kernels alone, for comparison with the reference's fused-ssim CUDA kernel (SSIM/ssim.cu, timed by the
    # reference arm as `fused_ssim_fwd_bwd_ms` on the same tensors)
    from lgdwt_b200 import fused_photometric_loss

    def ph_step():
        pred.grad = None
        l1, ss = fused_photometric_loss(pred, gt)
        (0.8 * l1 + 0.2 * (1.0 - ss)).backward()

    for _ in range(5):
        ph_step()
    ph_ms = timed_loop(lambda i: ph_step(), n, 1, device) / n
    chw = 3 * HEIGHT * WIDTH
    alg = (8 + 12) * chw + (8 + 12 + 20 + 4) * chw  # dwt fwd+bwd, photometric fwd (+maps) + bwd, SURVEY.md §8(d)
    cpu_pred = torch.from_numpy(pred_np).requires_grad_(True)
    cpu_gt = torch.from_numpy(gt_np)
    t0 = time.perf_counter()
    reps = 3
    for _ in range(reps):
        cpu_pred.grad = None
        d, p, _, _ = dwt_oracle.lgdwt_losses(cpu_pred, cpu_gt)
        l1, ss = photometric_oracle.photometric_terms(cpu_pred, cpu_gt)
        (0.8 * l1 + 0.2 * (1.0 - ss) + d + 0.1 * p).backward()
    cpu_ms = (time.perf_counter() - t0) / reps * 1e3
    return {"workload": "config 1: L1 + SSIM + 2-level DWT + patch-ELF loss, fwd+bwd, 3x%dx%d pair" % (HEIGHT, WIDTH),
            "gpu_fused_ms": round(ms, 4), "gpu_fused_device_ms": round(ms_dev, 4), "gpu_launches": loss_launches,
            "bound": "hbm", "algorithmic_bytes": alg,
            "achieved_GBps": round(alg / (ms_dev * 1e-3) / 1e9, 1), "peak_GBps": hbm_peak,
            "frac": round(alg / (ms_dev * 1e-3) / 1e9 / hbm_peak, 4),
            "timing": "gpu_fused_ms: isolated loop (the host's ~0.25 ms of Python per fwd+bwd paces it); gpu_fused_device_ms: "
                      "the same steps queued behind a 10 ms spin kernel, i.e. device time with the host running ahead as in a "
                      "training iteration; achieved / frac use the device time",
            "note": "one autograd node (lgdwt_b200.fused_image_loss), one library call per direction: 6 kernels forward, 2 backward (the wavelet gradient is added in place); host-bound at this size",
            "photometric_fwd_bwd_ms": round(ph_ms, 4), "photometric_fwd_bwd_device_ms": round(device_ms(ph_step), 4),
            "cpu_port_ms": round(cpu_ms, 2), "cpu_threads": torch.get_num_threads(), "cpu_kind": "port"}


def reference_fused_ssim_section(device):
    """the reference's own fused SSIM CUDA kernel (submodules/fused-ssim, built by its setup.py into baseline/_ref/site)
    forward + backward on the config-1 tensors, with the L1 term as LG/train.py computes it (torch) — the counterpart of
    `image_loss.photometric_fwd_bwd_ms` of this repo's arm"""
    if not os.path.isdir(os.path.join(REF_SITE, "fused_ssim")):
        return {"unavailable": "baseline/_ref/site/fused_ssim not installed"}
    if REF_SITE not in sys.path:
        sys.path.insert(0, REF_SITE)
    from fused_ssim import fused_ssim
    pred_np, gt_np = scenes.dwt_pair(3, HEIGHT, WIDTH, seed=0)
    pred = torch.from_numpy(pred_np).to(device).requires_grad_(True)
    gt = torch.from_numpy(gt_np).to(device)

    def step():
        pred.grad = None
        l1 = torch.abs(pred - gt).mean()                                   # l1_loss, LG/utils/loss_utils.py:40-41
        ss = fused_ssim(pred.unsqueeze(0), gt.unsqueeze(0))                # LG/train.py:182-183
        (0.8 * l1 + 0.2 * (1.0 - ss)).backward()

    for _ in range(5):
        step()
    n = 20
    ms = timed_loop(lambda i: step(), n, 1, device) / n
    # device time with the host running ahead (same protocol as image_loss.photometric_fwd_bwd_device_ms of this repo's arm)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(device)
    torch.cuda._sleep(int(2.0e7))
    a.record()
    for _ in range(n):
        step()
    b.record()
    torch.cuda.synchronize(device)
    return {"workload": "L1 (torch) + fused_ssim (reference CUDA kernel) forward+backward, 3x%dx%d pair" % (HEIGHT, WIDTH),
            "fused_ssim_fwd_bwd_ms": round(ms, 4), "fused_ssim_fwd_bwd_device_ms": round(a.elapsed_time(b) / n, 4)}


def train_iteration_section(impl, device, iters=40, warm=10):
    """BASELINE config 3: one LLFF-style training iteration (LG/train.py:105-288) at 1008x756 on the 500k-Gaussian
    slab scene — activations, render, L1 + SSIM + global/patch DWT loss, backward, densification statistics, Adam —
    plus one densify_and_prune call.  `ours`: raw parameters in the flat buffer, every stage a fused kernel of this
    repo (dp.RunningMeanLoss keeps the running-mean DWT scale on the device).  `reference`: the stock op chain (six
    nn.Parameters, torch activations, the stock reference rasterizer extension, the PyTorch loss chain with its
    host-side running-mean ratio, torch.optim.Adam) — nothing of this repo's package.  CUDA-event breakdown."""
    import math
    Wd, Hd, V = 1008, 756, 3
    sc = scenes.slab_scene(500_000, seed=2)
    fovy = 2 * math.atan(math.tan(0.525) * Hd / Wd)
    cams = [cam_dict(scenes.look_at_camera(Wd, Hd, 1.05, fovy, (0.25 * (k - 1), 0.0, 0.0), target=(0.0, 0.0, 5.0)), device)
            for k in range(V)]
    gen = torch.Generator(device=device).manual_seed(11)
    gts = [torch.rand((3, Hd, Wd), device=device, generator=gen) for _ in range(V)]
    bg = torch.zeros(3, device=device)
    phases = ("render", "loss", "backward", "stats", "adam")
    marks = []

    def ev():
        e = torch.cuda.Event(enable_timing=True)
        e.record()
        return e

    if impl == "ours":
        from lgdwt_b200 import densify, dp
        g = dp.FlatGaussians.from_scene(sc, device)
        stats = densify.DensifyStats(g.P, device)
        loss_fn = dp.RunningMeanLoss(device)
        cfg = dp.AdamConfig()

        def iteration(i, timed):
            cam, gt = cams[i % V], gts[i % V]
            t = [ev()]
            image, radii, vsp = dp.fused_render(g, cam, bg, accumulate=False)
            t.append(ev())
            loss = loss_fn(image, gt)
            t.append(ev())
            loss.backward()
            t.append(ev())
            densify.add_densification_stats(stats, vsp.grad, radii)
            t.append(ev())
            g.adam_step(cfg)
            t.append(ev())
            if timed:
                marks.append(t)
    else:
        from oracle import dwt_oracle, photometric_oracle
        if not os.path.isdir(os.path.join(REF_SITE, "dgr_3dgs")):
            return {"unavailable": "baseline/_ref/site/dgr_3dgs not installed"}
        if REF_SITE not in sys.path:
            sys.path.insert(0, REF_SITE)
        from dgr_3dgs import GaussianRasterizationSettings, GaussianRasterizer
        tt = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(device)
        P = sc.means3D.shape[0]
        op = tt(sc.opacities).clamp(1e-6, 1 - 1e-6)
        raw = {"xyz": tt(sc.means3D), "f_dc": tt(sc.shs[:, :1, :]), "f_rest": tt(sc.shs[:, 1:, :]),
               "opacity": torch.log(op / (1 - op)), "scaling": torch.log(tt(sc.scales)), "rotation": tt(sc.rotations)}
        raw = {k: torch.nn.Parameter(v.clone()) for k, v in raw.items()}
        lrs = dict(xyz=0.00016, f_dc=0.0025, f_rest=0.0025 / 20.0, opacity=0.025, scaling=0.005, rotation=0.001)
        opt = torch.optim.Adam([{"params": [raw[k]], "lr": lrs[k], "name": k} for k in raw], lr=0.0, eps=1e-15)
        accum, denom, maxr = torch.zeros((P, 1), device=device), torch.zeros((P, 1), device=device), torch.zeros(P, device=device)
        state = {"rm": 1.0}

        def iteration(i, timed):
            cam, gt = cams[i % V], gts[i % V]
            t = [ev()]
            shs = torch.cat((raw["f_dc"], raw["f_rest"]), dim=1)                    # gaussian_model.py:121-124
            vsp = torch.zeros_like(raw["xyz"], requires_grad=True)                    # gaussian_renderer:26-30
            vsp.retain_grad()
            rs = GaussianRasterizationSettings(Hd, Wd, cam["tanfovx"], cam["tanfovy"], bg, 1.0, cam["viewmatrix"],
                                               cam["projmatrix"], 3, cam["campos"], False, False, False)
            image, radii, _ = GaussianRasterizer(rs)(means3D=raw["xyz"], means2D=vsp, shs=shs,
                                                     opacities=torch.sigmoid(raw["opacity"]),
                                                     scales=torch.exp(raw["scaling"]),
                                                     rotations=torch.nn.functional.normalize(raw["rotation"]))
            image = image.clamp(0, 1)
            t.append(ev())
            l1, ssim = photometric_oracle.photometric_terms(image, gt)               # train.py:128,182-188
            dwt, patch, _, _ = dwt_oracle.lgdwt_losses(image, gt)                    # train.py:131-180
            base = 0.8 * l1 + 0.2 * (1.0 - ssim)
            ratio = (base / (dwt + 1e-8)).item()                                      # train.py:190-196 (host sync)
            state["rm"] = 0.95 * state["rm"] + 0.05 * ratio
            loss = base + min(max(state["rm"], 0.1), 10.0) * dwt + 0.1 * patch
            t.append(ev())
            loss.backward()
            t.append(ev())
            vis = radii > 0                                                           # train.py:268-269
            maxr[vis] = torch.max(maxr[vis], radii[vis].float())
            accum[vis] += torch.norm(vsp.grad[vis, :2], dim=-1, keepdim=True)
            denom[vis] += 1
            t.append(ev())
            opt.step()
            opt.zero_grad(set_to_none=True)
            t.append(ev())
            if timed:
                marks.append(t)

    for i in range(warm):
        iteration(i, False)
    ms_total = timed_loop(lambda i: iteration(warm + i, True), iters, 1, device)
    br = {name: round(float(np.mean([m[k].elapsed_time(m[k + 1]) for m in marks])), 4) for k, name in enumerate(phases)}
    out = {"workload": "config 3: one training iteration, 500k-Gaussian slab scene, 1008x756, 3 cameras, L1 + SSIM + "
                       "global/patch DWT loss, densification statistics, Adam", "ms_per_iteration": round(ms_total / iters, 4),
           "iterations_per_s": round(iters / (ms_total / 1e3), 2), "breakdown_ms": br}
    if impl == "ours":
        a0, b0 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        P0 = g.P
        saved = (g.data.clone(), g.exp_avg.clone(), g.exp_avg_sq.clone(), stats.xyz_gradient_accum.clone(),
                 stats.denom.clone())
        for rep in range(3):   # two warm calls (lazy module load, allocator), the third one timed
            g.replace(P0, saved[0].clone(), saved[1].clone(), saved[2].clone())
            stats.__init__(P0, device)
            stats.xyz_gradient_accum.copy_(saved[3])
            stats.denom.copy_(saved[4])
            torch.cuda.synchronize(device)
            a0.record()
            counts = densify.densify_and_prune(g, stats, 0.0002, 0.005, 6.0, None, generator=gen)
            b0.record()
            torch.cuda.synchronize(device)
        out["densify_and_prune"] = {"ms": round(a0.elapsed_time(b0), 3), "P_before": P0, "P_after": g.P,
                                    "cloned": counts["cloned"], "split": counts["split_parents"],
                                    "algorithmic_bytes": 36 * P0 + 2 * 3 * 4 * g.floats * g.P,
                                    "note": "plan + one host read-back of the counts + gather of parameters and both Adam moments"}
    return out


def cfg5_section(device, rank, world, steps=3):
    """BASELINE config 5 — Mip-NeRF360 scale: 6 M Gaussians, 1920x1080, a 32-view global batch per step, data-parallel
    over the ranks (rank r renders views r, r+N, ...; STRONG scaling: the batch is fixed, so N = 1 renders all 32).
    The whole fused training step: activations, render, L1 + SSIM + DWT + patch loss, backward into the flat bucket,
    one exchange + Adam per step (fused peer-memory kernel at N > 1); two views in flight on two CUDA streams
    (ViewParallelTrainer(view_streams=2))."""
    import math
    from lgdwt_b200 import dp
    P, Wd, Hd, batch = 6_000_000, 1920, 1080, 32
    sc = scenes.trained_like_scene(P, seed=5, sigma_xyz=1.2, clip=3.0, log_scale_mean=math.log(0.006))
    fovy = 2 * math.atan(math.tan(0.5) * Hd / Wd)
    cams = [cam_dict(scenes.look_at_camera(Wd, Hd, 1.0, fovy, (5.0 * math.sin(2 * math.pi * k / batch), 0.0,
                                                                 -5.0 * math.cos(2 * math.pi * k / batch))), device)
            for k in range(batch)]
    gen = torch.Generator(device=device).manual_seed(3)
    mine = set(dp.views_of_rank(batch, rank, world))
    gts = [torch.rand((3, Hd, Wd), device=device, generator=gen) if k in mine else None for k in range(batch)]
    bg = torch.zeros(3, device=device)
    g = dp.FlatGaussians.from_scene(sc, device)
    del sc
    streams = env_int("LGDWT_CFG5_STREAMS", CFG5_VIEW_STREAMS)
    tr = dp.ViewParallelTrainer(g, loss_fn=dp.RunningMeanLoss(device), exchange="peer" if world > 1 else "nccl",
                                view_streams=streams)
    tr.step(cams, gts, bg)
    ms = timed_loop(lambda i: tr.step(cams, gts, bg), steps, world, device) / steps
    ok = tr.replicas_in_sync()
    out = {"workload": "config 5: 6M Gaussians, 1920x1080, 32-view global batch per step, fused training step "
                       "(render + L1/SSIM/DWT/patch loss + backward + exchange + Adam)", "scaling": "strong",
           "ms_per_step": round(ms, 3), "views_per_s": round(batch / (ms * 1e-3), 2), "views_per_rank": len(mine),
           "exchange": (tr.peer.backend if tr.peer is not None else (tr.peer_unavailable or "none (1 rank)")),
           "replicas_in_sync": bool(ok), "steps": steps, "view_streams": streams}
    del tr, g
    torch.cuda.empty_cache()
    return out


STAGE_TIMING_EVERY = 3   # rasterizer calls per timed one (co-prime with the 4 views of a step: all positions sampled)


def measure(stepper, cam_devs, host_cams, host_gts, K, W, world, rank, device, sampler=None, stage_timing=None,
            launch_counter=None):
    """(ms per resident step, ms per end-to-end step, kernels launched in the timed resident region) of one Stepper:
    W warm-up steps, then exactly K timed ones, for each of the two measurements"""
    V = VIEWS_PER_RANK
    pick = lambda i: [cam_devs[((i * world + rank) * V + v) % len(cam_devs)] for v in range(V)]
    for i in range(W):
        stepper.step_resident(pick(i))
    if stage_timing is not None:
        # CUDA events around the kernels of every third rasterizer call of the timed region (an event record between two
        # kernels serialises the stream for ~3 us; ten per call on every call cost 3 % of the step)
        stage_timing[0](min(K * V, 256), STAGE_TIMING_EVERY)
    if sampler is not None:
        sampler.mark()
    n0 = launch_counter() if launch_counter is not None else 0
    ms_res = timed_loop(lambda i: stepper.step_resident(pick(i)), K, world, device, stepper.mode + "/resident") / K
    launches = (launch_counter() - n0) if launch_counter is not None else None
    if stage_timing is not None:
        stage_timing[1]()   # read the ring and disarm it: the end-to-end loop below runs without stage events
    ring = lambda i: [((i * world + rank) * V + v) % len(host_cams) for v in range(V)]
    e2e_fn = lambda i: stepper.step_e2e([host_cams[j] for j in ring(i)], [host_gts[j] for j in ring(i)])
    for i in range(W):
        e2e_fn(i)
    ms_e2e = timed_loop(e2e_fn, K, world, device, stepper.mode + "/e2e") / K
    return ms_res, ms_e2e, launches


def host_inputs(cams):
    host_cams = []
    for c in cams:
        d = cam_dict(c, "cpu")
        d["packed"] = torch.cat([d["viewmatrix"].reshape(-1), d["projmatrix"].reshape(-1), d["campos"].reshape(-1)]).pin_memory()
        host_cams.append(d)
    rng = np.random.default_rng(7)
    host_gts = [torch.from_numpy(rng.random((3, HEIGHT, WIDTH)).astype(np.float32)).pin_memory() for _ in range(len(host_cams))]
    return host_cams, host_gts


def workload_text(V, world, exchange):
    return ("1M-Gaussian trained-like synthetic scene (seed 1), 800x800, SH degree 3, %d views per rank per step (global "
            "view batch %d), rasterize forward+backward with the gradients of the rank's views accumulated in one flat "
            "bucket" % (V, V * world) + (" + one all-reduce of the 236 B/Gaussian bucket per step (%s)" % exchange
                                         if world > 1 else ""))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-train-iteration", action="store_true")
    ap.add_argument("--no-cfg5", action="store_true")
    ap.add_argument("--no-sustained", action="store_true")
    args = ap.parse_args()
    rank, world, local = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)
    W = max(args.warmup, 3)
    K = args.steps
    if args.impl == "reference":
        if rank != 0:
            return 0  # rank 0 alone runs the reference arm
        return reference_arm(args, K, W, local)
    if not torch.cuda.is_available():
        print(json.dumps({"error": "no CUDA device: this benchmark has no CPU fallback", "impl": args.impl}))
        return 1
    device = torch.device("cuda", local)
    torch.cuda.set_device(device)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    sys.path.insert(0, PKG)
    from lgdwt_b200 import _lib   # raises when the CUDA library is missing: there is no fallback

    V = VIEWS_PER_RANK
    sc, cams, params = make_workload(device)
    cam_devs = [cam_dict(c, device) for c in cams]
    host_cams, host_gts = host_inputs(cams)
    stepper = Stepper("sinks", params, device, world)
    sampler = ClockSampler(local)
    sampler.start()
    rows = []

    def read_stage_rows():
        n_timed = (K * V + STAGE_TIMING_EVERY - 1) // STAGE_TIMING_EVERY
        rows.extend(_lib.read_stage_times(s) for s in range(min(n_timed, 256)))
        _lib.stage_timing(0)

    ms_step, ms_e2e, launches = measure(stepper, cam_devs, host_cams, host_gts, K, W, world, rank, device, sampler,
                                        (_lib.stage_timing, read_stage_rows), _lib.lib.lg_launch_count)
    clocks = sampler.stop()
    value = world * V / (ms_step / 1e3)

    # ---- per-stage times recorded during the timed resident region (rows)
    mean_ms = {k: float(np.mean([r[k] for r in rows if r[k] >= 0])) for k in _lib.STAGES}
    import diff_gaussian_rasterization as dgr
    seen = []   # work terms of the metric camera (step 0's view on rank 0)
    dgr.set_inspection_hook(seen.append)
    stepper.step_resident([cam_devs[0]])
    dgr.set_inspection_hook(None)
    lc = seen[-1]
    counts = torch.zeros(2, dtype=torch.int64, device=device)
    _lib.check(_lib.lib.lg_blend_work_count(lc["P"], lc["channels"], lc["W"], lc["H"], lc["binning_capacity"],
                                            lc["geom"].data_ptr(), lc["binning"].data_ptr(), lc["img"].data_ptr(),
                                            counts.data_ptr(), _lib.stream_ptr(device)))
    n_contrib = torch.zeros(WIDTH * HEIGHT, dtype=torch.int32, device=device)
    _lib.check(_lib.lib.lg_state_read(b"n_contrib", lc["P"], lc["channels"], lc["W"], lc["H"], lc["binning_capacity"],
                                      lc["geom"].data_ptr(), lc["binning"].data_ptr(), lc["img"].data_ptr(),
                                      n_contrib.data_ptr(), n_contrib.numel() * 4, _lib.stream_ptr(device)))
    torch.cuda.synchronize(device)
    n_eval, n_hit = int(counts[0]), int(counts[1])
    n_trav = int(n_contrib.long().sum())
    R, P = lc["num_rendered"], lc["P"]
    del seen, lc
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    hbm_src = "MEASURED_PEAKS.json hbm_gbs (measured)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    simt = _lib.simt_peaks(device)   # FFMA / MUFU / LDS / SHFL ceilings measured on THIS box, now (csrc/microbench.cu)
    fp32_peak = simt["ffma_tflops"] if simt["ffma_tflops"] > 1.0 else FP32_SIMT_NOMINAL_TFLOPS
    T = ((WIDTH + 15) // 16) * ((HEIGHT + 15) // 16)
    alg = {  # algorithmic work per launch, SURVEY.md §8(d)
        "preprocess": ("hbm", P * (44 + 12 * 16) + 8 * P + 67 * P),
        "binning": ("hbm", 20 * P + 12 * R + 24 * R + 8 * R + 8 * T),
        "blend_forward": ("fp32_simt", 15 * n_eval + 15 * n_hit),
        "blend_backward": ("fp32_simt", 15 * n_trav + 80 * n_hit),
        "preprocess_backward": ("hbm", 560 * P),
    }
    stages = {}
    for k in _lib.STAGES:
        bound, work = alg[k]
        ach = work / (mean_ms[k] * 1e-3) / (1e9 if bound == "hbm" else 1e12)
        peak = hbm_peak if bound == "hbm" else fp32_peak
        stages[k] = {"ms": round(mean_ms[k], 4), "bound": bound, "achieved": round(ach, 2),
                     "unit": "GB/s" if bound == "hbm" else "TFLOP/s", "peak": round(peak, 2),
                     "frac": round(ach / peak, 4)}
    dom = max(_lib.STAGES, key=lambda k: mean_ms[k])
    roofline = dict(stages[dom])
    traffic = None  # DRAM bytes per launch of the dominant kernel from the committed ncu --set full capture
    for prof in ("r2_ncu_traffic.json", "r1_ncu_traffic.json"):
        try:
            tr = json.load(open(os.path.join(ROOT, "profiles", prof)))["kernels"]
            for name, recs in tr.items():
                if name.startswith(dom + "_kernel"):
                    traffic = int(recs[0]["dram_read_bytes"] + recs[0]["dram_write_bytes"])
                    roofline["ncu_issue_slots_busy"] = round(recs[0].get("issue_active_pct", 0.0) / 100.0, 3)
                    roofline["traffic_source"] = "profiles/" + prof
            if traffic is not None:
                break
        except Exception:
            pass
    roofline.update({"kernel": dom, "traffic": traffic,
                     "timing": "CUDA events on the launching stream around this kernel in every %d-th rasterizer call of "
                               "the timed region (mean of %d launches)" % (STAGE_TIMING_EVERY, len(rows)), "share_of_step": round(mean_ms[dom] * V / ms_step, 3),
                     "peak_source": hbm_src if roofline["bound"] == "hbm" else
                     "FFMA rate measured on this box by csrc/microbench.cu (lg_simt_peaks) at the start of this run; "
                     "nominal 148 SM x 128 lanes x 2 x 1.965 GHz = %.1f TFLOP/s" % FP32_SIMT_NOMINAL_TFLOPS,
                     "work_terms": {"num_rendered": R, "n_eval_fwd": n_eval, "n_trav": n_trav, "n_hit": n_hit}})
    if dom.startswith("blend"):
        # the other ceilings this kernel runs against (per hit-warp 12 SHFL + per evaluated warp-entry 4 broadcast LDS.128
        # go through the MIO pipe: SHFL 1 / clk / SM, LDS.128 1 / 2.6 clk / SM)
        roofline["simt_peaks"] = {k: round(v, 1) for k, v in simt.items()}

    exchange = exchange_section(stepper, device, world) if world > 1 else None
    exch_name = ("peer-memory kernel over NVLink" if stepper.use_peer else "NCCL")
    del stepper

    h2d = V * (3 * HEIGHT * WIDTH * 4 + (16 + 16 + 3) * 4)
    e2e = {"value": round(world * V / (ms_e2e / 1e3), 3), "unit": "views/s", "ms_per_step": round(ms_e2e, 4),
           "ms_per_view": round(ms_e2e / V, 4), "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4}
    line = {
        "metric": "train views/sec (rasterize fwd+bwd per view @1M Gaussians 800x800 SH3)", "value": round(value, 3),
        "unit": "views/s", "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": round(ms_step, 4),
        "ms_per_view": round(ms_step / V, 4),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_text(V, world, exch_name),
                   "views_per_rank_per_step": V, "global_view_batch": V * world, "gaussians": P_GAUSSIANS, "width": WIDTH,
                   "height": HEIGHT, "sh_degree": SH_DEGREE, "cameras": N_CAMERAS, "parallelism": "view-parallel dp%d" % world,
                   "gradient_path": "GradSinks: the backward kernel accumulates the views of a step in one flat bucket "
                                    "(extension of the operator; `dropin` below is the unchanged-caller path)",
                   "l2_policy": "inputs larger than L2 (236 MB of Gaussian parameters + 72 MB of binning state per view)"},
        "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "stages": stages,
    }
    # diagnostic only: per-step device times of the two timed regions (a mean far above the median = a host hiccup)
    line["step_ms"] = {k: {"median": round(float(np.median(v)), 4), "min": round(min(v), 4), "max": round(max(v), 4)}
                       for k, v in STEP_MS.items()}
    if exchange is not None:
        line["exchange"] = exchange
    if world == 1:
        # the same metric through the operator exactly as the unchanged LG/train.py uses it: no grad_sinks, fresh
        # gradient tensors per view, autograd sums the views of a step
        d_step, d_e2e, _ = measure(Stepper("dropin", params, device, 1), cam_devs, host_cams, host_gts, K, W, 1, 0, device)
        line["dropin"] = {"value": round(V / (d_step / 1e3), 3), "unit": "views/s", "ms_per_view": round(d_step / V, 4),
                          "e2e_value": round(V / (d_e2e / 1e3), 3), "e2e_ms_per_view": round(d_e2e / V, 4),
                          "note": "plain GaussianRasterizer(...) call as in LG/gaussian_renderer/__init__.py:98-110; "
                                  "gradients returned as new tensors and accumulated by autograd"}
        # two views in flight: consecutive views of a step on alternating CUDA streams (Stepper.overlap).  Reported next to
        # the headline, which stays single-stream so that its per-kernel event times are kernel durations
        o_step, o_e2e, _ = measure(Stepper("sinks", params, device, 1, view_streams=2), cam_devs, host_cams, host_gts, K, W,
                                   1, 0, device)
        line["view_overlap"] = {"value": round(V / (o_step / 1e3), 3), "unit": "views/s", "ms_per_view": round(o_step / V, 4),
                                "e2e_value": round(V / (o_e2e / 1e3), 3), "e2e_ms_per_view": round(o_e2e / V, 4),
                                "note": "the headline step with two CUDA streams: the latency-bound head of view v+1 "
                                        "(preprocess, tile scan, scatter, tile sort) runs under the issue-bound blend "
                                        "backward of view v; the bucket is still written in view order (bit-identical "
                                        "step); also ViewParallelTrainer(view_streams=2)"}
    if world == 1 and not args.no_sustained:
        # the headline's timed region lasts ~0.1 s at boost clocks; the same resident step back to back for ~3 s, with
        # the clocks sampled during it, says what holds under sustained load
        st = Stepper("sinks", params, device, 1)
        pick = lambda i: [cam_devs[(i * V + v) % len(cam_devs)] for v in range(V)]
        n_sus = max(K, int(3000.0 / ms_step))
        for i in range(3):
            st.step_resident(pick(i))
        sus_sampler = ClockSampler(local)
        sus_sampler.start()
        time.sleep(0.3)
        sus_sampler.mark()
        ms_sus = timed_loop(lambda i: st.step_resident(pick(i)), n_sus, 1, device) / n_sus
        line["sustained"] = {"value": round(V / (ms_sus / 1e3), 3), "unit": "views/s", "ms_per_view": round(ms_sus / V, 4),
                             "steps": n_sus, "seconds": round(ms_sus * n_sus / 1e3, 2), "clocks": sus_sampler.stop()}
        del st
    del params
    torch.cuda.empty_cache()
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(sc, cams[0])
        line["image_loss"] = image_loss_section(device, hbm_peak)
    if rank == 0 and world == 1 and not args.no_train_iteration:
        line["train_iteration"] = train_iteration_section("ours", device)
    if not args.no_cfg5:
        try:
            line["cfg5"] = cfg5_section(device, rank, world)
        except Exception as e:   # never lose the headline to the extra section
            line["cfg5"] = {"error": "%s: %s" % (type(e).__name__, e)}
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


def reference_arm(args, K, W, local):
    """`--impl reference`: the UNMODIFIED reference on the same metric, config and harness.  The reference has no CPU
    implementation of this path, so its own implementation — the CUDA extension dgr_3dgs built by its own setup.py for
    sm_100a into baseline/_ref/site — is what is timed, through its own GaussianRasterizationSettings /
    GaussianRasterizer; this process never imports this repo's package or library.  Fallbacks: oracle/_ref (the
    reference's .cu files behind a ctypes mirror of its glue), then the CPU oracle port."""
    if not torch.cuda.is_available():
        return reference_cpu_port(args, K, W)
    mode = None
    if os.path.isdir(os.path.join(REF_SITE, "dgr_3dgs")):
        mode = "stock"
    else:
        from oracle import ref_cuda
        if ref_cuda.load_ref() is not None:
            mode = "mirror"
    if mode is None:
        return reference_cpu_port(args, K, W)
    device = torch.device("cuda", local)
    torch.cuda.set_device(device)
    V = VIEWS_PER_RANK
    sc, cams, params = make_workload(device)
    cam_devs = [cam_dict(c, device) for c in cams]
    host_cams, host_gts = host_inputs(cams)
    stepper = Stepper(mode, params, device, 1)
    sampler = ClockSampler(local)
    if not os.environ.get("BENCH_NO_SAMPLER"):   # debugging aid only: the driver's runs always sample
        sampler.start()
    ms_step, ms_e2e, _ = measure(stepper, cam_devs, host_cams, host_gts, K, W, 1, 0, device, sampler)
    clocks = sampler.stop()
    del stepper, params
    torch.cuda.empty_cache()
    value = V / (ms_step / 1e3)
    h2d = V * (3 * HEIGHT * WIDTH * 4 + (16 + 16 + 3) * 4)
    kind = ("the reference's stock CUDA extension dgr_3dgs (diff-gaussian-rasterization built by its own setup.py for "
            "sm_100a, baseline/_ref/site) through its own Python API" if mode == "stock" else
            "reference CUDA rasterizer compiled for sm_100a from its own sources (oracle/_ref/libref_dgr.so) behind a "
            "line-for-line mirror of its torch glue")
    line = {
        "impl": "reference", "metric": "train views/sec (rasterize fwd+bwd per view @1M Gaussians 800x800 SH3)",
        "value": round(value, 3), "unit": "views/s", "n_gpus": 1, "steps": K, "warmup": W, "ms_per_step": round(ms_step, 4),
        "ms_per_view": round(ms_step / V, 4), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_text(V, 1, ""), "views_per_rank_per_step": V, "global_view_batch": V,
                   "gaussians": P_GAUSSIANS, "width": WIDTH, "height": HEIGHT, "sh_degree": SH_DEGREE, "cameras": N_CAMERAS,
                   "parallelism": "single GPU (the reference has no multi-GPU path)",
                   "gradient_path": "autograd accumulation of the views of a step (the reference's only path)",
                   "l2_policy": "inputs larger than L2 (236 MB of Gaussian parameters per view)"},
        "clocks": clocks, "gpu_launches": None, "reference_kind": kind,
        "e2e": {"value": round(V / (ms_e2e / 1e3), 3), "unit": "views/s", "ms_per_step": round(ms_e2e, 4),
                "ms_per_view": round(ms_e2e / V, 4), "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4},
        "cpu_baseline": {"value": round(value, 3), "unit": "views/s", "cores": 0, "kind": "reference",
                         "sample": "the reference has no CPU implementation of this path; this arm times its own CUDA "
                                   "implementation on the same GPU (cores = 0 host threads)"},
    }
    if not args.no_train_iteration:
        line["train_iteration"] = train_iteration_section("reference", device)
        line["image_loss"] = reference_fused_ssim_section(device)
    print(json.dumps(line))
    return 0


def reference_cpu_port(args, K, W):
    """last resort when neither baseline/_ref nor oracle/_ref was prebuilt: the CPU oracle port, bounded sample"""
    sc = scenes.trained_like_scene(P_GAUSSIANS, seed=1)
    cb = cpu_baseline(sc, scenes.metric_camera(WIDTH, HEIGHT))
    line = {"impl": "reference", "metric": "train views/sec (rasterize fwd+bwd per view @1M Gaussians 800x800 SH3)",
            "value": round(cb["value"], 5), "unit": "views/s", "n_gpus": 1, "steps": 1, "warmup": 0,
            "ms_per_step": round(1e3 / cb["value"], 2), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "cpu_baseline": cb,
            "config": {"workload": "CPU oracle port, one view of the 1M-Gaussian 800x800 SH3 scene"},
            "e2e": {"value": round(cb["value"], 5), "unit": "views/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))
    return 0


if __name__ == "__main__":
    sys.exit(main())
