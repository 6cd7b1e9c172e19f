"""CPU tier: the C-ABI library loads and exports every symbol include/*.h declares (no compute calls), host-side
validation of the operator surface, view sharding / gradient bucket / Adam logic under gloo with world_size 2."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "lgdwt_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(lg_[a-z0-9_]+)\s*\(", hdr)) - {"lg_alloc_fn"}
    assert len(declared) >= 18, declared
    lib = ctypes.CDLL(os.path.join(ROOT, "sparse-view-3dgs-pack_b200", "lib", "liblgdwt_b200.so"))
    missing = [s for s in sorted(declared) if not hasattr(lib, s)]
    assert not missing, missing
    lib.lg_abi_version.restype = ctypes.c_int
    assert lib.lg_abi_version() == 1
    lib.lg_geometry_state_bytes.restype = ctypes.c_size_t
    lib.lg_binning_state_bytes.restype = ctypes.c_size_t
    assert lib.lg_geometry_state_bytes(1000, 3) > 1000 * 95
    assert lib.lg_binning_state_bytes(0, 800, 800) < lib.lg_binning_state_bytes(100000, 800, 800)


def test_invalid_arguments_are_rejected_without_a_gpu():
    from lgdwt_b200 import _lib
    n = ctypes.c_int(0)
    cb = _lib.ALLOC_FN(lambda ctx, nbytes: 0)
    rc = _lib.lib.lg_rasterize_forward(cb, None, cb, None, cb, None, -1, 3, 16, 3, None, 8, 8, None, None, None, None,
                                       None, 1.0, None, None, None, None, None, 1.0, 1.0, 0, None, None, 0, None, 0, None,
                                       ctypes.byref(n))
    assert rc == _lib.LG_ERR_INVALID_ARGUMENT and b"invalid" in _lib.lib.lg_last_error()
    with pytest.raises(_lib.LgdwtError):
        _lib.check(rc)
    assert _lib.lib.lg_knn_mean_dist2(0, None, None, None, 0, None) == _lib.LG_OK       # P = 0 short-circuits
    assert _lib.lib.lg_mark_visible(-3, None, None, None, None, None) == _lib.LG_ERR_INVALID_ARGUMENT


def test_operator_surface_validation_and_no_cpu_fallback():
    from diff_gaussian_rasterization import GaussianRasterizationSettings, GaussianRasterizer
    import diff_gaussian_rasterization as dgr
    assert GaussianRasterizationSettings._fields == (
        "image_height", "image_width", "tanfovx", "tanfovy", "bg", "scale_modifier", "viewmatrix", "projmatrix",
        "sh_degree", "campos", "prefiltered", "debug", "antialiasing")          # DGR/dgr_3dgs/__init__.py:143-156
    assert not hasattr(dgr, "SparseGaussianAdam")                               # must stay absent
    rs = GaussianRasterizationSettings(8, 8, 1.0, 1.0, torch.zeros(3), 1.0, torch.eye(4), torch.eye(4), 3,
                                       torch.zeros(3), False, False, False)
    r = GaussianRasterizer(rs)
    z = lambda *s: torch.zeros(s)
    with pytest.raises(Exception, match="excatly one"):
        r(z(2, 3), z(2, 3), z(2, 1), shs=z(2, 16, 3), colors_precomp=z(2, 3), scales=z(2, 3), rotations=z(2, 4))
    with pytest.raises(Exception, match="exactly one of either scale/rotation"):
        r(z(2, 3), z(2, 3), z(2, 1), shs=z(2, 16, 3), scales=z(2, 3), rotations=z(2, 4), cov3D_precomp=z(2, 6))
    with pytest.raises(RuntimeError, match="no CPU path"):
        r(z(2, 3), z(2, 3), z(2, 1), shs=z(2, 16, 3), scales=z(2, 3), rotations=z(2, 4))
    from simple_knn._C import distCUDA2
    with pytest.raises(RuntimeError, match="no CPU path"):
        distCUDA2(z(5, 3))
    from lgdwt_b200 import fused_dwt_loss
    with pytest.raises(RuntimeError, match="no CPU path"):
        fused_dwt_loss(z(3, 8, 8), z(3, 8, 8))


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "sparse-view-3dgs-pack_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
                assert "liboracle" not in src and "libref_dgr" not in src, f


def test_scene_generators_are_deterministic_and_sane():
    from lgdwt_b200 import scenes
    a, b = scenes.trained_like_scene(1000, seed=1), scenes.trained_like_scene(1000, seed=1)
    assert all(np.array_equal(getattr(a, k), getattr(b, k)) for k in ("means3D", "scales", "rotations", "opacities", "shs"))
    np.testing.assert_allclose(np.linalg.norm(a.rotations, axis=1), 1.0, atol=1e-6)
    cam = scenes.metric_camera()
    np.testing.assert_allclose(cam.campos, [0, 0, -4.03], atol=1e-6)
    p = np.array([0.0, 0.0, 0.0, 1.0], np.float32) @ cam.viewmatrix        # row-vector convention
    np.testing.assert_allclose(p[:3], [0, 0, 4.03], atol=1e-6)


_DP_WORKER = r'''
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, os.path.join(%(root)r, "sparse-view-3dgs-pack_b200"))
sys.path.insert(0, %(root)r)
sys.path.insert(0, os.path.join(%(root)r, "tests"))
from lgdwt_b200 import dp
import helpers
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo", rank=rank, world_size=world)
torch.manual_seed(0)
P = 64
def make():
    g = dp.FlatGaussians(P, torch.device("cpu"))
    gen = torch.Generator().manual_seed(1)
    g.data.copy_(torch.randn(g.data.shape, generator=gen) * 0.1)
    return g
def render(act, cam, bg):   # differentiable stand-in for the rasterizer: every parameter group takes part
    P = act["means3D"].shape[0]
    vsp = torch.zeros((P, 3), requires_grad=True)   # "screen-space points": their .grad feeds the densification stats
    s = (act["means3D"].sum(1) + act["shs"].sum((1, 2)) + act["opacities"].sum(1) + act["scales"].sum(1)
         + act["rotations"].sum(1) + (vsp * cam["k"]).sum(1))
    radii = ((torch.arange(P) + cam["v"]) %% 3 != 0).int() * 5   # every view sees a different two thirds
    return (s[:, None] * cam["w"][None, :]).sum(0).view(1, 1, -1), radii, vsp
loss = lambda img, gt: ((img - gt) ** 2).mean()
cams = [{"w": torch.linspace(0.1 * (v + 1), 1.0, 16), "k": 1e-3 * (v + 1), "v": v} for v in range(6)]
dcfg = dp.DensifyConfig(densify_from_iter=1, densify_until_iter=100, densification_interval=2, opacity_reset_interval=3,
                        densify_grad_threshold=1e-4, cameras_extent=4.0)
kw = dict(render_fn=render, loss_fn=loss, densify=dcfg, densify_fn=helpers.oracle_densify_fn,
          stats_fn=helpers.oracle_stats_fn, reset_opacity_fn=helpers.oracle_reset_opacity_fn, seed=5)
gts = [torch.full((1, 1, 16), 0.3 * v) for v in range(6)]
g = make()
tr = dp.ViewParallelTrainer(g, **kw)
assert dp.views_of_rank(6, rank, world) == list(range(rank, 6, world))
sizes = []
for _ in range(5):    # iterations 2 and 4 densify (statistics all-reduced first), iteration 3 resets the opacities
    tr.step(cams, gts, None)
    sizes.append(g.P)
    assert tr.replicas_in_sync()
assert sizes[1] != P and sizes[0] == P, sizes
if rank == 0:   # single-process accumulation over the same 6 views must give the same parameters
    dist_backup = dist.is_initialized
    g1 = make()
    t1 = dp.ViewParallelTrainer(g1, **kw)
    t1.distributed, t1.rank, t1.world = False, 0, 1
    for _ in range(5):
        t1.step(cams, gts, None)
    assert g1.P == g.P, (g1.P, g.P)
    err = (g1.data - g.data).abs().max().item()
    assert err < 1e-6, err
    print("DP_OK", err)
dist.barrier()
dist.destroy_process_group()
'''


def test_view_parallel_trainer_gloo_world2(tmp_path):
    script = tmp_path / "dp_worker.py"
    script.write_text(_DP_WORKER % {"root": ROOT})
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29631", WORLD_SIZE="2")
    procs = [subprocess.Popen([sys.executable, str(script)], env=dict(env, RANK=str(r)), stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=240)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    assert "DP_OK" in outs[0]


def test_flat_gaussians_adam_matches_torch_optim():
    from lgdwt_b200 import dp
    g = dp.FlatGaussians(10, torch.device("cpu"))
    g.data.copy_(torch.randn(g.data.shape, generator=torch.Generator().manual_seed(2)))
    cfg = dp.AdamConfig()
    ref_params = {n: g.field(n).clone().requires_grad_(True) for n in dp.GROUPS}
    lrs = dict(xyz=cfg.lr_xyz, f_dc=cfg.lr_f_dc, f_rest=cfg.lr_f_rest, opacity=cfg.lr_opacity, scaling=cfg.lr_scaling,
               rotation=cfg.lr_rotation)
    opt = torch.optim.Adam([{"params": [ref_params[n]], "lr": lrs[n]} for n in dp.GROUPS], lr=0.0, eps=1e-15)
    for it in range(4):
        grads = torch.randn(g.data.shape, generator=torch.Generator().manual_seed(10 + it))
        g.grad.copy_(grads)
        for n in dp.GROUPS:
            ref_params[n].grad = g.field(n, grads).clone()
        opt.step()
        g.adam_step(cfg)
    for n in dp.GROUPS:
        torch.testing.assert_close(g.field(n), ref_params[n].detach(), rtol=1e-5, atol=1e-7)


def test_flat_gaussians_layout_padding_and_group_views():
    """slabs start at (floats before) x stride with stride = P rounded up to 4; f_dc / f_rest are column ranges of the
    SH slab whose rows are the rasterizer's (M, 3) layout; Adam segments follow the slabs"""
    from lgdwt_b200 import dp
    g = dp.FlatGaussians(10, torch.device("cpu"))
    assert g.stride == 12 and g.floats == 59 and g.data.numel() == 59 * 12 == g.grad.numel()
    offs = {n: g._slices[n][0] for n, _ in g.fields}
    assert offs == {"xyz": 0, "shs": 3 * 12, "opacity": 51 * 12, "scaling": 52 * 12, "rotation": 55 * 12}
    assert all(o % 4 == 0 for o in offs.values())
    g.slab("shs").copy_(torch.arange(10 * 48, dtype=torch.float32).view(10, 48))
    shs = g.slab("shs").view(10, 16, 3)
    assert torch.equal(g.field("f_dc"), shs[:, 0, :]) and torch.equal(g.field("f_rest"), shs[:, 1:, :].reshape(10, 45))
    g.field("f_dc").add_(1000.0)           # group views write through to the flat buffer
    assert float(g.data[offs["shs"]]) == 1000.0
    assert not g.data[offs["shs"] + 10 * 48: offs["opacity"]].any()   # padding rows untouched
    ends, lr_a, lr_b, width, split = g.adam_segments(dp.AdamConfig())
    assert list(ends) == [3 * 12, 51 * 12, 52 * 12, 55 * 12, 59 * 12]
    assert list(width) == [1, 48, 1, 1, 1] and list(split) == [0, 3, 0, 0, 0]
    assert abs(lr_a[1] - 0.0025) < 1e-9 and abs(lr_b[1] - 0.0025 / 20) < 1e-9 and lr_a[0] == lr_b[0]
    for P in (0, 1, 4, 5):
        h = dp.FlatGaussians(P, torch.device("cpu"))
        assert h.stride == (P + 3) // 4 * 4 and h.slab("rotation").shape == (P, 4)
    g2 = dp.FlatGaussians(7, torch.device("cpu"), sh_degree=1)
    assert g2.floats == 11 + 12 and g2.M == 4 and g2.slab("shs").shape == (7, 12)


def test_density_control_schedule_follows_train_py():
    """which iterations densify / reset opacity / use the size threshold (LG/train.py:265-276), and that the Adam
    update of a densification iteration is dropped"""
    from lgdwt_b200 import dp
    g = dp.FlatGaussians(8, torch.device("cpu"))
    calls = []
    cfg = dp.DensifyConfig(densify_from_iter=4, densify_until_iter=13, densification_interval=3, opacity_reset_interval=6)

    def render(act, cam, bg):
        return (act["means3D"].sum() + act["opacities"].sum()).view(1, 1, 1), torch.ones(8, dtype=torch.int32)

    tr = dp.ViewParallelTrainer(g, render_fn=render, loss_fn=lambda img, gt: img.sum(), densify=cfg,
                                densify_fn=lambda g_, st, c, mss, gen: calls.append(("densify", tr.iteration, mss)),
                                stats_fn=lambda *a: None, reset_opacity_fn=lambda g_: calls.append(("reset", tr.iteration)))
    steps_before = []
    for _ in range(15):
        before = g.step_count
        tr.step([{}], [None], None)
        steps_before.append(g.step_count - before)
    assert [c for c in calls if c[0] == "densify"] == [("densify", 6, None), ("densify", 9, 20.0), ("densify", 12, 20.0)]
    assert [c for c in calls if c[0] == "reset"] == [("reset", 6), ("reset", 12)]
    # iterations 6, 9 and 12 rebuilt the set: no optimizer step there (the reference's new Parameters have no .grad)
    assert [i + 1 for i, s in enumerate(steps_before) if s == 0] == [6, 9, 12]


def test_white_background_adds_the_reset_at_densify_from_iter():
    """LG/train.py:275: `iteration % opacity_reset_interval == 0 or (white_background and iteration == densify_from_iter)`"""
    from lgdwt_b200 import dp
    for white, want in ((False, [6, 12]), (True, [4, 6, 12])):
        g = dp.FlatGaussians(8, torch.device("cpu"))
        calls = []
        cfg = dp.DensifyConfig(densify_from_iter=4, densify_until_iter=13, densification_interval=3,
                               opacity_reset_interval=6, white_background=white)
        render = lambda act, cam, bg: ((act["means3D"].sum() + act["opacities"].sum()).view(1, 1, 1),
                                       torch.ones(8, dtype=torch.int32))
        tr = dp.ViewParallelTrainer(g, render_fn=render, loss_fn=lambda img, gt: img.sum(), densify=cfg,
                                    densify_fn=lambda *a: None, stats_fn=lambda *a: None,
                                    reset_opacity_fn=lambda g_: calls.append(tr.iteration))
        for _ in range(15):
            tr.step([{}], [None], None)
        assert calls == want, (white, calls)


def test_train_schedule_matches_reference_lr_functions_and_sh_schedule():
    """TrainSchedule.position_lr / depth_l1_weight against a line-by-line restatement of get_expon_lr_func
    (LG/utils/general_utils.py:29-63) as GaussianModel.training_setup (gaussian_model.py:203-206) and LG/train.py:69
    call it; oneupSHdegree every `oneup_sh_every` iterations from degree 0 (LG/train.py:101-103); the position learning
    rate of iteration i is what the Adam step of iteration i uses (update_learning_rate is called first, :99)."""
    import math
    from lgdwt_b200 import dp

    def expon(lr_init, lr_final, lr_delay_steps=0, lr_delay_mult=1.0, max_steps=1000000):
        def helper(step):
            if step < 0 or (lr_init == 0.0 and lr_final == 0.0):
                return 0.0
            if lr_delay_steps > 0:
                delay_rate = lr_delay_mult + (1 - lr_delay_mult) * np.sin(0.5 * np.pi * np.clip(step / lr_delay_steps, 0, 1))
            else:
                delay_rate = 1.0
            t = np.clip(step / max_steps, 0, 1)
            return delay_rate * np.exp(np.log(lr_init) * (1 - t) + np.log(lr_final) * t)
        return helper

    s = dp.TrainSchedule(iterations=7000, position_lr_max_steps=3000, spatial_lr_scale=4.7)
    ref_xyz = expon(s.position_lr_init * 4.7, s.position_lr_final * 4.7, lr_delay_mult=s.position_lr_delay_mult,
                    max_steps=3000)
    ref_depth = expon(s.depth_l1_weight_init, s.depth_l1_weight_final, max_steps=7000)
    for it in (1, 2, 17, 1500, 2999, 3000, 3001, 6999, 7000, 9000):
        assert math.isclose(s.position_lr(it), float(ref_xyz(it)), rel_tol=1e-12), it
        assert math.isclose(s.depth_l1_weight(it), float(ref_depth(it)), rel_tol=1e-12), it

    g = dp.FlatGaussians(8, torch.device("cpu"))
    render = lambda act, cam, bg: ((act["means3D"].sum() + act["opacities"].sum()).view(1, 1, 1),
                                   torch.ones(8, dtype=torch.int32))
    sched = dp.TrainSchedule(oneup_sh_every=4, position_lr_max_steps=10, spatial_lr_scale=2.0)
    tr = dp.ViewParallelTrainer(g, render_fn=render, loss_fn=lambda img, gt: img.sum(), schedule=sched)
    assert g.active_sh_degree == 0 and g.sh_degree == 3
    degrees, lrs = [], []
    for _ in range(18):
        tr.step([{}], [None], None)
        degrees.append(g.active_sh_degree)
        lrs.append(tr.adam.lr_xyz)
    assert degrees == [0] * 3 + [1] * 4 + [2] * 4 + [3] * 7          # raised at iterations 4, 8, 12, capped at 3
    assert all(math.isclose(lr, sched.position_lr(i + 1), rel_tol=1e-12) for i, lr in enumerate(lrs))
    assert lrs[0] > lrs[5] > lrs[9] and math.isclose(lrs[9], lrs[17])   # decays until max_steps, then constant


def test_binning_capacity_hint_bookkeeping():
    """the operator's speculative binning capacity: no hint before the first call of a (device, P, W, H) shape,
    afterwards 1.25 x the largest entry count of the last eight calls"""
    import diff_gaussian_rasterization as dgr
    dgr._CAPACITY_HISTORY.clear()
    key = ("cuda:0", 1000, 64, 48)
    assert dgr._capacity_hint(*key) == 0
    for r in (100, 400, 300):
        dgr._record_num_rendered(*key, r)
    assert dgr._capacity_hint(*key) == int(1.25 * 400) + 1024
    for r in range(8):
        dgr._record_num_rendered(*key, 50)
    assert dgr._capacity_hint(*key) == int(1.25 * 50) + 1024          # the large value aged out
    assert dgr._capacity_hint("cuda:0", 2000, 64, 48) == 0            # another shape: no history
    dgr._CAPACITY_HISTORY.clear()


def test_one_call_image_loss_entry_points_reject_bad_arguments_without_a_gpu():
    """lg_image_loss_forward / _backward and the scaled gradient entry points validate before they launch anything"""
    from lgdwt_b200 import _lib
    w = (ctypes.c_float * 8)(*([1.0] * 8))
    rc = _lib.lib.lg_image_loss_forward(None, None, 3, 8, 8, w, 0, 0.5, 1.0, 1.0, None, 0.2, 0.1, 1, None, None, 0, None, 0,
                                        None, 0, 1, None)
    assert rc == _lib.LG_ERR_INVALID_ARGUMENT and b"lg_image_loss_forward" in _lib.lib.lg_last_error()
    rc = _lib.lib.lg_image_loss_backward(None, None, 3, 8, 8, w, 0, 1.0, 1.0, None, None, None, None, None, None)
    assert rc == _lib.LG_ERR_INVALID_ARGUMENT and b"lg_image_loss_backward" in _lib.lib.lg_last_error()
    assert _lib.lib.lg_photometric_loss_backward_scaled(None, None, 3, 8, 8, None, None, None, None, None,
                                                        None) == _lib.LG_ERR_INVALID_ARGUMENT
    assert _lib.lib.lg_dwt_loss_backward_scaled(None, None, 3, 8, 8, w, 0, 1.0, 1.0, None, None, None, None, None, None, 1,
                                                None) == _lib.LG_ERR_INVALID_ARGUMENT


def test_view_streams_argument_of_the_trainer():
    """view_streams is validated, and a trainer without CUDA buffers (or with a custom render_fn) stays single-stream"""
    import numpy as np
    import torch
    from lgdwt_b200 import dp, scenes
    g = dp.FlatGaussians.from_scene(scenes.trained_like_scene(64, seed=3), torch.device("cpu"))
    with pytest.raises(ValueError):
        dp.ViewParallelTrainer(g, render_fn=lambda act, cam, bg: None, view_streams=3)
    tr = dp.ViewParallelTrainer(g, render_fn=lambda act, cam, bg: None, view_streams=2)
    assert tr.view_streams is None
