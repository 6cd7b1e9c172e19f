"""GPU tier for the trainer-side kernels around the rasterizer: density control on the
flat parameter buffer against the reference class's golden vectors and the CPU oracle, fused activations and their
in-kernel chain rule against torch autograd, and the fused training step against the generic autograd step."""
import os

import numpy as np
import pytest
import torch

import golden_inputs
import helpers
from lgdwt_b200 import densify, dp, scenes
from oracle import densify_oracle

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
dev = "cuda"


def T(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(dev)


def _stats_from(d):
    st = densify.DensifyStats(d["denom"].shape[0], dev)
    st.xyz_gradient_accum.copy_(T(d["xyz_gradient_accum"]).reshape(-1))
    st.denom.copy_(T(d["denom"]).reshape(-1))
    st.max_radii2D.copy_(T(d["max_radii2D"]).reshape(-1))
    return st


@pytest.mark.parametrize("name", list(golden_inputs.DENSIFY_CASES))
def test_densify_and_prune_matches_reference_golden(name):
    """same rows in the same order as GaussianModel.densify_and_prune, values to 1e-6, moments bit-equal"""
    gold = np.load(os.path.join(GOLD, "densify_reference.npz"))
    d, cfg = golden_inputs.densify_case(name)
    g = helpers.flat_from_groups(dp, d, dev)
    st = _stats_from(d)
    counts = densify.densify_and_prune(g, st, cfg["max_grad"], cfg["min_opacity"], cfg["extent"],
                                       cfg["max_screen_size"], percent_dense=cfg["percent_dense"],
                                       eps=T(gold[name + "/eps"]))
    assert counts["P"] == g.P == gold[name + "/xyz"].shape[0]
    assert 2 * counts["split_parents"] == gold[name + "/eps"].shape[0]
    p, m, v = helpers.groups_from_flat(g), helpers.groups_from_flat(g, g.exp_avg), helpers.groups_from_flat(g, g.exp_avg_sq)
    for k in helpers.DENSIFY_GROUPS:
        np.testing.assert_allclose(p[k].numpy(), gold["%s/%s" % (name, k)], rtol=1e-6, atol=1e-6, err_msg=k)
        np.testing.assert_array_equal(m[k].numpy(), gold["%s/exp_avg/%s" % (name, k)], err_msg=k)
        np.testing.assert_array_equal(v[k].numpy(), gold["%s/exp_avg_sq/%s" % (name, k)], err_msg=k)
    assert st.denom.numel() == g.P and not st.denom.any() and not st.max_radii2D.any()
    assert g.grad.numel() == g.data.numel() and not g.grad.any()


def test_densify_matches_oracle_at_scale_and_edge_sizes():
    """sizes around the 1024-row block of the plan kernels, more than one scan chunk, P = 0 and P = 1"""
    for P, seed in ((0, 1), (1, 2), (1023, 3), (1025, 4), (300_000, 5), (1_100_000, 6)):
        rng = np.random.default_rng(seed)
        f32 = lambda a: np.ascontiguousarray(a, dtype=np.float32)
        d = {"xyz": f32(rng.normal(0, 1, (P, 3))), "f_dc": f32(rng.normal(0, 1, (P, 1, 3))),
             "f_rest": f32(rng.normal(0, .1, (P, 15, 3))), "opacity": f32(rng.normal(-2, 3, (P, 1))),
             "scaling": f32(rng.normal(np.log(0.04), 0.7, (P, 3))), "rotation": f32(rng.normal(0, 1, (P, 4)))}
        for k in list(d):
            d["exp_avg/" + k] = f32(rng.normal(0, 1e-3, d[k].shape))
            d["exp_avg_sq/" + k] = f32(rng.random(d[k].shape) * 1e-6)
        denom = rng.integers(0, 4, (P, 1)).astype(np.float32)
        accum = f32(rng.random((P, 1)) * 0.0012 * np.maximum(denom, 1))
        accum[denom == 0] = 0
        g = helpers.flat_from_groups(dp, d, dev)
        st = densify.DensifyStats(P, dev)
        st.xyz_gradient_accum.copy_(T(accum).reshape(-1))
        st.denom.copy_(T(denom).reshape(-1))
        t = lambda a: torch.from_numpy(a)
        pp = {k: t(d[k]) for k in densify_oracle.GROUPS}
        _, _, _, parents = densify_oracle.plan(pp, t(accum), t(denom), 0.0002, 0.005, 4.0, 20, 0.01)
        eps = torch.randn((2 * parents.numel(), 3), generator=torch.Generator().manual_seed(seed))
        op, om, ov, oc = densify_oracle.densify_and_prune(
            pp, {k: t(d["exp_avg/" + k]) for k in densify_oracle.GROUPS},
            {k: t(d["exp_avg_sq/" + k]) for k in densify_oracle.GROUPS}, t(accum), t(denom), eps, 0.0002, 0.005, 4.0,
            20, 0.01)
        counts = densify.densify_and_prune(g, st, 0.0002, 0.005, 4.0, 20, eps=eps.to(dev))
        assert counts == oc, (P, counts, oc)
        p, m = helpers.groups_from_flat(g), helpers.groups_from_flat(g, g.exp_avg)
        for k in helpers.DENSIFY_GROUPS:
            torch.testing.assert_close(p[k], op[k], rtol=2e-6, atol=2e-6, msg=lambda s: "%s P=%d %s" % (k, P, s))
            assert torch.equal(m[k], om[k]), (k, P)


def test_densify_is_deterministic_and_generator_driven():
    """two replicas with the same generator seed produce bit-identical sets (what keeps DP ranks in sync)"""
    d, cfg = golden_inputs.densify_case("mixed")
    outs = []
    for _ in range(2):
        g = helpers.flat_from_groups(dp, d, dev)
        gen = torch.Generator(device=dev).manual_seed(77)
        densify.densify_and_prune(g, _stats_from(d), cfg["max_grad"], cfg["min_opacity"], cfg["extent"], None,
                                  generator=gen)
        outs.append(g.data.clone())
    assert torch.equal(outs[0], outs[1])


def test_densification_stats_and_reset_opacity_match_reference_golden():
    gold = np.load(os.path.join(GOLD, "densify_reference.npz"))
    s = golden_inputs.stats_case()
    P = s["radii"].shape[0]
    st = densify.DensifyStats(P, dev)
    st.xyz_gradient_accum.copy_(T(s["xyz_gradient_accum"]).reshape(-1))
    st.denom.copy_(T(s["denom"]).reshape(-1))
    st.max_radii2D.copy_(T(s["max_radii2D"]))
    densify.add_densification_stats(st, T(s["grad2D"]), T(s["radii"]))
    np.testing.assert_allclose(st.xyz_gradient_accum.cpu().numpy(), gold["stats/xyz_gradient_accum"].reshape(-1), rtol=1e-6)
    np.testing.assert_array_equal(st.denom.cpu().numpy(), gold["stats/denom"].reshape(-1))
    np.testing.assert_array_equal(st.max_radii2D.cpu().numpy(), gold["stats/max_radii2D"])
    g = dp.FlatGaussians(P, torch.device(dev))
    g.slab("opacity").copy_(T(s["opacity"]))
    g.exp_avg.fill_(1.0)
    g.exp_avg_sq.fill_(2.0)
    before = g.data.clone()
    densify.reset_opacity(g)
    np.testing.assert_allclose(g.slab("opacity").cpu().numpy(), gold["reset/opacity"], rtol=1e-6)
    assert not g.slab("opacity", g.exp_avg).any() and not g.slab("opacity", g.exp_avg_sq).any()
    a, b, _ = g._slices["opacity"]
    keep = torch.ones_like(before, dtype=torch.bool)
    keep[a:b] = False
    assert torch.equal(g.data[keep], before[keep]) and bool((g.exp_avg[keep] == 1).all())


def test_fused_activations_match_torch():
    """lg_activate_forward vs torch.sigmoid / torch.exp / F.normalize on the same device, incl. a zero quaternion"""
    P = 100_003
    g = dp.FlatGaussians(P, torch.device(dev))
    g.data.copy_(torch.randn(g.data.shape, generator=torch.Generator().manual_seed(3)).to(dev) * 2.0)
    g.slab("rotation")[5].zero_()
    act = g.activate()
    o, s, r = densify_oracle.activations(g.slab("opacity"), g.slab("scaling"), g.slab("rotation"))
    torch.testing.assert_close(act["opacities"], o, rtol=2e-7, atol=1e-9)
    torch.testing.assert_close(act["scales"], s, rtol=2e-7, atol=0)
    torch.testing.assert_close(act["rotations"], r, rtol=3e-7, atol=1e-9)
    torch.testing.assert_close(act["rot_norm"], g.slab("rotation").norm(dim=1), rtol=3e-7, atol=0)
    assert act["means3D"].data_ptr() == g.slab("xyz").data_ptr() and act["shs"].data_ptr() == g.slab("shs").data_ptr()


def _small_training_set(P=20_000, n_views=3):
    sc = scenes.trained_like_scene(P, seed=5, log_scale_mean=np.log(0.02))
    g = dp.FlatGaussians.from_scene(sc, torch.device(dev))
    g.slab("rotation").mul_(torch.rand(P, 1, device=dev) * 2 + 0.5)  # raw quaternions are not unit length
    cams = [dp.camera_to_device(c, dev) for c in scenes.orbit_cameras(n_views, 160, 128)]
    gts = [torch.rand(3, 128, 160, device=dev, generator=torch.Generator(device=dev).manual_seed(v)) for v in range(n_views)]
    return g, cams, gts


def test_fused_step_gradients_equal_autograd_through_torch_activations():
    """raw-parameter gradients from the in-kernel chain rule (lg_rasterize_backward_raw + GradSinks accumulation over
    three views) == autograd through torch.sigmoid / exp / normalize around the same operator"""
    g, cams, gts = _small_training_set()
    bg = torch.zeros(3, device=dev)
    fused = dp.ViewParallelTrainer(g)                                   # render_fn=None -> fused path
    loss_f = fused.accumulate_views(cams, gts, bg)
    grad_fused = g.grad.clone()
    generic = dp.ViewParallelTrainer(g, render_fn=dp.default_render)    # autograd leaves + torch activations
    loss_g = generic.accumulate_views(cams, gts, bg)
    torch.testing.assert_close(loss_f, loss_g, rtol=1e-5, atol=1e-7)
    for name, _ in g.fields:
        a, b = g.slab(name, grad_fused), g.slab(name, g.grad)
        scale = float(b.abs().max())
        assert scale > 0, name
        err = float((a - b).abs().max()) / scale
        assert err < 1e-3, (name, err)


def test_trainer_single_gpu_with_density_control():
    """fused step + statistics + densify_and_prune + opacity reset on the schedule of LG/train.py:265-276"""
    g, cams, gts = _small_training_set(P=30_000, n_views=2)
    cfg = dp.DensifyConfig(densify_from_iter=2, densify_until_iter=50, densification_interval=3,
                           opacity_reset_interval=7, densify_grad_threshold=1e-7, cameras_extent=4.0)
    tr = dp.ViewParallelTrainer(g, densify=cfg, seed=3)
    bg = torch.zeros(3, device=dev)
    sizes, losses = [], []
    for _ in range(9):
        losses.append(float(tr.step(cams, gts, bg)))
        sizes.append(g.P)
    assert sizes[0] == sizes[1] == 30_000 and sizes[2] != 30_000 and sizes[5] != sizes[4], sizes
    assert torch.isfinite(g.data).all() and g.grad.numel() == g.data.numel() == g.floats * g.stride
    assert tr.stats.denom.numel() == g.P
    assert np.isfinite(losses).all()
    # after the reset at iteration 7 no opacity exceeds sigmoid^-1(0.01) by more than two Adam steps' worth
    assert float(torch.sigmoid(g.slab("opacity")).max()) < 0.05


def test_two_stream_view_overlap_computes_the_single_stream_step():
    """view_streams=2 (consecutive views of a step on alternating CUDA streams, the head of view k+1 under the blend
    backward of view k) keeps the loss state, the bucket and the densification statistics in view order.  The first
    step (nothing amplified by Adam yet) must give the single-stream loss, running mean and gradient bucket to the
    last-bit spread of the backward's atomics; six steps of four views must stay as close to a single-stream run as a
    second single-stream run does (the backward adds with atomics, so two runs of the SAME schedule differ too)."""
    def run(streams):
        torch.manual_seed(11)  # _small_training_set scales the raw quaternions with torch.rand
        g, cams, gts = _small_training_set(P=25_000, n_views=4)
        cfg = dp.DensifyConfig(densify_from_iter=100, densify_until_iter=200)  # statistics only, no edit
        tr = dp.ViewParallelTrainer(g, loss_fn=dp.RunningMeanLoss(torch.device(dev)), densify=cfg, view_streams=streams)
        assert (tr.view_streams is not None) == (streams == 2)
        bg = torch.zeros(3, device=dev)
        tr.begin_iteration()
        first_loss = float(tr.accumulate_views(cams, gts, bg))
        first = (first_loss, g.grad.clone(), float(tr.loss_fn.running_mean), tr.stats.denom.clone())
        tr.exchange_and_update(0.25)
        losses = [float(tr.step(cams, gts, bg)) for _ in range(5)]
        torch.cuda.synchronize()
        return first, losses, g.data.clone(), tr.stats.xyz_gradient_accum.clone(), tr.stats.denom.clone()

    def spread(x, y):
        return dict(loss=float(np.max(np.abs(np.array(x[1]) - np.array(y[1])) / np.abs(np.array(x[1])))),
                    far=float(((x[2] - y[2]).abs() > 1e-4 + 1e-4 * x[2].abs()).float().mean()),
                    accum=float((x[3] - y[3]).abs().max() / x[3].abs().max()))

    one, again, two = run(1), run(1), run(2)
    (l1, g1, r1, d1), (l2, g2, r2, d2) = one[0], two[0]
    assert abs(l1 - l2) <= 2e-6 * abs(l1), (l1, l2)
    assert abs(r1 - r2) <= 2e-6 * abs(r1), (r1, r2)
    assert torch.equal(d1, d2)
    assert float((g2 - g1).abs().max()) <= 2e-5 * float(g1.abs().max())
    noise, diff = spread(one, again), spread(one, two)
    assert torch.equal(one[4], two[4])
    for k in noise:
        assert diff[k] <= 4.0 * noise[k] + 1e-5, (k, diff, noise)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs with NVLink / PCIe peer access")
def test_peer_exchange_two_gpus():
    """fused reduce-scatter + Adam + all-gather over peer memory == NCCL all-reduce + local Adam (tests/peer_worker.py)"""
    import subprocess
    import sys
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29641", os.path.join(os.path.dirname(os.path.abspath(__file__)), "peer_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "PEER_KERNELS_OK" in r.stdout and "PEER_TRAINER_OK" in r.stdout, r.stdout + r.stderr


def test_adam_scalar_fallback_for_unaligned_buffers():
    """lg_adam_step on a buffer whose size / segment ends are not multiples of 4 (the float4 kernel cannot be used)"""
    import ctypes
    from lgdwt_b200 import _lib
    n, cut = 1003, 501
    gen = torch.Generator().manual_seed(4)
    p0, g0 = torch.randn(n, generator=gen), torch.randn(n, generator=gen)
    ref_a, ref_b = p0[:cut].clone().requires_grad_(True), p0[cut:].clone().requires_grad_(True)
    opt = torch.optim.Adam([{"params": [ref_a], "lr": 0.01}, {"params": [ref_b], "lr": 0.002}], lr=0.0, eps=1e-15)
    p, m, v = p0.to(dev), torch.zeros(n, device=dev), torch.zeros(n, device=dev)
    ends = (ctypes.c_longlong * 2)(cut, n)
    lrs = (ctypes.c_float * 2)(0.01, 0.002)
    for step in range(1, 4):
        g = (g0 * step).to(dev)
        ref_a.grad, ref_b.grad = (g0 * step)[:cut].clone(), (g0 * step)[cut:].clone()
        opt.step()
        _lib.check(_lib.lib.lg_adam_step(p.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr(), n, 2, ends, lrs, 0.9,
                                         0.999, 1e-15, step, 1.0, _lib.stream_ptr(torch.device(dev))))
    torch.testing.assert_close(p.cpu(), torch.cat([ref_a.detach(), ref_b.detach()]), rtol=2e-5, atol=1e-7)


@pytest.mark.parametrize("shape", [(3, 800, 800), (3, 37, 53), (1, 4, 64, 64), (7,)])
def test_fused_l1_loss_matches_torch(shape):
    from lgdwt_b200 import fused_l1_loss
    gen = torch.Generator(device=dev).manual_seed(sum(shape))
    gt = torch.rand(shape, device=dev, generator=gen)
    pred = (gt + 0.1 * torch.randn(shape, device=dev, generator=gen)).requires_grad_(True)
    with torch.no_grad():
        pred[..., :2] = gt[..., :2]          # exact ties: sign(0) = 0 in both
    ref = pred.detach().clone().requires_grad_(True)
    (3.0 * fused_l1_loss(pred, gt)).backward()
    (3.0 * (ref - gt).abs().mean()).backward()
    torch.testing.assert_close(fused_l1_loss(pred, gt), (ref - gt).abs().mean(), rtol=2e-6, atol=1e-8)
    torch.testing.assert_close(pred.grad, ref.grad, rtol=1e-6, atol=0)


def test_fused_ssim_dropin_matches_pytorch_ssim():
    """the optional `fused_ssim` import of LG/train.py:36-40 resolved to this repo: value and gradient equal the
    reference's PyTorch ssim (restated in oracle/photometric_oracle.py, pinned on loss_utils.ssim)"""
    from fused_ssim import fused_ssim
    from oracle import photometric_oracle
    pred_np, gt_np = golden_inputs.photometric_case_inputs("smooth_odd", 3, 67, 131)
    gt = T(gt_np).unsqueeze(0)
    a = T(pred_np).unsqueeze(0).requires_grad_(True)
    b = T(pred_np).unsqueeze(0).requires_grad_(True)
    v = fused_ssim(a, gt)
    r = photometric_oracle.ssim(b, gt)
    torch.testing.assert_close(v, r, rtol=1e-5, atol=1e-7)
    (1.0 - v).backward()
    (1.0 - r).backward()
    torch.testing.assert_close(a.grad, b.grad, rtol=1e-4, atol=1e-8)
    two = fused_ssim(torch.cat([a.detach(), gt]), torch.cat([gt, gt]), train=False)
    torch.testing.assert_close(two, 0.5 * (v.detach() + 1.0), rtol=1e-5, atol=1e-6)
    with pytest.raises(NotImplementedError):
        fused_ssim(a, gt, padding="valid")


def test_fused_trainer_reproduces_the_reference_iteration():
    """N = 1, one view per step: 30 iterations of the fused trainer with a TrainSchedule and the stateful loss against
    the STOCK op chain of LG/train.py:99-288 — six nn.Parameters, torch activations (gaussian_model.py:102-130), the
    rasterizer as an ordinary autograd op (the reference CUDA when oracle/_ref is built), the PyTorch losses, the
    host-side running-mean DWT scale (:190-196), oneupSHdegree (:101-103, every 10 iterations here), the exponential
    position learning rate (gaussian_model.py:203-223), the depth-L1 term (:204-216) and torch.optim.Adam.
    Adam turns a gradient into a step of size ~lr whatever its magnitude, so a parameter whose gradient is summation
    noise around zero can move differently in the two chains; the bar is therefore 99.9 % of the parameters within
    1e-4 (absolute, after 30 steps of lr up to 2.5e-2), every parameter within 30 x its learning rate x 2, and the
    per-iteration loss within 1e-4 relative."""
    from oracle import dwt_oracle, photometric_oracle, ref_cuda
    sc = scenes.trained_like_scene(20_000, seed=41, log_scale_mean=np.log(0.03))
    W, H, iters = 272, 200, 30
    cams_np = scenes.orbit_cameras(5, W, H)
    cams = [dp.camera_to_device(c, dev) for c in cams_np]
    gen = torch.Generator(device=dev).manual_seed(5)
    gts = [torch.rand((3, H, W), device=dev, generator=gen) for _ in cams]
    depth_targets = [(torch.rand((1, H, W), device=dev, generator=gen), (torch.rand((1, H, W), device=dev, generator=gen) > 0.3).float())
                     for _ in cams]
    bg = torch.zeros(3, device=dev)
    sched = dp.TrainSchedule(iterations=60, position_lr_max_steps=40, spatial_lr_scale=2.5, oneup_sh_every=10)

    # ---- fused trainer
    g = dp.FlatGaussians.from_scene(sc, dev)
    tr = dp.ViewParallelTrainer(g, loss_fn=dp.RunningMeanLoss(dev), schedule=sched)
    fused_losses = []
    for it in range(iters):
        k = it % len(cams)
        fused_losses.append(float(tr.step([cams[k]], [gts[k]], bg, depths=[depth_targets[k]])))
    assert g.active_sh_degree == 3 and tr.iteration == iters

    # ---- stock op chain
    g0 = dp.FlatGaussians.from_scene(sc, dev)
    P = g0.P
    raw = {k: torch.nn.Parameter(g0.field(k).clone().reshape(P, *shape)) for k, shape in
           (("xyz", (3,)), ("f_dc", (1, 3)), ("f_rest", (15, 3)), ("opacity", (1,)), ("scaling", (3,)), ("rotation", (4,)))}
    a = dp.AdamConfig()
    lrs = dict(xyz=sched.position_lr_init * sched.spatial_lr_scale, f_dc=a.lr_f_dc, f_rest=a.lr_f_rest,
               opacity=a.lr_opacity, scaling=a.lr_scaling, rotation=a.lr_rotation)
    opt = torch.optim.Adam([{"params": [raw[k]], "lr": lrs[k], "name": k} for k in raw], lr=0.0, eps=1e-15)
    use_ref = ref_cuda.load_ref() is not None
    from diff_gaussian_rasterization import GaussianRasterizationSettings, GaussianRasterizer
    rm, active, stock_losses = 1.0, 0, []
    for it in range(1, iters + 1):
        for group in opt.param_groups:                       # update_learning_rate (gaussian_model.py:213-223)
            if group["name"] == "xyz":
                group["lr"] = sched.position_lr(it)
        if it % sched.oneup_sh_every == 0:
            active = min(active + 1, 3)
        k = (it - 1) % len(cams)
        cam, gt = cams[k], gts[k]
        shs = torch.cat((raw["f_dc"], raw["f_rest"]), dim=1)
        vsp = torch.zeros_like(raw["xyz"], requires_grad=True)
        args = (raw["xyz"], vsp, shs, torch.sigmoid(raw["opacity"]), torch.exp(raw["scaling"]),
                torch.nn.functional.normalize(raw["rotation"]))
        if use_ref:
            image, radii, invd = ref_cuda.RefRasterize.apply(*args, cam, bg, active)
        else:
            rs = GaussianRasterizationSettings(H, W, cam["tanfovx"], cam["tanfovy"], bg, 1.0, cam["viewmatrix"],
                                               cam["projmatrix"], active, cam["campos"], False, False, False)
            image, radii, invd = GaussianRasterizer(rs)(means3D=args[0], means2D=vsp, shs=shs, opacities=args[3],
                                                         scales=args[4], rotations=args[5])
        image = image.clamp(0, 1)
        l1, ssim = photometric_oracle.photometric_terms(image, gt)
        dwt, patch, _, _ = dwt_oracle.lgdwt_losses(image, gt)
        base = 0.8 * l1 + 0.2 * (1.0 - ssim)
        rm = 0.95 * rm + 0.05 * (base / (dwt + 1e-8)).item()
        loss = base + float(max(0.1, min(10.0, rm))) * dwt + 0.1 * patch
        mono, mask = depth_targets[k]
        loss = loss + sched.depth_l1_weight(it) * torch.abs((invd - mono) * mask).mean()
        loss.backward()
        stock_losses.append(float(loss))
        opt.step()
        opt.zero_grad(set_to_none=True)

    rel = np.abs(np.array(fused_losses) - np.array(stock_losses)) / np.abs(np.array(stock_losses))
    assert rel.max() <= 1e-4, "per-iteration loss: max rel diff %.3g at iteration %d" % (rel.max(), int(rel.argmax()) + 1)
    assert abs(float(tr.loss_fn.running_mean) - rm) <= 1e-5 * abs(rm)
    worst = {}
    for k in raw:
        mine, ref = g.field(k).reshape(-1), raw[k].detach().reshape(-1)
        d = (mine - ref).abs()
        frac = float((d <= 1e-4).float().mean())
        worst[k] = (frac, float(d.max()))
        assert frac >= 0.999, "%s: only %.5f of the parameters within 1e-4" % (k, frac)
        assert float(d.max()) <= 2 * iters * max(lrs[k], sched.position_lr(1)), "%s: max diff %g" % (k, float(d.max()))
    print("fused trainer vs stock chain after %d iterations: loss rel %.2g; per group (fraction within 1e-4, max): %s"
          % (iters, rel.max(), worst))


def test_peer_exchange_falls_back_when_sh_row_is_not_16_byte_sized():
    """SH degree 0 / 2 (rows of 3 / 27 floats) cannot take the 16-byte peer Adam path: the trainer must say so and
    keep the NCCL exchange instead of failing at the first step (single process: the decision is local)"""
    for deg, ok in ((0, False), (1, True), (2, False), (3, True)):
        g = dp.FlatGaussians(1000, torch.device(dev), sh_degree=deg)
        assert ((3 * g.M) % 4 == 0) == ok
