"""GPU tier, no oracle/_ref needed: the CUDA path called through the C ABI / public Python surface against
(a) the committed golden vectors produced by the reference itself and (b) the CPU oracle on the same seeded inputs,
plus size-independent properties at the benchmark's full size."""
import os

import numpy as np
import pytest
import torch

import golden_inputs
import helpers
from lgdwt_b200 import DWTLossConfig, fused_dwt_loss, scenes
from oracle import dwt_oracle, oracle

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
dev = "cuda"


def T(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(dev)


def _run_case(name):
    sc, cam, bg, ex = golden_inputs.raster_case(name)
    t, c = helpers.scene_to_torch(sc), helpers.cam_to_torch(cam)
    kw = dict(antialiasing=ex["aa"])
    if ex["mode"] == "precomp":
        kw.update(colors_precomp=T(ex["colors_precomp"]), cov3D_precomp=T(ex["cov3D_precomp"]))
    f = helpers.run_ours(t, c, cam, T(bg), **kw)
    g = helpers.backward_ours(t, c, cam, T(bg), f, T(ex["dL_dpix"]), T(ex["dL_dinvd"]), **kw)
    return sc, cam, bg, ex, f, g


@pytest.mark.parametrize("name", list(golden_inputs.RASTER_CASES))
def test_rasterizer_matches_reference_golden(name):
    gold = np.load(os.path.join(GOLD, "rasterizer_reference.npz"))
    sc, cam, bg, ex, f, g = _run_case(name)
    G = lambda k: gold["%s/%s" % (name, k)]
    assert f["num_rendered"] == int(G("num_rendered"))
    for k in ("radii", "tiles_touched", "point_offsets", "point_list_keys", "point_list", "ranges", "n_contrib"):
        np.testing.assert_array_equal(f[k].cpu().numpy().reshape(-1), G(k).reshape(-1), err_msg=k)   # bit-exact tier
    np.testing.assert_array_equal(f["final_T"].cpu().numpy().view(np.uint32), G("final_T").view(np.uint32))
    assert np.abs(f["color"].cpu().numpy() - G("color")).max() <= 1e-5                                 # image tier
    assert np.abs(f["invdepth"].cpu().numpy() - G("invdepth")).max() <= 1e-5
    for k, mine in g.items():                                                                         # gradient tier
        key = "%s/%s" % (name, k)
        if mine is None or key not in gold:
            continue
        ref = gold[key].reshape(tuple(mine.shape))
        m = mine.cpu().numpy()
        if k == "dL_dconic":
            m, ref = m[:, [0, 1, 3]], ref[:, [0, 1, 3]]
        rel = np.abs(m.astype(np.float64) - ref).max() / max(np.abs(ref).max(), 1e-12)
        assert rel <= 1e-3, "%s relative error %g" % (k, rel)


def test_public_operator_autograd_matches_oracle():
    from diff_gaussian_rasterization import GaussianRasterizationSettings, GaussianRasterizer
    sc, cam, bg, ex = golden_inputs.raster_case("g_sh")
    p = {k: T(getattr(sc, k)).requires_grad_(True) for k in ("means3D", "shs", "opacities", "scales", "rotations")}
    rs = GaussianRasterizationSettings(cam.image_height, cam.image_width, cam.tanfovx, cam.tanfovy, T(bg), 1.0,
                                       T(cam.viewmatrix), T(cam.projmatrix), 3, T(cam.campos), False, False, False)
    means2D = (torch.zeros_like(p["means3D"], requires_grad=True) + 0)
    means2D.retain_grad()
    color, radii, invd = GaussianRasterizer(rs)(means3D=p["means3D"], means2D=means2D, shs=p["shs"],
                                                 opacities=p["opacities"], scales=p["scales"], rotations=p["rotations"])
    assert color.shape == (3, cam.image_height, cam.image_width) and invd.shape == (1, cam.image_height, cam.image_width)
    assert radii.dtype == torch.int32
    ((color * T(ex["dL_dpix"])).sum() + (invd * T(ex["dL_dinvd"])).sum()).backward()
    kw = oracle.scene_kwargs(sc, cam, bg)
    o = oracle.rasterize_forward(**kw)
    go = oracle.rasterize_backward(o, dL_dpix=ex["dL_dpix"], dL_dinvdepth_pix=ex["dL_dinvd"], **kw)
    np.testing.assert_array_equal(radii.cpu().numpy(), o["radii"])
    pairs = (("dL_dmean3D", p["means3D"].grad), ("dL_dmean2D", means2D.grad), ("dL_dsh", p["shs"].grad),
             ("dL_dopacity", p["opacities"].grad), ("dL_dscale", p["scales"].grad), ("dL_drot", p["rotations"].grad))
    for k, mine in pairs:
        ref = go[k].reshape(tuple(mine.shape))
        rel = np.abs(mine.cpu().numpy().astype(np.float64) - ref).max() / max(np.abs(ref).max(), 1e-12)
        assert rel <= 2e-3, "%s relative error vs oracle %g" % (k, rel)


def test_grad_sinks_overwrite_and_accumulate():
    """GradSinks extension: the backward writes the parameter gradients into caller-owned slices of one flat bucket;
    accumulate=False overwrites (garbage in the bucket is ignored), accumulate=True adds a second view on top — equal to
    what autograd's own accumulation of the plain operator gives."""
    from diff_gaussian_rasterization import GaussianRasterizationSettings, GaussianRasterizer, GradSinks
    sc, cam, bg, ex = golden_inputs.raster_case("g_sh")
    names = ("means3D", "shs", "opacities", "scales", "rotations")
    p = {k: T(getattr(sc, k)).requires_grad_(True) for k in names}
    P = p["means3D"].shape[0]
    rs = GaussianRasterizationSettings(cam.image_height, cam.image_width, cam.tanfovx, cam.tanfovy, T(bg), 1.0,
                                       T(cam.viewmatrix), T(cam.projmatrix), 3, T(cam.campos), False, False, False)
    dL = [T(ex["dL_dpix"]), T(ex["dL_dpix"][::-1].copy() * 0.5)]

    def run(sinks, d):
        m2 = torch.zeros((P, 3), device=dev, requires_grad=True)
        color, _, _ = GaussianRasterizer(rs)(means3D=p["means3D"], means2D=m2, shs=p["shs"], opacities=p["opacities"],
                                             scales=p["scales"], rotations=p["rotations"], grad_sinks=sinks)
        color.backward(d)
        return m2.grad

    for t in p.values():
        t.grad = None
    m2_plain = run(None, dL[0])
    run(None, dL[1])
    want = {k: p[k].grad.clone() for k in names}      # autograd summed the two calls
    for t in p.values():
        t.grad = None

    bucket = torch.full((59 * P,), float("nan"), device=dev)  # overwrite mode must not read the old contents
    widths, views, off = (3, 48, 1, 3, 4), {}, 0
    for k, w in zip(names, widths):
        views[k] = bucket[off * P:(off + w) * P].view(p[k].shape)
        off += w
    sinks = GradSinks(views["means3D"], views["shs"], views["opacities"], views["scales"], views["rotations"])
    m2_sink = run(sinks, dL[0])
    assert all(p[k].grad is None for k in names)      # nothing handed to autograd a second time
    assert torch.isfinite(bucket).all()
    rel = (m2_sink - m2_plain).abs().max() / m2_plain.abs().max()
    assert rel <= 1e-4
    sinks.accumulate = True
    run(sinks, dL[1])
    for k in names:
        rel = float((views[k] - want[k]).abs().max() / want[k].abs().max().clamp_min(1e-12))
        assert rel <= 1e-4, "%s: accumulated sink differs from autograd accumulation by %g" % (k, rel)


def test_four_channel_rgb_nir_render_matches_oracle():
    """N-channel generalisation: one 4-channel pass == the oracle's 4-channel blend (and == two reference-style
    3-channel passes on the shared channels)."""
    sc, cam, bg, ex = golden_inputs.raster_case("g_sh")
    rng = np.random.default_rng(3)
    P = sc.means3D.shape[0]
    colors4 = rng.random((P, 4)).astype(np.float32)
    bg4 = np.array([0.1, 0.2, 0.3, 0.4], np.float32)
    t, c = helpers.scene_to_torch(sc), helpers.cam_to_torch(cam)
    f4 = helpers.run_ours(t, c, cam, T(bg4), colors_precomp=T(colors4))
    f3 = helpers.run_ours(t, c, cam, T(bg4[:3]), colors_precomp=T(colors4[:, :3].copy()))
    assert f4["color"].shape[0] == 4
    assert torch.equal(f4["color"][:3], f3["color"]) and torch.equal(f4["n_contrib"], f3["n_contrib"])
    kw = oracle.scene_kwargs(sc, cam, bg4)
    kw.update(shs=None, colors_precomp=colors4)
    o = oracle.rasterize_forward(**kw)
    err = np.abs(f4["color"].cpu().numpy() - o["color"])
    assert np.quantile(err, 0.999) <= 2e-5
    dL = rng.standard_normal((4, cam.image_height, cam.image_width)).astype(np.float32)
    g = helpers.backward_ours(t, c, cam, T(bg4), f4, T(dL), None, colors_precomp=T(colors4))
    go = oracle.rasterize_backward(o, dL_dpix=dL, **kw)
    for k in ("dL_dcolor", "dL_dmean3D", "dL_dopacity", "dL_dscale"):
        ref = go[k].reshape(tuple(g[k].shape))
        rel = np.abs(g[k].cpu().numpy().astype(np.float64) - ref).max() / max(np.abs(ref).max(), 1e-12)
        assert rel <= 2e-3, "%s %g" % (k, rel)


def test_empty_and_all_culled_inputs():
    from diff_gaussian_rasterization import GaussianRasterizationSettings, GaussianRasterizer
    cam = scenes.metric_camera(64, 48)
    bg = torch.tensor([0.2, 0.3, 0.4], device=dev)
    rs = GaussianRasterizationSettings(48, 64, cam.tanfovx, cam.tanfovy, bg, 1.0, T(cam.viewmatrix), T(cam.projmatrix),
                                       3, T(cam.campos), False, False, False)
    z = lambda *s: torch.zeros(s, device=dev)
    color, radii, invd = GaussianRasterizer(rs)(means3D=z(0, 3), means2D=z(0, 3), shs=z(0, 16, 3), opacities=z(0, 1),
                                                 scales=z(0, 3), rotations=z(0, 4))
    assert color.shape == (3, 48, 64) and radii.shape == (0,) and float(color.abs().max()) == 0.0  # reference: fill value 0
    # every Gaussian behind the camera: num_rendered == 0, image == background
    P = 100
    m = z(P, 3)
    m[:, 2] = -10.0
    rot = z(P, 4)
    rot[:, 0] = 1
    color, radii, invd = GaussianRasterizer(rs)(means3D=m.requires_grad_(True), means2D=z(P, 3), shs=z(P, 16, 3),
                                                 opacities=z(P, 1) + 0.5, scales=z(P, 3) + 0.1, rotations=rot)
    assert int(radii.abs().sum()) == 0
    assert torch.allclose(color, bg[:, None, None].expand_as(color))
    color.sum().backward()
    assert float(m.grad.abs().max()) == 0.0


def test_error_behaviour_matches_reference():
    from diff_gaussian_rasterization import GaussianRasterizationSettings, GaussianRasterizer
    cam = scenes.metric_camera(32, 32)
    rs = GaussianRasterizationSettings(32, 32, cam.tanfovx, cam.tanfovy, torch.zeros(3, device=dev), 1.0,
                                       T(cam.viewmatrix), T(cam.projmatrix), 3, T(cam.campos), False, False, False)
    r = GaussianRasterizer(rs)
    z = lambda *s: torch.zeros(s, device=dev)
    with pytest.raises(Exception, match="excatly one of either SHs or precomputed colors"):
        r(means3D=z(4, 3), means2D=z(4, 3), opacities=z(4, 1), scales=z(4, 3), rotations=z(4, 4))
    with pytest.raises(Exception, match="scale/rotation pair or precomputed 3D covariance"):
        r(means3D=z(4, 3), means2D=z(4, 3), opacities=z(4, 1), shs=z(4, 16, 3), scales=z(4, 3))
    with pytest.raises(RuntimeError, match="means3D must have dimensions"):
        r(means3D=z(4, 2), means2D=z(4, 3), opacities=z(4, 1), shs=z(4, 16, 3), scales=z(4, 3), rotations=z(4, 4))
    with pytest.raises(RuntimeError, match="For non-RGB, provide precomputed"):
        from lgdwt_b200 import _lib
        import ctypes
        b = [_lib.ResizableBuffer(torch.device(dev)) for _ in range(3)]
        n = ctypes.c_int(0)
        o = z(4, 32, 32)
        rc = _lib.lib.lg_rasterize_forward(b[0].callback, None, b[1].callback, None, b[2].callback, None, 4, 3, 16, 4,
                                           z(4).data_ptr(), 32, 32, z(4, 3).data_ptr(), z(4, 16, 3).data_ptr(), None,
                                           z(4, 1).data_ptr(), z(4, 3).data_ptr(), 1.0, z(4, 4).data_ptr(), None,
                                           T(cam.viewmatrix).data_ptr(), T(cam.projmatrix).data_ptr(),
                                           T(cam.campos).data_ptr(), cam.tanfovx, cam.tanfovy, 0, o.data_ptr(), None, 0,
                                           None, 0, None, ctypes.byref(n))
        _lib.check(rc, RuntimeError)


def test_full_size_properties():
    """1 M Gaussians, 800x800 (BASELINE metric size): size-independent invariants of the binning and the blend."""
    sc = scenes.trained_like_scene(1_000_000, seed=1)
    cam = scenes.metric_camera()
    t, c = helpers.scene_to_torch(sc), helpers.cam_to_torch(cam)
    bg = torch.zeros(3, device=dev)
    f = helpers.run_ours(t, c, cam, bg)
    R = f["num_rendered"]
    keys, plist, ranges = f["point_list_keys"], f["point_list"].long(), f["ranges"].view(-1, 2).long()
    assert R == int(f["tiles_touched"].long().sum()) == int(f["point_offsets"][-1])
    assert bool((keys[1:] >= keys[:-1]).all())                                   # sortedness
    tile = keys >> 32
    depth_bits = (keys & 0xFFFFFFFF).int()
    assert torch.equal(depth_bits, f["depths"].view(torch.int32)[plist])         # key low word = depth bits of value
    same = keys[1:] == keys[:-1]
    assert bool((plist[1:][same] > plist[:-1][same]).all())                      # stability: ties in ascending id
    counts = torch.bincount(tile, minlength=ranges.shape[0])
    assert torch.equal(ranges[:, 1] - ranges[:, 0], counts)                      # ranges partition the list
    nz = counts > 0
    assert torch.equal(ranges[nz, 0], (torch.cumsum(counts, 0) - counts)[nz])
    assert torch.equal(torch.bincount(plist, minlength=t["means3D"].shape[0]).int(), f["tiles_touched"])  # multiset kept
    # every pixel's last contributor lies inside its tile's list; transmittance in (0, 1]
    gx = (cam.image_width + 15) // 16
    ys, xs = torch.meshgrid(torch.arange(cam.image_height, device=dev), torch.arange(cam.image_width, device=dev), indexing="ij")
    tid = ((ys // 16) * gx + xs // 16).reshape(-1)
    assert bool((f["n_contrib"].long() <= counts[tid]).all())
    assert float(f["final_T"].min()) > 0.0 and float(f["final_T"].max()) <= 1.0
    # determinism of the forward, linearity of the backward in the upstream gradient
    f2 = helpers.run_ours(t, c, cam, bg, want_state=False)
    assert torch.equal(f["color"], f2["color"])
    gen = torch.Generator(device=dev).manual_seed(0)
    dL = torch.randn((3, cam.image_height, cam.image_width), device=dev, generator=gen)
    g1 = helpers.backward_ours(t, c, cam, bg, f, dL)
    g2 = helpers.backward_ours(t, c, cam, bg, f, 2.0 * dL)
    for k in ("dL_dmean3D", "dL_dsh", "dL_dopacity", "dL_dscale", "dL_drot"):
        assert helpers.rel_err(g2[k], 2.0 * g1[k]) <= 1e-3, k
    invisible = f["radii"] == 0
    assert float(g1["dL_dsh"][invisible].abs().max()) == 0.0 and float(g1["dL_dmean3D"][invisible].abs().max()) == 0.0


def test_distcuda2_matches_golden_and_oracle():
    from simple_knn._C import distCUDA2
    pts = golden_inputs.knn_points()
    d = distCUDA2(T(pts)).cpu().numpy()
    gold = np.load(os.path.join(GOLD, "knn_reference.npz"))["dist2"]
    np.testing.assert_allclose(d, gold, rtol=1e-6, atol=1e-12)
    for P in (1, 2, 3, 4, 7, 1023, 1024, 1025, 5000):
        sub = pts[:P]
        mine = distCUDA2(T(sub)).cpu().numpy()
        ref = oracle.knn_mean_dist2(sub)
        np.testing.assert_allclose(mine, ref, rtol=1e-6, atol=1e-12, err_msg="P=%d" % P)
    assert distCUDA2(torch.zeros((0, 3), device=dev)).shape == (0,)
    big = np.random.default_rng(1).uniform(-1.3, 1.3, (200_000, 3)).astype(np.float32)
    np.testing.assert_allclose(distCUDA2(T(big)).cpu().numpy(), oracle.knn_mean_dist2(big), rtol=1e-6, atol=1e-12)


DWT_CASES = ["even_6patch", "all_bands", "odd_sizes", "four_channels", "smaller_than_patch"]


@pytest.mark.parametrize("name", DWT_CASES)
def test_fused_dwt_loss_matches_reference_golden_and_oracle(name):
    gold = np.load(os.path.join(GOLD, "dwt_reference.npz"))
    cfg = gold[name + "/cfg"]
    C, H, W, ps = (int(v) for v in cfg[:4])
    pct, w_lh, w_hl, g_dwt, g_patch = (float(v) for v in cfg[4:9])
    wts = tuple(float(v) for v in cfg[9:17])
    pred, gt = golden_inputs.dwt_case_inputs(name, C, H, W)
    p = T(pred).requires_grad_(True)
    lc = DWTLossConfig(wts, ps, pct, w_lh, w_hl, True)
    dwt, patch, details = fused_dwt_loss(p, T(gt), lc)
    np.testing.assert_allclose(details[2:10].cpu().numpy(), gold[name + "/band_l1"], rtol=2e-6)   # <= 1e-6-ish relative
    np.testing.assert_allclose(float(dwt), gold[name + "/dwt_loss"], rtol=2e-6)
    np.testing.assert_allclose(float(patch), gold[name + "/patch_loss"], rtol=2e-6, atol=1e-12)
    (g_dwt * dwt + g_patch * patch).backward()
    grad = p.grad.cpu().numpy()
    np.testing.assert_allclose(grad[:, ::3, ::5], gold[name + "/grad_sub"], atol=1e-9, rtol=1e-4)
    st = gold[name + "/grad_stats"]
    np.testing.assert_allclose([grad.astype(np.float64).sum(), np.abs(grad).astype(np.float64).sum()], st[:2], rtol=1e-5, atol=1e-9)
    # patch selection identical to the oracle's
    _, _, _, mask = dwt_oracle.lgdwt_losses(torch.from_numpy(pred), torch.from_numpy(gt), wts, ps, pct, w_lh, w_hl)
    if mask is not None:
        assert int(details[10]) == int(mask.sum())
    # ELF map: the selection threshold the kernel found is the k-th smallest of the reference's own per-patch ELF means
    # (compute_elf_map + unfold + mean, stored by make_dwt_golden.py), and the selected count follows from it
    if name + "/patch_elf_means" in gold.files:
        means = gold[name + "/patch_elf_means"].astype(np.float64)
        k = min(max(1, int(means.size * (1.0 - pct))), means.size)
        thr = np.sort(means)[k - 1]
        assert abs(float(details[11]) - thr) <= 2e-5 * abs(thr), (float(details[11]), thr)
        margin = np.abs(means - thr) > 1e-5 * abs(thr)      # patches not within rounding of the threshold
        assert int((means[margin] >= thr).sum()) <= int(details[10]) <= int((means[margin] >= thr).sum() + (~margin).sum())


@pytest.mark.parametrize("name", list(golden_inputs.PHOTOMETRIC_CASES))
def test_fused_photometric_loss_matches_reference_golden(name):
    """L1 + SSIM terms of the base loss: fused kernels vs outputs / autograd gradients of the reference's own
    loss_utils.l1_loss / ssim (tolerances: 2e-6 relative on the scalars, 1e-8 + 1e-4 relative on the gradient)."""
    from lgdwt_b200 import fused_photometric_loss
    gold = np.load(os.path.join(GOLD, "photometric_reference.npz"))
    C, H, W = golden_inputs.PHOTOMETRIC_CASES[name]
    pred, gt = golden_inputs.photometric_case_inputs(name, C, H, W)
    p = T(pred).requires_grad_(True)
    l1, s = fused_photometric_loss(p, T(gt))
    np.testing.assert_allclose(float(l1), gold[name + "/l1"], rtol=2e-6)
    np.testing.assert_allclose(float(s), gold[name + "/ssim"], rtol=2e-6)
    (0.8 * l1 + 0.2 * (1.0 - s)).backward()
    grad = p.grad.cpu().numpy()
    np.testing.assert_allclose(grad[:, ::3, ::5], gold[name + "/grad_sub"], atol=1e-8, rtol=1e-4)
    st = gold[name + "/grad_stats"]
    np.testing.assert_allclose([grad.astype(np.float64).sum(), np.abs(grad).astype(np.float64).sum()], st[:2],
                               rtol=1e-5, atol=1e-9)


def test_fused_photometric_loss_full_size_vs_oracle_fp64():
    """3x800x800 (the benchmark image size) and a (1,C,H,W) input: fused kernels vs the fp64 oracle."""
    from lgdwt_b200 import fused_photometric_loss
    from oracle import photometric_oracle
    pred, gt = scenes.dwt_pair(3, 800, 800, seed=3)
    p = T(pred).unsqueeze(0).requires_grad_(True)
    l1, s = fused_photometric_loss(p, T(gt).unsqueeze(0))
    (0.8 * l1 + 0.2 * (1.0 - s)).backward()
    po = torch.from_numpy(pred).double().requires_grad_(True)
    l1o, so = photometric_oracle.photometric_terms(po, torch.from_numpy(gt).double())
    (0.8 * l1o + 0.2 * (1.0 - so)).backward()
    assert abs(float(l1) - float(l1o)) <= 1e-6 * float(l1o)
    assert abs(float(s) - float(so)) <= 1e-5 * abs(float(so))
    g, go = p.grad[0].cpu().double(), po.grad
    assert float((g - go).abs().max()) <= 1e-9 + 1e-4 * float(go.abs().max())
    # no gradient requested -> the forward skips the derivative maps and still returns the same scalars
    with torch.no_grad():
        l1n, sn = fused_photometric_loss(T(pred), T(gt))
    assert float(l1n) == float(l1) and float(sn) == float(s)


def test_fused_dwt_loss_full_size_config1():
    """BASELINE config 1: 3x800x800 pair, default weights — fused kernel vs the oracle (fp32 and fp64)."""
    pred, gt = scenes.dwt_pair(3, 800, 800, seed=0)
    p = T(pred).requires_grad_(True)
    dwt, patch, details = fused_dwt_loss(p, T(gt))
    (dwt + 0.1 * patch).backward()
    po = torch.from_numpy(pred).double().requires_grad_(True)
    d64, p64, _, mask = dwt_oracle.lgdwt_losses(po, torch.from_numpy(gt).double())
    (d64 + 0.1 * p64).backward()
    assert abs(float(dwt) - float(d64)) <= 2e-6 * float(d64)
    assert abs(float(patch) - float(p64)) <= 2e-6 * float(p64)
    assert int(details[10]) == int(mask.sum()) == 9
    # image gradient: identical to the fp32 oracle (same roundings => same signs) ...
    p32 = torch.from_numpy(pred).requires_grad_(True)
    d32, q32, _, _ = dwt_oracle.lgdwt_losses(p32, torch.from_numpy(gt))
    (d32 + 0.1 * q32).backward()
    assert np.abs(p.grad.cpu().numpy() - p32.grad.numpy()).max() <= 1e-6 * 1e-3
    # ... and equal to the fp64 oracle except where an fp32 sub-band difference rounds across zero (sign flip)
    d = np.abs(p.grad.cpu().numpy() - po.grad.numpy())
    assert (d > 1e-9).mean() <= 1e-4 and d.max() <= 4.2e-6


def test_pytorch_wavelets_compat_module():
    from pytorch_wavelets import DWTForward
    x = torch.randn(2, 3, 37, 50, device=dev, requires_grad=True)
    yl, yh = DWTForward(J=2, mode="symmetric", wave="db1").to(dev)(x)
    xo = x.detach().cpu().requires_grad_(True)
    ol, oh = dwt_oracle.DWTForward(J=2, mode="symmetric", wave="db1")(xo)
    assert yl.shape == ol.shape and [t.shape for t in yh] == [t.shape for t in oh]
    torch.testing.assert_close(yl.cpu(), ol, atol=1e-6, rtol=1e-6)
    for a, b in zip(yh, oh):
        torch.testing.assert_close(a.cpu(), b, atol=1e-6, rtol=1e-6)
    (yl.sum() + sum((t * t).sum() for t in yh)).backward()
    (ol.sum() + sum((t * t).sum() for t in oh)).backward()
    torch.testing.assert_close(x.grad.cpu(), xo.grad, atol=1e-5, rtol=1e-5)
    with pytest.raises(NotImplementedError):
        DWTForward(J=1, mode="zero", wave="db1")


def test_dp_trainer_single_gpu_step_changes_parameters_like_reference_iteration():
    from lgdwt_b200 import dp
    sc = scenes.trained_like_scene(20_000, seed=5, log_scale_mean=np.log(0.02))
    g = dp.FlatGaussians.from_scene(sc, torch.device(dev))
    cams = [dp.camera_to_device(c, dev) for c in scenes.orbit_cameras(2, 160, 128)]
    gts = [torch.rand(3, 128, 160, device=dev) for _ in cams]
    tr = dp.ViewParallelTrainer(g)
    before = g.data.clone()
    l0 = float(tr.step(cams, gts, torch.zeros(3, device=dev)))
    for _ in range(5):
        l1 = float(tr.step(cams, gts, torch.zeros(3, device=dev)))
    assert torch.isfinite(g.data).all() and not torch.equal(before, g.data)
    assert l1 < l0, (l0, l1)


def test_fused_adam_matches_torch_optim_adam():
    """lg_adam_step_split on the flat (59 x P) buffer == torch.optim.Adam with one param group per field (the reference's
    optimiser set-up, LG/scene/gaussian_model.py:178-211), over several steps, plus the folded 1/views gradient scale."""
    from lgdwt_b200 import dp
    P = 1237
    g = dp.FlatGaussians(P, torch.device(dev))
    g.data.copy_(torch.randn(g.data.shape, generator=torch.Generator().manual_seed(2)).to(dev))
    cfg = dp.AdamConfig()
    ref_params = {n: g.field(n).clone().requires_grad_(True) for n in dp.GROUPS}
    lrs = dict(xyz=cfg.lr_xyz, f_dc=cfg.lr_f_dc, f_rest=cfg.lr_f_rest, opacity=cfg.lr_opacity, scaling=cfg.lr_scaling,
               rotation=cfg.lr_rotation)
    opt = torch.optim.Adam([{"params": [ref_params[n]], "lr": lrs[n]} for n in dp.GROUPS], lr=0.0, eps=1e-15)
    for it in range(5):
        grads = torch.randn(g.data.shape, generator=torch.Generator().manual_seed(10 + it)).to(dev)
        scale = 0.25 if it == 3 else 1.0
        g.grad.copy_(grads)
        for n in dp.GROUPS:
            ref_params[n].grad = (g.field(n, grads) * scale).clone()
        opt.step()
        g.adam_step(cfg, grad_scale=scale)
    for n in dp.GROUPS:
        torch.testing.assert_close(g.field(n), ref_params[n].detach(), rtol=2e-5, atol=1e-7)


@pytest.mark.parametrize("hint_factor", [0.3, 1.0, 4.0])
def test_speculative_binning_capacity_gives_identical_results(hint_factor):
    """lg_rasterize_forward_hinted: the whole forward is queued for a guessed binning capacity before the host learns
    num_rendered.  Whatever the guess (too small -> the tail is queued again at the exact size; exact; generous), the
    outputs, the state and the gradients are those of the synchronous call."""
    from lgdwt_b200 import scenes
    sc = scenes.trained_like_scene(40_000, seed=31, log_scale_mean=np.log(0.02))
    cam = scenes.look_at_camera(400, 304, 0.6911, 0.6911 * 304 / 400, (0.2, -0.1, -4.03))
    t, c = helpers.scene_to_torch(sc), helpers.cam_to_torch(cam)
    bg = torch.tensor([0.3, 0.2, 0.1], device="cuda")
    base = helpers.run_ours(t, c, cam, bg)
    R = base["num_rendered"]
    assert R > 1000 and base["binning_capacity"] == R
    hinted = helpers.run_ours(t, c, cam, bg, capacity_hint=max(1, int(R * hint_factor)))
    assert hinted["num_rendered"] == R
    assert hinted["binning_capacity"] == (R if hint_factor < 1.0 else int(R * hint_factor))
    for k in ("radii", "tiles_touched", "ranges", "point_list_keys", "point_list", "n_contrib", "final_T", "color",
              "invdepth"):
        assert torch.equal(hinted[k], base[k]), k
    dL = torch.randn((3, cam.image_height, cam.image_width), device="cuda",
                     generator=torch.Generator(device="cuda").manual_seed(2))
    g0 = helpers.backward_ours(t, c, cam, bg, base, dL, None)
    g1 = helpers.backward_ours(t, c, cam, bg, hinted, dL, None)
    # the backward adds the per-warp partial sums with floating-point atomics whose order differs from launch to launch:
    # two runs of the SAME state differ by ~1e-5 of the tensor's maximum, so that is the bar here (the state is bit-equal)
    for k in ("dL_dmean2D", "dL_dmean3D", "dL_dsh", "dL_dopacity", "dL_dscale", "dL_drot"):
        assert helpers.rel_err(g1[k], g0[k]) <= 1e-4, k


@pytest.mark.parametrize("shape", [(3, 264, 392), (3, 800, 800), (4, 301, 257)])
def test_fused_image_loss_equals_the_reference_loss_assembly(shape):
    """fused_image_loss (one autograd node) against the loss assembly of LG/train.py:128-202 written out with the CPU
    oracle's PyTorch restatement of the reference terms: value, the running-mean update over three consecutive
    iterations, and the image gradient."""
    from lgdwt_b200 import fused_image_loss
    from oracle import photometric_oracle
    C, H, W = shape
    gen = torch.Generator(device="cuda").manual_seed(3)
    gt = torch.rand((C, H, W), device="cuda", generator=gen)
    rm_dev = torch.ones((), device="cuda")
    rm_host = 1.0
    for it in range(3):
        pred = (gt + 0.05 * (it + 1) * torch.randn((C, H, W), device="cuda", generator=gen)).clamp(0, 1).requires_grad_(True)
        loss, terms = fused_image_loss(pred, gt, rm_dev)
        loss.backward()
        ref_pred = pred.detach().clone().requires_grad_(True)
        l1, ssim = photometric_oracle.photometric_terms(ref_pred, gt)
        dwt, patch, _, _ = dwt_oracle.lgdwt_losses(ref_pred, gt)
        base = 0.8 * l1 + 0.2 * (1.0 - ssim)
        rm_host = 0.95 * rm_host + 0.05 * (base / (dwt + 1e-8)).item()          # LG/train.py:193-195
        ref = base + float(max(0.1, min(10.0, rm_host))) * dwt + 0.1 * patch
        ref.backward()
        assert abs(float(loss) - float(ref)) <= 2e-6 * abs(float(ref)), (it, float(loss), float(ref))
        assert abs(float(rm_dev) - rm_host) <= 1e-6 * abs(rm_host), (it, float(rm_dev), rm_host)
        np.testing.assert_allclose(terms[:4].cpu().numpy(), [float(l1), float(ssim), float(dwt), float(patch)], rtol=2e-6, atol=1e-8)
        err = (pred.grad - ref_pred.grad).abs().max()
        assert float(err) <= 1e-5 * float(ref_pred.grad.abs().max()) + 1e-9, (it, float(err))
    # update_running_mean=False leaves the scalar alone and uses it as given
    before = float(rm_dev)
    pred = gt.clone().mul_(0.9).requires_grad_(True)
    fused_image_loss(pred, gt, rm_dev, update_running_mean=False)[0].backward()
    assert float(rm_dev) == before
    # an upstream gradient other than 1 reaches both gradient kernels (each scales its coefficients by it in-kernel)
    g1 = pred.grad.clone()
    pred.grad = None
    (2.5 * fused_image_loss(pred, gt, rm_dev, update_running_mean=False)[0]).backward()
    torch.testing.assert_close(pred.grad, 2.5 * g1, rtol=2e-6, atol=1e-12)
