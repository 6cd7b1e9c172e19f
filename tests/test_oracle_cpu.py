"""CPU tier: pins the oracle (oracle/*.c, oracle/dwt_oracle.py) against golden vectors produced by the reference
itself — the reference CUDA rasterizer / simple-knn run on a B200 (tests/golden/make_rasterizer_golden.py) and the
reference's LG/utils/loss_utils.py run in the build container (tests/golden/make_dwt_golden.py)."""
import os

import numpy as np
import pytest
import torch

import golden_inputs
from oracle import dwt_oracle, oracle

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def rgold():
    return np.load(os.path.join(GOLD, "rasterizer_reference.npz"))


def _kwargs(name):
    sc, cam, bg, ex = golden_inputs.raster_case(name)
    kw = oracle.scene_kwargs(sc, cam, bg)
    kw["antialiasing"] = ex["aa"]
    if ex["mode"] == "precomp":
        kw.update(shs=None, scales=None, rotations=None, colors_precomp=ex["colors_precomp"],
                  cov3D_precomp=ex["cov3D_precomp"])
    return kw, ex


@pytest.mark.parametrize("name", list(golden_inputs.RASTER_CASES))
def test_oracle_forward_matches_reference_cuda(rgold, name):
    kw, ex = _kwargs(name)
    o = oracle.rasterize_forward(**kw)
    g = lambda k: rgold["%s/%s" % (name, k)]
    assert o["num_rendered"] == int(g("num_rendered")) > 0
    # integer / bit-exact tier
    np.testing.assert_array_equal(o["radii"], g("radii"))
    np.testing.assert_array_equal(o["tiles_touched"], g("tiles_touched").view(np.uint32))
    np.testing.assert_array_equal(o["point_offsets"], g("point_offsets").view(np.uint32))
    vis = g("radii") > 0
    for k, w in (("depths", 1), ("means2D", 2), ("conic_opacity", 4)):
        a, b = o[k].reshape(-1, w)[vis], g(k).reshape(-1, w)[vis]
        np.testing.assert_array_equal(a.view(np.uint32), b.view(np.uint32), err_msg=k)
    if ex["mode"] == "sh":
        np.testing.assert_array_equal(o["cov3D"][vis].view(np.uint32), g("cov3D").reshape(-1, 6)[vis].view(np.uint32))
        np.testing.assert_array_equal(o["clamped"][vis], g("clamped").reshape(-1, 3)[vis])
        np.testing.assert_array_equal(o["rgb"][vis].view(np.uint32), g("rgb").reshape(-1, 3)[vis].view(np.uint32))
    np.testing.assert_array_equal(o["point_list_keys"], g("point_list_keys").view(np.uint64))
    np.testing.assert_array_equal(o["point_list"], g("point_list").view(np.uint32))
    np.testing.assert_array_equal(o["ranges"].reshape(-1), g("ranges").view(np.uint32))
    # blend tier: exp() differs from MUFU.EX2 by an ulp or two, so thresholds can flip on isolated pixels
    same = o["n_contrib"] == g("n_contrib").view(np.uint32)
    assert same.mean() >= 0.999, "n_contrib agreement %.5f" % same.mean()
    err = np.abs(o["color"] - g("color"))
    assert np.quantile(err, 0.999) <= 2e-5 and err.max() <= 2e-2, (np.quantile(err, 0.999), err.max())
    errd = np.abs(o["invdepth"] - g("invdepth"))
    assert np.quantile(errd, 0.999) <= 2e-5
    assert np.abs(o["final_T"][same] - g("final_T")[same]).max() <= 1e-5


@pytest.mark.parametrize("name", list(golden_inputs.RASTER_CASES))
def test_oracle_backward_matches_reference_cuda(rgold, name):
    kw, ex = _kwargs(name)
    o = oracle.rasterize_forward(**kw)
    gr = oracle.rasterize_backward(o, dL_dpix=ex["dL_dpix"], dL_dinvdepth_pix=ex["dL_dinvd"], **kw)
    for k, mine in gr.items():
        key = "%s/%s" % (name, k)
        if mine is None or key not in rgold:
            continue
        ref = rgold[key].reshape(mine.shape)
        if k == "dL_dconic":
            mine, ref = mine[:, [0, 1, 3]], ref[:, [0, 1, 3]]
        rel = np.abs(mine.astype(np.float64) - ref).max() / max(np.abs(ref).max(), 1e-12)
        assert rel <= 2e-3, "%s relative error %g" % (k, rel)


def test_higher_msb_is_bit_length():
    """rasterizer_impl.cu:35-50 restated in the oracle == the closed form the CUDA side uses"""
    for n in list(range(1, 5000)) + [8159, 8160, 8192, 65535, 65536, 2 ** 20 + 3, 2 ** 31 - 1]:
        assert oracle.higher_msb(n) == int(n).bit_length(), n


def test_mark_visible_oracle():
    sc, cam, bg, ex = golden_inputs.raster_case("g_sh")
    pts = sc.means3D.copy()
    pts[::3, 2] -= 6.0
    vis = oracle.mark_visible(pts, cam.viewmatrix)
    depth = pts @ cam.viewmatrix[:3, 2] + cam.viewmatrix[3, 2]
    assert 0 < vis.sum() < len(vis)
    margin = np.abs(depth - 0.2) > 1e-4
    np.testing.assert_array_equal(vis[margin], (depth > 0.2)[margin])


def test_knn_oracle_matches_reference_cuda_and_bruteforce():
    pts = golden_inputs.knn_points()
    gold = np.load(os.path.join(GOLD, "knn_reference.npz"))["dist2"]
    mine = oracle.knn_mean_dist2(pts)
    np.testing.assert_allclose(mine, gold, rtol=1e-6, atol=1e-12)
    assert (mine[: len(pts) // 50] >= 0).all()
    small = pts[:1500]
    np.testing.assert_allclose(oracle.knn_mean_dist2(small), oracle.knn_mean_dist2(small, brute_force=True), rtol=1e-6)


def test_knn_oracle_edge_cases():
    assert oracle.knn_mean_dist2(np.zeros((0, 3), np.float32)).shape == (0,)
    three = np.array([[0, 0, 0], [1, 0, 0], [0, 2, 0], [0, 0, 3]], np.float32)
    np.testing.assert_allclose(oracle.knn_mean_dist2(three)[0], (1 + 4 + 9) / 3.0, rtol=1e-6)


DWT_CASES = ["even_6patch", "all_bands", "odd_sizes", "four_channels", "smaller_than_patch"]


@pytest.mark.parametrize("name", DWT_CASES)
def test_dwt_oracle_matches_reference_loss_utils(name):
    gold = np.load(os.path.join(GOLD, "dwt_reference.npz"))
    cfg = gold[name + "/cfg"]
    C, H, W, ps = (int(v) for v in cfg[:4])
    pct, w_lh, w_hl, g_dwt, g_patch = cfg[4:9]
    wts = tuple(float(v) for v in cfg[9:17])
    pred, gt = golden_inputs.dwt_case_inputs(name, C, H, W)
    p = torch.from_numpy(pred).requires_grad_(True)
    dwt, patch, band_l1, mask = dwt_oracle.lgdwt_losses(p, torch.from_numpy(gt), wts, ps, pct, w_lh, w_hl)
    np.testing.assert_allclose([float(band_l1[n]) for n in dwt_oracle.BAND_NAMES], gold[name + "/band_l1"], rtol=1e-6)
    np.testing.assert_allclose(float(dwt), gold[name + "/dwt_loss"], rtol=1e-6)
    np.testing.assert_allclose(float(patch), gold[name + "/patch_loss"], rtol=1e-6, atol=1e-12)
    (g_dwt * dwt + g_patch * patch).backward()
    np.testing.assert_allclose(p.grad.numpy()[:, ::3, ::5], gold[name + "/grad_sub"], atol=1e-9, rtol=1e-5)
    # the ELF map itself (compute_elf_map, LG/utils/loss_utils.py:336-366) and its per-patch means, not only what follows
    elf = dwt_oracle.compute_elf_map(torch.from_numpy(gt).unsqueeze(0))
    np.testing.assert_allclose(float(elf.mean()), gold[name + "/elf_mean"], rtol=1e-6)
    if name + "/patch_elf_means" in gold.files:
        means = dwt_oracle.patch_selection(elf, ps, float(pct))[2].view(-1).numpy()
        np.testing.assert_allclose(means, gold[name + "/patch_elf_means"], rtol=1e-5, atol=1e-7)


@pytest.mark.parametrize("name", list(golden_inputs.PHOTOMETRIC_CASES))
def test_photometric_oracle_matches_reference_loss_utils(name):
    from oracle import photometric_oracle
    gold = np.load(os.path.join(GOLD, "photometric_reference.npz"))
    C, H, W = golden_inputs.PHOTOMETRIC_CASES[name]
    pred, gt = golden_inputs.photometric_case_inputs(name, C, H, W)
    p = torch.from_numpy(pred).requires_grad_(True)
    l1, s = photometric_oracle.photometric_terms(p, torch.from_numpy(gt))
    np.testing.assert_allclose(float(l1), gold[name + "/l1"], rtol=1e-6)
    np.testing.assert_allclose(float(s), gold[name + "/ssim"], rtol=1e-6)
    (0.8 * l1 + 0.2 * (1.0 - s)).backward()
    np.testing.assert_allclose(p.grad.numpy()[:, ::3, ::5], gold[name + "/grad_sub"], atol=1e-9, rtol=1e-5)


def test_haar_level_closed_form():
    x = torch.arange(2 * 3 * 6 * 8, dtype=torch.float32).reshape(2, 3, 6, 8).sin()
    LL, (LH, HL, HH) = dwt_oracle.haar_dwt_level(x)
    a, b, c, d = x[..., 0::2, 0::2], x[..., 0::2, 1::2], x[..., 1::2, 0::2], x[..., 1::2, 1::2]
    torch.testing.assert_close(LL, 0.5 * (a + b + c + d), atol=1e-6, rtol=0)
    torch.testing.assert_close(LH, 0.5 * ((a + b) - (c + d)), atol=1e-6, rtol=0)
    torch.testing.assert_close(HL, 0.5 * ((a - b) + (c - d)), atol=1e-6, rtol=0)
    torch.testing.assert_close(HH, 0.5 * (a - b - c + d), atol=1e-6, rtol=0)
    # orthonormal transform: energy is preserved for even sizes
    e = (LL ** 2).sum() + (LH ** 2).sum() + (HL ** 2).sum() + (HH ** 2).sum()
    torch.testing.assert_close(e, (x ** 2).sum(), rtol=1e-5, atol=0)


# ---------------------------------------------------------------- density control (oracle vs the reference class)
@pytest.fixture(scope="module")
def dgold():
    return np.load(os.path.join(GOLD, "densify_reference.npz"))


@pytest.mark.parametrize("name", list(golden_inputs.DENSIFY_CASES))
def test_densify_oracle_matches_reference_gaussian_model(dgold, name):
    from oracle import densify_oracle
    d, cfg = golden_inputs.densify_case(name)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a))
    p = {k: t(d[k]) for k in densify_oracle.GROUPS}
    m = {k: t(d["exp_avg/" + k]) for k in densify_oracle.GROUPS}
    v = {k: t(d["exp_avg_sq/" + k]) for k in densify_oracle.GROUPS}
    eps = t(dgold[name + "/eps"])
    np_, nm, nv, counts = densify_oracle.densify_and_prune(
        p, m, v, t(d["xyz_gradient_accum"]), t(d["denom"]), eps, cfg["max_grad"], cfg["min_opacity"], cfg["extent"],
        cfg["max_screen_size"], cfg["percent_dense"])
    assert counts["P"] == dgold[name + "/xyz"].shape[0]
    assert 2 * counts["split_parents"] == eps.shape[0]
    for k in densify_oracle.GROUPS:
        np.testing.assert_allclose(np_[k].numpy(), dgold["%s/%s" % (name, k)], rtol=1e-6, atol=1e-6, err_msg=k)
        np.testing.assert_array_equal(nm[k].numpy(), dgold["%s/exp_avg/%s" % (name, k)], err_msg=k)
        np.testing.assert_array_equal(nv[k].numpy(), dgold["%s/exp_avg_sq/%s" % (name, k)], err_msg=k)
    # densification_postfix: statistics restart from zero at the new size
    assert dgold[name + "/denom"].shape[0] == counts["P"] and not dgold[name + "/denom"].any()
    assert not dgold[name + "/max_radii2D"].any() and not dgold[name + "/xyz_gradient_accum"].any()


def test_densify_stats_and_reset_opacity_oracle_match_reference(dgold):
    from oracle import densify_oracle
    s = golden_inputs.stats_case()
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a))
    a, dn, mr = densify_oracle.add_densification_stats(t(s["xyz_gradient_accum"]), t(s["denom"]), t(s["max_radii2D"]),
                                                       t(s["grad2D"]), t(s["radii"]))
    np.testing.assert_allclose(a.numpy(), dgold["stats/xyz_gradient_accum"].reshape(-1), rtol=1e-6, atol=0)
    np.testing.assert_array_equal(dn.numpy(), dgold["stats/denom"].reshape(-1))
    np.testing.assert_array_equal(mr.numpy(), dgold["stats/max_radii2D"])
    np.testing.assert_allclose(densify_oracle.reset_opacity(t(s["opacity"])).numpy(), dgold["reset/opacity"], rtol=1e-6)
