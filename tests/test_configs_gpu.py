"""Every BASELINE.json rasterizer config at FULL size against the UNMODIFIED reference CUDA (oracle/_ref) executed live on
the same GPU: metric scene (1 M, 800x800), cfg3 (500 k, 1008x756, 44-bit keys), cfg4 (1 M, 1296x964, 45-bit keys, plus
the RGB+NIR 4-channel pass against two reference passes), cfg5 (6 M, 1920x1080, 13-bit tile ids).

Bars: integer / bit tier identical; image <= 1e-5 absolute; gradients, against the mean of three (atomically
accumulated, hence nondeterministic) reference runs,
  * per tensor   max|d| <= 2e-4 * max|ref|                      (north_star asks for 1e-3)
  * per element  |d| <= 1e-3 * |ref| + 2e-6 * max|ref|          for at least 99.9 % of the elements — an element far
    below the tensor's maximum cannot be wrong and hide behind the maximum — and
                 |d| <= 1e-2 * |ref| + 2e-4 * max|ref|          for every element.
The per-element absolute terms are what float32 summation order costs: a gradient is a sum over up to thousands of
pixels, the reference adds them with atomics in arbitrary order, and the reference's own run-to-run spread is checked
against the same bars first, so a failure here is a real difference and not noise.
"""
import json
import os

import pytest
import torch

import helpers
from lgdwt_b200 import scenes

pytestmark = pytest.mark.gpu

from helpers import assert_grad_close, grad_stats  # noqa: E402  (the three gradient bars described above)

IMG_ATOL = 1e-5
BIT_EXACT = ("radii", "tiles_touched", "point_offsets", "point_list_keys", "point_list", "ranges", "n_contrib", "final_T")
GRADS = ("dL_dmean2D", "dL_dconic", "dL_dopacity", "dL_dmean3D", "dL_dsh", "dL_dscale", "dL_drot")
STATS_OUT = os.path.join(helpers.ROOT, "gpurun_out", "configs_gpu_stats.json")


def _need_ref():
    if helpers.load_ref() is None:
        pytest.skip("oracle/_ref/libref_dgr.so not prebuilt (needs /root/reference at build time)")


def _record(case, stats):
    try:
        os.makedirs(os.path.dirname(STATS_OUT), exist_ok=True)
        data = json.load(open(STATS_OUT)) if os.path.exists(STATS_OUT) else {}
        data[case] = stats
        json.dump(data, open(STATS_OUT, "w"), indent=1)
    except OSError:
        pass


def _bit_equal(a, b):
    if a.dtype == torch.float32:
        a, b = a.view(torch.int32), b.view(torch.int32)
    return a.shape == b.shape and bool((a == b).all())


@pytest.mark.parametrize("name", ["metric", "cfg3", "cfg4", "cfg5"])
def test_baseline_config_full_size_vs_reference(name):
    _need_ref()
    sc, cam = scenes.baseline_config(name)
    t, c = helpers.scene_to_torch(sc), helpers.cam_to_torch(cam)
    bg = torch.tensor([0.0, 0.3, 0.6], device="cuda")
    ours = helpers.run_ours(t, c, cam, bg)
    ref = helpers.run_ref(t, c, cam, bg)
    assert ours["num_rendered"] == ref["num_rendered"] > 0
    bad = [k for k in BIT_EXACT if not _bit_equal(ours[k], ref[k])]
    assert not bad, "%s: not bit-identical with the reference: %s" % (name, bad)
    vis = ref["radii"] > 0
    for k, w in (("depths", 1), ("means2D", 2), ("conic_opacity", 4), ("rgb", 3), ("clamped", 3)):
        assert _bit_equal(ours[k].view(-1, w)[vis], ref[k].view(-1, w)[vis]), k
    assert float((ours["color"] - ref["color"]).abs().max()) <= IMG_ATOL
    assert float((ours["invdepth"] - ref["invdepth"]).abs().max()) <= IMG_ATOL
    R, n_vis = ours["num_rendered"], int(vis.sum())
    for k in list(ours):
        if k not in ("radii", "num_rendered", "binning_capacity", "geom", "binning", "img", "C", "M"):
            del ours[k]
    for k in list(ref):
        if k not in ("radii", "num_rendered", "binning_capacity", "geom", "binning", "img", "C", "M"):
            del ref[k]
    torch.cuda.empty_cache()
    gen = torch.Generator(device="cuda").manual_seed(3)
    dL = torch.randn((3, cam.image_height, cam.image_width), device="cuda", generator=gen)
    dLd = torch.randn((1, cam.image_height, cam.image_width), device="cuda", generator=gen)
    mine = helpers.backward_ours(t, c, cam, bg, ours, dL, dLd)
    refs = [helpers.backward_ref(t, c, cam, bg, ref, dL, dLd) for _ in range(3)]
    stats = {"P": int(sc.means3D.shape[0]), "num_rendered": R, "visible": n_vis, "grads": {}, "reference_spread": {}}
    for g in GRADS + ("dL_dinvdepth",):
        mean = torch.stack([r[g] for r in refs]).double().mean(0)
        a, b = mine[g], mean
        if g == "dL_dconic":   # element [2] is unused by both sides
            a, b = a[:, [0, 1, 3]], b[:, [0, 1, 3]]
        # the reference against itself (run 0 vs the mean): the noise floor the bars have to sit above
        r0 = refs[0][g][:, [0, 1, 3]] if g == "dL_dconic" else refs[0][g]
        stats["reference_spread"][g] = grad_stats(r0, b)
        stats["grads"][g] = assert_grad_close(g, a, b)
    _record(name, stats)


def test_rgb_nir_one_four_channel_pass_vs_two_reference_passes():
    """cfg4: RGB+NIR.  The reference renders twice with 3-channel `colors_precomp` (render + render_nir and takes one
    channel of the second); here ONE 4-channel pass must give the same four images and the same summed gradients."""
    _need_ref()
    sc, cam = scenes.baseline_config("cfg4")
    t, c = helpers.scene_to_torch(sc), helpers.cam_to_torch(cam)
    P = sc.means3D.shape[0]
    gen = torch.Generator(device="cuda").manual_seed(5)
    col4 = torch.rand((P, 4), device="cuda", generator=gen)
    bg4 = torch.tensor([0.1, 0.2, 0.3, 0.4], device="cuda")
    dL4 = torch.randn((4, cam.image_height, cam.image_width), device="cuda", generator=gen)
    ours = helpers.run_ours(t, c, cam, bg4, colors_precomp=col4, want_state=False)
    rgb = col4[:, :3].contiguous()
    nir = col4[:, 3:4].repeat(1, 3).contiguous()            # MS/gaussian_renderer/__init__.py:196-197
    bg_nir = bg4[3:4].repeat(3).contiguous()
    ref_rgb = helpers.run_ref(t, c, cam, bg4[:3].contiguous(), colors_precomp=rgb, want_state=False)
    ref_nir = helpers.run_ref(t, c, cam, bg_nir, colors_precomp=nir, want_state=False)
    assert torch.equal(ours["radii"], ref_rgb["radii"])
    assert float((ours["color"][:3] - ref_rgb["color"]).abs().max()) <= IMG_ATOL
    assert float((ours["color"][3] - ref_nir["color"][0]).abs().max()) <= IMG_ATOL
    mine = helpers.backward_ours(t, c, cam, bg4, ours, dL4, None, colors_precomp=col4)
    dL_nir = torch.zeros((3, cam.image_height, cam.image_width), device="cuda")
    dL_nir[0] = dL4[3]                                      # rendered_image[0:1] is all the caller keeps (:255)
    stats = {}
    acc = None
    for _ in range(3):
        g1 = helpers.backward_ref(t, c, cam, bg4[:3].contiguous(), ref_rgb, dL4[:3].contiguous(), None, colors_precomp=rgb)
        g2 = helpers.backward_ref(t, c, cam, bg_nir, ref_nir, dL_nir, None, colors_precomp=nir)
        tot = {k: g1[k].double() + g2[k].double() for k in ("dL_dmean2D", "dL_dopacity", "dL_dmean3D", "dL_dscale", "dL_drot")}
        tot["dL_dcolor"] = torch.cat([g1["dL_dcolor"].double(), g2["dL_dcolor"].double().sum(1, keepdim=True)], 1)
        acc = tot if acc is None else {k: acc[k] + tot[k] for k in tot}
    for k, v in acc.items():
        stats[k] = assert_grad_close(k, mine[k], v / 3.0)
    _record("cfg4_rgb_nir", stats)


def test_prefiltered_flag_matches_reference():
    """`prefiltered=True` (SURVEY §8b optional paths): legal only when no point is behind the near plane — the
    reference traps otherwise (auxiliary.h:168-172) and so does this library.  Results are those of prefiltered=False."""
    _need_ref()
    sc = scenes.trained_like_scene(20_000, seed=12, log_scale_mean=-4.0)
    cam = scenes.metric_camera(320, 240)
    t, c = helpers.scene_to_torch(sc), helpers.cam_to_torch(cam)
    bg = torch.zeros(3, device="cuda")
    depth = t["means3D"] @ c["viewmatrix"][:3, 2] + c["viewmatrix"][3, 2]
    assert float(depth.min()) > 0.2                      # every point in front of the near plane
    ours = helpers.run_ours(t, c, cam, bg, prefiltered=True)
    ref = helpers.run_ref(t, c, cam, bg, prefiltered=True)
    base = helpers.run_ours(t, c, cam, bg, prefiltered=False)
    for k in BIT_EXACT:
        assert _bit_equal(ours[k], ref[k]), k
        assert _bit_equal(ours[k], base[k]), k
    assert torch.equal(ours["color"], base["color"])
    assert float((ours["color"] - ref["color"]).abs().max()) <= IMG_ATOL


def test_prefiltered_violation_traps_like_the_reference(tmp_path):
    """a culled point with prefiltered=True: the reference prints and `__trap()`s (auxiliary.h:168-172), which kills
    the CUDA context — run in a child process and expect it to fail with a CUDA error."""
    import subprocess
    import sys
    code = (
        "import sys, torch\n"
        "sys.path[:0] = [%r, %r, %r]\n"
        "import helpers\n"
        "from lgdwt_b200 import scenes\n"
        "sc = scenes.trained_like_scene(2000, seed=12, log_scale_mean=-4.0)\n"
        "cam = scenes.metric_camera(160, 120)\n"
        "t, c = helpers.scene_to_torch(sc), helpers.cam_to_torch(cam)\n"
        "t['means3D'][7, 2] = -50.0\n"
        "try:\n"
        "    helpers.run_ours(t, c, cam, torch.zeros(3, device='cuda'), prefiltered=True, want_state=False)\n"
        "    torch.cuda.synchronize()\n"
        "except Exception as e:\n"
        "    print('RAISED', type(e).__name__, e); sys.exit(3)\n"
        "sys.exit(0)\n" % (helpers.ROOT, os.path.join(helpers.ROOT, "sparse-view-3dgs-pack_b200"),
                           os.path.join(helpers.ROOT, "tests")))
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert r.returncode != 0, "a filtered point under prefiltered=True must fail as in the reference\n" + r.stdout
    assert "Point is filtered although prefiltered is set" in r.stdout + r.stderr
