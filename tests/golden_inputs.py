"""Seeded inputs shared by the golden-vector generators and the tests that consume the fixtures."""
import numpy as np


def dwt_case_inputs(name, C, H, W):
    rng = np.random.default_rng(sum(map(ord, name)))
    gt = rng.random((C, H, W)).astype(np.float32)
    pred = np.clip(gt + 0.05 * rng.standard_normal((C, H, W)), 0, 1).astype(np.float32)
    return pred, gt


PHOTOMETRIC_CASES = {  # name: (C, H, W) — sizes around / below the 32x16 tile and the 11-tap window, odd sizes, C = 4
    "noise": (3, 72, 104),
    "smooth_odd": (3, 67, 131),
    "tiny": (3, 7, 9),
    "four_channels": (4, 48, 80),
}


def photometric_case_inputs(name, C, H, W):
    """render-like pair: smooth structure + texture (pure noise for the "noise" case), pred = perturbed gt in [0,1]"""
    rng = np.random.default_rng(1000 + sum(map(ord, name)))
    if name == "noise":
        gt = rng.random((C, H, W)).astype(np.float32)
    else:
        yy, xx = np.mgrid[0:H, 0:W].astype(np.float64)
        gt = np.stack([0.5 + 0.35 * np.sin(0.11 * (c + 1) * xx + 0.3 * c) * np.cos(0.07 * (c + 2) * yy) for c in range(C)])
        gt = np.clip(gt + 0.08 * rng.standard_normal((C, H, W)), 0, 1).astype(np.float32)
    pred = np.clip(0.9 * gt + 0.04 + 0.05 * rng.standard_normal((C, H, W)), 0, 1).astype(np.float32)
    pred[:, : H // 3, : W // 4] = gt[:, : H // 3, : W // 4]  # an exactly matching region: sign(0) = 0 in the L1 gradient
    return pred, gt


RASTER_CASES = {
    # name: (P, seed, spacing_scale, W, H, antialiasing, mode)
    "g_sh": (800, 11, 0.11, 120, 72, False, "sh"),
    "g_sh_aa": (800, 12, 0.09, 104, 88, True, "sh"),
    "g_precomp": (600, 13, 0.12, 96, 64, False, "precomp"),
}


def raster_case(name):
    """(scene, camera, bg, extras) for a golden rasterizer case; extras holds colors_precomp / cov3D_precomp and the
    upstream gradients."""
    import math
    from lgdwt_b200 import scenes
    P, seed, spacing, W, H, aa, mode = RASTER_CASES[name]
    rng = np.random.default_rng(seed)
    base = scenes.blender_init_scene(P, seed=seed, spacing_scale=spacing)
    # anisotropic scales + random rotations so that no gradient is analytically zero
    scales = (base.scales * np.exp(rng.normal(0, 0.4, (P, 3)))).astype(np.float32)
    q = rng.standard_normal((P, 4))
    rots = (q / np.linalg.norm(q, axis=1, keepdims=True)).astype(np.float32)
    opac = (1.0 / (1.0 + np.exp(-rng.normal(0.0, 1.5, (P, 1))))).astype(np.float32)
    sc = scenes.Scene(base.means3D, scales, rots, opac, base.shs, 3)
    cam = scenes.look_at_camera(W, H, 0.6911, 0.6911 * H / W, (0.4, -0.3, -4.03))
    bg = np.array([0.1, 0.4, 0.7], np.float32)
    ex = {"aa": aa, "mode": mode,
          "dL_dpix": rng.standard_normal((3, H, W)).astype(np.float32),
          "dL_dinvd": rng.standard_normal((1, H, W)).astype(np.float32)}
    if mode == "precomp":
        ex["colors_precomp"] = rng.random((P, 3)).astype(np.float32)
        r, x, y, z = rots[:, 0], rots[:, 1], rots[:, 2], rots[:, 3]
        Rm = np.stack([1 - 2 * (y * y + z * z), 2 * (x * y - r * z), 2 * (x * z + r * y),
                       2 * (x * y + r * z), 1 - 2 * (x * x + z * z), 2 * (y * z - r * x),
                       2 * (x * z - r * y), 2 * (y * z + r * x), 1 - 2 * (x * x + y * y)], 1).reshape(-1, 3, 3)
        Lm = Rm * scales[:, None, :]
        Sg = Lm @ Lm.transpose(0, 2, 1)
        ex["cov3D_precomp"] = np.stack([Sg[:, 0, 0], Sg[:, 0, 1], Sg[:, 0, 2], Sg[:, 1, 1], Sg[:, 1, 2], Sg[:, 2, 2]],
                                       1).astype(np.float32)
    return sc, cam, bg, ex


def knn_points(P=6000, seed=5):
    rng = np.random.default_rng(seed)
    pts = rng.normal(0, 1.0, (P, 3)).astype(np.float32)
    pts[: P // 50] = pts[P // 50: 2 * (P // 50)]  # exact duplicates -> zero distances, like COLMAP clouds have
    return pts


DENSIFY_CASES = {
    # name: (P, seed, extent, max_grad, min_opacity, max_screen_size or 0 for None, log-scale centre)
    "mixed": (1200, 21, 4.0, 0.0002, 0.005, 0, np.log(0.04)),
    "size_threshold": (1000, 22, 4.0, 0.0002, 0.005, 20, np.log(0.12)),
    "nothing_selected": (600, 23, 4.0, 1.0e3, 0.005, 0, np.log(0.04)),
    "identity": (300, 24, 4.0, 1.0e3, 0.0, 0, np.log(0.04)),
    "all_split": (400, 25, 1.0, 0.0, 0.005, 0, np.log(0.2)),
}
PERCENT_DENSE = 0.01  # LG/arguments/__init__.py:91


def densify_case(name):
    """Raw Gaussian parameters + Adam moments + densification statistics for one densify_and_prune call
    (LG/scene/gaussian_model.py:456-476).  Everything in the reference's own tensor shapes."""
    P, seed, extent, max_grad, min_opacity, mss, ls = DENSIFY_CASES[name]
    rng = np.random.default_rng(seed)
    f32 = lambda a: np.ascontiguousarray(a, dtype=np.float32)
    d = {"xyz": f32(rng.normal(0, 1.0, (P, 3))), "f_dc": f32(rng.normal(0, 1.0, (P, 1, 3))),
         "f_rest": f32(rng.normal(0, 0.15, (P, 15, 3))), "opacity": f32(rng.normal(-2.0, 3.0, (P, 1))),
         "scaling": f32(rng.normal(ls, 0.7, (P, 3))), "rotation": f32(rng.normal(0, 1.0, (P, 4)))}
    for k in list(d):
        d["exp_avg/" + k] = f32(rng.normal(0, 1e-3, d[k].shape))
        d["exp_avg_sq/" + k] = f32(rng.random(d[k].shape) * 1e-6)
    denom = rng.integers(0, 4, (P, 1)).astype(np.float32)  # zeros give 0/0 = nan -> 0 (gaussian_model.py:458)
    accum = f32(rng.random((P, 1)) * 0.0012 * np.maximum(denom, 1.0))
    accum[denom == 0] = 0.0
    d["xyz_gradient_accum"], d["denom"] = accum, denom
    d["max_radii2D"] = f32(rng.integers(0, 60, P))
    d["radii"] = rng.integers(0, 40, P).astype(np.int32)
    cfg = dict(extent=extent, max_grad=max_grad, min_opacity=min_opacity, max_screen_size=(mss or None),
               percent_dense=PERCENT_DENSE, seed=seed)
    return d, cfg


def stats_case(P=1000, seed=31):
    """add_densification_stats + max_radii2D update inputs (LG/train.py:268-269, gaussian_model.py:478-480)"""
    rng = np.random.default_rng(seed)
    radii = (rng.integers(0, 50, P) * (rng.random(P) < 0.6)).astype(np.int32)
    return {"grad2D": rng.normal(0, 1e-3, (P, 3)).astype(np.float32), "radii": radii,
            "xyz_gradient_accum": (rng.random((P, 1)) * 1e-2).astype(np.float32),
            "denom": rng.integers(0, 9, (P, 1)).astype(np.float32),
            "max_radii2D": rng.integers(0, 50, P).astype(np.float32),
            "opacity": rng.normal(-2.0, 3.0, (P, 1)).astype(np.float32)}
