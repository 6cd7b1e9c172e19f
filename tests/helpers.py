"""Shared test plumbing: run this repo's operator and (when oracle/_ref/libref_dgr.so was prebuilt) the unmodified
reference CUDA rasterizer on the same tensors, and read both sides' internal state for bit-exact comparison."""
import ctypes
import os

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

STATE_DTYPES = {
    "depths": (torch.float32, lambda P, C, N, T, R: P),
    "means2D": (torch.float32, lambda P, C, N, T, R: 2 * P),
    "cov3D": (torch.float32, lambda P, C, N, T, R: 6 * P),
    "conic_opacity": (torch.float32, lambda P, C, N, T, R: 4 * P),
    "rgb": (torch.float32, lambda P, C, N, T, R: C * P),
    "clamped": (torch.uint8, lambda P, C, N, T, R: 3 * P),
    "tiles_touched": (torch.int32, lambda P, C, N, T, R: P),
    "point_offsets": (torch.int32, lambda P, C, N, T, R: P),
    "final_T": (torch.float32, lambda P, C, N, T, R: N),
    "n_contrib": (torch.int32, lambda P, C, N, T, R: N),
    "ranges": (torch.int32, lambda P, C, N, T, R: 2 * T),
    "point_list": (torch.int32, lambda P, C, N, T, R: R),
    "point_list_keys": (torch.int64, lambda P, C, N, T, R: R),
}


def scene_to_torch(scene, device="cuda"):
    return {k: torch.from_numpy(np.ascontiguousarray(getattr(scene, k))).to(device)
            for k in ("means3D", "scales", "rotations", "opacities", "shs")}


def cam_to_torch(cam, device="cuda"):
    return {"viewmatrix": torch.from_numpy(cam.viewmatrix).to(device),
            "projmatrix": torch.from_numpy(cam.projmatrix).to(device),
            "campos": torch.from_numpy(cam.campos).to(device)}


def _ptr(t):
    return None if t is None or t.numel() == 0 else t.data_ptr()


class _Buf:
    def __init__(self, fn_type, device):
        self.tensor = torch.empty(0, dtype=torch.uint8, device=device)
        self.device = device
        self.cb = fn_type(self._alloc)

    def _alloc(self, _ctx, n):
        self.tensor = torch.empty(int(n), dtype=torch.uint8, device=self.device)
        return self.tensor.data_ptr()


def run_ours(t, c, cam, bg, sh_degree=3, colors_precomp=None, cov3D_precomp=None, antialiasing=False,
             scale_modifier=1.0, debug=False, want_state=True, with_invdepth=True, prefiltered=False, capacity_hint=0):
    """Low-level call through the C ABI (no autograd).  Returns dict of outputs, state tensors and raw buffers.
    capacity_hint > 0 takes the speculative entry point (lg_rasterize_forward_hinted)."""
    from lgdwt_b200 import _lib
    dev = t["means3D"].device
    P = t["means3D"].shape[0]
    H, W = cam.image_height, cam.image_width
    C = 3 if colors_precomp is None else colors_precomp.shape[-1]
    color = torch.zeros((C, H, W), device=dev)
    invd = torch.zeros((1, H, W), device=dev)
    radii = torch.zeros((P,), dtype=torch.int32, device=dev)
    geom, binning, img = (_lib.ResizableBuffer(dev) for _ in range(3))
    R = ctypes.c_int(0)
    shs = None if colors_precomp is not None else t["shs"]
    M = 0 if shs is None else shs.shape[1]
    scales = None if cov3D_precomp is not None else t["scales"]
    rots = None if cov3D_precomp is not None else t["rotations"]
    cap = ctypes.c_int(0)
    rc = _lib.lib.lg_rasterize_forward_hinted(
        geom.callback, None, binning.callback, None, img.callback, None, P, sh_degree, M, C, _ptr(bg), W, H,
        _ptr(t["means3D"]), _ptr(shs), _ptr(colors_precomp), _ptr(t["opacities"]), _ptr(scales), scale_modifier,
        _ptr(rots), _ptr(cov3D_precomp), _ptr(c["viewmatrix"]), _ptr(c["projmatrix"]), _ptr(c["campos"]),
        cam.tanfovx, cam.tanfovy, int(prefiltered), _ptr(color), _ptr(invd) if with_invdepth else None, int(antialiasing),
        _ptr(radii), int(debug), _lib.stream_ptr(dev), int(capacity_hint), ctypes.byref(R), ctypes.byref(cap))
    _lib.check(rc)
    out = {"color": color, "invdepth": invd, "radii": radii, "num_rendered": R.value, "binning_capacity": cap.value,
           "geom": geom.tensor,
           "binning": binning.tensor, "img": img.tensor, "C": C, "M": M}
    if want_state:
        N, T = W * H, ((W + 15) // 16) * ((H + 15) // 16)
        for name, (dt, nfn) in STATE_DTYPES.items():
            if name == "rgb" and colors_precomp is not None:
                continue
            n = nfn(P, C, N, T, R.value)
            dst = torch.zeros(max(nfn(P, C, N, T, cap.value), 1), dtype=dt, device=dev)  # laid out for the capacity
            rc = _lib.lib.lg_state_read(name.encode(), P, C, W, H, cap.value, _ptr(geom.tensor), _ptr(binning.tensor),
                                        _ptr(img.tensor), dst.data_ptr(), dst.numel() * dst.element_size(),
                                        _lib.stream_ptr(dev))
            _lib.check(rc)
            out[name] = dst[:n]
    torch.cuda.synchronize()
    return out


def backward_ours(t, c, cam, bg, fwd, dL_dpix, dL_dinvd=None, sh_degree=3, colors_precomp=None, cov3D_precomp=None,
                  antialiasing=False, scale_modifier=1.0, debug=False):
    from lgdwt_b200 import _lib
    dev = t["means3D"].device
    P = t["means3D"].shape[0]
    H, W = cam.image_height, cam.image_width
    C, M = fwd["C"], fwd["M"]
    shs = None if colors_precomp is not None else t["shs"]
    scales = None if cov3D_precomp is not None else t["scales"]
    rots = None if cov3D_precomp is not None else t["rotations"]
    new = lambda *s: torch.full(s, float("nan"), device=dev)  # NaN-filled: every element must be written
    g = {"dL_dmean2D": new(P, 3), "dL_dconic": new(P, 4), "dL_dopacity": new(P, 1), "dL_dcolor": new(P, C),
         "dL_dinvdepth": new(P, 1), "dL_dmean3D": new(P, 3), "dL_dcov3D": new(P, 6),
         "dL_dsh": new(P, M, 3) if shs is not None else None, "dL_dscale": new(P, 3) if scales is not None else None,
         "dL_drot": new(P, 4) if scales is not None else None}
    rc = _lib.lib.lg_rasterize_backward(
        P, sh_degree, M, fwd["binning_capacity"], C, _ptr(bg), W, H, _ptr(t["means3D"]), _ptr(shs), _ptr(colors_precomp),
        _ptr(t["opacities"]), _ptr(scales), scale_modifier, _ptr(rots), _ptr(cov3D_precomp), _ptr(c["viewmatrix"]),
        _ptr(c["projmatrix"]), _ptr(c["campos"]), cam.tanfovx, cam.tanfovy, _ptr(fwd["radii"]), _ptr(fwd["geom"]),
        _ptr(fwd["binning"]), _ptr(fwd["img"]), _ptr(dL_dpix), _ptr(dL_dinvd), _ptr(g["dL_dmean2D"]),
        _ptr(g["dL_dconic"]), _ptr(g["dL_dopacity"]), _ptr(g["dL_dcolor"]),
        _ptr(g["dL_dinvdepth"]) if dL_dinvd is not None else None, _ptr(g["dL_dmean3D"]), _ptr(g["dL_dcov3D"]),
        _ptr(g["dL_dsh"]), _ptr(g["dL_dscale"]), _ptr(g["dL_drot"]), int(antialiasing), int(debug),
        _lib.stream_ptr(dev))
    _lib.check(rc)
    torch.cuda.synchronize()
    if dL_dinvd is None:
        g["dL_dinvdepth"] = None
    return g


from oracle.ref_cuda import REF_LIB, backward_ref, load_ref, run_ref  # noqa: E402,F401  (reference side)


def rel_err(a, b):
    """max |a-b| / max(|b|_inf, eps): the per-tensor relative error of SURVEY §8c."""
    a, b = a.double(), b.double()
    return float((a - b).abs().max() / max(float(b.abs().max()), 1e-12))


# ---------------------------------------------------------------- gradient bars (tests/test_configs_gpu.py docstring)
TENSOR_RTOL = 2e-4
ELEM_RTOL, ELEM_ATOL_FRAC, ELEM_QUANTILE = 1e-3, 2e-6, 0.999
HARD_RTOL, HARD_ATOL_FRAC = 1e-2, 2e-4


def grad_stats(mine, ref):
    """the three gradient measures of this file, as numbers (shared with test_rasterizer_vs_reference_gpu.py)"""
    mine, ref = mine.double().reshape(-1), ref.double().reshape(-1)
    scale = max(float(ref.abs().max()), 1e-30)
    d = (mine - ref).abs()
    soft = d <= ELEM_RTOL * ref.abs() + ELEM_ATOL_FRAC * scale
    hard = d <= HARD_RTOL * ref.abs() + HARD_ATOL_FRAC * scale
    return {"tensor_rel": float(d.max()) / scale, "elem_ok_frac": float(soft.double().mean()),
            "hard_violations": int((~hard).sum()), "scale": scale}


def assert_grad_close(name, mine, ref, floor=0.0):
    s = grad_stats(mine, ref)
    if s["scale"] <= floor:      # analytically zero tensor: both sides hold rounding noise
        assert float((mine.double() - ref.double()).abs().max()) <= TENSOR_RTOL * floor, name
        return s
    assert s["tensor_rel"] <= TENSOR_RTOL, "%s: per-tensor relative error %.3g" % (name, s["tensor_rel"])
    assert s["elem_ok_frac"] >= ELEM_QUANTILE, "%s: only %.5f of the elements within rtol %.0e" % (
        name, s["elem_ok_frac"], ELEM_RTOL)
    assert s["hard_violations"] == 0, "%s: %d elements off by more than 1 %%" % (name, s["hard_violations"])
    return s



# ---------------------------------------------------------------- flat Gaussian buffer <-> reference-shaped groups
DENSIFY_GROUPS = ("xyz", "f_dc", "f_rest", "opacity", "scaling", "rotation")


def flat_from_groups(dp, d, device, prefix=""):
    """dp.FlatGaussians filled from reference-shaped arrays (tests/golden_inputs.densify_case)"""
    P = d[prefix + "xyz"].shape[0]
    g = dp.FlatGaussians(P, torch.device(device))
    for buf, pre in ((g.data, ""), (g.exp_avg, "exp_avg/"), (g.exp_avg_sq, "exp_avg_sq/")):
        for k in DENSIFY_GROUPS:
            dst = g.field(k, buf)
            dst.copy_(torch.as_tensor(d[prefix + pre + k]).reshape(P, dst.shape[1]))
    return g


def groups_from_flat(g, buf=None):
    shapes = {"xyz": (3,), "f_dc": (1, 3), "f_rest": (g.M - 1, 3), "opacity": (1,), "scaling": (3,), "rotation": (4,)}
    return {k: g.field(k, buf).detach().cpu().reshape(g.P, *shapes[k]).clone() for k in DENSIFY_GROUPS}


def oracle_densify_fn(g, stats, cfg, max_screen_size, generator):
    """densify_fn for dp.ViewParallelTrainer on CPU tensors, backed by the oracle (tests only)"""
    from oracle import densify_oracle
    p, m, v = groups_from_flat(g), groups_from_flat(g, g.exp_avg), groups_from_flat(g, g.exp_avg_sq)
    _, _, _, parents = densify_oracle.plan(p, stats.xyz_gradient_accum, stats.denom, cfg.densify_grad_threshold,
                                           cfg.min_opacity, cfg.cameras_extent, max_screen_size, cfg.percent_dense)
    eps = torch.randn((2 * parents.numel(), 3), generator=generator)
    np_, nm, nv, counts = densify_oracle.densify_and_prune(p, m, v, stats.xyz_gradient_accum, stats.denom, eps,
                                                           cfg.densify_grad_threshold, cfg.min_opacity,
                                                           cfg.cameras_extent, max_screen_size, cfg.percent_dense)
    P_new = counts["P"]
    bufs = [torch.zeros(g.floats * g.padded(P_new)) for _ in range(3)]
    g.replace(P_new, *bufs)
    for buf, src in ((g.data, np_), (g.exp_avg, nm), (g.exp_avg_sq, nv)):
        for k in DENSIFY_GROUPS:
            g.field(k, buf).copy_(src[k].reshape(P_new, g.field(k, buf).shape[1]))
    return counts


def oracle_stats_fn(stats, grad2D, radii):
    from oracle import densify_oracle
    a, d, m = densify_oracle.add_densification_stats(stats.xyz_gradient_accum, stats.denom, stats.max_radii2D, grad2D,
                                                     radii)
    stats.xyz_gradient_accum, stats.denom, stats.max_radii2D = a, d, m


def oracle_reset_opacity_fn(g):
    from oracle import densify_oracle
    g.slab("opacity").copy_(densify_oracle.reset_opacity(g.slab("opacity")))
    g.slab("opacity", g.exp_avg).zero_()
    g.slab("opacity", g.exp_avg_sq).zero_()
