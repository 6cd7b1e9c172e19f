"""TEST INFRASTRUCTURE ONLY — ctypes driver of oracle/_ref/libref_dgr.so, the UNMODIFIED reference CUDA rasterizer
and simple-knn compiled from the reference's own sources by oracle/build_ref.sh.  Used by tests/ (parity by
execution on the same GPU) and by bench.py's reference arm; never imported by the product package.

`RefRasterize` mirrors the reference's Python autograd glue (DGR/dgr_3dgs/__init__.py:44-141) and torch C++ glue
(DGR/rasterize_points.cu:35-223) allocation for allocation — zero-filled outputs, nine zero-filled gradient
tensors, materialised inverse-depth gradients — so that the reference arm pays what the reference really pays.
"""
import ctypes
import os

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF_LIB = os.path.join(HERE, "_ref", "libref_dgr.so")

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


def _ptr(t):
    return None if t is None or t.numel() == 0 else t.data_ptr()


class _Buf:
    """resizable uint8 buffer + ctypes callback (closure over a list: no reference cycle, see lgdwt_b200/_lib.py)"""

    def __init__(self, fn_type, device):
        holder = [torch.empty(0, dtype=torch.uint8, device=device)]

        def _alloc(_ctx, n, _holder=holder, _device=device):
            _holder[0] = torch.empty(int(n), dtype=torch.uint8, device=_device)
            return _holder[0].data_ptr()

        self._holder = holder
        self.cb = fn_type(_alloc)

    @property
    def tensor(self):
        return self._holder[0]


_REF = None
REF_ALLOC = ctypes.CFUNCTYPE(ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t)


def load_ref():
    """ctypes handle of the prebuilt reference shim, or None when oracle/_ref was not built (no /root/reference)."""
    global _REF
    if _REF is None and os.path.exists(REF_LIB):
        lib = ctypes.CDLL(REF_LIB)
        lib.ref_last_error.restype = ctypes.c_char_p
        i, f, p = ctypes.c_int, ctypes.c_float, ctypes.c_void_p
        lib.ref_rasterize_forward.restype = i
        lib.ref_rasterize_forward.argtypes = ([REF_ALLOC, p] * 3 + [i, i, i, p, i, i] + [p] * 5 + [f] + [p] * 5 +
                                              [f, f, i, p, p, i, p, i, ctypes.POINTER(i)])
        lib.ref_rasterize_backward.restype = i
        lib.ref_rasterize_backward.argtypes = ([i, i, i, i, p, i, i] + [p] * 5 + [f] + [p] * 5 + [f, f] + [p] * 16 +
                                               [i, i])
        lib.ref_state_read.restype = i
        lib.ref_state_read.argtypes = [ctypes.c_char_p, i, i, i, i, p, p, p, p, ctypes.c_size_t]
        lib.ref_mark_visible.restype = i
        lib.ref_mark_visible.argtypes = [i, p, p, p, p]
        lib.ref_knn_mean_dist2.restype = i
        lib.ref_knn_mean_dist2.argtypes = [i, p, p]
        _REF = lib
    return _REF


def _ref_check(lib, rc):
    if rc != 0:
        raise RuntimeError("reference shim: " + lib.ref_last_error().decode())


def run_ref(t, c, cam, bg, sh_degree=3, colors_precomp=None, cov3D_precomp=None, antialiasing=False,
            scale_modifier=1.0, debug=False, want_state=True, prefiltered=False):
    lib = load_ref()
    dev = t["means3D"].device
    P = t["means3D"].shape[0]
    H, W = cam.image_height, cam.image_width
    color = torch.zeros((3, H, W), device=dev)
    invd = torch.zeros((1, H, W), device=dev)
    radii = torch.zeros((P,), dtype=torch.int32, device=dev)
    geom, binning, img = (_Buf(REF_ALLOC, dev) for _ in range(3))
    R = ctypes.c_int(0)
    shs = None if colors_precomp is not None else t["shs"]
    M = 0 if shs is None else shs.shape[1]
    scales = None if cov3D_precomp is not None else t["scales"]
    rots = None if cov3D_precomp is not None else t["rotations"]
    torch.cuda.synchronize()
    rc = lib.ref_rasterize_forward(
        geom.cb, None, binning.cb, None, img.cb, None, P, sh_degree, M, _ptr(bg), W, H, _ptr(t["means3D"]), _ptr(shs),
        _ptr(colors_precomp), _ptr(t["opacities"]), _ptr(scales), scale_modifier, _ptr(rots), _ptr(cov3D_precomp),
        _ptr(c["viewmatrix"]), _ptr(c["projmatrix"]), _ptr(c["campos"]), cam.tanfovx, cam.tanfovy, int(prefiltered),
        _ptr(color), _ptr(invd), int(antialiasing), _ptr(radii), int(debug), ctypes.byref(R))
    _ref_check(lib, rc)
    torch.cuda.synchronize()
    out = {"color": color, "invdepth": invd, "radii": radii, "num_rendered": R.value, "geom": geom.tensor,
           "binning": binning.tensor, "img": img.tensor, "C": 3, "M": M}
    if want_state:
        N, T = W * H, ((W + 15) // 16) * ((H + 15) // 16)
        for name, (dt, nfn) in STATE_DTYPES.items():
            if name == "rgb" and colors_precomp is not None:
                continue
            n = nfn(P, 3, N, T, R.value)
            dst = torch.zeros(max(n, 1), dtype=dt, device=dev)
            rc = lib.ref_state_read(name.encode(), P, W, H, R.value, _ptr(geom.tensor), _ptr(binning.tensor),
                                    _ptr(img.tensor), dst.data_ptr(), dst.numel() * dst.element_size())
            _ref_check(lib, rc)
            out[name] = dst[:n]
    torch.cuda.synchronize()
    return out


def backward_ref(t, c, cam, bg, fwd, dL_dpix, dL_dinvd=None, sh_degree=3, colors_precomp=None, cov3D_precomp=None,
                 antialiasing=False, scale_modifier=1.0, debug=False):
    lib = load_ref()
    dev = t["means3D"].device
    P = t["means3D"].shape[0]
    H, W = cam.image_height, cam.image_width
    M = fwd["M"]
    shs = None if colors_precomp is not None else t["shs"]
    scales = None if cov3D_precomp is not None else t["scales"]
    rots = None if cov3D_precomp is not None else t["rotations"]
    z = lambda *s: torch.zeros(s, device=dev)  # the reference requires zero-initialised gradient buffers
    g = {"dL_dmean2D": z(P, 3), "dL_dconic": z(P, 4), "dL_dopacity": z(P, 1), "dL_dcolor": z(P, 3),
         "dL_dinvdepth": z(P, 1), "dL_dmean3D": z(P, 3), "dL_dcov3D": z(P, 6), "dL_dsh": z(P, max(M, 1), 3),
         "dL_dscale": z(P, 3), "dL_drot": z(P, 4)}
    torch.cuda.synchronize()
    rc = lib.ref_rasterize_backward(
        P, sh_degree, M, fwd["num_rendered"], _ptr(bg), W, H, _ptr(t["means3D"]), _ptr(shs), _ptr(colors_precomp),
        _ptr(t["opacities"]), _ptr(scales), scale_modifier, _ptr(rots), _ptr(cov3D_precomp), _ptr(c["viewmatrix"]),
        _ptr(c["projmatrix"]), _ptr(c["campos"]), cam.tanfovx, cam.tanfovy, _ptr(fwd["radii"]), _ptr(fwd["geom"]),
        _ptr(fwd["binning"]), _ptr(fwd["img"]), _ptr(dL_dpix), _ptr(dL_dinvd), _ptr(g["dL_dmean2D"]),
        _ptr(g["dL_dconic"]), _ptr(g["dL_dopacity"]), _ptr(g["dL_dcolor"]),
        _ptr(g["dL_dinvdepth"]) if dL_dinvd is not None else None, _ptr(g["dL_dmean3D"]), _ptr(g["dL_dcov3D"]),
        _ptr(g["dL_dsh"]), _ptr(g["dL_dscale"]), _ptr(g["dL_drot"]), int(antialiasing), int(debug))
    _ref_check(lib, rc)
    torch.cuda.synchronize()
    if shs is None:
        g["dL_dsh"] = None
    if dL_dinvd is None:
        g["dL_dinvdepth"] = None
    return g




class RefRasterize(torch.autograd.Function):
    """reference `_RasterizeGaussians` (DGR/dgr_3dgs/__init__.py:44-141) on top of the shim"""

    @staticmethod
    def forward(ctx, means3D, means2D, sh, opacities, scales, rotations, cam, bg, sh_degree):
        lib = load_ref()
        dev = means3D.device
        P, H, W = means3D.shape[0], cam["H"], cam["W"]
        color = torch.full((3, H, W), 0.0, device=dev)              # rasterize_points.cu:69
        invd = torch.full((1, H, W), 0.0, device=dev)               # :73
        radii = torch.full((P,), 0, dtype=torch.int32, device=dev)  # :76
        geom, binning, img = (_Buf(REF_ALLOC, dev) for _ in range(3))
        R = ctypes.c_int(0)
        rc = lib.ref_rasterize_forward(
            geom.cb, None, binning.cb, None, img.cb, None, P, sh_degree, sh.shape[1], _ptr(bg), W, H, _ptr(means3D),
            _ptr(sh), None, _ptr(opacities), _ptr(scales), 1.0, _ptr(rotations), None, _ptr(cam["viewmatrix"]),
            _ptr(cam["projmatrix"]), _ptr(cam["campos"]), cam["tanfovx"], cam["tanfovy"], 0, _ptr(color), _ptr(invd), 0,
            _ptr(radii), 0, ctypes.byref(R))
        _ref_check(lib, rc)
        ctx.cam, ctx.bg, ctx.R, ctx.sh_degree = cam, bg, R.value, sh_degree
        ctx.save_for_backward(means3D, scales, rotations, radii, sh, opacities, geom.tensor, binning.tensor, img.tensor)
        return color, radii, invd

    @staticmethod
    def backward(ctx, g_color, _g_radii, g_invd):
        lib = load_ref()
        means3D, scales, rotations, radii, sh, opacities, geom, binning, img = ctx.saved_tensors
        cam, bg = ctx.cam, ctx.bg
        dev = means3D.device
        P, H, W, M = means3D.shape[0], cam["H"], cam["W"], sh.shape[1]
        z = lambda *s: torch.zeros(s, device=dev)                   # rasterize_points.cu:163-178
        d3, d2, dc, dcon, dop = z(P, 3), z(P, 3), z(P, 3), z(P, 2, 2), z(P, 1)
        dcov, dsh, dsc, drot, dinv = z(P, 6), z(P, M, 3), z(P, 3), z(P, 4), z(P, 1)
        rc = lib.ref_rasterize_backward(
            P, ctx.sh_degree, M, ctx.R, _ptr(bg), W, H, _ptr(means3D), _ptr(sh), None, _ptr(opacities), _ptr(scales), 1.0,
            _ptr(rotations), None, _ptr(cam["viewmatrix"]), _ptr(cam["projmatrix"]), _ptr(cam["campos"]), cam["tanfovx"],
            cam["tanfovy"], _ptr(radii), _ptr(geom), _ptr(binning), _ptr(img), _ptr(g_color.contiguous()),
            _ptr(g_invd.contiguous()), _ptr(d2), _ptr(dcon), _ptr(dop), _ptr(dc), _ptr(dinv), _ptr(d3), _ptr(dcov),
            _ptr(dsh), _ptr(dsc), _ptr(drot), 0, 0)
        _ref_check(lib, rc)
        return d3, d2, dsh, dop, dsc, drot, None, None, None
