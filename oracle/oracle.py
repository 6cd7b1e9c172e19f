"""TEST INFRASTRUCTURE ONLY — numpy/ctypes front-end of the CPU oracle (oracle/_build/liboracle.so).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference-fallback legs may import this module;
the product package (sparse-view-3dgs-pack_b200/) never does.
"""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "_build", "liboracle.so")
_lib = None


def build():
    subprocess.run(["make", "-C", HERE, "_build/liboracle.so"], check=True, capture_output=True)


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB):
            build()
        _lib = ctypes.CDLL(LIB)
        _lib.oracle_preprocess.restype = ctypes.c_int
        _lib.oracle_higher_msb.restype = ctypes.c_uint32
        _lib.oracle_higher_msb.argtypes = [ctypes.c_uint32]
        _lib.oracle_num_threads.restype = ctypes.c_int
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def _f32(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float32)


def num_threads():
    return lib().oracle_num_threads()


def higher_msb(n):
    return int(lib().oracle_higher_msb(int(n)))


def rasterize_forward(means3D, opacities, viewmatrix, projmatrix, campos, tanfovx, tanfovy, W, H, bg, shs=None,
                      sh_degree=3, colors_precomp=None, scales=None, rotations=None, cov3D_precomp=None,
                      scale_modifier=1.0, antialiasing=False, stop_after=None):
    """CPU restatement of CudaRasterizer::Rasterizer::forward.  Returns a dict with every intermediate the reference
    keeps in its opaque buffers.  stop_after in {None, 'preprocess', 'binning'}."""
    L = lib()
    means3D, opacities = _f32(means3D), _f32(opacities)
    shs, colors_precomp, scales, rotations, cov3D_precomp = map(_f32, (shs, colors_precomp, scales, rotations, cov3D_precomp))
    V, PM, cp, bg = _f32(viewmatrix), _f32(projmatrix), _f32(campos), _f32(bg)
    P = means3D.shape[0]
    C = 3 if colors_precomp is None else colors_precomp.shape[-1]
    M = 0 if shs is None else shs.shape[1]
    o = {"radii": np.zeros(P, np.int32), "means2D": np.zeros((P, 2), np.float32), "depths": np.zeros(P, np.float32),
         "cov3D": np.zeros((P, 6), np.float32), "rgb": np.zeros((P, 3), np.float32),
         "conic_opacity": np.zeros((P, 4), np.float32), "clamped": np.zeros((P, 3), np.uint8),
         "tiles_touched": np.zeros(P, np.uint32), "point_offsets": np.zeros(P, np.uint32), "C": C, "M": M}
    R = L.oracle_preprocess(P, int(sh_degree), M, W, H, _p(means3D), _p(shs), _p(colors_precomp), _p(opacities),
                            _p(scales), ctypes.c_float(scale_modifier), _p(rotations), _p(cov3D_precomp), _p(V), _p(PM),
                            _p(cp), ctypes.c_float(tanfovx), ctypes.c_float(tanfovy), int(antialiasing), _p(o["radii"]),
                            _p(o["means2D"]), _p(o["depths"]), _p(o["cov3D"]), _p(o["rgb"]), _p(o["conic_opacity"]),
                            _p(o["clamped"]), _p(o["tiles_touched"]), _p(o["point_offsets"]))
    o["num_rendered"] = R
    if cov3D_precomp is not None:
        o["cov3D"] = cov3D_precomp.copy()
    if stop_after == "preprocess":
        return o
    T = ((W + 15) // 16) * ((H + 15) // 16)
    o["point_list_keys"] = np.zeros(R, np.uint64)
    o["point_list"] = np.zeros(R, np.uint32)
    o["ranges"] = np.zeros((T, 2), np.uint32)
    L.oracle_binning(P, R, W, H, _p(o["radii"]), _p(o["means2D"]), _p(o["depths"]), _p(o["point_offsets"]),
                     _p(o["point_list_keys"]), _p(o["point_list"]), _p(o["ranges"]))
    if stop_after == "binning":
        return o
    feats = colors_precomp if colors_precomp is not None else o["rgb"]
    o["features"] = feats
    o["final_T"] = np.zeros(W * H, np.float32)
    o["n_contrib"] = np.zeros(W * H, np.uint32)
    o["color"] = np.zeros((C, H, W), np.float32)
    o["invdepth"] = np.zeros((1, H, W), np.float32)
    L.oracle_render_forward(C, W, H, _p(o["ranges"]), _p(o["point_list"]), _p(o["means2D"]), _p(feats),
                            _p(o["conic_opacity"]), _p(o["depths"]), _p(bg), _p(o["final_T"]), _p(o["n_contrib"]),
                            _p(o["color"]), _p(o["invdepth"]))
    return o


def rasterize_backward(fwd, means3D, opacities, viewmatrix, projmatrix, campos, tanfovx, tanfovy, W, H, bg, dL_dpix,
                       dL_dinvdepth_pix=None, shs=None, sh_degree=3, colors_precomp=None, scales=None, rotations=None,
                       cov3D_precomp=None, scale_modifier=1.0, antialiasing=False):
    """CPU restatement of CudaRasterizer::Rasterizer::backward on the state returned by rasterize_forward."""
    L = lib()
    means3D, opacities = _f32(means3D), _f32(opacities)
    shs, colors_precomp, scales, rotations, cov3D_precomp = map(_f32, (shs, colors_precomp, scales, rotations, cov3D_precomp))
    V, PM, cp, bg = _f32(viewmatrix), _f32(projmatrix), _f32(campos), _f32(bg)
    dL_dpix, dL_dinvdepth_pix = _f32(dL_dpix), _f32(dL_dinvdepth_pix)
    P, C, M = means3D.shape[0], fwd["C"], fwd["M"]
    g = {"dL_dmean2D": np.zeros((P, 3), np.float32), "dL_dconic": np.zeros((P, 4), np.float32),
         "dL_dopacity": np.zeros((P, 1), np.float32), "dL_dcolor": np.zeros((P, C), np.float32),
         "dL_dinvdepth": np.zeros((P, 1), np.float32) if dL_dinvdepth_pix is not None else None,
         "dL_dmean3D": np.zeros((P, 3), np.float32), "dL_dcov3D": np.zeros((P, 6), np.float32),
         "dL_dsh": np.zeros((P, M, 3), np.float32) if shs is not None else None,
         "dL_dscale": np.zeros((P, 3), np.float32) if scales is not None else None,
         "dL_drot": np.zeros((P, 4), np.float32) if scales is not None else None}
    L.oracle_render_backward(P, C, W, H, _p(fwd["ranges"]), _p(fwd["point_list"]), _p(bg), _p(fwd["means2D"]),
                             _p(fwd["conic_opacity"]), _p(fwd["features"]), _p(fwd["depths"]), _p(fwd["final_T"]),
                             _p(fwd["n_contrib"]), _p(dL_dpix), _p(dL_dinvdepth_pix), _p(g["dL_dmean2D"]),
                             _p(g["dL_dconic"]), _p(g["dL_dopacity"]), _p(g["dL_dcolor"]), _p(g["dL_dinvdepth"]))
    cov = cov3D_precomp if cov3D_precomp is not None else fwd["cov3D"]
    L.oracle_preprocess_backward(P, int(sh_degree), M, W, H, _p(means3D), _p(fwd["radii"]), _p(shs), _p(fwd["clamped"]),
                                 _p(opacities), _p(scales), _p(rotations), ctypes.c_float(scale_modifier), _p(cov), _p(V),
                                 _p(PM), _p(cp), ctypes.c_float(tanfovx), ctypes.c_float(tanfovy), int(antialiasing),
                                 _p(g["dL_dmean2D"]), _p(g["dL_dconic"]), _p(g["dL_dopacity"]), _p(g["dL_dcolor"]),
                                 _p(g["dL_dinvdepth"]), _p(g["dL_dmean3D"]), _p(g["dL_dcov3D"]), _p(g["dL_dsh"]),
                                 _p(g["dL_dscale"]), _p(g["dL_drot"]))
    return g


def mark_visible(means3D, viewmatrix):
    means3D, V = _f32(means3D), _f32(viewmatrix)
    out = np.zeros(means3D.shape[0], np.uint8)
    lib().oracle_mark_visible(means3D.shape[0], _p(means3D), _p(V), _p(out))
    return out.astype(bool)


def knn_mean_dist2(points, brute_force=False):
    pts = _f32(points)
    out = np.zeros(pts.shape[0], np.float32)
    fn = lib().oracle_knn_bruteforce if brute_force else lib().oracle_knn_mean_dist2
    fn(pts.shape[0], _p(pts), _p(out))
    return out


def scene_kwargs(scene, cam, bg):
    """keyword arguments for rasterize_forward/backward from lgdwt_b200.scenes objects"""
    return dict(means3D=scene.means3D, opacities=scene.opacities, viewmatrix=cam.viewmatrix, projmatrix=cam.projmatrix,
                campos=cam.campos, tanfovx=cam.tanfovx, tanfovy=cam.tanfovy, W=cam.image_width, H=cam.image_height,
                bg=np.asarray(bg, np.float32), shs=scene.shs, sh_degree=scene.sh_degree, scales=scene.scales,
                rotations=scene.rotations)
