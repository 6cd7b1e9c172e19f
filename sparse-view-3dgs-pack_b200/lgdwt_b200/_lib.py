"""ctypes binding of lib/liblgdwt_b200.so (the C ABI declared in include/lgdwt_b200.h).

There is deliberately no fallback: if the shared library is missing or a call fails, importing / calling raises.
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_PKG_ROOT = os.path.dirname(_HERE)
LIB_PATH = os.path.join(_PKG_ROOT, "lib", os.environ.get("LGDWT_LIBNAME", "liblgdwt_b200.so"))

LG_OK = 0
LG_ERR_INVALID_ARGUMENT = 1
LG_ERR_CUDA = 2
LG_ERR_ALLOC = 3
LG_ERR_UNSUPPORTED = 4

ALLOC_FN = ctypes.CFUNCTYPE(ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t)

_c_float_p = ctypes.c_void_p  # raw device pointers travel as integers
_P = ctypes.c_void_p


class LgdwtError(RuntimeError):
    pass


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "liblgdwt_b200.so not found at %s — build it with `python sparse-view-3dgs-pack_b200/build.py` "
            "(there is no CPU or PyTorch fallback for this operator)" % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    lib.lg_abi_version.restype = ctypes.c_int
    lib.lg_last_error.restype = ctypes.c_char_p
    for name in ("lg_geometry_state_bytes", "lg_image_state_bytes", "lg_binning_state_bytes", "lg_knn_workspace_bytes",
                 "lg_dwt_workspace_bytes", "lg_photometric_workspace_bytes"):
        getattr(lib, name).restype = ctypes.c_size_t
    lib.lg_geometry_state_bytes.argtypes = [ctypes.c_int, ctypes.c_int]
    lib.lg_image_state_bytes.argtypes = [ctypes.c_int, ctypes.c_int]
    lib.lg_binning_state_bytes.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int]
    lib.lg_knn_workspace_bytes.argtypes = [ctypes.c_int]
    lib.lg_dwt_workspace_bytes.argtypes = [ctypes.c_int] * 4

    i, f = ctypes.c_int, ctypes.c_float
    lib.lg_rasterize_forward.restype = i
    lib.lg_rasterize_forward.argtypes = (
        [ALLOC_FN, _P, ALLOC_FN, _P, ALLOC_FN, _P, i, i, i, i, _P, i, i] + [_P] * 5 + [f, _P, _P, _P, _P, _P, f, f, i,
                                                                                      _P, _P, i, _P, i, _P,
                                                                                      ctypes.POINTER(i)])
    lib.lg_rasterize_forward_hinted.restype = i
    lib.lg_rasterize_forward_hinted.argtypes = (
        [ALLOC_FN, _P, ALLOC_FN, _P, ALLOC_FN, _P, i, i, i, i, _P, i, i] + [_P] * 5 + [f, _P, _P, _P, _P, _P, f, f, i,
                                                                                      _P, _P, i, _P, i, _P, i,
                                                                                      ctypes.POINTER(i), ctypes.POINTER(i)])
    lib.lg_rasterize_backward.restype = i
    lib.lg_rasterize_backward.argtypes = (
        [i, i, i, i, i, _P, i, i] + [_P] * 5 + [f, _P, _P, _P, _P, _P, f, f] + [_P] * 16 + [i, i, _P])
    lib.lg_rasterize_backward_ex.restype = i
    lib.lg_rasterize_backward_ex.argtypes = lib.lg_rasterize_backward.argtypes + [i]
    lib.lg_rasterize_backward_raw.restype = i
    lib.lg_rasterize_backward_raw.argtypes = lib.lg_rasterize_backward_ex.argtypes + [_P]
    lib.lg_activate_forward.restype = i
    lib.lg_activate_forward.argtypes = [i] + [_P] * 8
    lib.lg_mark_visible.restype = i
    lib.lg_mark_visible.argtypes = [i, _P, _P, _P, _P, _P]
    lib.lg_state_read.restype = i
    lib.lg_state_read.argtypes = [ctypes.c_char_p, i, i, i, i, i, _P, _P, _P, _P, ctypes.c_size_t, _P]
    lib.lg_knn_mean_dist2.restype = i
    lib.lg_knn_mean_dist2.argtypes = [i, _P, _P, _P, ctypes.c_size_t, _P]
    lib.lg_dwt_loss_forward.restype = i
    lib.lg_dwt_loss_forward.argtypes = [_P, _P, i, i, i, ctypes.POINTER(f), i, ctypes.c_double, f, f, _P, _P, _P,
                                        ctypes.c_size_t, _P]
    lib.lg_dwt_loss_backward.restype = i
    lib.lg_dwt_loss_backward.argtypes = [_P, _P, i, i, i, ctypes.POINTER(f), i, f, f, _P, _P, _P, _P, _P, _P]
    lib.lg_photometric_workspace_bytes.argtypes = [i, i, i]
    lib.lg_photometric_loss_forward.restype = i
    lib.lg_photometric_loss_forward.argtypes = [_P, _P, i, i, i, _P, _P, ctypes.c_size_t, i, _P]
    lib.lg_photometric_loss_backward.restype = i
    lib.lg_photometric_loss_backward.argtypes = [_P, _P, i, i, i, _P, _P, _P, _P, _P]
    lib.lg_adam_step.restype = i
    lib.lg_adam_step.argtypes = [_P, _P, _P, _P, ctypes.c_longlong, i, ctypes.POINTER(ctypes.c_longlong),
                                 ctypes.POINTER(f), f, f, f, i, f, _P]
    lib.lg_adam_step_split.restype = i
    lib.lg_adam_step_split.argtypes = [_P, _P, _P, _P, ctypes.c_longlong, i, ctypes.POINTER(ctypes.c_longlong),
                                       ctypes.POINTER(f), ctypes.POINTER(f), ctypes.POINTER(i), ctypes.POINTER(i),
                                       f, f, f, i, f, _P]
    lib.lg_densify_scratch_bytes.restype = ctypes.c_size_t
    lib.lg_densify_scratch_bytes.argtypes = [i]
    lib.lg_densify_plan.restype = i
    lib.lg_densify_plan.argtypes = [i, _P, _P, _P, _P, f, f, f, f, f, _P, _P, _P, _P, ctypes.c_size_t, _P]
    lib.lg_densify_apply.restype = i
    lib.lg_densify_apply.argtypes = [i, i, i, i, i] + [_P] * 9 + [i, _P]
    lib.lg_densify_stats.restype = i
    lib.lg_densify_stats.argtypes = [i] + [_P] * 6
    lib.lg_reset_opacity.restype = i
    lib.lg_reset_opacity.argtypes = [i] + [_P] * 4
    lib.lg_peer_alloc.restype = i
    lib.lg_peer_alloc.argtypes = [ctypes.c_size_t, ctypes.POINTER(ctypes.c_void_p)]
    lib.lg_peer_free.restype = i
    lib.lg_peer_free.argtypes = [_P]
    lib.lg_peer_export.restype = i
    lib.lg_peer_export.argtypes = [_P, ctypes.c_char_p]
    lib.lg_peer_open.restype = i
    lib.lg_peer_open.argtypes = [ctypes.c_char_p, ctypes.POINTER(ctypes.c_void_p)]
    lib.lg_peer_close.restype = i
    lib.lg_peer_close.argtypes = [_P]
    lib.lg_peer_barrier.restype = i
    lib.lg_peer_barrier.argtypes = [i, i, ctypes.POINTER(ctypes.c_void_p), ctypes.c_uint, _P]
    lib.lg_peer_check.restype = i
    lib.lg_peer_check.argtypes = [_P]
    lib.lg_peer_reduce_adam.restype = i
    lib.lg_peer_reduce_adam.argtypes = [i, i, ctypes.POINTER(ctypes.c_void_p), ctypes.POINTER(ctypes.c_void_p), _P, _P,
                                        ctypes.c_longlong, i, ctypes.POINTER(ctypes.c_longlong), ctypes.POINTER(f),
                                        ctypes.POINTER(f), ctypes.POINTER(i), ctypes.POINTER(i), f, f, f, i, f, _P]
    lib.lg_peer_allreduce.restype = i
    lib.lg_peer_allreduce.argtypes = [i, i, ctypes.POINTER(ctypes.c_void_p), ctypes.c_longlong, f, _P]
    lib.lg_peer_reduce_adam_mc.restype = i
    lib.lg_peer_reduce_adam_mc.argtypes = [i, i, _P, _P, _P, _P, _P, ctypes.c_longlong, i,
                                           ctypes.POINTER(ctypes.c_longlong), ctypes.POINTER(f), ctypes.POINTER(f),
                                           ctypes.POINTER(i), ctypes.POINTER(i), f, f, f, i, f, _P]
    lib.lg_peer_allreduce_mc.restype = i
    lib.lg_peer_allreduce_mc.argtypes = [i, i, _P, ctypes.c_longlong, f, _P]
    lib.lg_l1_loss_forward.restype = i
    lib.lg_l1_loss_forward.argtypes = [_P, _P, ctypes.c_longlong, _P, _P, _P]
    lib.lg_l1_loss_backward.restype = i
    lib.lg_l1_loss_backward.argtypes = [_P, _P, ctypes.c_longlong, _P, _P, _P]
    lib.lg_image_loss_combine.restype = i
    lib.lg_image_loss_combine.argtypes = [_P, _P, _P, f, f, i, _P, _P, _P]
    lib.lg_image_loss_backward_coefs.restype = i
    lib.lg_image_loss_backward_coefs.argtypes = [_P, _P, _P, _P]
    lib.lg_image_loss_add.restype = i
    lib.lg_image_loss_add.argtypes = [_P, _P, ctypes.c_longlong, _P]
    lib.lg_image_loss_forward.restype = i
    lib.lg_image_loss_forward.argtypes = [_P, _P, i, i, i, ctypes.POINTER(f), i, ctypes.c_double, f, f, _P, f, f, i, _P,
                                          _P, ctypes.c_size_t, _P, ctypes.c_size_t, _P, ctypes.c_size_t, i, _P]
    lib.lg_image_loss_backward.restype = i
    lib.lg_image_loss_backward.argtypes = [_P, _P, i, i, i, ctypes.POINTER(f), i, f, f, _P, _P, _P, _P, _P, _P]
    lib.lg_photometric_loss_backward_scaled.restype = i
    lib.lg_photometric_loss_backward_scaled.argtypes = [_P, _P, i, i, i, _P, _P, _P, _P, _P, _P]
    lib.lg_dwt_loss_backward_scaled.restype = i
    lib.lg_dwt_loss_backward_scaled.argtypes = [_P, _P, i, i, i, ctypes.POINTER(f), i, f, f, _P, _P, _P, _P, _P, _P, i, _P]
    lib.lg_haar_dwt2_forward.restype = i
    lib.lg_haar_dwt2_forward.argtypes = [_P, i, i, i, _P, _P, _P]
    lib.lg_haar_dwt2_backward.restype = i
    lib.lg_haar_dwt2_backward.argtypes = [_P, _P, i, i, i, _P, _P]
    lib.lg_blend_work_count.restype = i
    lib.lg_blend_work_count.argtypes = [i, i, i, i, i, _P, _P, _P, _P, _P]
    lib.lg_launch_count.restype = ctypes.c_ulonglong
    lib.lg_stage_timing_enable.restype = i
    lib.lg_stage_timing_enable.argtypes = [i]
    lib.lg_stage_timing_sample.restype = i
    lib.lg_stage_timing_sample.argtypes = [i]
    lib.lg_stage_timing_read.restype = i
    lib.lg_stage_timing_read.argtypes = [i, ctypes.POINTER(f), i]
    lib.lg_simt_peaks.restype = i
    lib.lg_simt_peaks.argtypes = [ctypes.POINTER(f), i, _P]
    if lib.lg_abi_version() != 1:
        raise ImportError("liblgdwt_b200.so has ABI version %d, expected 1" % lib.lg_abi_version())
    return lib


lib = _load()


def check(rc, exc=LgdwtError):
    if rc != LG_OK:
        msg = lib.lg_last_error().decode("utf-8", "replace")
        raise exc(msg)


def ptr(t):
    """Device pointer of a tensor, or NULL for None / empty tensors (the reference passes torch.Tensor([]) for
    absent optional inputs and its C++ side sees a null data pointer, DGR/dgr_3dgs/__init__.py:184-194)."""
    if t is None or t.numel() == 0:
        return None
    return t.data_ptr()


def stream_ptr(device=None):
    return torch.cuda.current_stream(device).cuda_stream


STAGES = ("preprocess", "binning", "blend_forward", "blend_backward", "preprocess_backward")


def stage_timing(slots, every=1):
    """arm (slots > 0) or disarm (0) the per-stage CUDA-event ring; `every` = time one rasterizer call in that many"""
    check(lib.lg_stage_timing_sample(int(every)))
    check(lib.lg_stage_timing_enable(int(slots)))


def read_stage_times(slot):
    """dict stage -> milliseconds recorded in ring slot `slot` (-1 = stage did not run)"""
    buf = (ctypes.c_float * len(STAGES))()
    check(lib.lg_stage_timing_read(int(slot), buf, len(STAGES)))
    return {name: float(buf[k]) for k, name in enumerate(STAGES)}


SIMT_PEAKS = ("ffma_tflops", "mufu_ex2_gops", "mufu_rcp_gops", "lds128_broadcast_gwarpinst", "shfl_gwarpinst",
              "alu_gwarpinst")


def simt_peaks(device=None):
    """on-box FFMA / MUFU / LDS / SHFL / issue ceilings (csrc/microbench.cu), a few milliseconds"""
    buf = (ctypes.c_float * len(SIMT_PEAKS))()
    with torch.cuda.device(device):
        check(lib.lg_simt_peaks(buf, len(SIMT_PEAKS), stream_ptr(device)))
    return {name: float(buf[k]) for k, name in enumerate(SIMT_PEAKS)}


class ResizableBuffer:
    """A torch.uint8 CUDA tensor the library can grow through an lg_alloc_fn callback — the counterpart of the
    reference's resizeFunctional lambdas (DGR/rasterize_points.cu:27-33).

    The ctypes callback closes over a plain list, not over `self`: a bound-method callback would form a reference
    cycle (buffer -> callback -> buffer) and the hundreds of megabytes of state would then only be released by the
    cyclic garbage collector instead of by reference counting."""

    def __init__(self, device):
        holder = [torch.empty(0, dtype=torch.uint8, device=device)]

        def _alloc(_ctx, nbytes, _holder=holder, _device=device):
            try:
                _holder[0] = torch.empty(int(nbytes), dtype=torch.uint8, device=_device)
                return _holder[0].data_ptr()
            except Exception:  # surfaced as LG_ERR_ALLOC by the library
                return 0

        self._holder = holder
        self.callback = ALLOC_FN(_alloc)

    @property
    def tensor(self):
        return self._holder[0]
