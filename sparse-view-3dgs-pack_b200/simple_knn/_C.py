"""`simple_knn._C.distCUDA2` backed by the B200-native library (replaces KNN/spatial.cu:16-26)."""
import torch

from lgdwt_b200 import _lib


def distCUDA2(points):
    """Mean squared distance to the 3 nearest neighbours.  points: (P, 3) fp32 CUDA tensor -> (P,) fp32."""
    if not points.is_cuda:
        raise RuntimeError("distCUDA2 (B200-native): points must be a CUDA tensor; there is no CPU path")
    pts = points.contiguous().float()
    P = pts.size(0)
    means = torch.zeros((P,), dtype=torch.float32, device=pts.device)
    if P == 0:
        return means
    nbytes = _lib.lib.lg_knn_workspace_bytes(P)
    workspace = torch.empty(nbytes, dtype=torch.uint8, device=pts.device)
    with torch.cuda.device(pts.device):
        rc = _lib.lib.lg_knn_mean_dist2(P, pts.data_ptr(), means.data_ptr(), workspace.data_ptr(), nbytes,
                                        _lib.stream_ptr(pts.device))
    _lib.check(rc, RuntimeError)
    return means
