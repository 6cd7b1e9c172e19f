"""Generates tests/golden/rasterizer_reference.npz and knn_reference.npz by running the UNMODIFIED reference CUDA
(oracle/_ref/libref_dgr.so, compiled from the reference's own sources by oracle/build_ref.sh) on a B200:

    gpurun -- python tests/golden/make_rasterizer_golden.py      # writes gpurun_out/golden/*.npz
    cp gpurun_out/golden/*.npz tests/golden/

The fixtures pin the CPU oracle (oracle/rasterizer_oracle.c, oracle/knn_oracle.c) in the CPU-only test tier and
give the GPU tier a reference that does not need oracle/_ref at run time.  Inputs are regenerated from seeds
(tests/golden_inputs.py), only outputs are stored.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, "sparse-view-3dgs-pack_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import golden_inputs  # noqa: E402
import helpers  # noqa: E402

out_dir = os.path.join(ROOT, "gpurun_out", "golden")
os.makedirs(out_dir, exist_ok=True)
lib = helpers.load_ref()
assert lib is not None, "oracle/_ref/libref_dgr.so missing"

out = {}
for name in golden_inputs.RASTER_CASES:
    sc, cam, bg, ex = golden_inputs.raster_case(name)
    t, c = helpers.scene_to_torch(sc), helpers.cam_to_torch(cam)
    bg_t = torch.from_numpy(bg).cuda()
    kw = dict(antialiasing=ex["aa"])
    if ex["mode"] == "precomp":
        kw["colors_precomp"] = torch.from_numpy(ex["colors_precomp"]).cuda()
        kw["cov3D_precomp"] = torch.from_numpy(ex["cov3D_precomp"]).cuda()
    f = helpers.run_ref(t, c, cam, bg_t, **kw)
    for k in ("radii", "tiles_touched", "point_offsets", "point_list_keys", "point_list", "ranges", "n_contrib",
              "final_T", "color", "invdepth", "depths", "means2D", "conic_opacity", "cov3D", "clamped", "rgb"):
        if k in f:
            out["%s/%s" % (name, k)] = f[k].cpu().numpy()
    out["%s/num_rendered" % name] = np.int64(f["num_rendered"])
    dL, dLd = torch.from_numpy(ex["dL_dpix"]).cuda(), torch.from_numpy(ex["dL_dinvd"]).cuda()
    runs = [helpers.backward_ref(t, c, cam, bg_t, f, dL, dLd, **kw) for _ in range(5)]
    for k in runs[0]:
        if runs[0][k] is not None:
            out["%s/%s" % (name, k)] = torch.stack([r[k] for r in runs]).double().mean(0).float().cpu().numpy()
np.savez_compressed(os.path.join(out_dir, "rasterizer_reference.npz"), **out)

pts = torch.from_numpy(golden_inputs.knn_points()).cuda()
d = torch.zeros(pts.shape[0], device="cuda")
lib.ref_knn_mean_dist2(pts.shape[0], pts.data_ptr(), d.data_ptr())
torch.cuda.synchronize()
np.savez_compressed(os.path.join(out_dir, "knn_reference.npz"), dist2=d.cpu().numpy())
print("golden written:", sorted(os.listdir(out_dir)), {k: os.path.getsize(os.path.join(out_dir, k)) for k in os.listdir(out_dir)})
