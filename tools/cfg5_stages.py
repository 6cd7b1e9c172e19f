"""Per-stage device times of the rasterizer at BASELINE config 5 size (6 M Gaussians, 1920x1080) for one camera of
bench.cfg5_section, next to the metric scene — do the kernels scale with P / pixels / entries as their bounds say?"""
import math, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "sparse-view-3dgs-pack_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import helpers
from lgdwt_b200 import scenes, _lib
import diff_gaussian_rasterization as dgr

def run(name, sc, cam, iters=6):
    t, c = helpers.scene_to_torch(sc), helpers.cam_to_torch(cam)
    bg = torch.zeros(3, device="cuda")
    dL = torch.randn((3, cam.image_height, cam.image_width), device="cuda")
    seen = []
    dgr.set_inspection_hook(None)
    for _ in range(2):
        f = helpers.run_ours(t, c, cam, bg, want_state=False)
        helpers.backward_ours(t, c, cam, bg, f, dL, None)
    _lib.stage_timing(iters, 1)
    for _ in range(iters):
        f = helpers.run_ours(t, c, cam, bg, want_state=False)
        helpers.backward_ours(t, c, cam, bg, f, dL, None)
    torch.cuda.synchronize()
    rows = [_lib.read_stage_times(s) for s in range(iters)]
    _lib.stage_timing(0)
    R = int(f["num_rendered"]) if isinstance(f, dict) and "num_rendered" in f else -1
    ms = {k: float(np.median([r[k] for r in rows])) for k in _lib.STAGES}
    print(name, "P", sc.means3D.shape[0], "R", R, {k: round(v, 4) for k, v in ms.items()}, "sum", round(sum(ms.values()), 4))

run("metric", scenes.trained_like_scene(1_000_000, seed=1), scenes.metric_camera())
Wd, Hd = 1920, 1080
fovy = 2 * math.atan(math.tan(0.5) * Hd / Wd)
sc5 = scenes.trained_like_scene(6_000_000, seed=5, sigma_xyz=1.2, clip=3.0, log_scale_mean=math.log(0.006))
for k in (0, 8):
    cam = scenes.look_at_camera(Wd, Hd, 1.0, fovy, (5.0 * math.sin(2 * math.pi * k / 32), 0.0, -5.0 * math.cos(2 * math.pi * k / 32)))
    run("cfg5 cam%d" % k, sc5, cam)
