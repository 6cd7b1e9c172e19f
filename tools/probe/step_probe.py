"""forward / backward of a small scene with progress markers (run under `timeout -s KILL`)"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, "sparse-view-3dgs-pack_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import helpers
from lgdwt_b200 import scenes
P = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
sc = scenes.trained_like_scene(P, seed=1, log_scale_mean=-4.0)
cam = scenes.metric_camera(320, 240)
t, c = helpers.scene_to_torch(sc), helpers.cam_to_torch(cam)
bg = torch.zeros(3, device="cuda")
dL = torch.randn((3, cam.image_height, cam.image_width), device="cuda")
print("start", flush=True)
f = helpers.run_ours(t, c, cam, bg, want_state=False)
print("forward ok R=%d" % f["num_rendered"], flush=True)
g = helpers.backward_ours(t, c, cam, bg, f, dL, None)
print("backward ok", float(g["dL_dsh"].abs().sum()), flush=True)
