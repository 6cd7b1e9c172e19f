"""One forward+backward of ours and of the reference on the metric scene, for `ncu` launch lists."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "sparse-view-3dgs-pack_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import helpers  # noqa: E402
from lgdwt_b200 import scenes  # noqa: E402

P = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
which = sys.argv[2] if len(sys.argv) > 2 else "both"
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 2
sc = scenes.trained_like_scene(P, seed=1)
cam = scenes.metric_camera()
t, c = helpers.scene_to_torch(sc), helpers.cam_to_torch(cam)
bg = torch.zeros(3, device="cuda")
dL = torch.randn((3, cam.image_height, cam.image_width), device="cuda")
for _ in range(iters):
    if which in ("both", "ours"):
        f = helpers.run_ours(t, c, cam, bg, want_state=False)
        helpers.backward_ours(t, c, cam, bg, f, dL, None)
    if which in ("both", "ref") and helpers.load_ref() is not None:
        f = helpers.run_ref(t, c, cam, bg, want_state=False)
        helpers.backward_ref(t, c, cam, bg, f, dL, dL[:1].contiguous())
torch.cuda.synchronize()
print("done")
