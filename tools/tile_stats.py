"""tile-list length statistics of a BASELINE config + un-profiled CUDA-event time of the forward stages"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "sparse-view-3dgs-pack_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import helpers  # noqa: E402
from lgdwt_b200 import _lib, scenes  # noqa: E402

for name in sys.argv[1:] or ["metric"]:
    sc, cam = scenes.baseline_config(name)
    t, c = helpers.scene_to_torch(sc), helpers.cam_to_torch(cam)
    bg = torch.zeros(3, device="cuda")
    o = helpers.run_ours(t, c, cam, bg)
    r = o["ranges"].view(-1, 2).long()
    n = (r[:, 1] - r[:, 0]).cpu().numpy()
    print("%s: R=%d tiles=%d mean=%.0f p50=%d p90=%d p99=%d max=%d  >4096: %d  >8192: %d" % (
        name, o["num_rendered"], n.size, n.mean(), np.percentile(n, 50), np.percentile(n, 90), np.percentile(n, 99),
        n.max(), (n > 4096).sum(), (n > 8192).sum()))
    _lib.stage_timing(32)
    for _ in range(32):
        helpers.run_ours(t, c, cam, bg, want_state=False)
    rows = [_lib.read_stage_times(s) for s in range(8, 32)]
    _lib.stage_timing(0)
    print("   stages (ms, mean of 24): " + ", ".join("%s %.4f" % (k, np.mean([r_[k] for r_ in rows])) for k in ("preprocess", "binning", "blend_forward")))
