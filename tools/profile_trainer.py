"""A few fused training iterations + one densify_and_prune at the metric size (1M Gaussians, 800x800), for ncu:
the trainer-side kernels around the rasterizer (activate, image loss, statistics, Adam, density control)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "sparse-view-3dgs-pack_b200")):
    sys.path.insert(0, p)
from lgdwt_b200 import densify, dp, scenes  # noqa: E402

P = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 3
dev = torch.device("cuda", 0)
sc = scenes.trained_like_scene(P, seed=1)
g = dp.FlatGaussians.from_scene(sc, dev)
cam = dp.camera_to_device(scenes.metric_camera(800, 800), dev)
gt = torch.rand(3, 800, 800, device=dev)
cfg = dp.DensifyConfig(densify_from_iter=0, densify_until_iter=10**9, densification_interval=iters,
                       opacity_reset_interval=10**9, densify_grad_threshold=2e-6, cameras_extent=4.0)
tr = dp.ViewParallelTrainer(g, densify=cfg, seed=1)
bg = torch.zeros(3, device=dev)
for _ in range(iters):   # the last iteration densifies
    tr.step([cam], [gt], bg)
torch.cuda.synchronize()
print("done P", g.P)
