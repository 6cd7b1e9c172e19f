"""BASELINE config 5 at full size under data parallelism (run with torchrun on 2+ GPUs): 6 M Gaussians, 1920x1080,
4 views per rank per step through dp.ViewParallelTrainer (fused path) with exchange = peer and nccl: step time,
exchange time, replicas in sync.   torchrun --nproc-per-node N tools/dp_check.py [P]"""
import math
import os
import sys
import time

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "sparse-view-3dgs-pack_b200")):
    sys.path.insert(0, p)
from lgdwt_b200 import dp, scenes  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
dev = torch.device("cuda", local)
torch.cuda.set_device(dev)
dist.init_process_group("nccl", device_id=dev)
P = int(sys.argv[1]) if len(sys.argv) > 1 else 6_000_000
W, H, V = 1920, 1080, 4
sc = scenes.trained_like_scene(P, seed=5, sigma_xyz=1.2, clip=3.0, log_scale_mean=math.log(0.006))
fovy = 2 * math.atan(math.tan(0.5) * H / W)
cams = [dp.camera_to_device(scenes.look_at_camera(W, H, 1.0, fovy, (5.0 * math.sin(0.3 * k), 0.0, -5.0 * math.cos(0.3 * k))), dev)
        for k in range(V * world)]
gen = torch.Generator(device=dev).manual_seed(3)
gts = [torch.rand((3, H, W), device=dev, generator=gen) for _ in range(V * world)]
bg = torch.zeros(3, device=dev)
for mode in ("peer", "nccl"):
    g = dp.FlatGaussians.from_scene(sc, dev)
    tr = dp.ViewParallelTrainer(g, exchange=mode)
    for _ in range(2):
        tr.step(cams, gts, bg)
    dist.barrier()
    torch.cuda.synchronize()
    a, b, c = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    steps, t_ex = 3, 0.0
    t0 = time.perf_counter()
    for _ in range(steps):
        tr.iteration += 1
        a.record()
        tr.accumulate_views(cams, gts, bg)
        b.record()
        tr.exchange_and_update(1.0 / len(cams))
        c.record()
        torch.cuda.synchronize()
        t_ex += b.elapsed_time(c)
    wall = (time.perf_counter() - t0) / steps * 1e3
    ok = tr.replicas_in_sync()
    if tr.peer is not None:
        tr.peer.check()
    t = torch.tensor([wall, t_ex / steps], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        print("DP_CHECK P=%d %dx%d world=%d exchange=%s (%s): %.2f ms/step (%d views), exchange+Adam %.3f ms, %.1f views/s, "
              "replicas in sync: %s" % (P, W, H, world, mode, tr.peer.backend if tr.peer is not None else tr.peer_unavailable or "nccl",
                                        float(t[0]), V * world, float(t[1]), V * world / float(t[0]) * 1e3, ok), flush=True)
    del tr, g
    torch.cuda.empty_cache()
dist.barrier()
dist.destroy_process_group()
