"""Write a synthetic training scene in the formats the reference's loaders read: ground-truth images are
rendered with this repo's rasterizer from a seeded "true" Gaussian scene, the initial point cloud is a noisy subsample
of its centres.  Lets LG/train.py (Blender or COLMAP loader) run on a box that has no datasets.

    python tools/make_synthetic_dataset.py OUT_DIR [--format blender|colmap] [--views 24] [--size 400x300] [--gaussians 50000]
"""
import argparse
import math
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "sparse-view-3dgs-pack_b200")):
    sys.path.insert(0, p)
from diff_gaussian_rasterization import GaussianRasterizationSettings, GaussianRasterizer  # noqa: E402
from lgdwt_b200 import io, scenes  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("out")
    ap.add_argument("--format", default="blender", choices=["blender", "colmap"])
    ap.add_argument("--views", type=int, default=24)
    ap.add_argument("--size", default="400x300")
    ap.add_argument("--gaussians", type=int, default=50_000)
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--nir", action="store_true", help="also write nir/<name>.png (grey-scale near-infrared stand-in: a "
                    "fixed mix of the rendered channels), where MS/utils/camera_utils.py:65-84 looks for it")
    args = ap.parse_args()
    W, H = (int(v) for v in args.size.lower().split("x"))
    dev = torch.device("cuda", 0)
    sc = scenes.trained_like_scene(args.gaussians, seed=args.seed, log_scale_mean=math.log(0.02))
    cams = scenes.orbit_cameras(args.views, W, H)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    p = {k: t(getattr(sc, k)) for k in ("means3D", "shs", "opacities", "scales", "rotations")}
    bg = torch.zeros(3, device=dev)
    images = []
    with torch.no_grad():
        for cam in cams:
            rs = GaussianRasterizationSettings(H, W, cam.tanfovx, cam.tanfovy, bg, 1.0, t(cam.viewmatrix), t(cam.projmatrix),
                                               3, t(cam.campos), False, False, False)
            color, radii, _ = GaussianRasterizer(rs)(means3D=p["means3D"], means2D=torch.zeros_like(p["means3D"]),
                                                     shs=p["shs"], opacities=p["opacities"], scales=p["scales"],
                                                     rotations=p["rotations"])
            images.append(color.clamp(0, 1).cpu().numpy())
    rng = np.random.default_rng(args.seed)
    pick = rng.choice(args.gaussians, size=min(args.gaussians, 20_000), replace=False)
    xyz = (sc.means3D[pick] + rng.normal(0, 0.01, (pick.size, 3))).astype(np.float32)
    rgb = np.clip((0.5 + 0.28209479177387814 * sc.shs[pick, 0, :]) * 255.0, 0, 255).astype(np.uint8)   # SH2RGB of f_dc
    if args.format == "blender":
        counts = io.write_blender_dataset(args.out, cams, images, split_test_every=8, points=(xyz, rgb))
    else:
        counts = {"images": io.write_colmap_dataset(args.out, cams, images, (xyz, rgb))}
    if args.nir:
        from PIL import Image
        os.makedirs(os.path.join(args.out, "nir"), exist_ok=True)
        names = (["view_%04d.png" % k for k in range(len(images))] if args.format == "colmap" else
                 ["r_%d.png" % k for k in range(len(images))])
        for name, im in zip(names, images):
            nir = np.clip(0.6 * im[0] + 0.1 * im[1] + 0.3 * (1.0 - im[2]) * im[1], 0, 1)
            Image.fromarray((nir * 255.0 + 0.5).astype(np.uint8), "L").save(os.path.join(args.out, "nir", name))
    print("wrote %s dataset to %s: %s, %dx%d, %d initial points, mean image %.3f" %
          (args.format, args.out, counts, W, H, pick.size, float(np.mean([im.mean() for im in images]))))


if __name__ == "__main__":
    main()
