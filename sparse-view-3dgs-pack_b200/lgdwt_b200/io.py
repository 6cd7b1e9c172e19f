"""File formats either side of the training path: Gaussian point-cloud PLY in the reference's exact
property order, the Blender-format ("NeRF synthetic") and COLMAP-format dataset layouts its loaders read, the initial
point cloud `points3d.ply`, and the `cfg_args` file `render.py` looks for — so that a trained `dp.FlatGaussians` can be opened by
the reference's viewer / render.py and a synthetic dataset can drive the reference's train.py on a box without data.

Mirrors: GaussianModel.save_ply / load_ply / construct_list_of_attributes (LG/scene/gaussian_model.py:225-314),
storePly / fetchPly (LG/scene/dataset_readers.py:163-186), readCamerasFromTransforms (:331-374), the cfg_args dump of
LG/train.py:305-306.  PLY access goes through `plyfile` (the real package if installed, else compat/plyfile.py).
"""
import json
import math
import os
import sys

import numpy as np

try:
    from plyfile import PlyData, PlyElement
except ImportError:  # the stand-in shipped with this repo
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "compat"))
    from plyfile import PlyData, PlyElement


def gaussian_ply_attributes(sh_degree=3):
    """property order of construct_list_of_attributes (LG/scene/gaussian_model.py:225-237)"""
    rest = 3 * ((sh_degree + 1) ** 2 - 1)
    return (["x", "y", "z", "nx", "ny", "nz"] + ["f_dc_%d" % i for i in range(3)] +
            ["f_rest_%d" % i for i in range(rest)] + ["opacity"] + ["scale_%d" % i for i in range(3)] +
            ["rot_%d" % i for i in range(4)])


def save_gaussians_ply(path, g):
    """GaussianModel.save_ply for a dp.FlatGaussians: raw parameters, f_rest stored channel-major
    (`transpose(1, 2).flatten`, :245-246)"""
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    P, M = g.P, g.M
    shs = g.slab("shs").detach().cpu().numpy().reshape(P, M, 3)
    cols = [g.slab("xyz").detach().cpu().numpy(), np.zeros((P, 3), np.float32), shs[:, 0, :],
            shs[:, 1:, :].transpose(0, 2, 1).reshape(P, -1), g.slab("opacity").detach().cpu().numpy(),
            g.slab("scaling").detach().cpu().numpy(), g.slab("rotation").detach().cpu().numpy()]
    flat = np.concatenate(cols, axis=1).astype(np.float32)
    names = gaussian_ply_attributes(g.sh_degree)
    assert flat.shape[1] == len(names)
    elements = np.empty(P, dtype=[(n, "f4") for n in names])
    for k, n in enumerate(names):
        elements[n] = flat[:, k]
    PlyData([PlyElement.describe(elements, "vertex")]).write(path)


def load_gaussians_ply(path, device, sh_degree=3):
    """GaussianModel.load_ply (:263-314) into a dp.FlatGaussians"""
    import torch
    from . import dp
    v = PlyData.read(path).elements[0]
    names = [p.name for p in v.properties]
    P = len(v["x"])
    M = (sh_degree + 1) ** 2
    rest_names = sorted((n for n in names if n.startswith("f_rest_")), key=lambda s: int(s.split("_")[-1]))
    if len(rest_names) != 3 * M - 3:
        raise ValueError("%s holds %d f_rest properties, SH degree %d needs %d" % (path, len(rest_names), sh_degree, 3 * M - 3))
    col = lambda ns: np.stack([np.asarray(v[n], dtype=np.float32) for n in ns], axis=1)
    shs = np.zeros((P, M, 3), np.float32)
    shs[:, 0, :] = col(["f_dc_0", "f_dc_1", "f_dc_2"])
    shs[:, 1:, :] = col(rest_names).reshape(P, 3, M - 1).transpose(0, 2, 1)
    g = dp.FlatGaussians(P, device, sh_degree)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(device)
    g.slab("xyz").copy_(t(col(["x", "y", "z"])))
    g.slab("shs").copy_(t(shs.reshape(P, -1)))
    g.slab("opacity").copy_(t(col(["opacity"])))
    g.slab("scaling").copy_(t(col(sorted((n for n in names if n.startswith("scale_")), key=lambda s: int(s.split("_")[-1])))))
    g.slab("rotation").copy_(t(col(sorted((n for n in names if n.startswith("rot")), key=lambda s: int(s.split("_")[-1])))))
    return g


def store_point_cloud_ply(path, xyz, rgb_u8):
    """storePly (LG/scene/dataset_readers.py:171-186): x y z nx ny nz (f4) red green blue (u1)"""
    xyz = np.asarray(xyz, np.float32)
    elements = np.empty(xyz.shape[0], dtype=[("x", "f4"), ("y", "f4"), ("z", "f4"), ("nx", "f4"), ("ny", "f4"),
                                             ("nz", "f4"), ("red", "u1"), ("green", "u1"), ("blue", "u1")])
    for k, n in enumerate(("x", "y", "z")):
        elements[n] = xyz[:, k]
    for n in ("nx", "ny", "nz"):
        elements[n] = 0.0
    rgb = np.asarray(rgb_u8)
    for k, n in enumerate(("red", "green", "blue")):
        elements[n] = rgb[:, k].astype(np.uint8)
    PlyData([PlyElement.describe(elements, "vertex")]).write(path)


def camera_to_blender_frame(cam, file_path):
    """inverse of readCamerasFromTransforms (:342-350): world-to-camera (COLMAP axes, row-vector viewmatrix as in
    lgdwt_b200.scenes.Camera) -> NeRF camera-to-world with OpenGL axes"""
    w2c = np.asarray(cam.viewmatrix, np.float64).T      # column-vector convention
    c2w = np.linalg.inv(w2c)
    c2w[:3, 1:3] *= -1
    return {"file_path": file_path, "transform_matrix": c2w.tolist()}


def write_blender_dataset(root, cameras, images, split_test_every=8, points=None):
    """Blender-format scene: transforms_train.json / transforms_test.json + PNGs (+ points3d.ply).
    cameras: lgdwt_b200.scenes.Camera list (all the same size / FoV x); images: (3|4, H, W) float arrays in [0, 1]."""
    from PIL import Image
    os.makedirs(os.path.join(root, "train"), exist_ok=True)
    os.makedirs(os.path.join(root, "test"), exist_ok=True)
    fovx = 2.0 * math.atan(cameras[0].tanfovx)
    frames = {"train": [], "test": []}
    for k, (cam, img) in enumerate(zip(cameras, images)):
        split = "test" if (split_test_every and k % split_test_every == split_test_every - 1) else "train"
        rel = "./%s/r_%d" % (split, k)
        a = np.clip(np.asarray(img, np.float32), 0, 1)
        if a.shape[0] == 3:
            a = np.concatenate([a, np.ones_like(a[:1])], axis=0)
        Image.fromarray((a.transpose(1, 2, 0) * 255.0 + 0.5).astype(np.uint8), "RGBA").save(os.path.join(root, rel + ".png"))
        frames[split].append(camera_to_blender_frame(cam, rel))
    for split in ("train", "test"):
        with open(os.path.join(root, "transforms_%s.json" % split), "w") as f:
            json.dump({"camera_angle_x": fovx, "frames": frames[split]}, f, indent=1)
    if points is not None:
        xyz, rgb_u8 = points
        store_point_cloud_ply(os.path.join(root, "points3d.ply"), xyz, rgb_u8)
    return {k: len(v) for k, v in frames.items()}


def write_cfg_args(model_path, **kwargs):
    """the `Namespace(...)` text file train.py leaves next to a model and render.py evals (LG/train.py:305-306,
    LG/arguments/__init__.py:126-139)"""
    from argparse import Namespace
    os.makedirs(model_path, exist_ok=True)
    base = dict(sh_degree=3, source_path="", model_path=model_path, images="images", depths="", resolution=-1,
                white_background=False, train_test_exp=False, data_device="cuda", eval=False)
    base.update(kwargs)
    with open(os.path.join(model_path, "cfg_args"), "w") as f:
        f.write(str(Namespace(**base)))


# ---------------------------------------------------------------- COLMAP sparse model (LLFF / Mip-NeRF360 style scenes)
def _rotmat_to_qvec(R):
    """COLMAP quaternion (w, x, y, z) of a rotation matrix — the inverse of qvec2rotmat (LG/scene/colmap_loader.py:43-53),
    largest-component branch for numerical stability"""
    R = np.asarray(R, np.float64)
    t = np.trace(R)
    if t > 0:
        s = math.sqrt(t + 1.0) * 2
        q = [0.25 * s, (R[2, 1] - R[1, 2]) / s, (R[0, 2] - R[2, 0]) / s, (R[1, 0] - R[0, 1]) / s]
    elif R[0, 0] > R[1, 1] and R[0, 0] > R[2, 2]:
        s = math.sqrt(1.0 + R[0, 0] - R[1, 1] - R[2, 2]) * 2
        q = [(R[2, 1] - R[1, 2]) / s, 0.25 * s, (R[0, 1] + R[1, 0]) / s, (R[0, 2] + R[2, 0]) / s]
    elif R[1, 1] > R[2, 2]:
        s = math.sqrt(1.0 + R[1, 1] - R[0, 0] - R[2, 2]) * 2
        q = [(R[0, 2] - R[2, 0]) / s, (R[0, 1] + R[1, 0]) / s, 0.25 * s, (R[1, 2] + R[2, 1]) / s]
    else:
        s = math.sqrt(1.0 + R[2, 2] - R[0, 0] - R[1, 1]) * 2
        q = [(R[1, 0] - R[0, 1]) / s, (R[0, 2] + R[2, 0]) / s, (R[1, 2] + R[2, 1]) / s, 0.25 * s]
    q = np.array(q)
    return q if q[0] >= 0 else -q


def write_colmap_dataset(root, cameras, images, points, image_dir="images"):
    """COLMAP-format scene as readColmapSceneInfo expects it (LG/scene/dataset_readers.py:188-260): images/*.png and
    sparse/0/{cameras,images,points3D}.bin in COLMAP's binary layout (LG/scene/colmap_loader.py:113-142,167-229): one
    PINHOLE camera per view (fx, fy, cx, cy), world-to-camera (qvec, tvec) per image with no 2-D observations, points
    with empty tracks.  cameras: lgdwt_b200.scenes.Camera; images: (3, H, W) floats in [0, 1]; points: (xyz, rgb_u8)."""
    import struct
    from PIL import Image
    sparse = os.path.join(root, "sparse", "0")
    os.makedirs(sparse, exist_ok=True)
    os.makedirs(os.path.join(root, image_dir), exist_ok=True)
    with open(os.path.join(sparse, "cameras.bin"), "wb") as f:
        f.write(struct.pack("<Q", len(cameras)))
        for k, cam in enumerate(cameras):
            W, H = cam.image_width, cam.image_height
            fx, fy = W / (2.0 * cam.tanfovx), H / (2.0 * cam.tanfovy)
            f.write(struct.pack("<iiQQ", k + 1, 1, W, H))                 # model 1 = PINHOLE
            f.write(struct.pack("<dddd", fx, fy, W / 2.0, H / 2.0))
    with open(os.path.join(sparse, "images.bin"), "wb") as f:
        f.write(struct.pack("<Q", len(cameras)))
        for k, (cam, img) in enumerate(zip(cameras, images)):
            w2c = np.asarray(cam.viewmatrix, np.float64).T                # column-vector world-to-camera
            q, t = _rotmat_to_qvec(w2c[:3, :3]), w2c[:3, 3]
            name = "view_%04d.png" % k
            f.write(struct.pack("<idddddddi", k + 1, q[0], q[1], q[2], q[3], t[0], t[1], t[2], k + 1))
            f.write(name.encode("utf-8") + b"\x00")
            f.write(struct.pack("<Q", 0))
            a = np.clip(np.asarray(img, np.float32), 0, 1)[:3]
            Image.fromarray((a.transpose(1, 2, 0) * 255.0 + 0.5).astype(np.uint8), "RGB").save(os.path.join(root, image_dir, name))
    xyz, rgb = np.asarray(points[0], np.float64), np.asarray(points[1]).astype(np.uint8)
    with open(os.path.join(sparse, "points3D.bin"), "wb") as f:
        f.write(struct.pack("<Q", xyz.shape[0]))
        for i in range(xyz.shape[0]):
            f.write(struct.pack("<QdddBBBd", i + 1, xyz[i, 0], xyz[i, 1], xyz[i, 2], int(rgb[i, 0]), int(rgb[i, 1]),
                                int(rgb[i, 2]), 0.0))
            f.write(struct.pack("<Q", 0))
    return len(cameras)
