"""Deterministic synthetic scenes and cameras (numpy only, so CPU oracle, tests and GPU runs see identical bytes).

Generators follow SURVEY.md §8(d): `blender_init_scene` mirrors the reference's random-point Blender
initialisation (LG/scene/dataset_readers.py:398-409 + GaussianModel.create_from_pcd), `trained_like_scene` is the
"1 M Gaussians, 800x800" metric scene.  Cameras follow LG/scene/cameras.py:80-89 and
LG/utils/graphics_utils.py:38-71 (row-vector convention: matrices are stored transposed).
"""
import math
from typing import NamedTuple

import numpy as np


class Camera(NamedTuple):
    image_width: int
    image_height: int
    tanfovx: float
    tanfovy: float
    viewmatrix: np.ndarray   # (4,4) float32 = world_view_transform (transposed W2C)
    projmatrix: np.ndarray   # (4,4) float32 = full_proj_transform
    campos: np.ndarray       # (3,) float32


class Scene(NamedTuple):
    means3D: np.ndarray      # (P,3)
    scales: np.ndarray       # (P,3) activated (exp applied)
    rotations: np.ndarray    # (P,4) normalised quaternions (r,x,y,z)
    opacities: np.ndarray    # (P,1) activated (sigmoid applied)
    shs: np.ndarray          # (P,16,3)
    sh_degree: int


def _projection(znear, zfar, fovx, fovy):
    t, r = math.tan(fovy / 2) * znear, math.tan(fovx / 2) * znear
    P = np.zeros((4, 4), dtype=np.float32)
    P[0, 0] = 2.0 * znear / (2 * r)
    P[1, 1] = 2.0 * znear / (2 * t)
    P[3, 2] = 1.0
    P[2, 2] = zfar / (zfar - znear)
    P[2, 3] = -(zfar * znear) / (zfar - znear)
    return P


def look_at_camera(width, height, fovx, fovy, eye, target=(0.0, 0.0, 0.0), up=(0.0, -1.0, 0.0), znear=0.01, zfar=100.0):
    """Camera at `eye` looking at `target` (+z forward, COLMAP-style), in the reference's storage convention."""
    eye = np.asarray(eye, dtype=np.float64)
    fwd = np.asarray(target, dtype=np.float64) - eye
    fwd /= np.linalg.norm(fwd)
    right = np.cross(fwd, np.asarray(up, dtype=np.float64))
    if np.linalg.norm(right) < 1e-8:
        right = np.cross(fwd, np.array([1.0, 0.0, 0.0]))
    right /= np.linalg.norm(right)
    down = np.cross(fwd, right)
    w2c = np.eye(4)
    w2c[:3, :3] = np.stack([right, down, fwd], axis=0)
    w2c[:3, 3] = -w2c[:3, :3] @ eye
    view = np.float32(w2c).T.copy()                       # world_view_transform
    proj = _projection(znear, zfar, fovx, fovy).T.copy()  # projection_matrix (transposed)
    full = (view.astype(np.float32) @ proj.astype(np.float32)).astype(np.float32)
    campos = np.linalg.inv(view.astype(np.float64))[3, :3].astype(np.float32)
    return Camera(int(width), int(height), math.tan(fovx * 0.5), math.tan(fovy * 0.5), view, full, campos)


def orbit_cameras(n, width, height, fov=0.6911, radius=4.03, elevation=0.3, phase=0.0):
    cams = []
    for i in range(n):
        a = phase + 2.0 * math.pi * i / max(n, 1)
        eye = (radius * math.cos(a) * math.cos(elevation), -radius * math.sin(elevation),
               radius * math.sin(a) * math.cos(elevation))
        cams.append(look_at_camera(width, height, fov, fov * height / width if width != height else fov, eye))
    return cams


def _random_unit_quaternions(rng, n):
    q = rng.standard_normal((n, 4))
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    return q.astype(np.float32)


def _sigmoid(x):
    return 1.0 / (1.0 + np.exp(-x))


def blender_init_scene(P=100_000, seed=0, spacing_scale=None):
    """Config 2: xyz ~ U[-1.3,1.3]^3, isotropic scale ~ 3-NN spacing, identity rotation, opacity 0.1."""
    rng = np.random.default_rng(seed)
    xyz = (rng.random((P, 3)) * 2.6 - 1.3).astype(np.float32)
    if spacing_scale is None:
        spacing_scale = 0.55 * (2.6 ** 3 / P) ** (1.0 / 3.0)  # ~ sqrt(mean 3-NN dist^2) for a uniform cloud
    scales = np.full((P, 3), spacing_scale, dtype=np.float32)
    rots = np.zeros((P, 4), dtype=np.float32)
    rots[:, 0] = 1.0
    opac = np.full((P, 1), 0.1, dtype=np.float32)
    shs = np.zeros((P, 16, 3), dtype=np.float32)
    shs[:, 0, :] = rng.normal(0.0, 0.3, (P, 3))
    shs[:, 1:, :] = rng.normal(0.0, 0.05, (P, 15, 3))
    return Scene(xyz, scales, rots, opac, shs.astype(np.float32), 3)


def trained_like_scene(P=1_000_000, seed=1, sigma_xyz=0.6, clip=1.5, log_scale_mean=math.log(0.005),
                       log_scale_std=0.6, opacity_logit_std=2.0):
    """Metric scene: anisotropic, randomly rotated, wide opacity distribution, strong view-dependent colour."""
    rng = np.random.default_rng(seed)
    xyz = np.clip(rng.normal(0.0, sigma_xyz, (P, 3)), -clip, clip).astype(np.float32)
    scales = np.exp(rng.normal(log_scale_mean, log_scale_std, (P, 3))).astype(np.float32)
    rots = _random_unit_quaternions(rng, P)
    opac = _sigmoid(rng.normal(0.0, opacity_logit_std, (P, 1))).astype(np.float32)
    shs = np.zeros((P, 16, 3), dtype=np.float32)
    shs[:, 0, :] = rng.normal(0.0, 1.0, (P, 3))
    shs[:, 1:, :] = rng.normal(0.0, 0.15, (P, 15, 3))
    return Scene(xyz, scales, rots, opac, shs, 3)


def slab_scene(P=500_000, seed=2, fov=1.05):
    """Config 3: forward-facing slab z in [2, 8] seen by LLFF-style cameras near the origin."""
    rng = np.random.default_rng(seed)
    z = rng.uniform(2.0, 8.0, P)
    half = np.tan(fov / 2) * z * 1.3
    xyz = np.stack([rng.uniform(-1, 1, P) * half, rng.uniform(-1, 1, P) * half * 0.75, z], axis=1).astype(np.float32)
    scales = np.exp(rng.normal(math.log(0.01), 0.6, (P, 3))).astype(np.float32)
    rots = _random_unit_quaternions(rng, P)
    opac = _sigmoid(rng.normal(0.0, 2.0, (P, 1))).astype(np.float32)
    shs = np.zeros((P, 16, 3), dtype=np.float32)
    shs[:, 0, :] = rng.normal(0.0, 1.0, (P, 3))
    shs[:, 1:, :] = rng.normal(0.0, 0.15, (P, 15, 3))
    return Scene(xyz, scales, rots, opac, shs, 3)


def metric_camera(width=800, height=800, fov=0.6911, radius=4.03):
    """The camera the headline metric is quoted on: distance 4.03, FoV 0.6911 rad, looking at 0."""
    return look_at_camera(width, height, fov, fov, (0.0, 0.0, -radius))


def dwt_pair(C=3, H=800, W=800, seed=0, noise=0.05):
    """Config 1: gt = rand, pred = clamp(gt + noise * randn, 0, 1)."""
    rng = np.random.default_rng(seed)
    gt = rng.random((C, H, W)).astype(np.float32)
    pred = np.clip(gt + noise * rng.standard_normal((C, H, W)), 0.0, 1.0).astype(np.float32)
    return pred, gt


def baseline_config(name):
    """(scene, camera) of a BASELINE.json config at FULL size:
    metric = 1 M trained-like Gaussians, 800x800; cfg2 = 100 k Blender-style init, 800x800; cfg3 = LLFF-style slab,
    500 k, 1008x756; cfg4 = RGB+NIR scale, 1 M, 1296x964; cfg5 = Mip-NeRF360 scale, 6 M, 1920x1080."""
    if name == "metric":
        return trained_like_scene(1_000_000, seed=1), metric_camera(800, 800)
    if name == "cfg2":
        return blender_init_scene(100_000, seed=0), metric_camera(800, 800)
    if name == "cfg3":
        return slab_scene(500_000, seed=2), look_at_camera(1008, 756, 1.05, 2 * math.atan(math.tan(0.525) * 756 / 1008),
                                                           (0.0, 0.0, 0.0), target=(0.0, 0.0, 5.0))
    if name == "cfg4":
        return trained_like_scene(1_000_000, seed=4), look_at_camera(1296, 964, 0.8, 2 * math.atan(math.tan(0.4) * 964 / 1296),
                                                                     (0.0, 0.0, -4.03))
    if name == "cfg5":
        sc = trained_like_scene(6_000_000, seed=5, sigma_xyz=1.2, clip=3.0, log_scale_mean=math.log(0.006))
        return sc, look_at_camera(1920, 1080, 1.0, 2 * math.atan(math.tan(0.5) * 1080 / 1920), (0.0, 0.0, -5.0))
    raise ValueError("unknown BASELINE config %r" % (name,))
