"""Generates tests/golden/densify_reference.npz by executing the REFERENCE's own LG/scene/gaussian_model.py
(imported from /root/reference, unmodified) on CPU in the build container: `densify_and_prune`
(:456-476 -> densify_and_clone :437-454, densify_and_split :411-435, prune_points :347-363, the optimiser-state
surgery :331-409), `add_densification_stats` (:478-480), the max_radii2D update of LG/train.py:268 and `reset_opacity`
(:258-261).

The module hard-codes device="cuda"; torch's factory functions are wrapped for the duration of this script so that
those requests land on the CPU (nothing else is changed).  Its un-importable dependencies (`plyfile`, the CUDA-only
`simple_knn._C`) are stubbed: neither is touched by the functions exercised here.  The normal samples of
densify_and_split come from torch's global CPU generator; `torch.normal(mean, std)` draws `randn * std + mean`, so the
unit normals are recovered by re-seeding and drawing `torch.randn` of the same shape (asserted below) and stored with
the fixture — the product takes them as an input (the caller owns the generator).

Run:  python tests/golden/make_densify_golden.py     (needs /root/reference; not run on the GPU box)
"""
import os
import sys
import types
from argparse import Namespace

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import golden_inputs  # noqa: E402

REF_LG = "/root/reference/fs3dgs_benchmark/LGDWT-GS"

for fname in ("zeros", "ones", "empty", "tensor", "full", "eye", "rand", "randn"):
    def _wrap(fn):
        def inner(*a, **kw):
            if str(kw.get("device", "")).startswith("cuda"):
                kw["device"] = "cpu"
            return fn(*a, **kw)
        return inner
    setattr(torch, fname, _wrap(getattr(torch, fname)))

ply = types.ModuleType("plyfile")
ply.PlyData = ply.PlyElement = object
sys.modules["plyfile"] = ply
knn = types.ModuleType("simple_knn")
knn_c = types.ModuleType("simple_knn._C")
knn_c.distCUDA2 = lambda pts: (_ for _ in ()).throw(RuntimeError("not used here"))
knn._C = knn_c
sys.modules["simple_knn"], sys.modules["simple_knn._C"] = knn, knn_c
sys.path.insert(0, REF_LG)
from scene.gaussian_model import GaussianModel  # noqa: E402  (the reference class itself)

GROUPS = ("xyz", "f_dc", "f_rest", "opacity", "scaling", "rotation")
ATTR = {"xyz": "_xyz", "f_dc": "_features_dc", "f_rest": "_features_rest", "opacity": "_opacity",
        "scaling": "_scaling", "rotation": "_rotation"}
OPT = Namespace(percent_dense=golden_inputs.PERCENT_DENSE, position_lr_init=0.00016, position_lr_final=0.0000016,
                position_lr_delay_mult=0.01, position_lr_max_steps=30000, feature_lr=0.0025, opacity_lr=0.025,
                scaling_lr=0.005, rotation_lr=0.001, exposure_lr_init=0.01, exposure_lr_final=0.001,
                exposure_lr_delay_steps=0, exposure_lr_delay_mult=0.0, iterations=30000)


def build_model(d):
    gm = GaussianModel(3)
    for k in GROUPS:
        setattr(gm, ATTR[k], torch.nn.Parameter(torch.from_numpy(d[k].copy()).requires_grad_(True)))
    gm.spatial_lr_scale = 1.0
    gm._exposure = torch.nn.Parameter(torch.eye(3, 4)[None].clone())
    gm.training_setup(OPT)
    for group in gm.optimizer.param_groups:
        k = group["name"]
        gm.optimizer.state[group["params"][0]] = {"step": torch.tensor(7.0),
                                                  "exp_avg": torch.from_numpy(d["exp_avg/" + k].copy()),
                                                  "exp_avg_sq": torch.from_numpy(d["exp_avg_sq/" + k].copy())}
    return gm


def dump(gm, out, prefix):
    for group in gm.optimizer.param_groups:
        k = group["name"]
        p = group["params"][0]
        assert p is getattr(gm, ATTR[k])
        out["%s/%s" % (prefix, k)] = p.detach().numpy().copy()
        st = gm.optimizer.state[p]
        out["%s/exp_avg/%s" % (prefix, k)] = st["exp_avg"].numpy().copy()
        out["%s/exp_avg_sq/%s" % (prefix, k)] = st["exp_avg_sq"].numpy().copy()
    out[prefix + "/xyz_gradient_accum"] = gm.xyz_gradient_accum.numpy().copy()
    out[prefix + "/denom"] = gm.denom.numpy().copy()
    out[prefix + "/max_radii2D"] = gm.max_radii2D.numpy().copy()


out = {}
for name in golden_inputs.DENSIFY_CASES:
    d, cfg = golden_inputs.densify_case(name)
    gm = build_model(d)
    gm.xyz_gradient_accum = torch.from_numpy(d["xyz_gradient_accum"].copy())
    gm.denom = torch.from_numpy(d["denom"].copy())
    gm.max_radii2D = torch.from_numpy(d["max_radii2D"].copy())
    # how many points densify_and_split will draw for: recorded by wrapping torch.normal (pure observation)
    seen = {}
    real_normal = torch.normal

    def spy(*a, **kw):
        r = real_normal(*a, **kw)
        seen["std"], seen["samples"] = kw["std"].clone(), r.clone()
        return r
    torch.normal = spy
    torch.manual_seed(cfg["seed"])
    with torch.no_grad():
        pass
    gm.densify_and_prune(cfg["max_grad"], cfg["min_opacity"], cfg["extent"], cfg["max_screen_size"],
                         torch.from_numpy(d["radii"].copy()))
    torch.normal = real_normal
    torch.manual_seed(cfg["seed"])
    eps = torch.randn(seen["std"].shape)
    assert torch.equal(eps * seen["std"], seen["samples"]), "torch.normal(mean=0, std) is not randn * std here"
    dump(gm, out, name)
    out[name + "/eps"] = eps.numpy().copy()
    print(name, "P", d["xyz"].shape[0], "->", gm._xyz.shape[0], "split parents", eps.shape[0] // 2)

# statistics update + opacity reset
s = golden_inputs.stats_case()
P = s["radii"].shape[0]
d, _ = golden_inputs.densify_case("mixed")
d = {k: v[:P] if v.shape[0] >= P else np.resize(v, (P,) + v.shape[1:]) for k, v in d.items()}
d["opacity"] = s["opacity"]
gm = build_model(d)
gm.xyz_gradient_accum = torch.from_numpy(s["xyz_gradient_accum"].copy())
gm.denom = torch.from_numpy(s["denom"].copy())
gm.max_radii2D = torch.from_numpy(s["max_radii2D"].copy())
radii = torch.from_numpy(s["radii"].copy())
vis = radii > 0                                                   # LG/gaussian_renderer/__init__.py visibility_filter
vsp = torch.zeros((P, 3), requires_grad=True)
vsp.grad = torch.from_numpy(s["grad2D"].copy())
gm.max_radii2D[vis] = torch.max(gm.max_radii2D[vis], radii[vis])  # LG/train.py:268
gm.add_densification_stats(vsp, vis)                              # LG/train.py:269
out["stats/xyz_gradient_accum"] = gm.xyz_gradient_accum.numpy().copy()
out["stats/denom"] = gm.denom.numpy().copy()
out["stats/max_radii2D"] = gm.max_radii2D.numpy().copy()
gm.reset_opacity()                                                # LG/train.py:276
st = gm.optimizer.state[gm._opacity]
out["reset/opacity"] = gm._opacity.detach().numpy().copy()
assert float(st["exp_avg"].abs().max()) == 0.0 and float(st["exp_avg_sq"].abs().max()) == 0.0

path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "densify_reference.npz")
np.savez_compressed(path, **out)
print("wrote", path, os.path.getsize(path) // 1024, "KiB")
