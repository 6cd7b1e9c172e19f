"""Test helper (run as a subprocess by tests/test_callers_gpu.py with the arm's PYTHONPATH and cwd = the reference's
MS tree): imports the reference's UNCHANGED multispectral modules, loads a saved model with use_nir=True — which is
where the reference creates the NIR albedo (MS/scene/gaussian_model.py:411-426) — and calls
MS/gaussian_renderer.render(), i.e. the RGB pass plus render_nir() (`colors_precomp` path, :151-258), on the first
views; also backpropagates a fixed cotangent so that both passes' backward kernels run.

    python ms_nir_probe.py MS_DIR MODEL_DIR ITERATION OUT.npz
"""
import os
import sys
from argparse import ArgumentParser

import numpy as np
import torch

ms_dir, model_dir, iteration, out = sys.argv[1], sys.argv[2], int(sys.argv[3]), sys.argv[4]
sys.path.insert(0, ms_dir)
from arguments import ModelParams, PipelineParams  # noqa: E402
from gaussian_renderer import render  # noqa: E402
from scene import GaussianModel, Scene  # noqa: E402

parser = ArgumentParser()
mp, pp = ModelParams(parser, sentinel=True), PipelineParams(parser)
cfg = eval(open(os.path.join(model_dir, "cfg_args")).read(), {"Namespace": __import__("argparse").Namespace})
args = parser.parse_args([])
for k, v in vars(cfg).items():
    setattr(args, k, v)
args.model_path = model_dir
dataset, pipe = mp.extract(args), pp.extract(args)
assert dataset.use_nir
gaussians = GaussianModel(dataset.sh_degree, use_nir=True)
scene = Scene(dataset, gaussians, load_iteration=iteration, shuffle=False)
bg = torch.zeros(3, device="cuda")
renders, nirs, radii = [], [], []
cams = scene.getTrainCameras()[:3]
gen = torch.Generator(device="cpu").manual_seed(7)
for cam in cams:
    pkg = render(cam, gaussians, pipe, bg)
    assert "nir" in pkg, "render() did not produce the NIR image"
    renders.append(pkg["render"].detach().cpu().numpy())
    nirs.append(pkg["nir"].detach().cpu().numpy())
    radii.append(pkg["radii"].cpu().numpy())
    wr = torch.rand(pkg["render"].shape, generator=gen).cuda()
    wn = torch.rand(pkg["nir"].shape, generator=gen).cuda()
    ((pkg["render"] * wr).sum() + (pkg["nir"] * wn).sum()).backward()
np.savez(out, render=np.stack(renders), nir=np.stack(nirs), radii=np.stack(radii),
         grad_xyz=gaussians._xyz.grad.cpu().numpy(), grad_nir=gaussians._nir_albedo.grad.cpu().numpy())
print("ms_nir_probe: %d views, NIR mean %.4f" % (len(cams), float(np.mean(nirs))))
