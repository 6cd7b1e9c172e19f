"""Generates tests/golden/photometric_reference.npz by executing the REFERENCE's own LG/utils/loss_utils.py
(`l1_loss`, `ssim`; imported from /root/reference, unmodified) on seeded inputs in the build container, with the
autograd gradient of the reference's base loss (1 - lambda) * L1 + lambda * (1 - SSIM) (LG/train.py:188).
The module imports `pytorch_wavelets` at load time; the compat transform is injected as for the DWT goldens (it is
not exercised by the functions used here).

Run:  python tests/golden/make_photometric_golden.py     (needs /root/reference; not run on the GPU box)
"""
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import dwt_oracle  # noqa: E402
import golden_inputs  # noqa: E402

REF_LG = "/root/reference/fs3dgs_benchmark/LGDWT-GS"
shim = types.ModuleType("pytorch_wavelets")
shim.DWTForward = dwt_oracle.DWTForward
sys.modules["pytorch_wavelets"] = shim
sys.path.insert(0, REF_LG)
from utils import loss_utils as ref  # noqa: E402  (the reference module itself)

torch.set_num_threads(4)
LAMBDA = 0.2  # lambda_dssim, LG/arguments/__init__.py
out = {}
for name, (C, H, W) in golden_inputs.PHOTOMETRIC_CASES.items():
    pred, gt = golden_inputs.photometric_case_inputs(name, C, H, W)
    p = torch.from_numpy(pred).requires_grad_(True)
    g = torch.from_numpy(gt)
    l1 = ref.l1_loss(p, g)                     # LG/train.py:128
    s = ref.ssim(p, g)                         # LG/train.py:185
    loss = (1.0 - LAMBDA) * l1 + LAMBDA * (1.0 - s)
    loss.backward()
    grad = p.grad.numpy().astype(np.float32)
    out[name + "/l1"] = np.float64(float(l1))
    out[name + "/ssim"] = np.float64(float(s))
    out[name + "/grad_sub"] = grad[:, ::3, ::5].copy()   # strided sample keeps the fixture small
    out[name + "/grad_stats"] = np.array([grad.astype(np.float64).sum(), np.abs(grad).astype(np.float64).sum(),
                                          float(np.abs(grad).max())])
path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "photometric_reference.npz")
np.savez_compressed(path, **out)
print("wrote", path, {k: (v.shape if hasattr(v, "shape") else v) for k, v in out.items() if k.endswith("/l1") or k.endswith("/ssim")})
