"""Generates tests/golden/dwt_reference.npz by executing the REFERENCE's own LG/utils/loss_utils.py (imported from
/root/reference, unmodified) on seeded inputs in the build container.  The reference's un-vendored dependency
`pytorch_wavelets` is absent, so oracle/dwt_oracle.DWTForward is injected under that name; what this pins is
therefore everything the reference itself implements on top of the transform (sub-band bookkeeping, L1s, ELF map,
bilinear upsampling, unfold/kthvalue patch selection, patch loss) and its autograd gradients.

Run:  python tests/golden/make_dwt_golden.py     (needs /root/reference; not run on the GPU box)
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
CASES = {  # name: (C, H, W, patch_size, percentile, band weights)
    "even_6patch": (3, 264, 392, 128, 0.2, (1, 1, 1, 0, 0, 0, 0, 0)),
    "all_bands": (3, 256, 256, 64, 0.3, (1.0, 0.5, 0.25, 2.0, 0.7, 0.3, 0.9, 1.1)),
    "odd_sizes": (3, 131, 77, 32, 0.2, (1, 1, 1, 1, 1, 1, 1, 1)),
    "four_channels": (4, 160, 192, 64, 0.5, (1, 1, 1, 0, 0, 0, 0, 0)),
    "smaller_than_patch": (3, 100, 90, 128, 0.2, (1, 1, 1, 0, 0, 0, 0, 0)),
}
NAMES = ("LL1", "LH1", "HL1", "HH1", "LL2", "LH2", "HL2", "HH2")
out = {}
for name, (C, H, W, ps, pct, wts) in CASES.items():
    pred, gt = golden_inputs.dwt_case_inputs(name, C, H, W)
    p = torch.from_numpy(pred).unsqueeze(0).requires_grad_(True)
    g = torch.from_numpy(gt).unsqueeze(0)
    pb, gb = ref.get_dwt_subbands(p), ref.get_dwt_subbands(g)          # LG/train.py:138-139
    band_l1 = [ref.l1_loss(pb[n], gb[n]) for n in NAMES]
    dwt = sum(w * l for w, l in zip(wts, band_l1) if w != 0.0)          # LG/train.py:142-164
    elf = ref.compute_elf_map(g)                                       # LG/train.py:173
    patch = ref.compute_patch_dwt_loss(p, g, elf, patch_size=ps, percentile=pct, lh1_weight=1.0, hl1_weight=0.5)
    g_dwt, g_patch = 1.7, 0.1                                          # dwt_scale, patch_dwt_weight (LG/train.py:196,202)
    loss = g_dwt * dwt + g_patch * patch
    loss.backward()
    # pred / gt are regenerated from the seed by the tests (tests/golden_inputs.py), not stored
    out[name + "/cfg"] = np.array([C, H, W, ps, pct, 1.0, 0.5, g_dwt, g_patch] + list(wts), dtype=np.float64)
    out[name + "/band_l1"] = np.array([float(v) for v in band_l1], dtype=np.float64)
    out[name + "/dwt_loss"] = np.float64(float(dwt))
    out[name + "/patch_loss"] = np.float64(float(patch))
    out[name + "/elf_mean"] = np.float64(float(elf.mean()))
    grad = p.grad[0].numpy().astype(np.float32)
    out[name + "/grad_sub"] = grad[:, ::3, ::5].copy()          # strided sample keeps the fixture small
    out[name + "/grad_stats"] = np.array([grad.astype(np.float64).sum(), np.abs(grad).astype(np.float64).sum(),
                                          float(np.abs(grad).max())])
    if H >= ps and W >= ps:
        ep = torch.nn.functional.unfold(elf, kernel_size=ps, stride=ps).mean(dim=1).view(-1)
        out[name + "/patch_elf_means"] = ep.numpy().astype(np.float32)
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "dwt_reference.npz"), **out)
print("wrote dwt_reference.npz:", {k: (v.shape if hasattr(v, "shape") else v) for k, v in out.items() if "/cfg" in k})
