import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "sparse-view-3dgs-pack_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import golden_inputs
from oracle import dwt_oracle
from lgdwt_b200 import DWTLossConfig, fused_dwt_loss
gold = np.load(os.path.join(ROOT, "tests/golden/dwt_reference.npz"))
name = "odd_sizes"
cfg = gold[name + "/cfg"]
C, H, W, ps = (int(v) for v in cfg[:4])
pct, w_lh, w_hl, g_dwt, g_patch = (float(v) for v in cfg[4:9])
wts = tuple(float(v) for v in cfg[9:17])
pred, gt = golden_inputs.dwt_case_inputs(name, C, H, W)
for label, wt, pe in (("all", wts, True), ("lvl1 only", wts[:4] + (0, 0, 0, 0), False), ("lvl2 only", (0, 0, 0, 0) + wts[4:], False),
                      ("patch only", (0,) * 8, True)):
    p = torch.from_numpy(pred).cuda().requires_grad_(True)
    dwt, patch, det = fused_dwt_loss(p, torch.from_numpy(gt).cuda(), DWTLossConfig(wt, ps, pct, w_lh, w_hl, pe))
    (g_dwt * dwt + g_patch * patch).backward()
    po = torch.from_numpy(pred).requires_grad_(True)
    d2, p2, _, mask = dwt_oracle.lgdwt_losses(po, torch.from_numpy(gt), wt, ps, pct, w_lh, w_hl, pe)
    (g_dwt * d2 + g_patch * p2).backward()
    diff = (p.grad.cpu() - po.grad).abs()
    print(label, "max diff", float(diff.max()), "nsel", float(det[10]), None if mask is None else int(mask.sum()))
    rows = diff.sum((0, 2)); cols = diff.sum((0, 1))
    print("  rows with diff:", [(int(i), round(float(rows[i]), 6)) for i in torch.nonzero(rows > 1e-7).flatten()[:12]])
    print("  cols with diff:", [(int(i), round(float(cols[i]), 6)) for i in torch.nonzero(cols > 1e-7).flatten()[:12]])
