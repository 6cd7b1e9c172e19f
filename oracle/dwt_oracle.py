"""TEST INFRASTRUCTURE ONLY — CPU restatement (plain PyTorch ops) of the LGDWT-GS wavelet loss.

PARITY UNPINNED at the wavelet-transform boundary: the reference computes its sub-bands with
`pytorch_wavelets.DWTForward` (LG/utils/loss_utils.py:104), an external package that is neither vendored, pinned
nor installable here.  `haar_dwt_level` below restates that package's published algorithm for
wave='db1', mode='symmetric' (pywt db1 taps, afb1d = strided cross-correlation along W then H,
one-sample symmetric extension at the end for odd lengths, and the package's cropped backward).  Everything ABOVE
that boundary follows the reference's own code line by line and is pinned against it: tests/golden/dwt_reference.npz
holds outputs of the reference's LG/utils/loss_utils.py functions executed in the build container with this
transform injected as `pytorch_wavelets` (script: tests/golden/make_dwt_golden.py).
"""
import torch
import torch.nn.functional as F

S = 0.7071067811865476  # pywt.Wavelet('db1').dec_lo[0]; pytorch_wavelets stores the filters as fp32


def _pad_last(x, dim):
    """one-sample symmetric (half-sample) extension at the end; detached = the package's cropped backward"""
    idx = [slice(None)] * x.dim()
    idx[dim] = slice(-1, None)
    return torch.cat([x, x[tuple(idx)].detach()], dim=dim)


def haar_dwt_level(x):
    """pytorch_wavelets.DWTForward(J=1, wave='db1', mode='symmetric')(x) -> (LL, (LH, HL, HH)).
    x: (N, C, H, W).  Filters h0 = [s, s], h1 = [s, -s] (pywt dec_lo / dec_hi reversed for cross-correlation);
    LH = low-pass along W, high-pass along H."""
    s = torch.tensor(S, dtype=x.dtype)
    if x.shape[-1] % 2:
        x = _pad_last(x, -1)
    if x.shape[-2] % 2:
        x = _pad_last(x, -2)
    a, b = x[..., :, 0::2], x[..., :, 1::2]          # along W
    lo, hi = s * a + s * b, s * a - s * b
    lo_t, lo_b = lo[..., 0::2, :], lo[..., 1::2, :]  # along H
    hi_t, hi_b = hi[..., 0::2, :], hi[..., 1::2, :]
    LL, LH = s * lo_t + s * lo_b, s * lo_t - s * lo_b
    HL, HH = s * hi_t + s * hi_b, s * hi_t - s * hi_b
    return LL, (LH, HL, HH)


class DWTForward(torch.nn.Module):
    """Stand-in with the pytorch_wavelets call signature, used to drive the reference's loss_utils.py when the
    golden vectors are generated."""

    def __init__(self, J=1, wave="db1", mode="zero"):
        super().__init__()
        assert wave in ("db1", "haar") and mode == "symmetric"
        self.J = J

    def forward(self, x):
        yh = []
        ll = x
        for _ in range(self.J):
            ll, (lh, hl, hh) = haar_dwt_level(ll)
            yh.append(torch.stack([lh, hl, hh], dim=2))
        return ll, yh


def get_dwt_subbands(x):
    """LG/utils/loss_utils.py:106-153 (the discarded J=2 transform at :121-135 is not repeated)."""
    LL1, (LH1, HL1, HH1) = haar_dwt_level(x)          # :140-144
    LL2, (LH2, HL2, HH2) = haar_dwt_level(LL1)        # :147-148
    return {"LL1": LL1, "LH1": LH1, "HL1": HL1, "HH1": HH1, "LL2": LL2, "LH2": LH2, "HL2": HL2, "HH2": HH2}


def l1_loss(a, b):
    """LG/utils/loss_utils.py:40-41"""
    return torch.abs(a - b).mean()


def compute_elf_map(image):
    """LG/utils/loss_utils.py:336-366"""
    bands = get_dwt_subbands(image)
    l1 = lambda t: torch.sum(torch.abs(t), dim=1, keepdim=True)
    LL, HF = l1(bands["LL1"]), l1(bands["LH1"]) + l1(bands["HL1"]) + l1(bands["HH1"])
    elf_low = LL / (LL + HF + 1e-8)
    H, W = image.shape[-2:]
    return F.interpolate(elf_low, size=(H, W), mode="bilinear", align_corners=False)


def patch_selection(elf_map, patch_size, percentile):
    """LG/utils/loss_utils.py:396-417 -> (mask (N, L) bool, threshold, patch means)"""
    elf_patches = F.unfold(elf_map, kernel_size=patch_size, stride=patch_size)
    means = elf_patches.mean(dim=1)
    flat = means.view(-1)
    k = int(flat.numel() * (1.0 - percentile))
    k = max(1, k)
    k = min(k, flat.numel())
    thr, _ = torch.kthvalue(flat, k)
    return means >= thr, thr, means


def compute_patch_dwt_loss(pred, gt, elf_map, patch_size=128, percentile=0.2, lh1_weight=1.0, hl1_weight=1.0):
    """LG/utils/loss_utils.py:368-442"""
    N, C, H, W = pred.shape
    if H < patch_size or W < patch_size:
        return torch.tensor(0.0)
    pp = F.unfold(pred, kernel_size=patch_size, stride=patch_size)
    gp = F.unfold(gt, kernel_size=patch_size, stride=patch_size)
    L = pp.shape[2]
    mask, _, _ = patch_selection(elf_map, patch_size, percentile)
    if mask.sum() == 0:
        return torch.tensor(0.0, device=pred.device)
    pp = pp.view(N, C, patch_size, patch_size, L).permute(0, 4, 1, 2, 3)
    gp = gp.view(N, C, patch_size, patch_size, L).permute(0, 4, 1, 2, 3)
    pb, gb = get_dwt_subbands(pp[mask]), get_dwt_subbands(gp[mask])
    l_lh, l_hl, l_hh = l1_loss(pb["LH1"], gb["LH1"]), l1_loss(pb["HL1"], gb["HL1"]), l1_loss(pb["HH1"], gb["HH1"])
    return (lh1_weight * l_lh) + (hl1_weight * l_hl) + (0.5 * (lh1_weight + hl1_weight) * l_hh)


BAND_NAMES = ("LL1", "LH1", "HL1", "HH1", "LL2", "LH2", "HL2", "HH2")


def lgdwt_losses(pred, gt, band_weights=(1, 1, 1, 0, 0, 0, 0, 0), patch_size=128, percentile=0.2, w_lh=1.0, w_hl=1.0,
                 patch_enable=True):
    """The DWT part of one LGDWT-GS iteration (LG/train.py:131-180) on (C,H,W) or (1,C,H,W) tensors.
    Returns (dwt_loss, patch_loss, band_l1 dict, mask or None)."""
    pb = pred.unsqueeze(0) if pred.dim() == 3 else pred
    gb = gt.unsqueeze(0) if gt.dim() == 3 else gt
    p_bands, g_bands = get_dwt_subbands(pb), get_dwt_subbands(gb)
    band_l1 = {n: l1_loss(p_bands[n], g_bands[n]) for n in BAND_NAMES}
    total = 0.0
    for n, w in zip(BAND_NAMES, band_weights):
        if w != 0.0:
            total = total + w * band_l1[n]
    patch_loss, mask = torch.tensor(0.0, device=pb.device), None
    if patch_enable:
        elf = compute_elf_map(gb)
        patch_loss = compute_patch_dwt_loss(pb, gb, elf, patch_size, percentile, w_lh, w_hl)
        if pb.shape[-2] >= patch_size and pb.shape[-1] >= patch_size:
            mask = patch_selection(elf, patch_size, percentile)[0]
    return total, patch_loss, band_l1, mask
