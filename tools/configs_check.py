"""BASELINE configs 3/4/5 at full size: ours vs the reference shim (bit-exact state, image error, fwd/bwd ms).
Not part of the test-suite; run on a GPU box:  python tools/configs_check.py [cfg3|cfg4|cfg5|all]"""
import math
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "sparse-view-3dgs-pack_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import helpers  # noqa: E402
from lgdwt_b200 import scenes  # noqa: E402
from bringup import timeit  # noqa: E402


cfg = scenes.baseline_config


def main():
    names = sys.argv[1:] or ["all"]
    if names == ["all"]:
        names = ["cfg2", "cfg3", "cfg4", "cfg5"]
    for name in names:
        sc, cam = cfg(name)
        t, c = helpers.scene_to_torch(sc), helpers.cam_to_torch(cam)
        bg = torch.zeros(3, device="cuda")
        dL = torch.randn((3, cam.image_height, cam.image_width), device="cuda")
        ours = helpers.run_ours(t, c, cam, bg, want_state=True)
        print("== %s: P=%d %dx%d num_rendered=%d visible=%d" % (name, sc.means3D.shape[0], cam.image_width,
                                                               cam.image_height, ours["num_rendered"],
                                                               int((ours["radii"] > 0).sum())))
        have_ref = helpers.load_ref() is not None
        if have_ref:
            ref = helpers.run_ref(t, c, cam, bg, want_state=True)
            bad = []
            for k in ("radii", "tiles_touched", "point_list_keys", "point_list", "ranges", "n_contrib", "final_T"):
                a, b = ours[k], ref[k]
                if a.dtype == torch.float32:
                    a, b = a.view(torch.int32), b.view(torch.int32)
                if a.shape != b.shape or not bool((a == b).all()):
                    bad.append(k)
            print("   bit-exact vs reference:", "ALL EQUAL" if not bad else "DIFFERS: %s" % bad,
                  "| image max abs err %.3g" % float((ours["color"] - ref["color"]).abs().max()))
            del ref
        f_ours = lambda: helpers.run_ours(t, c, cam, bg, want_state=False)
        fo = f_ours()
        print("   ours fwd %.3f ms, bwd %.3f ms" % (timeit(f_ours)[0], timeit(lambda: helpers.backward_ours(t, c, cam, bg, fo, dL, None))[0]))
        if have_ref:
            f_ref = lambda: helpers.run_ref(t, c, cam, bg, want_state=False)
            fr = f_ref()
            g_o = helpers.backward_ours(t, c, cam, bg, fo, dL, None)
            g_r = helpers.backward_ref(t, c, cam, bg, fr, dL, None)
            worst, detail = 0.0, []
            for k in ("dL_dmean3D", "dL_dsh", "dL_dopacity", "dL_dscale", "dL_drot"):
                if g_o.get(k) is None or g_r.get(k) is None:
                    continue
                a, b = g_o[k].double().reshape(-1), g_r[k].double().reshape(-1)
                e = float((a - b).abs().max() / b.abs().max().clamp_min(1e-12))
                detail.append("%s %.2g (|ref|max %.2g)" % (k[3:], e, float(b.abs().max())))
                if float(b.abs().max()) > 1e-6:  # e.g. dL/drot of isotropic Gaussians is rounding noise on both sides
                    worst = max(worst, e)
            print("   per-tensor rel err:", "; ".join(detail))
            print("   ref  fwd %.3f ms, bwd %.3f ms | worst gradient rel err %.3g" % (
                timeit(f_ref)[0], timeit(lambda: helpers.backward_ref(t, c, cam, bg, fr, dL, dL[:1].contiguous()))[0], worst))
        if name == "cfg4":  # RGB+NIR: one 4-channel pass of ours vs two 3-channel passes of the reference (render + render_nir)
            P = sc.means3D.shape[0]
            col4 = torch.rand((P, 4), device="cuda")
            bg4 = torch.zeros(4, device="cuda")
            dL4 = torch.randn((4, cam.image_height, cam.image_width), device="cuda")
            f4 = lambda: helpers.run_ours(t, c, cam, bg4, colors_precomp=col4, want_state=False)
            fo4 = f4()
            t_f = timeit(f4)[0]
            t_b = timeit(lambda: helpers.backward_ours(t, c, cam, bg4, fo4, dL4, None, colors_precomp=col4))[0]
            print("   RGB+NIR ours, ONE 4-channel pass: fwd %.3f ms, bwd %.3f ms" % (t_f, t_b))
            if have_ref:
                col3 = col4[:, :3].contiguous()
                fr3 = lambda: helpers.run_ref(t, c, cam, bg, colors_precomp=col3, want_state=False)
                fr = fr3()
                r_f = timeit(fr3)[0]
                r_b = timeit(lambda: helpers.backward_ref(t, c, cam, bg, fr, dL, dL[:1].contiguous(), colors_precomp=col3))[0]
                print("   RGB+NIR reference, TWO 3-channel passes: fwd %.3f ms, bwd %.3f ms" % (2 * r_f, 2 * r_b))
        del ours, fo
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
