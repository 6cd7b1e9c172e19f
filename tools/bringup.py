"""Bring-up / quick timing: ours vs the reference shim on the metric scene.  Not part of the test-suite."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "sparse-view-3dgs-pack_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import helpers  # noqa: E402
from lgdwt_b200 import scenes  # noqa: E402


def timeit(fn, warm=3, iters=10):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2], ts[0]


def main():
    P = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
    sc = scenes.trained_like_scene(P, seed=1)
    cam = scenes.metric_camera()
    t, c = helpers.scene_to_torch(sc), helpers.cam_to_torch(cam)
    bg = torch.zeros(3, device="cuda")
    dL = torch.randn((3, cam.image_height, cam.image_width), device="cuda")
    dLd = torch.randn((1, cam.image_height, cam.image_width), device="cuda")
    ours = helpers.run_ours(t, c, cam, bg, want_state=True)
    print("ours: num_rendered", ours["num_rendered"], "visible", int((ours["radii"] > 0).sum()),
          "n_trav", int(ours["n_contrib"].sum()))
    have_ref = helpers.load_ref() is not None
    if have_ref:
        ref = helpers.run_ref(t, c, cam, bg, want_state=True)
        print("ref : num_rendered", ref["num_rendered"])
        for k in ("radii", "tiles_touched", "point_offsets", "point_list_keys", "point_list", "ranges", "n_contrib", "final_T"):
            a, b = ours[k], ref[k]
            if a.dtype == torch.float32:
                a, b = a.view(torch.int32), b.view(torch.int32)
            print("  %-16s equal=%s ndiff=%d/%d" % (k, bool((a == b).all()) if a.shape == b.shape else "SHAPE", int((a != b).sum()) if a.shape == b.shape else -1, a.numel()))
        print("  image max abs err", float((ours["color"] - ref["color"]).abs().max()))
    f_ours = lambda: helpers.run_ours(t, c, cam, bg, want_state=False)
    print("ours fwd  ms (median,min):", timeit(f_ours))
    fo = f_ours()
    print("ours bwd  ms:", timeit(lambda: helpers.backward_ours(t, c, cam, bg, fo, dL, dLd)))
    print("ours bwd (no invdepth) ms:", timeit(lambda: helpers.backward_ours(t, c, cam, bg, fo, dL, None)))
    if have_ref:
        f_ref = lambda: helpers.run_ref(t, c, cam, bg, want_state=False)
        print("ref  fwd  ms:", timeit(f_ref))
        fr = f_ref()
        print("ref  bwd  ms:", timeit(lambda: helpers.backward_ref(t, c, cam, bg, fr, dL, dLd)))


if __name__ == "__main__":
    main()
