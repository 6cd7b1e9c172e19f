"""Why is bench.py's e2e slower than tools/e2e_variants.py's?  Calls bench.measure() the way main() does, piece by piece."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "sparse-view-3dgs-pack_b200")):
    sys.path.insert(0, p)
import bench
from lgdwt_b200 import _lib
dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
sc, cams, params = bench.make_workload(dev)
cam_devs = [bench.cam_dict(c, dev) for c in cams]
host_cams, host_gts = bench.host_inputs(cams)
K, W, V = 20, 5, bench.VIEWS_PER_RANK
st = bench.Stepper("sinks", params, dev, 1)
for label, timing in (("plain", None), ("stage timing as in main()", "yes"), ("plain again", None)):
    rows = []
    def rd():
        _lib.stage_timing(0)
    t = (_lib.stage_timing, rd) if timing else None
    a, b, n = bench.measure(st, cam_devs, host_cams, host_gts, K, W, 1, 0, dev, None, t, _lib.lib.lg_launch_count)
    print("%-30s resident %.4f  e2e %.4f ms/view" % (label, a / V, b / V), flush=True)
st2 = bench.Stepper("sinks", params, dev, 1)
a, b, n = bench.measure(st2, cam_devs, host_cams, host_gts, K, W, 1, 0, dev, None, None, None)
print("%-30s resident %.4f  e2e %.4f ms/view" % ("second Stepper", a / V, b / V), flush=True)
# e2e alone, long
ring = lambda i: [((i) * V + v) % len(host_cams) for v in range(V)]
fn = lambda i: st2.step_e2e([host_cams[j] for j in ring(i)], [host_gts[j] for j in ring(i)])
for K2 in (20, 60):
    ms = bench.timed_loop(fn, K2, 1, dev) / K2 / V
    print("e2e alone K=%d: %.4f ms/view" % (K2, ms), flush=True)
