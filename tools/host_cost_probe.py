"""Host cost of one view (forward + backward through the public operator with GradSinks) when the GPU work is negligible:
a 2000-Gaussian scene at 64x64.  If this is close to the device time per view of the real workload, the host paces the
step and no stream overlap can help."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "sparse-view-3dgs-pack_b200")):
    sys.path.insert(0, p)
import bench
from lgdwt_b200 import scenes

dev = torch.device("cuda", 0)
bench.P_GAUSSIANS, bench.WIDTH, bench.HEIGHT = 2000, 64, 64
sc = scenes.trained_like_scene(2000, seed=1)
cams = scenes.orbit_cameras(4, 64, 64)
t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
params = {k: t(getattr(sc, k)).requires_grad_(True) for k in ("means3D", "shs", "opacities", "scales", "rotations")}
cam_devs = [bench.cam_dict(c, dev) for c in cams]
for mode, vs in (("sinks", 0), ("sinks", 2), ("dropin", 0)):
    st = bench.Stepper(mode, params, dev, 1, view_streams=vs)
    for _ in range(20):
        st.step_resident(cam_devs)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    n = 200
    for _ in range(n):
        st.step_resident(cam_devs)
    torch.cuda.synchronize()
    print("%s view_streams=%d: %.1f us of host time per view" % (mode, vs, (time.perf_counter() - t0) / (4 * n) * 1e6))
