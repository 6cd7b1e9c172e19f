"""Wall-clock breakdown of one step (host + device) for ours and the reference mirror."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "sparse-view-3dgs-pack_b200")):
    sys.path.insert(0, p)
import bench
dev = torch.device("cuda", 0)
sc, cams, params = bench.make_workload(dev)
cam_devs = [bench.cam_dict(c, dev) for c in cams]
for impl in ("ours", "ref"):
    st = bench.Stepper(impl, params, dev, 1)
    for fixed in (True, False):
        for i in range(6):
            st.step_resident(cam_devs[0 if fixed else i % 8])
        torch.cuda.synchronize()
        tf = tb = 0.0
        n = 16
        t_all0 = time.perf_counter()
        for i in range(n):
            cam = cam_devs[0 if fixed else i % 8]
            st.zero_grads()
            torch.cuda.synchronize(); t0 = time.perf_counter()
            color, radii, invd = st.render(cam)
            torch.cuda.synchronize(); t1 = time.perf_counter()
            color.backward(st.dL)
            torch.cuda.synchronize(); t2 = time.perf_counter()
            tf += t1 - t0; tb += t2 - t1
        print(impl, "fixed_cam" if fixed else "cycling_cams", "fwd %.3f ms  bwd %.3f ms (wall, synced)" % (1e3 * tf / n, 1e3 * tb / n))
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for i in range(n):
            st.step_resident(cam_devs[0 if fixed else i % 8])
        torch.cuda.synchronize(); t1 = time.perf_counter()
        print(impl, "fixed_cam" if fixed else "cycling_cams", "pipelined step %.3f ms" % (1e3 * (t1 - t0) / n))
print(torch.cuda.memory_summary(abbreviated=True)[:1500])
