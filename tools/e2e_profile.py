"""Host-side timeline of the bench's e2e step (where does the wall time of one step go?)."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "sparse-view-3dgs-pack_b200")):
    sys.path.insert(0, p)
import bench
dev = torch.device("cuda", 0)
sc, cams, params = bench.make_workload(dev)
host_cams = []
for c in cams:
    d = bench.cam_dict(c, "cpu")
    for k in ("viewmatrix", "projmatrix", "campos"):
        d[k] = d[k].pin_memory()
    host_cams.append(d)
rng = np.random.default_rng(7)
host_gts = [torch.from_numpy(rng.random((3, bench.HEIGHT, bench.WIDTH)).astype(np.float32)).pin_memory() for _ in range(2)]
dev_gt = torch.empty((3, bench.HEIGHT, bench.WIDTH), device=dev)
for impl in ("ours", "ref"):
    st = bench.Stepper(impl, params, dev, 1)
    for i in range(5):
        st.step_e2e(host_cams[i % 8], host_gts[i % 2], dev_gt)
    torch.cuda.synchronize()
    names = ["zero_grads", "h2d_issue", "render_call", "loss_fwd", "backward_call", "item"]
    acc = np.zeros(len(names))
    n = 20
    t_start = time.perf_counter()
    for i in range(n):
        host_cam, host_gt = host_cams[i % 8], host_gts[i % 2]
        t = [time.perf_counter()]
        st.zero_grads(); t.append(time.perf_counter())
        cam = dict(host_cam)
        for k in ("viewmatrix", "projmatrix", "campos"):
            cam[k] = host_cam[k].to(dev, non_blocking=True)
        dev_gt.copy_(host_gt, non_blocking=True); t.append(time.perf_counter())
        color, radii, invd = st.render(cam); t.append(time.perf_counter())
        loss = (color - dev_gt).abs().mean(); t.append(time.perf_counter())
        loss.backward(); t.append(time.perf_counter())
        v = float(loss.item()); t.append(time.perf_counter())
        acc += np.diff(t)
    total = (time.perf_counter() - t_start) / n
    print(impl, "step %.3f ms:" % (1e3 * total), "  ".join("%s %.3f" % (k, 1e3 * a / n) for k, a in zip(names, acc)))
