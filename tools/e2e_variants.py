"""Where does the e2e step differ from the resident step?  Times harness variants of bench.Stepper (ours only)."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "sparse-view-3dgs-pack_b200")):
    sys.path.insert(0, p)
import bench
dev = torch.device("cuda", 0)
sc, cams, params = bench.make_workload(dev)
V = bench.VIEWS_PER_RANK
cam_devs = [bench.cam_dict(c, dev) for c in cams]
host_cams = []
for c in cams:
    d = bench.cam_dict(c, "cpu")
    d["packed"] = torch.cat([d["viewmatrix"].reshape(-1), d["projmatrix"].reshape(-1), d["campos"].reshape(-1)]).pin_memory()
    host_cams.append(d)
rng = np.random.default_rng(7)
host_gts = [torch.from_numpy(rng.random((3, bench.HEIGHT, bench.WIDTH)).astype(np.float32)).pin_memory() for _ in range(8)]
st = bench.Stepper("sinks", params, dev, 1)
ring = lambda i: [(i * V + v) % 8 for v in range(V)]
dev_gts = [g.to(dev) for g in host_gts]
state = {"n": 0}

def variant(name, upload_cam, upload_gt, loss_kind, read, spin_us=0):
    def step(i):
        st.begin_step()
        main = torch.cuda.current_stream(dev)
        total = torch.zeros((), device=dev) if loss_kind == "l1" else None
        for v, j in enumerate(ring(i)):
            if upload_cam:
                cam = dict(host_cams[j])
                pk = host_cams[j]["packed"].to(dev, non_blocking=True)
                cam["viewmatrix"], cam["projmatrix"], cam["campos"] = pk[:16].view(4, 4), pk[16:32].view(4, 4), pk[32:35]
            else:
                cam = cam_devs[j]
            gt = dev_gts[j]
            if upload_gt:
                gt = st.dev_gt[state["n"] % 2][v]
                with torch.cuda.stream(st.copy_stream):
                    gt.copy_(host_gts[j], non_blocking=True)
                    st.copy_done[v].record(st.copy_stream)
            if spin_us:
                t_end = time.perf_counter() + spin_us * 1e-6
                while time.perf_counter() < t_end:
                    pass
            color, radii, invd = st.render(cam, v == 0)
            if upload_gt:
                main.wait_event(st.copy_done[v])
            if loss_kind == "l1":
                loss = st.l1(color, gt)
                loss.backward()
                total += loss.detach()
            else:
                color.backward(st.dL)
        if read and total is not None:
            slot = state["n"] % 2
            st.host_loss[slot].copy_(total, non_blocking=True)
            st.loss_ready[slot].record(main)
            if state["n"] >= 1:
                st.loss_ready[1 - slot].synchronize()
        state["n"] += 1
    for i in range(4):
        step(i)
    ms = bench.timed_loop(step, 15, 1, dev) / 15 / V
    print("%-60s %.4f ms/view" % (name, ms), flush=True)

variant("resident: dev cam, backward(dL)", False, False, "dl", False)
variant("e2e", True, True, "l1", True)
variant("e2e again", True, True, "l1", True)
smp = bench.ClockSampler(0)
smp.start()
time.sleep(0.5)
variant("resident with nvidia-smi -lms 100 running", False, False, "dl", False)
variant("e2e with nvidia-smi -lms 100 running", True, True, "l1", True)
variant("e2e with nvidia-smi -lms 100 running, again", True, True, "l1", True)
print(smp.stop())
