"""Multi-GPU worker (launched by tests/test_trainer_gpu.py::test_peer_exchange_two_gpus and by hand:
   python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29641 tests/peer_worker.py [--time])
Checks the fused peer-memory exchange (csrc/peer.cu) against NCCL all-reduce + the local fused Adam, at kernel level and
through dp.ViewParallelTrainer (incl. a densification), and optionally times both at the benchmark's bucket size."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "sparse-view-3dgs-pack_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
from lgdwt_b200 import dp, scenes  # noqa: E402
from lgdwt_b200.peer import PeerExchange  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", device_id=dev)
    say = (lambda *a: print(*a, flush=True)) if rank == 0 else (lambda *a: None)

    # ---- kernel level: all-reduce and reduce+Adam+gather vs NCCL + local Adam, both ways of sharing the block
    P = 10_001
    cfg = dp.AdamConfig()
    for backend in ("nvls", "ipc"):
      g_ref, g_peer = dp.FlatGaussians(P, dev), dp.FlatGaussians(P, dev)
      init = torch.randn(g_ref.data.shape, device=dev, generator=torch.Generator(device=dev).manual_seed(1))
      g_ref.data.copy_(init)
      g_peer.data.copy_(init)
      ex, why = PeerExchange.create(g_peer.data.numel(), dev, backends=(backend,))
      if ex is None:
          say("PEER_BACKEND_UNAVAILABLE", why)
          assert backend == "nvls", why   # NVLS needs an NVSwitch; plain IPC must work wherever there is peer access
          continue
      g_peer.adopt(ex.param, ex.grad)
      for step in range(1, 4):
          grad = torch.randn(init.shape, device=dev, generator=torch.Generator(device=dev).manual_seed(100 * step + rank))
          # all-reduce alone
          ex.grad.copy_(grad)
          ex.allreduce(0.5)
          want = grad.clone()
          dist.all_reduce(want)
          torch.testing.assert_close(ex.grad, want * 0.5, rtol=1e-6, atol=1e-6)
          # fused exchange + optimizer
          g_ref.grad.copy_(grad)
          dist.all_reduce(g_ref.grad)
          g_ref.adam_step(cfg, grad_scale=0.25)
          ex.grad.copy_(grad)
          g_peer.step_count += 1
          ex.reduce_adam(g_peer.exp_avg, g_peer.exp_avg_sq, g_peer.adam_segments(cfg), cfg, g_peer.step_count, 0.25)
          ex.check()
          torch.testing.assert_close(g_peer.data, g_ref.data, rtol=1e-6, atol=1e-7)
      # every replica holds bit-identical parameters (each element is computed by exactly one rank)
      lo, hi = g_peer.data.clone(), g_peer.data.clone()
      dist.all_reduce(lo, op=dist.ReduceOp.MIN)
      dist.all_reduce(hi, op=dist.ReduceOp.MAX)
      assert torch.equal(lo, hi)
      say("PEER_KERNELS_OK world", world, "backend", ex.backend)

    # ---- trainer level: nccl vs peer exchange on the same views, with a densification in between
    sc = scenes.trained_like_scene(20_000, seed=5, log_scale_mean=np.log(0.02))
    cams = [dp.camera_to_device(c, dev) for c in scenes.orbit_cameras(4, 160, 128)]
    gts = [torch.rand(3, 128, 160, device=dev, generator=torch.Generator(device=dev).manual_seed(v)) for v in range(4)]
    dcfg = dp.DensifyConfig(densify_from_iter=1, densify_until_iter=50, densification_interval=3,
                            opacity_reset_interval=1000, densify_grad_threshold=1e-7, cameras_extent=4.0)
    bg = torch.zeros(3, device=dev)
    out = {}
    for mode in ("nccl", "peer"):
        g = dp.FlatGaussians.from_scene(sc, dev)
        tr = dp.ViewParallelTrainer(g, densify=dcfg, seed=3, exchange=mode)
        assert (tr.peer is not None) == (mode == "peer"), tr.peer_unavailable
        for _ in range(5):
            tr.step(cams, gts, bg)
            assert tr.replicas_in_sync()
        if tr.peer is not None:
            tr.peer.check()
        out[mode] = (g.P, g.data[: g.floats * g.stride].clone())
    assert out["nccl"][0] == out["peer"][0] != 20_000, (out["nccl"][0], out["peer"][0])
    err = float((out["nccl"][1] - out["peer"][1]).abs().max())
    assert err < 5e-4, err   # Adam's sign-like first steps amplify last-bit differences of the summed gradients
    say("PEER_TRAINER_OK P", out["peer"][0], "max |param diff| nccl vs peer", err)

    if "--time" in sys.argv:
      n = 59 * 1_000_000
      bucket = torch.randn(n, device=dev)
      for backend in ("nvls", "ipc"):
        g = dp.FlatGaussians(1_000_000, dev)
        ex, why = PeerExchange.create(g.data.numel(), dev, backends=(backend,))
        if ex is None:
            continue
        g.adopt(ex.param, ex.grad)

        def timeit(fn, reps=20):
            for _ in range(3):
                fn()
            dist.barrier()
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(reps):
                fn()
            b.record()
            torch.cuda.synchronize()
            t = torch.tensor([a.elapsed_time(b) / reps], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t)

        def nccl_path():
            dist.all_reduce(g.grad)
            g.adam_step(cfg)

        def peer_path():
            g.step_count += 1
            ex.reduce_adam(g.exp_avg, g.exp_avg_sq, g.adam_segments(cfg), cfg, g.step_count, 1.0)

        from lgdwt_b200 import _lib

        def kernel_only():   # the fused kernel without its two barriers (timing only; ranks are not ordered)
            ends, lr_a, lr_b, width, split = g.adam_segments(cfg)
            _lib.lib.lg_peer_reduce_adam(ex.rank, ex.world, ex._grads, ex._params, g.exp_avg.data_ptr(),
                                         g.exp_avg_sq.data_ptr(), ex.n, len(ends), ends, lr_a, lr_b, width, split,
                                         cfg.beta1, cfg.beta2, cfg.eps, 5, 1.0, _lib.stream_ptr(dev))

        t_bar = timeit(lambda: (ex.barrier(), ex.barrier()))
        t_ker = timeit(kernel_only)
        say("PEER_PARTS two barriers %.4f ms | fused kernel alone %.4f ms" % (t_bar, t_ker))
        t_ar = timeit(lambda: dist.all_reduce(bucket))
        t_par = timeit(lambda: ex.allreduce(1.0))
        t_nccl = timeit(nccl_path)
        t_peer = timeit(peer_path)
        ex.check()
        say("PEER_TIMING %s world %d bucket %.0f MB: nccl all-reduce %.3f ms | peer all-reduce %.3f ms | nccl all-reduce + "
            "Adam %.3f ms | fused peer reduce+Adam+gather %.3f ms" % (backend, world, n * 4 / 1e6, t_ar, t_par, t_nccl, t_peer))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
