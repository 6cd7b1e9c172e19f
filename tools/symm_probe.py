"""probe: does torch's symmetric memory (multicast / NVLS mapping) work on this box?"""
import os, torch, torch.distributed as dist
rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
dev = torch.device("cuda", local); torch.cuda.set_device(dev)
dist.init_process_group("nccl", device_id=dev)
try:
    import torch.distributed._symmetric_memory as symm
    t = symm.empty(1024 * 1024, dtype=torch.float32, device=dev)
    h = symm.rendezvous(t, dist.group.WORLD.group_name)
    print(rank, "OK buffer_ptrs", [hex(p) for p in h.buffer_ptrs], "multicast", hex(h.multicast_ptr) if h.multicast_ptr else None,
          "signal_pad", [hex(p) for p in h.signal_pad_ptrs][:2], "attrs", [a for a in dir(h) if not a.startswith("_")], flush=True)
except Exception as e:
    import traceback; traceback.print_exc()
    print(rank, "FAILED", type(e).__name__, e, flush=True)
dist.barrier(); dist.destroy_process_group()
