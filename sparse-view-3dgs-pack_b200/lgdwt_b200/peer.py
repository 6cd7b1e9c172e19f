"""The data-parallel exchange step over NVLink peer memory (host side of csrc/peer.cu).

One process per GPU.  Every rank allocates one cudaMalloc block [flags | gradient bucket | parameters], exports it
with CUDA IPC, and maps the blocks of all other ranks of the node; `PeerExchange.reduce_adam` then runs the fused
reduce-scatter + Adam + all-gather kernel between two flag barriers.  The NCCL all-reduce + local Adam path of
`dp.ViewParallelTrainer` stays available (`exchange="nccl"`) and is what runs when peer mapping is impossible
(different nodes, IPC disabled) — `PeerExchange.create` then returns None and says why.
"""
import ctypes

import torch
import torch.distributed as dist

from . import _lib

_FLAG_BYTES = 256


class _DeviceArray:
    """__cuda_array_interface__ view of raw device memory (keeps the owning exchange alive)"""

    def __init__(self, ptr, n_floats, owner):
        self.__cuda_array_interface__ = {"shape": (int(n_floats),), "typestr": "<f4", "data": (int(ptr), False),
                                         "version": 2}
        self._owner = owner


class _Block:
    """owns the cudaMalloc block and the peer mappings; freed when the last tensor view dies"""

    def __init__(self):
        self.base, self.mapped = None, []

    def __del__(self):
        try:
            for m in self.mapped:
                _lib.lib.lg_peer_close(m)
            if self.base:
                _lib.lib.lg_peer_free(self.base)
        except Exception:
            pass


class PeerExchange:
    def __init__(self, n_floats, device, group=None):
        if n_floats % 4:
            raise ValueError("PeerExchange: the flat buffer must hold a multiple of 4 floats")
        self.n, self.device, self.group = int(n_floats), device, group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.epoch = 0
        blk = self._blk = _Block()
        nbytes = _FLAG_BYTES + 2 * 4 * self.n
        base = ctypes.c_void_p()
        with torch.cuda.device(device):
            _lib.check(_lib.lib.lg_peer_alloc(nbytes, ctypes.byref(base)), RuntimeError)
            blk.base = base.value
            handle = ctypes.create_string_buffer(64)
            _lib.check(_lib.lib.lg_peer_export(blk.base, handle), RuntimeError)
            handles = [None] * self.world
            dist.all_gather_object(handles, bytes(handle.raw), group=group)
            bases = []
            for r, h in enumerate(handles):
                if r == self.rank:
                    bases.append(blk.base)
                    continue
                m = ctypes.c_void_p()
                _lib.check(_lib.lib.lg_peer_open(h, ctypes.byref(m)), RuntimeError)
                blk.mapped.append(m.value)
                bases.append(m.value)
        arr = lambda vals: (ctypes.c_void_p * self.world)(*vals)
        self._flags = arr(bases)
        self._grads = arr([b + _FLAG_BYTES for b in bases])
        self._params = arr([b + _FLAG_BYTES + 4 * self.n for b in bases])
        self.grad = torch.as_tensor(_DeviceArray(blk.base + _FLAG_BYTES, self.n, blk), device=device)
        self.param = torch.as_tensor(_DeviceArray(blk.base + _FLAG_BYTES + 4 * self.n, self.n, blk), device=device)
        dist.barrier(group=group)  # every rank has mapped every block before anyone uses (or frees) one

    @classmethod
    def create(cls, n_floats, device, group=None):
        """-> (exchange or None, reason).  Collective: every rank of the group must call it; all ranks agree on the
        outcome (a rank that cannot map its peers makes everyone fall back)."""
        ex, why = None, ""
        try:
            ex = cls(n_floats, device, group)
        except Exception as e:  # IPC unavailable / no peer access: all ranks then take the NCCL path
            why = "%s: %s" % (type(e).__name__, e)
        ok = torch.tensor([1 if ex is not None else 0], device=device)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
        if int(ok.item()) == 0:
            return None, why or "a peer rank could not map the shared buffers"
        return ex, ""

    def barrier(self):
        self.epoch += 1
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib.lg_peer_barrier(self.rank, self.world, self._flags, self.epoch,
                                                _lib.stream_ptr(self.device)), RuntimeError)

    def check(self):
        """host sync: raises if a barrier timed out since the last check"""
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib.lg_peer_check(_lib.stream_ptr(self.device)), RuntimeError)

    def reduce_adam(self, exp_avg, exp_avg_sq, segments, cfg, step, grad_scale=1.0):
        """barrier | fused reduce-scatter + Adam (this rank's shard) + parameter all-gather | barrier.
        segments = (ends, lr_a, lr_b, width, split) ctypes arrays as dp.FlatGaussians.adam_segments builds them."""
        ends, lr_a, lr_b, width, split = segments
        self.barrier()
        with torch.cuda.device(self.device):
            rc = _lib.lib.lg_peer_reduce_adam(self.rank, self.world, self._grads, self._params, exp_avg.data_ptr(),
                                              exp_avg_sq.data_ptr(), self.n, len(ends), ends, lr_a, lr_b, width, split,
                                              cfg.beta1, cfg.beta2, cfg.eps, int(step), float(grad_scale),
                                              _lib.stream_ptr(self.device))
        _lib.check(rc, RuntimeError)
        self.barrier()

    def allreduce(self, grad_scale=1.0):
        """barrier | every rank sums its shard over all ranks and stores it into every gradient buffer | barrier"""
        self.barrier()
        with torch.cuda.device(self.device):
            rc = _lib.lib.lg_peer_allreduce(self.rank, self.world, self._grads, self.n, float(grad_scale),
                                            _lib.stream_ptr(self.device))
        _lib.check(rc, RuntimeError)
        self.barrier()
