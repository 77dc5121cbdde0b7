"""Multi-GPU plumbing: vehicles are independent, so a population is split into contiguous index
ranges, one process and one agf_batch per GPU, with NO communication inside the step loop.  The
only collective is the Monte-Carlo statistics read-out, and it lives in the C ABI
(agf_batch_reduce_stats_nccl: stats kernel + ONE ncclAllGather of 16 doubles per rank + combine, on
the batch's stream, replayed as a CUDA graph).  torch.distributed is the plumbing only: StatsComm
uses it to hand rank 0's NCCL unique id to the other ranks; combine_stats is the same reduction
through torch collectives for the CPU (gloo) tests of the host logic.
"""
import ctypes as C

import numpy as np

from . import _abi as abi

N_SUM = 14  # entries [0, 14) combine with SUM, [14, 16) with MAX (include/agrifly_b200.h AGF_ST_*)


def shard_range(n_total, rank, world):
    """Contiguous [first, first+count) of `rank`; the first n_total % world ranks get one more."""
    base, rem = divmod(int(n_total), int(world))
    first = rank * base + min(rank, rem)
    return first, base + (1 if rank < rem else 0)


def combine_stats(vec, dist=None, group=None):
    """All-reduce a stats vector (torch tensor, float64[16]) in place across ranks."""
    if dist is None or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return vec
    s = vec[:N_SUM].clone()
    m = vec[N_SUM:].clone()
    dist.all_reduce(s, op=dist.ReduceOp.SUM, group=group)
    dist.all_reduce(m, op=dist.ReduceOp.MAX, group=group)
    vec[:N_SUM] = s
    vec[N_SUM:] = m
    return vec


def combine_gathered(gathered):
    """What the library's combine kernel computes from the all-gathered vectors [nranks][16]: entries [0, 14) summed
    in rank order, [14, 16) maximised."""
    g = np.asarray(gathered, dtype=np.float64).reshape(-1, abi.STATS_LEN)
    out = g[0].copy()
    for q in range(1, len(g)):
        out[:N_SUM] = out[:N_SUM] + g[q, :N_SUM]
        out[N_SUM:] = np.maximum(out[N_SUM:], g[q, N_SUM:])
    return out


def exchange_unique_id(make_id, dist, group=None):
    """Rank 0 calls make_id() -> bytes; every rank returns those bytes (broadcast through torch.distributed, any backend)."""
    obj = [make_id() if dist.get_rank(group) == 0 else None]
    dist.broadcast_object_list(obj, src=0, group=group)
    return obj[0]


class StatsComm:
    """The library's own NCCL communicator for a batch: created from a unique id that rank 0 draws through the C ABI and
    torch.distributed carries to the other ranks.  One per process (= per GPU)."""

    def __init__(self, batch, dist, device, group=None):
        self.b, self.L = batch, batch.L
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)

        def make_id():
            buf = (C.c_uint8 * abi.NCCL_UNIQUE_ID_BYTES)()
            rc = self.L.agf_nccl_get_unique_id(buf)
            if rc != 0:
                raise RuntimeError("agf_nccl_get_unique_id: " + self.L.agf_last_error_string().decode(errors="replace"))
            return bytes(buf)

        uid = exchange_unique_id(make_id, dist, group)
        self.comm = C.c_void_p()
        buf = (C.c_uint8 * abi.NCCL_UNIQUE_ID_BYTES)(*uid)
        rc = self.L.agf_nccl_comm_init_rank(buf, self.world, self.rank, int(device), C.byref(self.comm))
        if rc != 0:
            raise RuntimeError("agf_nccl_comm_init_rank: " + self.L.agf_last_error_string().decode(errors="replace"))

    def reduce_device(self, dev_ptr):
        self.b.stats_nccl_device(self.comm, dev_ptr)

    def reduce_host(self):
        return self.b.stats_nccl(self.comm)

    def close(self):
        if self.comm:
            self.L.agf_nccl_comm_destroy(self.comm)
            self.comm = C.c_void_p()


def summarize_stats(v):
    """Human-readable Monte-Carlo summary from a (combined) stats vector."""
    v = np.asarray(v, dtype=np.float64)
    n = max(v[0] - v[9], 1.0)
    return dict(vehicles=int(v[0]), mean_error=[v[1] / n, v[2] / n, v[3] / n], rms_error=float(np.sqrt(v[4] / n)),
                mean_abs_error=v[5] / n, max_error=v[14], n_panic=int(v[6]), n_killed=int(v[7]),
                n_autonomous=int(v[8]), n_nonfinite=int(v[9]), mean_speed=v[10] / n,
                mean_estimator_error=v[11] / n, max_estimator_error=v[15])
