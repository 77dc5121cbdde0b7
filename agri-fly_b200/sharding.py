"""Multi-GPU plumbing: vehicles are independent, so a population is split into contiguous index
ranges, one process and one agf_batch per GPU, with NO communication inside the step loop.  The
only collective is the final Monte-Carlo statistics reduction: one all-reduce(SUM) and one
all-reduce(MAX) on the 16-double vector produced by the stats kernel (agf_batch_reduce_stats_device).
torch.distributed is used for the plumbing only (NCCL on GPUs, gloo in the CPU tests).
"""
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


def summarize_stats(v):
    """Human-readable Monte-Carlo summary from a (combined) stats vector."""
    v = np.asarray(v, dtype=np.float64)
    n = max(v[0] - v[9], 1.0)
    return dict(vehicles=int(v[0]), mean_error=[v[1] / n, v[2] / n, v[3] / n], rms_error=float(np.sqrt(v[4] / n)),
                mean_abs_error=v[5] / n, max_error=v[14], n_panic=int(v[6]), n_killed=int(v[7]),
                n_autonomous=int(v[8]), n_nonfinite=int(v[9]), mean_speed=v[10] / n,
                mean_estimator_error=v[11] / n, max_estimator_error=v[15])
