"""Host-side multi-GPU logic on CPU: world_size-2 gloo run of the sharding + statistics reduction."""
import os
import socket
import sys

import numpy as np
import pytest

from conftest import ROOT


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_total, q):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    import agrifly_b200
    from agrifly_b200 import sharding
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    first, count = sharding.shard_range(n_total, rank, world)
    # a fake per-shard statistics vector, as the stats kernel would produce it
    e = np.arange(first, first + count, dtype=np.float64) * 1e-3
    v = np.zeros(16)
    v[0] = count
    v[4] = np.sum(e * e)
    v[5] = np.sum(e)
    v[14] = e.max()
    v[15] = 2 * e.max()
    t = torch.from_numpy(v.copy())
    sharding.combine_stats(t, dist)
    # the read-out as the C ABI does it on the GPUs: ONE all-gather of the 16 doubles of every rank, combined in rank order
    # (agf_batch_reduce_stats_nccl); and the hand-over of rank 0's communicator id through torch.distributed
    g = [torch.zeros(16, dtype=torch.float64) for _ in range(world)]
    dist.all_gather(g, torch.from_numpy(v.copy()))
    gathered = sharding.combine_gathered(np.stack([x.numpy() for x in g]))
    uid = sharding.exchange_unique_id(lambda: bytes(range(128)), dist)
    q.put((rank, first, count, t.numpy().copy(), gathered, uid))
    dist.barrier()
    dist.destroy_process_group()


def test_shard_ranges_partition_the_population(agf):
    from agrifly_b200 import sharding
    for n, w in [(1 << 20, 8), (1000, 3), (7, 8), (4096, 1)]:
        parts = [sharding.shard_range(n, r, w) for r in range(w)]
        assert parts[0][0] == 0 and sum(c for _, c in parts) == n
        for (f0, c0), (f1, _) in zip(parts, parts[1:]):
            assert f0 + c0 == f1
        assert max(c for _, c in parts) - min(c for _, c in parts) <= 1


def test_stats_allreduce_gloo_world2(agf):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    n_total = 1001
    ps = [ctx.Process(target=_worker, args=(r, 2, port, n_total, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = [q.get(timeout=120) for _ in ps]
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    e = np.arange(n_total, dtype=np.float64) * 1e-3
    for rank, first, count, v, gathered, uid in res:
        assert np.array_equal(gathered, v)      # one all-gather + rank-order combine == the two all-reduces
        assert uid == bytes(range(128))         # every rank holds rank 0's id
        assert v[0] == n_total
        assert abs(v[4] - np.sum(e * e)) < 1e-9 and abs(v[5] - np.sum(e)) < 1e-9
        assert v[14] == e.max() and v[15] == 2 * e.max()
    from agrifly_b200 import sharding
    s = sharding.summarize_stats(res[0][3])
    assert s["vehicles"] == n_total and abs(s["rms_error"] - np.sqrt(np.mean(e * e))) < 1e-12
