"""CPU, world_size 2, gloo: trial sharding + outcome gather + statistics reduction."""
import os
import socket
import sys

import numpy as np
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_total, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from tortoisesat.jl_b200 import host, parallel
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = parallel.shard_range(n_total, rank, world)
    out = np.zeros(hi - lo, dtype=host.OUTCOME_DTYPE)
    out["N"] = np.arange(lo, hi)
    out["status"] = rank
    out["J"] = 1.5 * np.arange(lo, hi)
    full = parallel.gather_outcomes(out)
    vec = np.arange(len(parallel.STAT_FIELDS), dtype=float) * (rank + 1)
    red = parallel.reduce_stats(vec)
    if rank == 0:
        q.put((full["N"].tolist(), full["status"].tolist(), full["J"].tolist(), red.tolist()))
    dist.destroy_process_group()


def test_shard_gather_reduce_world2():
    world, n_total = 2, 10
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker, args=(r, world, port, n_total, q)) for r in range(world)]
    for p in ps:
        p.start()
    N, status, J, red = q.get(timeout=120)
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert N == list(range(10))
    assert status == [0] * 5 + [1] * 5
    assert J == [1.5 * i for i in range(10)]
    assert red == [3.0 * i for i in range(11)]


def test_shard_helpers():
    sys.path.insert(0, ROOT)
    from tortoisesat.jl_b200 import parallel
    cover = []
    for r in range(8):
        lo, hi = parallel.shard_range(65536 + 3, r, 8)
        cover += list(range(lo, hi))
    assert cover == list(range(65536 + 3))
    il = np.sort(np.concatenate([parallel.interleaved_shard(101, r, 4) for r in range(4)]))
    assert il.tolist() == list(range(101))
