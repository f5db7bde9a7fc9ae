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


def test_persistence_round_trip_and_retry_list(tmp_path):
    """results in the on-disk shape of monte_carlo.jl:334-343 (same container and dataset names, .npz
    instead of HDF5) and the `retry` list of monte_carlo.jl:269 (host logic only: no GPU needed)."""
    import numpy as np
    from tortoisesat.jl_b200 import host
    n = 4
    rng = np.random.default_rng(0)
    res = dict(A=rng.random((n, 6)), t_final=np.array([100., 200., 300., 50.]), slew_time=np.array([50., 200., 120., 50.]),
               fails=np.array([0., 1., 0., 1.]), outcomes=np.zeros(n, dtype=host.OUTCOME_DTYPE),
               sim_states=[rng.random((8, 5 + i)) for i in range(n)], sim_control_inputs=[rng.random((3, 5 + i)) for i in range(n)],
               B_ECI_total=[rng.random((10 + 2 * i, 3)) for i in range(n)], t_total=[np.arange(6 + i) * 0.2 for i in range(n)])
    retry = host.save_monte_carlo(res, str(tmp_path), prefix="100")
    assert retry.tolist() == [2, 4]                                     # findall(x -> x == 1., fails), 1-based
    for name in ("100_A", "100_states_1", "100_control_4", "100_B_N_2", "100_t_total_3", "100_summary"):
        assert (tmp_path / (name + ".npz")).exists()
    assert set(np.load(tmp_path / "100_states_2.npz").files) == {"one_state", "states"}
    back = host.load_monte_carlo(str(tmp_path), prefix="100")
    assert back["retry"].tolist() == [2, 4] and np.array_equal(back["A"], res["A"])
    for k in ("sim_states", "sim_control_inputs", "B_ECI_total", "t_total"):
        assert all(np.array_equal(a, b) for a, b in zip(back[k], res[k]))
