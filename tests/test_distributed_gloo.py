"""N>1 path on CPU: world_size-2 gloo run of the permutation-id sharding + all-gather of minima + epilogue."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from dual_threshold_optimization_b200.sharding import empirical_from_minima, gather_minima, shard_range


def test_shard_range_partitions_ids():
    for P in (0, 1, 7, 100, 100001):
        for world in (1, 2, 3, 8):
            got = []
            for r in range(world):
                first, cnt = shard_range(P, world, r)
                got.extend(range(first, first + cnt))
            assert got == list(range(P))
            sizes = [shard_range(P, world, r)[1] for r in range(world)]
            assert max(sizes) == -(-P // world) if P else max(sizes) == 0
    with pytest.raises(ValueError):
        shard_range(10, 2, 2)


def _fake_minp(ids):
    # deterministic stand-in for the per-permutation minimum (a pure function of the id, like the Philox path)
    x = (ids.astype(np.uint64) * np.uint64(2654435761)) % np.uint64(1 << 32)
    return (x.astype(np.float64) + 1.0) / float(1 << 32)


def _worker(rank, world, port, P, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    first, cnt = shard_range(P, world, rank)
    local = torch.from_numpy(_fake_minp(np.arange(first, first + cnt)))
    full = gather_minima(local, P)
    ok = np.array_equal(full.numpy(), _fake_minp(np.arange(P)))
    emp = empirical_from_minima(full.numpy(), 0.25)
    dist.barrier()
    dist.destroy_process_group()
    q.put((rank, ok, emp))


@pytest.mark.parametrize("P", [1001, 64])
def test_two_rank_gather_and_epilogue(P):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, P, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = empirical_from_minima(_fake_minp(np.arange(P)), 0.25)
    for rank, ok, emp in res:
        assert ok, f"rank {rank}: gathered minima out of id order"
        assert emp == want


def test_empirical_from_minima_rules():
    assert empirical_from_minima([], 0.1) == 1.0
    assert empirical_from_minima([0.10, 0.03, 0.07], 0.05) == 1 / 3
    assert empirical_from_minima([0.05], 0.05) == 1.0  # <= , no +1 correction
