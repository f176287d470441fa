"""CPU tests of the multi-rank host logic (gloo, world_size 2): shard partitioning, the
cell->GPU assignment and the final result gather."""
import os

import numpy as np
import pytest

from openair4g_b200 import sharding


def test_shard_range_partitions_exactly():
    for n in (0, 1, 7, 13, 64, 23680, 65536):
        for world in (1, 2, 3, 4, 8):
            parts = [sharding.shard_range(n, world, r) for r in range(world)]
            assert parts[0][0] == 0 and parts[-1][1] == n
            assert all(parts[i][1] == parts[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in parts]
            assert max(sizes) - min(sizes) <= 1


def test_assign_by_cell_keeps_cells_together_and_balances():
    rng = np.random.default_rng(0)
    cells = list(rng.integers(0, 64, size=5000))           # 64 cells, config 4 of BASELINE.json
    for world in (2, 4, 8):
        ranks = sharding.assign_by_cell(cells, world)
        owner = {}
        for c, r in zip(cells, ranks):
            assert owner.setdefault(c, r) == r
        load = np.bincount(ranks, minlength=world)
        assert load.max() - load.min() <= max(np.bincount(cells))


def test_assign_by_cell_follows_rank_weights():
    """ranks with a faster host link take proportionally more cells (the 8-GPU box: 4 x 23 GB/s, 4 x 35 GB/s)"""
    cells = [c for c in range(64) for _ in range(128)]
    w = [23.0] * 4 + [35.0] * 4
    ranks = sharding.assign_by_cell(cells, 8, weights=w)
    owner = {}
    for c, r in zip(cells, ranks):
        assert owner.setdefault(c, r) == r
    load = np.bincount(ranks, minlength=8) / 128
    assert load.sum() == 64 and set(load[:4]) <= {6.0, 7.0} and set(load[4:]) <= {9.0, 10.0}, load
    t = load / np.array(w)
    assert t.max() / t.min() < 1.25
    assert sharding.assign_by_cell(cells, 8) == sharding.assign_by_cell(cells, 8, weights=[2.0] * 8)


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n = 1001
    lo, hi = sharding.shard_range(n, world, rank)
    status = np.full(hi - lo, 7, dtype=np.uint8)
    status[: (hi - lo) // 3] = 2 + rank
    recs = sharding.gather_results(torch.from_numpy(status), (hi - lo) * 6144, dist)
    # max-over-ranks timing reduction used by bench.py
    t = torch.tensor([10.0 + rank], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    w = sharding.measured_weights(1.0 + rank, dist)
    assert w == [2.0 / 3.0, 4.0 / 3.0]
    q.put((rank, [r.tolist() for r in recs], float(t.item())))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_gather():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, recs, tmax in got:
        assert tmax == 11.0
        assert len(recs) == 2
        assert recs[0][0] + recs[1][0] == 1001
        assert recs[0][1] + recs[1][1] == 1001 * 6144
        assert recs[0][2 + 2] == 501 // 3 and recs[1][2 + 3] == 500 // 3
