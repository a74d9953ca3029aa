"""CPU suite: the N>1 host logic (frame sharding, max-over-ranks timing reduction) with world_size 2 on gloo."""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_units_cover_once(pkg):
    import importlib
    sh = importlib.import_module("svo_b200.sharding")
    for F in (1, 2, 3, 17, 100, 10001):
        for W in (1, 2, 4, 8):
            seen = []
            for r in range(W):
                f0, n = sh.shard_units(F, W, r)
                seen += list(range(f0, f0 + n - 1)) if n else []
                for b0, bn in sh.batches(f0, n, 7):
                    assert 1 <= bn <= 7
                assert sum(bn for _, bn in sh.batches(f0, n, 7)) == n
            assert seen == list(range(F - 1))  # every unit exactly once, in order; chunks overlap by one frame


def _gloo_worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    import __graft_entry__ as g
    g.load_package()
    import importlib
    sh = importlib.import_module("svo_b200.sharding")
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    f0, n = sh.shard_units(101, world, rank)
    units = torch.tensor([n - 1], dtype=torch.int64)
    dist.all_reduce(units)  # bookkeeping only -- the data path has no collective
    t = torch.tensor([10.0 + rank], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)  # bench.py: max over ranks
    dist.barrier()
    q.put((rank, f0, n, int(units[0]), float(t[0])))
    dist.destroy_process_group()


def test_gloo_world2():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 1000
    ps = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = sorted(q.get(timeout=120) for _ in ps)
    for p in ps:
        p.join(60)
        assert p.exitcode == 0
    assert res[0][1:3] == (0, 51) and res[1][1:3] == (50, 51)
    assert all(r[3] == 100 and r[4] == 11.0 for r in res)
