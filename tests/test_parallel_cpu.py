"""world_size-2 gloo tests of the N > 1 plumbing (island sharding, barrier, max-over-ranks timing, aggregate throughput)
and of the sharded CPU pipeline: two ranks stepping disjoint islands with the oracle give exactly the single-process result."""
import os
import sys
from pathlib import Path

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parent.parent


def _worker(rank, world, port, q):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
    from avian_b200 import parallel, scenes, plugins
    import oracle_lib
    info = parallel.init(backend="gloo")
    assert info.rank == rank and info.world == world
    mine = parallel.shard_islands(5, world, rank)
    # each island = one small independent pile; step the ones this rank owns
    out = {}
    for isl in mine:
        sc = scenes.cube_stack(2 + isl % 2, 2, 2, brick=True)
        w = plugins.World(sc, oracle_lib.oracle_plugins(), substeps=2)
        for _ in range(3):
            w.step()
        out[isl] = w.bodies.position.copy()
    parallel.barrier(info)
    t = parallel.reduce_max([1.0 + rank, 10.0 - rank], info)
    thr = parallel.aggregate_throughput(units_per_rank=len(mine) * 3, seconds_per_rank=0.5 * (rank + 1), info=info)
    q.put((rank, mine, out, t, thr))
    dist.destroy_process_group()


def test_two_rank_island_sharding_matches_single_process():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 1000)
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    results.sort(key=lambda r: r[0])
    owned = sorted(i for r in results for i in r[1])
    assert owned == [0, 1, 2, 3, 4]                                  # every island exactly once
    assert results[0][3] == results[1][3] == [2.0, 10.0]              # max over ranks
    assert results[0][4] == results[1][4] == pytest.approx(15 / 1.0)  # 15 island-steps / slowest rank's 1.0 s
    # single-process reference
    sys.path.insert(0, str(ROOT / "tests"))
    from avian_b200 import scenes, plugins
    import oracle_lib
    for r in results:
        for isl, pos in r[2].items():
            sc = scenes.cube_stack(2 + isl % 2, 2, 2, brick=True)
            w = plugins.World(sc, oracle_lib.oracle_plugins(), substeps=2)
            for _ in range(3):
                w.step()
            assert np.array_equal(pos, w.bodies.position)
