"""x-slab partition of one scene, CPU side: the partitioner and the merge, with the oracle as the per-slab
engine.  The per-slab pair lists concatenated in slab order must be the single-process list bit for bit — pairs, order, flags and
the new persistent interval order — for any number of slabs, with ties on min.x, a ground slab spanning every cut, existing pairs
and layers.  A 2-rank gloo run covers the all-gather."""
import os
import sys
from pathlib import Path

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
from avian_b200 import api, parallel  # noqa: E402
import oracle_lib  # noqa: E402


def random_aabbs(n, seed, scalar=np.float32, ground=True, ties=True, existing_frac=0.3, layers=True) -> api.Aabbs:
    rng = np.random.default_rng(seed)
    c = rng.uniform(0, 12, size=(n, 3))
    if ties:
        c[:, 0] = np.round(c[:, 0] * 2) / 2            # many identical min.x values (a grid stack has them too)
    he = rng.uniform(0.2, 1.2, size=(n, 3))
    mn, mx = c - he, c + he
    if ties:
        mn[:, 0] = c[:, 0] - 0.5
    if ground:
        mn[0], mx[0] = (-5, -2, -5), (20, 0.4, 20)     # one interval reaching across every cut
    perm = rng.permutation(n)                           # persistent order = arbitrary
    mn, mx = mn[perm].astype(scalar), mx[perm].astype(scalar)
    collider = (np.arange(n, dtype=np.uint32) * 3 + 7)[perm]
    body = (collider // 2).astype(np.uint32)
    flags = np.full(n, api.AABB_GENERATE_CONSTRAINTS, dtype=np.uint8)
    flags[rng.random(n) < 0.1] |= api.AABB_IS_INACTIVE
    flags[rng.random(n) < 0.1] |= api.AABB_CONTACT_EVENTS
    a = api.Aabbs(collider=collider, body=body, aabb_min=np.ascontiguousarray(mn), aabb_max=np.ascontiguousarray(mx), flags=flags,
                  order_out=np.zeros(n, dtype=np.uint32))
    if layers:
        a.memberships = rng.integers(1, 4, size=n, dtype=np.uint32)
        a.filters = rng.integers(1, 4, size=n, dtype=np.uint32)
    if existing_frac > 0:
        full = oracle_lib.broadphase(a)
        keep = rng.random(full.count) < existing_frac
        lo = np.minimum(full.collider1, full.collider2).astype(np.uint64)
        hi = np.maximum(full.collider1, full.collider2).astype(np.uint64)
        a.existing_pairs = np.ascontiguousarray(((lo << np.uint64(32)) | hi)[keep])
    return a


def assert_same_pairs(got: api.PairList, want: api.PairList):
    assert got.count == want.count
    for c in parallel.PAIR_COLUMNS:
        assert np.array_equal(getattr(got, c)[:got.count], getattr(want, c)[:want.count]), c


@pytest.mark.parametrize("scalar", [np.float32, np.float64])
@pytest.mark.parametrize("world", [1, 2, 3, 5, 8])
def test_slabs_concatenate_to_the_single_list(world, scalar):
    a = random_aabbs(700, seed=world, scalar=scalar)
    want = oracle_lib.broadphase(a)
    want_order = a.order_out.copy()
    assert want.count > 300
    cuts = parallel.slab_cuts(a.aabb_min[:, 0], world)
    assert cuts.shape == (world - 1,) and np.all(np.diff(cuts) >= 0)
    parts = [parallel.slab_broadphase_local(oracle_lib.broadphase, a, cuts, r) for r in range(world)]
    got, order = parallel.merge_slab_results(parts, a.collider)
    assert_same_pairs(got, want)
    assert np.array_equal(order, want_order)
    if world > 1:   # the work really is split: no slab emits everything, halos carry the flag and never start a pair
        assert max(p[0]["collider1"].shape[0] for p in parts) < want.count
        sh = parallel.shard_aabbs(a, cuts, 0)
        assert (sh.aabbs.flags[~sh.owned] & parallel.AABB_HALO).all() and not (sh.aabbs.flags[sh.owned] & parallel.AABB_HALO).any()
        # the ground reaches across every cut: its sweep is split, so no slab holds (nearly) the whole scene
        assert (sh.aabbs.flags[sh.owned] & parallel.AABB_SPLIT_I).sum() >= 1
        assert max(parallel.shard_aabbs(a, cuts, r).index.size for r in range(world)) < 0.85 * a.collider.shape[0]
        assert sum(p[2] for p in parts) > 0


def test_every_interval_is_owned_once_and_ties_stay_together():
    a = random_aabbs(500, seed=3, ties=True)
    cuts = parallel.slab_cuts(a.aabb_min[:, 0], 4)
    slab = parallel.slab_of(a.aabb_min[:, 0], cuts)
    assert slab.min() >= 0 and slab.max() <= 3
    for v in np.unique(a.aabb_min[:, 0]):
        assert np.unique(slab[a.aabb_min[:, 0] == v]).size == 1
    owned = np.zeros(500, dtype=int)
    for r in range(4):
        sh = parallel.shard_aabbs(a, cuts, r)
        owned[sh.index[sh.owned]] += 1
    assert (owned == 1).all()


def test_empty_and_degenerate_slabs():
    a = random_aabbs(40, seed=9, ground=False, existing_frac=0)
    a.aabb_min[:, 0] = 1.0                     # every interval starts at the same x: one slab owns everything
    a.aabb_max[:, 0] = 2.0
    want = oracle_lib.broadphase(a)
    cuts = parallel.slab_cuts(a.aabb_min[:, 0], 3)
    parts = [parallel.slab_broadphase_local(oracle_lib.broadphase, a, cuts, r) for r in range(3)]
    got, order = parallel.merge_slab_results(parts, a.collider)
    assert_same_pairs(got, want)
    assert sorted(p[1].shape[0] for p in parts) == [0, 0, 40]


def _worker(rank, world, port, q):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    info = parallel.init(backend="gloo")
    a = random_aabbs(900, seed=21)
    got, order = parallel.slab_broadphase(oracle_lib.broadphase, a, info)
    q.put((rank, got.count, {c: getattr(got, c).copy() for c in parallel.PAIR_COLUMNS}, order))
    dist.destroy_process_group()


def test_two_rank_gloo_all_gather_gives_every_rank_the_full_list():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 30500 + (os.getpid() % 1000)
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = sorted((q.get(timeout=180) for _ in range(world)), key=lambda r: r[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    a = random_aabbs(900, seed=21)
    want = oracle_lib.broadphase(a)
    for rank, count, cols, order in results:
        assert count == want.count
        for c in parallel.PAIR_COLUMNS:
            assert np.array_equal(cols[c], getattr(want, c)[:want.count]), (rank, c)
        assert np.array_equal(order, a.order_out)


@pytest.mark.parametrize("world", [2, 4, 7])
def test_several_far_reaching_intervals_owned_by_different_slabs(world):
    """Walls that start in different slabs and reach over several cuts, overlapping each other and the ground: every one is swept
    piecewise, wide-wide pairs included, and the merge still reproduces the single list."""
    a = random_aabbs(900, seed=40 + world, existing_frac=0.2)
    n = a.collider.shape[0]
    xs = np.sort(a.aabb_min[:, 0])
    rng = np.random.default_rng(world)
    walls = rng.choice(np.arange(1, n), size=6, replace=False)
    for k, w in enumerate(walls):
        start = xs[(k * n) // 7]
        a.aabb_min[w] = (start, -1.0, -1.0)
        a.aabb_max[w] = (start + rng.uniform(6, 14), 13.0, 13.0 if k % 2 else 2.0)
        a.flags[w] = api.AABB_GENERATE_CONSTRAINTS
    a.existing_pairs = None
    want = oracle_lib.broadphase(a)
    want_order = a.order_out.copy()
    cuts = parallel.slab_cuts(a.aabb_min[:, 0], world)
    slab = parallel.slab_of(a.aabb_min[:, 0], cuts)
    wide = parallel.wide_intervals(a, cuts, slab)
    assert wide.sum() >= 3 and np.unique(slab[wide]).size >= min(2, world - 1)          # far-reaching intervals with different owners
    parts = [parallel.slab_broadphase_local(oracle_lib.broadphase, a, cuts, r) for r in range(world)]
    assert sum(p[2] for p in parts) > 0
    got, order = parallel.merge_slab_results(parts, a.collider)
    assert_same_pairs(got, want)
    assert np.array_equal(order, want_order)


def _worker_degenerate(rank, world, port, q):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    info = parallel.init(backend="gloo")
    a = random_aabbs(60, seed=4, ground=False, existing_frac=0)
    a.aabb_min[:, 0] = 1.0            # one slab owns everything: the other rank contributes zero-length columns to every gather
    a.aabb_max[:, 0] = 2.0
    got, order = parallel.slab_broadphase(oracle_lib.broadphase, a, info)
    q.put((rank, got.count, got.collider1.copy(), got.collider2.copy(), order))
    dist.destroy_process_group()


def test_two_rank_gloo_with_an_empty_slab():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 35500 + (os.getpid() % 1000)
    procs = [ctx.Process(target=_worker_degenerate, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = sorted((q.get(timeout=180) for _ in range(world)), key=lambda r: r[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    a = random_aabbs(60, seed=4, ground=False, existing_frac=0)
    a.aabb_min[:, 0] = 1.0
    a.aabb_max[:, 0] = 2.0
    want = oracle_lib.broadphase(a)
    assert want.count > 20
    for rank, count, c1, c2, order in results:
        assert count == want.count and np.array_equal(c1, want.collider1[:count]) and np.array_equal(c2, want.collider2[:count])
        assert np.array_equal(order, a.order_out)
