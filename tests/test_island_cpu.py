"""Island sharding of ONE scene, CPU side with the oracle as the engine: connected components of dynamic
bodies, a balanced deterministic assignment, and — the point — the sharded step equals the unsharded step BIT FOR BIT (contacts,
joints, kinematic and static bodies, joint forces), also over a 2-rank gloo run."""
import os
import sys
from pathlib import Path

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
from avian_b200 import api, parallel, scenes  # noqa: E402
import oracle_lib  # noqa: E402
from helpers import advance_to_solver_input  # noqa: E402


def ragdoll_input(n=12, steps=25, substeps=4):
    """ragdolls that have fallen onto the ground: contacts with the static ground (which must not merge islands), joints inside"""
    _, (prm, b, m, j) = advance_to_solver_input(scenes.ragdoll_field(n, pitch=3.0, drop_height=0.2), steps=steps, substeps=substeps)
    for t in j.types.values():   # ask for the joint force / torque outputs
        t.force = np.zeros((t.count, 3), dtype=b.position.dtype)
        t.torque = np.zeros((t.count, 3), dtype=b.position.dtype)
    return prm, b, m, j


def assert_same_step(b1, m1, j1, b2, m2, j2):
    for k in parallel.BODY_OUTPUTS:
        assert np.array_equal(getattr(b1, k), getattr(b2, k)), k
    if m1 is not None:
        for k in parallel.POINT_OUTPUTS:
            assert np.array_equal(getattr(m1, k), getattr(m2, k)), k
    if j1 is not None:
        for t in j1.types:
            for k in parallel.JOINT_OUTPUTS:
                assert np.array_equal(getattr(j1.types[t], k), getattr(j2.types[t], k)), (t, k)


def test_islands_are_the_components_of_dynamic_bodies():
    prm, b, m, j = ragdoll_input()
    labels, n = parallel.find_islands(b, m, j)
    assert (labels[b.kind != api.BODY_DYNAMIC] == -1).all() and (labels[b.kind == api.BODY_DYNAMIC] >= 0).all()
    assert n == 12                                         # one island per ragdoll: the ground they all touch is static
    assert m.count > 12 and np.unique(labels[labels >= 0]).size == n
    for t in j.types.values():                             # a joint never spans two islands
        assert (labels[t.body1] == labels[t.body2]).all()
    d1, d2 = b.kind[m.body1] == api.BODY_DYNAMIC, b.kind[m.body2] == api.BODY_DYNAMIC
    both = d1 & d2
    assert (labels[m.body1[both]] == labels[m.body2[both]]).all()
    first = [int(np.nonzero(labels == i)[0][0]) for i in range(n)]
    assert first == sorted(first)                          # numbered by first body: independent of the search order


def test_assignment_is_balanced_and_deterministic():
    labels = np.repeat(np.arange(10), [50, 1, 1, 30, 20, 1, 1, 1, 40, 5])
    w = np.ones(labels.size)
    r = parallel.assign_islands(labels, w, 3)
    assert np.array_equal(r, parallel.assign_islands(labels, w, 3))
    load = np.bincount(r, weights=np.bincount(labels), minlength=3)
    assert load.max() - load.min() <= 10 and set(r) == {0, 1, 2}


@pytest.mark.parametrize("world", [2, 3, 5])
def test_island_sharded_step_is_bit_identical(world):
    prm, b, m, j = ragdoll_input()
    bo, mo, jo = b.copy(), m.copy(), j.copy()
    oracle_lib.solver_step(prm, bo, mo, jo)
    bs, ms, js = b.copy(), m.copy(), j.copy()
    shards = parallel.island_solver_step_local(oracle_lib.solver_step, prm, bs, ms, js, world)
    assert sum(int(sh.owned_body.sum()) for sh in shards) == b.count           # every body reported exactly once
    assert sum(sh.manifolds.count for sh in shards if sh.manifolds is not None) == m.count
    assert sum(sh.joints.count for sh in shards if sh.joints is not None) == j.count
    assert min(sh.bodies.count for sh in shards) > 17                          # the work really is spread
    assert_same_step(bs, ms, js, bo, mo, jo)
    assert np.abs(jo.types[api.JOINT_SPHERICAL].force).max() > 0               # the joint outputs are not trivially zero


def test_stack_piles_and_a_kinematic_body():
    """two piles 30 m apart + a kinematic body nobody touches: three islands' worth of work on 2 ranks, bit for bit"""
    from test_slab_solver_cpu import two_piles_input
    prm, b, m = two_piles_input()
    kin = api.Bodies(**{k: (None if v is None else np.concatenate([v, v[-1:]])) for k, v in b.__dict__.items()})
    kin.kind[-1] = api.BODY_KINEMATIC
    kin.position[-1] = (15.0, 40.0, 0.0)
    kin.linear_velocity[-1] = (1.0, 0.0, 0.0)
    bo, mo = kin.copy(), m.copy()
    oracle_lib.solver_step(prm, bo, mo)
    bs, ms = kin.copy(), m.copy()
    shards = parallel.island_solver_step_local(oracle_lib.solver_step, prm, bs, ms, None, 2)
    assert_same_step(bs, ms, None, bo, mo, None)
    assert bs.position[-1, 0] > 15.0                                           # the kinematic body moved (rank 0 stepped it)


def _worker(rank, world, port, q):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    info = parallel.init(backend="gloo")
    prm, b, m, j = ragdoll_input()
    parallel.island_solver_step(oracle_lib.solver_step, prm, b, m, j, info)
    q.put((rank, {k: getattr(b, k).copy() for k in parallel.BODY_OUTPUTS}, {k: getattr(m, k).copy() for k in parallel.POINT_OUTPUTS},
           {t: jt.force.copy() for t, jt in j.types.items()}))
    dist.destroy_process_group()


def test_two_rank_gloo_island_sharding_is_bit_identical():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 34500 + (os.getpid() % 1000)
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = sorted((q.get(timeout=240) for _ in range(world)), key=lambda r: r[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    prm, b, m, j = ragdoll_input()
    oracle_lib.solver_step(prm, b, m, j)
    for rank, bodies, points, forces in results:
        for k in parallel.BODY_OUTPUTS:
            assert np.array_equal(bodies[k], getattr(b, k)), (rank, k)
        for k in parallel.POINT_OUTPUTS:
            assert np.array_equal(points[k], getattr(m, k)), (rank, k)
        for t in forces:
            assert np.array_equal(forces[t], j.types[t].force), (rank, t)
