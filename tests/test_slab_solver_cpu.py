"""x-slab partition of the solver stage (include/avian_b200.h "one coupled scene over several GPUs"), CPU side, with the oracle's
resumable stage as the per-rank engine:
  * the partitioner: every body owned once, every constraint owned once, colour order kept, boundary slots consistent across ranks;
  * a scene whose constraints do not cross a cut is reproduced bit for bit;
  * a coupled stack stays within solver tolerance of the unpartitioned step and conserves what the exchange must conserve;
  * a 2-rank gloo run equals the in-process lockstep run bit for bit (the collective is the only difference)."""
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


def stack_input(nx=8, ny=4, nz=4, steps=2, substeps=4, scalar=np.float32):
    _, (prm, b, m, j) = advance_to_solver_input(scenes.cube_stack(nx, ny, nz, brick=True, scalar=scalar), steps=steps, substeps=substeps)
    return prm, b, m


def two_piles_input():
    """two stacks 30 m apart in x on separate static slabs: no constraint crosses the cut between them"""
    a = scenes.cube_stack(3, 3, 3, brick=True)
    prm, b1, m1 = stack_input(3, 3, 3)
    prm, b2, m2 = stack_input(3, 3, 3)
    b2.position[:, 0] += 30.0
    nb = b1.count
    bodies = api.Bodies(**{k: (None if v is None else np.concatenate([v, getattr(b2, k)])) for k, v in b1.__dict__.items()})
    # merge the two colour-major manifold lists colour by colour
    order1, order2 = [], []
    cols = {}
    co1, co2 = np.asarray(m1.color_offsets, dtype=np.int64), np.asarray(m2.color_offsets, dtype=np.int64)
    pick = []
    for c in range(24):
        pick += [(0, i) for i in range(co1[c], co1[c + 1])] + [(1, i) for i in range(co2[c], co2[c + 1])]
    src = [m1, m2]
    def mcol(k):
        return np.stack([getattr(src[s], k)[i] for s, i in pick])
    po = [0]
    pts = []
    for s, i in pick:
        a0, a1 = int(src[s].point_offsets[i]), int(src[s].point_offsets[i + 1])
        pts += [(s, p) for p in range(a0, a1)]
        po.append(len(pts))
    def pcol(k):
        return np.stack([getattr(src[s], k)[p] for s, p in pts])
    body1 = np.array([src[s].body1[i] + (nb if s else 0) for s, i in pick], dtype=np.int32)
    body2 = np.array([src[s].body2[i] + (nb if s else 0) for s, i in pick], dtype=np.int32)
    man = api.Manifolds(color_offsets=(co1 + co2).astype(np.uint32), body1=body1, body2=body2, normal=mcol("normal"), friction=mcol("friction"),
                        restitution=mcol("restitution"), point_offsets=np.array(po, dtype=np.uint32), anchor1=pcol("anchor1"), anchor2=pcol("anchor2"),
                        penetration=pcol("penetration"), normal_speed=pcol("normal_speed"), warm_start_normal_impulse=pcol("warm_start_normal_impulse"),
                        warm_start_tangent_impulse=pcol("warm_start_tangent_impulse"), normal_impulse=pcol("normal_impulse"))
    return prm, bodies, man


def test_partitioner_invariants():
    prm, b, m = stack_input()
    world = 3
    cuts = parallel.body_slab_cuts(b, world)
    shards = [parallel.shard_solver(b, m, cuts, r, world) for r in range(world)]
    owned = np.zeros(b.count, dtype=int)
    mowned = np.zeros(m.count, dtype=int)
    slot_owner = {}
    for r, sh in enumerate(shards):
        owned[sh.body_index[sh.owned_body]] += 1
        mowned[sh.manifold_index] += 1
        assert np.all(np.diff(sh.body_index) > 0) and np.all(np.diff(sh.manifold_index) > 0)       # reference order kept
        lm = sh.manifolds
        assert lm.color_offsets[0] == 0 and lm.color_offsets[-1] == lm.count
        col_global = np.searchsorted(np.asarray(m.color_offsets)[1:], sh.manifold_index, side="right")
        col_local = np.searchsorted(np.asarray(lm.color_offsets)[1:], np.arange(lm.count), side="right")
        assert np.array_equal(col_global, col_local)                                                # colours unchanged
        assert np.array_equal(sh.body_index[lm.body1], m.body1[sh.manifold_index])                  # indices remapped consistently
        assert np.array_equal(lm.anchor1, m.anchor1[sh.point_index])
        for lb, slot, own in zip(sh.bnd_body, sh.bnd_slot, sh.bnd_owner):
            g = int(sh.body_index[lb])
            assert slot_owner.setdefault(int(slot), (g, int(own))) == (g, int(own))                 # every rank agrees on slot -> (body, owner)
    static = b.kind == api.BODY_STATIC
    assert np.all(owned[~static] == 1) and np.all(owned[static] == 0) and np.all(mowned == 1)
    assert shards[0].slot_count == len(slot_owner) > 0


def test_uncoupled_piles_are_reproduced_bit_for_bit():
    prm, b, m = two_piles_input()
    bo, mo = b.copy(), m.copy()
    oracle_lib.solver_step(prm, bo, mo)
    bs, ms = b.copy(), m.copy()
    cuts = np.array([15.0], dtype=np.float32)
    shards = parallel.slab_solver_step_local(lambda r: oracle_lib.OracleSlabEngine(), prm, bs, ms, 2, cuts)
    assert shards[0].slot_count == 0 and min(sh.manifolds.count for sh in shards) > 0
    for k in parallel.BODY_OUTPUTS:
        assert np.array_equal(getattr(bs, k), getattr(bo, k)), k
    for k in parallel.POINT_OUTPUTS:
        assert np.array_equal(getattr(ms, k), getattr(mo, k)), k


@pytest.mark.parametrize("world", [2, 4])
def test_coupled_stack_stays_within_solver_tolerance(world):
    prm, b, m = stack_input(nx=10, ny=4, nz=4, steps=3, substeps=6)
    bo, mo = b.copy(), m.copy()
    oracle_lib.solver_step(prm, bo, mo)
    bs, ms = b.copy(), m.copy()
    shards = parallel.slab_solver_step_local(lambda r: oracle_lib.OracleSlabEngine(), prm, bs, ms, world)
    assert shards[0].slot_count > 0
    # impulses cross a cut with one substep of lag: the velocity error is bounded by a few g*h (h = dt / substeps = 2.8 ms ->
    # g*h = 0.027 m/s), positions by that times dt; the stack does not gain energy
    assert np.abs(bs.linear_velocity - bo.linear_velocity).max() < 0.06
    assert np.abs(bs.position - bo.position).max() < 1e-3
    assert np.abs(bs.linear_velocity).max() < np.abs(bo.linear_velocity).max() + 0.03
    # every copy of a boundary body ends in the same state: the owner's row was scattered, the ghosts must equal it
    for sh in shards:
        for k in parallel.BODY_OUTPUTS:
            assert np.array_equal(getattr(sh.bodies, k)[sh.bnd_body], getattr(bs, k)[sh.body_index[sh.bnd_body]]), k


def test_restitution_pass_is_exchanged():
    _, (prm, b, m, j) = advance_to_solver_input(scenes.cube_stack(6, 3, 3, brick=True, restitution=0.5), steps=2, substeps=3)
    b.linear_velocity[:, 1] -= 2.0          # approaching contacts so that the restitution pass has something to do
    bo, mo = b.copy(), m.copy()
    oracle_lib.solver_step(prm, bo, mo)
    bs, ms = b.copy(), m.copy()
    shards = parallel.slab_solver_step_local(lambda r: oracle_lib.OracleSlabEngine(), prm, bs, ms, 2)
    assert shards[0].slot_count > 0
    # a violent case (2 m/s impacts with e = 0.5 crossing the cut): only sanity is asserted — no energy gain, same order of magnitude as the
    # unpartitioned step — plus the consistency of the copies after the extra exchange that follows the restitution pass
    ke = lambda x: float((x.linear_velocity.astype(np.float64) ** 2).sum())
    assert ke(bs) <= ke(b) * 1.01 and np.isfinite(bs.linear_velocity).all()
    assert np.abs(bs.linear_velocity - bo.linear_velocity).max() < 2.0
    for sh in shards:
        assert np.array_equal(sh.bodies.linear_velocity[sh.bnd_body], bs.linear_velocity[sh.body_index[sh.bnd_body]])


def _worker(rank, world, port, q):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    info = parallel.init(backend="gloo")
    prm, b, m = stack_input()
    parallel.slab_solver_step(oracle_lib.OracleSlabEngine(), prm, b, m, info)
    q.put((rank, {k: getattr(b, k).copy() for k in parallel.BODY_OUTPUTS}, {k: getattr(m, k).copy() for k in parallel.POINT_OUTPUTS}))
    dist.destroy_process_group()


def test_two_rank_gloo_equals_the_lockstep_run():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 32500 + (os.getpid() % 1000)
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = sorted((q.get(timeout=180) for _ in range(world)), key=lambda r: r[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    prm, b, m = stack_input()
    parallel.slab_solver_step_local(lambda r: oracle_lib.OracleSlabEngine(), prm, b, m, world)
    for rank, bodies, points in results:
        for k in parallel.BODY_OUTPUTS:
            assert np.array_equal(bodies[k], getattr(b, k)), (rank, k)
        for k in parallel.POINT_OUTPUTS:
            assert np.array_equal(points[k], getattr(m, k)), (rank, k)


def test_empty_slab_and_constraint_free_rank():
    """Cuts that leave a slab without bodies (nothing between the two piles) and more slabs than piles: the empty ranks go through the
    same launch sequence with nothing to do, and the result is still the unpartitioned one bit for bit."""
    prm, b, m = two_piles_input()
    bo, mo = b.copy(), m.copy()
    oracle_lib.solver_step(prm, bo, mo)
    bs, ms = b.copy(), m.copy()
    cuts = np.array([10.0, 20.0, 100.0], dtype=np.float32)      # slabs: pile 1 | empty | pile 2 | empty
    shards = parallel.slab_solver_step_local(lambda r: oracle_lib.OracleSlabEngine(), prm, bs, ms, 4, cuts)
    assert [sh.bodies.count for sh in shards][1] == 0 and shards[3].bodies.count == 0
    assert shards[1].manifolds is None and shards[0].slot_count == 0
    for k in parallel.BODY_OUTPUTS:
        assert np.array_equal(getattr(bs, k), getattr(bo, k)), k
    for k in parallel.POINT_OUTPUTS:
        assert np.array_equal(getattr(ms, k), getattr(mo, k)), k
