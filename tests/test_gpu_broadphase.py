"""Broad phase parity: the CUDA sweep must reproduce the oracle's (= reference's) pair list BIT FOR BIT — same pairs,
same order, same flags — and the same persistent interval order."""
import numpy as np
import pytest

from avian_b200 import api, scenes, plugins

import oracle_lib
from helpers import oracle_world

pytestmark = pytest.mark.gpu


def random_aabbs(n, seed, scalar=np.float32, extent=30.0, size=1.5, with_filters=True):
    rng = np.random.default_rng(seed)
    s = np.dtype(scalar)
    c = rng.uniform(-extent, extent, size=(n, 3))
    c[:, 0] = np.round(c[:, 0] * 4) / 4          # many exact ties on min.x
    h = rng.uniform(0.1, size, size=(n, 3)); h[:, 0] = 0.5
    mn, mx = (c - h).astype(s), (c + h).astype(s)
    z = rng.random(n) < 0.05
    mn[z, 0] = np.where(rng.random(z.sum()) < 0.5, 0.0, -0.0)   # -0.0 == +0.0 for the reference's comparison
    body = np.arange(n, dtype=np.uint32)
    body[1::9] = body[0::9][: len(body[1::9])]    # two colliders on one body
    a = api.Aabbs(collider=(np.arange(n, dtype=np.uint32) * 3 + 1), body=body, aabb_min=mn, aabb_max=mx,
                  flags=rng.choice([4, 4, 4, 5, 6, 12, 20, 0], size=n).astype(np.uint8), order_out=np.zeros(n, dtype=np.uint32))
    if with_filters:
        a.memberships = rng.choice([1, 2, 3, 0xFFFFFFFF], size=n).astype(np.uint32)
        a.filters = rng.choice([1, 2, 3, 0xFFFFFFFF], size=n).astype(np.uint32)
    return a


def assert_pairs_equal(g: api.PairList, o: api.PairList):
    assert g.count == o.count, (g.count, o.count)
    for name in ("collider1", "collider2", "body1", "body2", "flags"):
        assert np.array_equal(getattr(g, name), getattr(o, name)), name


@pytest.mark.parametrize("n,seed", [(0, 0), (1, 1), (2, 2), (33, 3), (1000, 4), (5000, 5), (20000, 6)])
def test_random_aabbs(gpu_ctx, n, seed):
    a = random_aabbs(n, seed)
    ao = random_aabbs(n, seed)
    o = oracle_lib.broadphase(ao)
    g = gpu_ctx.broadphase(a)
    assert_pairs_equal(g, o)
    assert np.array_equal(a.order_out, ao.order_out)
    if n >= 1000:
        assert o.count > 0


def test_wide_intervals(gpu_ctx):
    """every interval overlaps every other one along x (a tower / a ground slab): the wide-interval path, multi-block per i"""
    n = 9000
    a, ao = random_aabbs(n, 41, extent=8.0, size=0.6, with_filters=False), random_aabbs(n, 41, extent=8.0, size=0.6, with_filters=False)
    rng = np.random.default_rng(1)
    for x in (a, ao):
        x.aabb_min[:, 0] = (-1.0 + 0.001 * np.arange(n)).astype(np.float32)
        x.aabb_max[:, 0] = 50.0
        x.aabb_min[::7, 0] = 20.0          # some narrow ones mixed in
        x.aabb_max[::7, 0] = 20.5
        x.flags[:] = 4
    o = oracle_lib.broadphase(ao)
    g = gpu_ctx.broadphase(a)
    assert o.count > 10000
    assert_pairs_equal(g, o)
    assert np.array_equal(a.order_out, ao.order_out)


def test_existing_pairs_and_joint_disabled(gpu_ctx):
    a, ao = random_aabbs(4000, 11), random_aabbs(4000, 11)
    full = oracle_lib.broadphase(random_aabbs(4000, 11))
    rng = np.random.default_rng(0)
    pick = rng.random(full.count) < 0.5
    key = lambda x, y: (np.minimum(x, y).astype(np.uint64) << np.uint64(32)) | np.maximum(x, y).astype(np.uint64)
    existing = key(full.collider1[pick], full.collider2[pick])
    jd = key(full.body1[~pick][::3], full.body2[~pick][::3])
    for x in (a, ao):
        x.existing_pairs = existing.copy(); x.joint_disabled_body_pairs = jd.copy()
    o = oracle_lib.broadphase(ao)
    g = gpu_ctx.broadphase(a)
    assert 0 < o.count < full.count
    assert_pairs_equal(g, o)


def test_persistent_order_across_steps(gpu_ctx):
    """feed the previous sorted order back in, move the boxes, sort again: the insertion sort's stability contract"""
    a, ao = random_aabbs(3000, 21, with_filters=False), random_aabbs(3000, 21, with_filters=False)
    rng = np.random.default_rng(5)
    for step in range(4):
        o = oracle_lib.broadphase(ao)
        g = gpu_ctx.broadphase(a)
        assert_pairs_equal(g, o)
        assert np.array_equal(a.order_out, ao.order_out)
        perm = ao.order_out.copy()
        d = rng.normal(size=(3000, 1)).astype(np.float32) * 0.3
        for x in (a, ao):
            for name in ("collider", "body", "aabb_min", "aabb_max", "flags"):
                setattr(x, name, np.ascontiguousarray(getattr(x, name)[perm]))
            x.aabb_min[:, :1] += d; x.aabb_max[:, :1] += d
            x.order_out = np.zeros(3000, dtype=np.uint32)


def test_f64_keys(gpu_ctx):
    a, ao = random_aabbs(3000, 31, scalar=np.float64), random_aabbs(3000, 31, scalar=np.float64)
    with api.Context(device=0, scalar=np.float64) as ctx:
        g = ctx.broadphase(a)
    o = oracle_lib.broadphase(ao)
    assert_pairs_equal(g, o)
    assert np.array_equal(a.order_out, ao.order_out)


def test_world_broadphase_on_stack(gpu_ctx):
    """the broad phase inside the stepping world: first frame (everything new) and a later frame (mostly existing)"""
    wo = oracle_world(scenes.cube_stack(8, 6, 8, brick=True), substeps=2)
    wg = plugins.World(scenes.cube_stack(8, 6, 8, brick=True), plugins.PhysicsPlugins().add(plugins.IntegratorPlugin())
                       .add(plugins.BroadPhasePlugin(gpu_ctx)).add(oracle_lib.OracleSolverPlugin()), substeps=2)
    for i in range(4):
        po, pg = wo.broad_phase(), wg.broad_phase()
        assert_pairs_equal(pg, po)
        if i == 0:
            assert po.count > 1000
        wo.narrow_phase(); wg.narrow_phase(); wo.solve(); wg.solve()


def test_large_sizes_properties(gpu_ctx):
    """BASELINE-size property checks (no oracle at this size): the 100k-cube stack's pair list is sorted by
    (rank i, rank j), has no duplicates, and every pair overlaps on all three axes."""
    sc = scenes.cube_stack(51, 40, 50, brick=True)
    w = plugins.World(sc, plugins.PhysicsPlugins(gpu_ctx), substeps=8)
    mn, mx = w.pipeline.update_aabbs(w.bodies, w.params.dt)
    aabbs = w.pipeline.intervals(w.bodies, mn, mx)
    pairs = gpu_ctx.broadphase(aabbs, capacity=4_000_000)
    assert pairs.count > 300_000
    rank = np.empty(sc.bodies.count, dtype=np.int64); rank[aabbs.collider[aabbs.order_out]] = np.arange(sc.bodies.count)
    ri, rj = rank[pairs.collider1], rank[pairs.collider2]
    assert (ri < rj).all()
    keys = ri * sc.bodies.count + rj
    assert (np.diff(keys) > 0).all()            # strictly increasing: ordered and duplicate-free
    a, b = pairs.collider1, pairs.collider2
    assert ((mn[a] <= mx[b]) & (mn[b] <= mx[a])).all()
    sx = mn[aabbs.collider[aabbs.order_out], 0]
    assert (np.diff(sx) >= 0).all()


def _poison(a, rows):
    vals = [np.nan, np.inf, -np.inf]
    for k, r in enumerate(rows):
        (a.aabb_min if k % 2 else a.aabb_max)[r, k % 3] = vals[k % 3]


@pytest.mark.parametrize("scalar", [np.float32, np.float64])
def test_nonfinite_aabbs_are_dropped(scalar):
    """update_aabb_intervals' retain (broad_phase.rs:243-245): an interval whose AABB is NaN / infinite leaves the list — no pairs, absent from
    the new persistent order.  The device flags it, the host compacts and reruns; the result equals the sweep of the finite intervals alone."""
    n = 3000
    rows = [5, 17, 400, 401, 2999, 1234]
    a, ao = random_aabbs(n, 77, scalar=scalar), random_aabbs(n, 77, scalar=scalar)
    _poison(a, rows); _poison(ao, rows)
    keep = np.setdiff1d(np.arange(n), rows)
    clean = random_aabbs(n, 77, scalar=scalar)
    clean = api.Aabbs(**{k: (v[keep] if isinstance(v, np.ndarray) and v.shape[0] == n else v) for k, v in clean.__dict__.items()})
    clean.order_out = np.zeros(keep.size, dtype=np.uint32)
    want = oracle_lib.broadphase(clean)
    o = oracle_lib.broadphase(ao)
    assert_pairs_equal(o, want)
    assert ao.retained_count == keep.size and np.array_equal(ao.order_out[:keep.size], keep[clean.order_out])
    with api.Context(device=0, scalar=scalar) as ctx:
        g = ctx.broadphase(a)
        assert_pairs_equal(g, want)
        assert a.retained_count == keep.size
        assert np.array_equal(a.order_out[:keep.size], keep[clean.order_out])
        # and a finite upload afterwards is untouched by the episode
        b, bo = random_aabbs(n, 78, scalar=scalar), random_aabbs(n, 78, scalar=scalar)
        assert_pairs_equal(ctx.broadphase(b), oracle_lib.broadphase(bo))
        assert b.retained_count == n


def test_repeated_runs_report_their_own_launch_count(gpu_ctx):
    a = random_aabbs(4000, 3)
    gpu_ctx.broadphase_upload(a)
    out = api.PairList.empty(1 << 16)
    counts = []
    for _ in range(3):
        gpu_ctx.broadphase_run()
        gpu_ctx.broadphase_download(out)
        counts.append(gpu_ctx.timings()["kernel_launches"])
    assert counts[0] == counts[1] == counts[2] and 0 < counts[0] < 64, counts
