"""Device narrow phase, geometry stage: avn_narrow_phase against the host fixture's generator — the same header
(csrc/narrow_math.hpp) compiled by g++ and by nvcc — bit for bit: point counts, normals, anchors, penetrations, normal speeds and the
disjoint flags, on random cuboid / sphere soups (face, edge and vertex contacts, deep overlaps, near misses) in f32 and f64."""
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
from avian_b200 import api, fixture  # noqa: E402

pytestmark = pytest.mark.gpu


def soup(n, seed, scalar, spheres=0.3, box=6.0):
    rng = np.random.default_rng(seed)
    pos = rng.uniform(0, box, size=(n, 3))
    q = rng.normal(size=(n, 4)); q /= np.linalg.norm(q, axis=1, keepdims=True)
    q[: n // 5] = (0, 0, 0, 1)                                   # some axis-aligned boxes: exact face-face and parallel-edge cases
    shape = (rng.random(n) < spheres).astype(np.uint8)
    dims = rng.uniform(0.2, 0.8, size=(n, 3))
    he = np.where(shape[:, None] == 1, dims[:, :1], np.abs(dims).max(axis=1, keepdims=True) * 1.8)
    cols = {"shape": shape, "dims": dims.astype(scalar), "position": pos.astype(scalar), "rotation": q.astype(scalar),
            "aabb_min": (pos - he).astype(scalar), "aabb_max": (pos + he).astype(scalar)}
    lv, av = rng.normal(0, 1.5, size=(n, 3)).astype(scalar), rng.normal(0, 2.0, size=(n, 3)).astype(scalar)
    # candidate pairs: everything within 2 units (plus some far ones for the disjoint flag)
    d = np.linalg.norm(pos[:, None] - pos[None], axis=2)
    i, j = np.nonzero(np.triu(d < 2.0, k=1))
    far = rng.integers(0, n, size=(50, 2))
    far = far[far[:, 0] != far[:, 1]]
    c1 = np.concatenate([i, far[:, 0]]).astype(np.uint32)
    c2 = np.concatenate([j, far[:, 1]]).astype(np.uint32)
    return cols, lv, av, (c1, c2, c1.copy(), c2.copy())


@pytest.mark.parametrize("scalar", [np.float32, np.float64])
@pytest.mark.parametrize("seed", [1, 2, 3])
def test_device_manifolds_equal_the_fixture(scalar, seed):
    cols, lv, av, pairs = soup(600, seed, scalar)
    want = fixture.raw_manifolds(scalar, 1.0 / 60.0, 0.005, pairs, cols, lv, av)
    with api.Context(device=0, scalar=scalar) as ctx:
        got = ctx.narrow_phase(1.0 / 60.0, 0.005, pairs, cols, lv, av)
    assert int((want["point_count"] > 0).sum()) > 300 and int((want["point_count"] == 4).sum()) > 10 and int(want["disjoint"].sum()) > 5
    for k in want:
        assert np.array_equal(got[k], want[k]), k


def test_without_aabbs_and_empty_input(gpu_ctx):
    cols, lv, av, pairs = soup(100, 9, np.float32)
    cols["aabb_min"] = cols["aabb_max"] = None
    want = fixture.raw_manifolds(np.float32, 1.0 / 60.0, 0.005, pairs, cols, lv, av)
    got = gpu_ctx.narrow_phase(1.0 / 60.0, 0.005, pairs, cols, lv, av)
    for k in want:
        assert np.array_equal(got[k], want[k]), k
    empty = tuple(np.zeros(0, dtype=np.uint32) for _ in range(4))
    assert gpu_ctx.narrow_phase(1.0 / 60.0, 0.005, empty, cols, lv, av)["point_count"].shape == (0,)
    bad = (np.array([1000], dtype=np.uint32),) * 4
    with pytest.raises(api.AvianError):
        gpu_ctx.narrow_phase(1.0 / 60.0, 0.005, bad, cols, lv, av)


# ---- the solver fed from edge-indexed storage --------------------------------------------------------------------------------
def test_solver_from_edge_indexed_manifolds_equals_the_csr_input(gpu_ctx):
    """avn_solver_upload_edges: the same manifolds scattered over ContactId-indexed rows (4 slots per edge, gaps, arbitrary ids) and
    listed colour by colour give the step of the CSR input bit for bit; impulses come back in the edges' slots, rows of edges that are
    not in the graph keep their values."""
    from avian_b200 import scenes
    from helpers import advance_to_solver_input
    _, (prm, b, m, j) = advance_to_solver_input(scenes.cube_stack(6, 4, 5, brick=True), steps=3, substeps=4)
    bs, ms = b.copy(), m.copy()
    gpu_ctx.solver_step(prm, bs, ms)
    M = m.count
    rng = np.random.default_rng(0)
    E = 3 * M + 17
    edge = rng.permutation(E)[:M].astype(np.uint32)            # arbitrary ContactIds with gaps
    s = b.position.dtype
    cnt = np.diff(m.point_offsets.astype(np.int64))
    slot = np.arange(4)[None, :] < cnt[:, None]
    edges = {"point_count": np.zeros(E, dtype=np.uint8), "normal": np.zeros((E, 3), dtype=s), "anchor1": np.zeros((E, 4, 3), dtype=s),
             "anchor2": np.zeros((E, 4, 3), dtype=s), "penetration": np.zeros((E, 4), dtype=s), "normal_speed": np.zeros((E, 4), dtype=s),
             "warm_start_normal_impulse": np.full((E, 4), 7.0, dtype=s), "warm_start_tangent_impulse": np.full((E, 4, 2), 7.0, dtype=s),
             "normal_impulse": np.full((E, 4), 7.0, dtype=s)}
    edges["point_count"][edge] = cnt
    edges["normal"][edge] = m.normal
    for k in ("anchor1", "anchor2", "penetration", "normal_speed", "warm_start_normal_impulse", "warm_start_tangent_impulse", "normal_impulse"):
        rows = edges[k][edge]
        rows[slot] = getattr(m, k)
        edges[k][edge] = rows
    graph = {"color_offsets": m.color_offsets, "edge": edge, "body1": m.body1, "body2": m.body2, "friction": m.friction, "restitution": m.restitution}
    be = b.copy()
    gpu_ctx.solver_step_edges(prm, be, graph, edges)
    for k in ("position", "rotation", "linear_velocity", "angular_velocity"):
        assert np.array_equal(getattr(be, k), getattr(bs, k)), k
    for k in ("warm_start_normal_impulse", "warm_start_tangent_impulse", "normal_impulse"):
        assert np.array_equal(edges[k][edge][slot], getattr(ms, k)), k
    untouched = np.ones(E, dtype=bool); untouched[edge] = False
    assert (edges["warm_start_normal_impulse"][untouched] == 7.0).all() and (edges["normal_impulse"][untouched] == 7.0).all()


# ---- the whole resident protocol on the device -----------------------------------------------------------------------------------
def _tumble(w):
    rng = np.random.default_rng(5)
    w.bodies.angular_velocity[1:] = rng.normal(0, 3.0, size=(w.bodies.count - 1, 3)).astype(w.bodies.angular_velocity.dtype)
    w.bodies.linear_velocity[1:] = rng.normal(0, 1.5, size=(w.bodies.count - 1, 3)).astype(w.bodies.linear_velocity.dtype)


@pytest.mark.parametrize("scene_fn,steps,substeps,kick", [
    (lambda: __import__("avian_b200.scenes", fromlist=["x"]).cubes_example(4), 120, 6, True),          # pairs come and go, ContactIds are reused
    (lambda: __import__("avian_b200.scenes", fromlist=["x"]).cube_stack(6, 5, 5, brick=True), 25, 4, False),
    (lambda: __import__("avian_b200.scenes", fromlist=["x"]).ragdoll_field(9, pitch=1.2, drop_height=0.5), 40, 4, False),
])
def test_device_resident_world_equals_the_ordinary_world(gpu_ctx, scene_fn, steps, substeps, kick):
    """Bodies of the device-resident pipeline vs the ordinary GPU world (host narrow phase, CSR upload), step after step, bit for bit;
    the resident rows' impulses equal the ordinary world's manifold columns; a few bytes per edge cross the bus."""
    from avian_b200 import plugins
    wa = plugins.World(scene_fn(), plugins.PhysicsPlugins(gpu_ctx), substeps=substeps)
    with api.Context(device=0, scalar=wa.scalar) as ctx_b:
        wb = plugins.DeviceResidentWorld(scene_fn(), plugins.PhysicsPlugins(ctx_b), ctx_b, substeps=substeps)
        if kick:
            _tumble(wa); _tumble(wb)
        for i in range(steps):
            wa.broad_phase(); wb.broad_phase()
            ma, gb = wa.narrow_phase(), wb.narrow_phase()
            assert ma.count == gb["edge"].shape[0], f"step {i}: manifold count"
            assert np.array_equal(ma.color_offsets, gb["color_offsets"]) and np.array_equal(ma.body1, gb["body1"]), f"step {i}: graph"
            wa.solve(); wb.solve()
            for k in ("position", "rotation", "linear_velocity", "angular_velocity"):
                assert np.array_equal(getattr(wa.bodies, k), getattr(wb.bodies, k)), f"step {i}: {k}"
            if ma.count:
                wn, wt, ni = ctx_b.contacts_download_impulses(wb.capacity)
                cnt = np.diff(ma.point_offsets.astype(np.int64))
                slot = np.arange(4)[None, :] < cnt[:, None]
                assert np.array_equal(wn[gb["edge"]][slot], ma.warm_start_normal_impulse), f"step {i}: impulses"
                assert np.array_equal(ni[gb["edge"]][slot], ma.normal_impulse), f"step {i}: total impulses"
        assert wb.bytes_to_host <= 2 * wb.capacity


@pytest.mark.parametrize("scene_fn,steps,substeps,kick", [
    (lambda: __import__("avian_b200.scenes", fromlist=["x"]).cubes_example(4), 80, 6, True),
    (lambda: __import__("avian_b200.scenes", fromlist=["x"]).cube_stack(6, 5, 5, brick=True), 25, 4, False),
])
def test_step_steady_on_the_device_equals_the_ordinary_world(gpu_ctx, scene_fn, steps, substeps, kick):
    """The bench's end-to-end arm: host AABB / body columns -> avn_broadphase -> avn_contacts_narrow_phase -> avn_solver_upload_graph (the resident
    colour-major list reused while no contact starts or stops touching) -> bodies, against the ordinary GPU world, bit for bit every step."""
    from avian_b200 import plugins
    wa = plugins.World(scene_fn(), plugins.PhysicsPlugins(gpu_ctx), substeps=substeps)
    with api.Context(device=0, scalar=wa.scalar) as ctx_b:
        wb = plugins.DeviceResidentWorld(scene_fn(), plugins.PhysicsPlugins(ctx_b), ctx_b, substeps=substeps)
        if kick:
            _tumble(wa); _tumble(wb)
        pairs_out = api.PairList.empty(1 << 16)
        fast = 0
        for i in range(steps):
            wa.step()
            mn, mx = wb.pipeline.update_aabbs(wb.bodies, wb.params.dt)
            aabbs = wb.pipeline.intervals(wb.bodies, mn, mx)
            aabbs.joint_disabled_body_pairs = wb.scene.joint_disabled_body_pairs
            if wb._colliders is None:
                wb.prepare_steady(aabbs)
            wb._colliders["aabb_min"], wb._colliders["aabb_max"] = mn, mx
            fast += bool(wb.step_steady(aabbs, pairs_out))
            for k in ("position", "rotation", "linear_velocity", "angular_velocity"):
                assert np.array_equal(getattr(wa.bodies, k), getattr(wb.bodies, k)), f"step {i}: {k}"
        if not kick:   # (a tumbling scene changes its touching set every step: the reuse path is for settled scenes)
            assert fast > 0, "the graph-reuse path never ran"
