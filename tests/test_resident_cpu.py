"""The resident protocol (plugins.ResidentWorld): per-edge geometry with no host state, counts to the host, graph updates, edge list back,
solver input gathered from edge-indexed arrays, impulses scattered back — stepped next to the ordinary World, the solver's input
manifolds and the bodies must stay identical bit for bit, step after step (new pairs, removed pairs, reused ContactIds, matching)."""
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
from avian_b200 import api, plugins, scenes  # noqa: E402
import oracle_lib  # noqa: E402

COLUMNS = ("color_offsets", "body1", "body2", "normal", "friction", "restitution", "point_offsets", "anchor1", "anchor2", "penetration",
           "normal_speed", "warm_start_normal_impulse", "warm_start_tangent_impulse")


def run_both(scene_fn, steps, substeps=4, kick=None):
    wa = plugins.World(scene_fn(), oracle_lib.oracle_plugins(threads=2), substeps=substeps)
    wb = plugins.ResidentWorld(scene_fn(), oracle_lib.oracle_plugins(threads=2), substeps=substeps)
    if kick is not None:
        for w in (wa, wb):
            kick(w)
    touched = 0
    for i in range(steps):
        for w in (wa, wb):
            w.broad_phase()
        ma, mb = wa.narrow_phase(), wb.narrow_phase()
        assert ma.count == mb.count, f"step {i}: manifold count"
        for k in COLUMNS:
            assert np.array_equal(getattr(ma, k), getattr(mb, k)), f"step {i}: {k}"
        touched += ma.count
        for w in (wa, wb):
            w.solve()
        assert np.array_equal(wa.bodies.position, wb.bodies.position) and np.array_equal(wa.bodies.linear_velocity, wb.bodies.linear_velocity), f"step {i}"
    return wa, wb, touched


def _tumble(w):
    rng = np.random.default_rng(5)
    w.bodies.angular_velocity[1:] = rng.normal(0, 3.0, size=(w.bodies.count - 1, 3)).astype(w.bodies.angular_velocity.dtype)
    w.bodies.linear_velocity[1:] = rng.normal(0, 1.5, size=(w.bodies.count - 1, 3)).astype(w.bodies.linear_velocity.dtype)


def test_tumbling_cubes_pairs_come_and_go():
    wa, wb, touched = run_both(lambda: scenes.cubes_example(4), 120, substeps=6, kick=_tumble)
    assert touched > 5000
    assert wb.bytes_to_host < 1000 and wb.bytes_to_device < 2000          # the protocol's traffic: bytes, not manifolds


def test_brick_stack_and_matching():
    wa, wb, touched = run_both(lambda: scenes.cube_stack(6, 5, 5, brick=True), 25)
    assert touched > 5000
    assert np.abs(wb.e_ws_n).max() > 0                                     # impulses really are carried in the edge arrays


@pytest.mark.parametrize("scalar", [np.float32, np.float64])
def test_spheres(scalar):
    run_both(lambda: scenes.falling_spheres(2000, seed=3, box=(12.0, 6.0, 12.0), scalar=scalar), 30)


def test_ragdolls_with_joints():
    run_both(lambda: scenes.ragdoll_field(9, pitch=1.2, drop_height=0.5), 50)


class MockContactStore:
    """CPU stand-in for the library's contact store + avn_solver_upload_graph (same calls as api.Context).  The rows are numpy arrays with
    the device's column layout, and the per-row work is the DEVICE's row function (csrc/contact_rows.hpp through avh_rows_narrow): what the
    kernel of csrc/contacts.cu runs one thread per row runs here in a loop.  The oracle stands in for the solver; like the device it reads the
    rows' `in` impulses and writes the `out` ones.  It lets DeviceResidentWorld run on the CPU against the ordinary World."""

    COLUMNS = {"c1": ((), np.uint32), "c2": ((), np.uint32), "b1": ((), np.uint32), "b2": ((), np.uint32), "live": ((), np.uint8), "count": ((), np.uint8),
               "disjoint": ((), np.uint8), "normal": ((3,), None), "anchor1": ((4, 3), None), "anchor2": ((4, 3), None), "penetration": ((4,), None),
               "normal_speed": ((4,), None), "prev_count": ((), np.uint8), "prev_a1": ((4, 3), np.float64), "prev_a2": ((4, 3), np.float64),
               "ws_n_in": ((4,), None), "ws_t_in": ((4, 2), None), "ws_n_out": ((4,), None), "ws_t_out": ((4, 2), None), "nimp_out": ((4,), None)}

    def __init__(self, scalar):
        from avian_b200 import fixture
        self.fixture, self.scalar, self.E = fixture, np.dtype(scalar), 0
        self.rows = {}

    def contacts_reserve(self, capacity):
        for k, (shape, dt) in self.COLUMNS.items():
            new = np.zeros((capacity,) + shape, dtype=dt or self.scalar)
            if k in self.rows:
                new[:self.E] = self.rows[k]
            self.rows[k] = new
        self.E = capacity

    def contacts_add(self, ids, c1, c2, b1, b2):        # edge_add_kernel
        r = self.rows
        r["c1"][ids], r["c2"][ids], r["b1"][ids], r["b2"][ids] = c1, c2, b1, b2
        r["live"][ids] = 1; r["count"][ids] = 0; r["prev_count"][ids] = 0

    def contacts_remove(self, ids):                      # edge_remove_kernel
        r = self.rows
        r["live"][ids] = 0; r["count"][ids] = 0; r["prev_count"][ids] = 0

    def contacts_narrow_phase(self, dt, tol, colliders, lv, av, capacity, match_contacts=True, length_unit=1.0):
        r, s = self.rows, self.scalar
        assert capacity == self.E
        cols = {k: (None if colliders.get(k) is None else np.ascontiguousarray(colliders[k], dtype=(np.uint8 if k == "shape" else s)))
                for k in ("shape", "dims", "position", "rotation", "aabb_min", "aabb_max")}
        lv, av = np.ascontiguousarray(lv, dtype=s), np.ascontiguousarray(av, dtype=s)
        p = lambda a: None if a is None else a.ctypes.data
        self.fixture._load().avh_rows_narrow(
            32 if s == np.float32 else 64, self.E, p(r["c1"]), p(r["c2"]), p(r["b1"]), p(r["b2"]), p(r["live"]), p(r["count"]), p(r["disjoint"]), p(r["normal"]),
            p(r["anchor1"]), p(r["anchor2"]), p(r["penetration"]), p(r["normal_speed"]), p(r["prev_count"]), p(r["prev_a1"]), p(r["prev_a2"]), p(r["ws_n_in"]),
            p(r["ws_t_in"]), p(r["ws_n_out"]), p(r["ws_t_out"]), p(cols["shape"]), p(cols["dims"]), p(cols["position"]), p(cols["rotation"]), p(lv), p(av),
            p(cols["aabb_min"]), p(cols["aabb_max"]), float(dt), float(tol), float(length_unit), 1 if match_contacts else 0)
        return r["count"].copy(), r["disjoint"].copy()

    # the broad phase of step_steady (the oracle stands in for the device)
    def broadphase_upload(self, aabbs): self._aabbs = aabbs
    def broadphase_run(self): pass
    def broadphase_download(self, out):
        p = oracle_lib.broadphase(self._aabbs)
        for k in ("collider1", "collider2", "body1", "body2", "flags"):
            getattr(out, k)[:p.count] = getattr(p, k)
        out.count = p.count
        return out

    def solver_step_graph(self, params, bodies, graph, joints=None, reuse_graph=False):
        from avian_b200 import api
        if reuse_graph:      # the library keeps the previous list on the device; so does the mock
            assert np.array_equal(graph["edge"], self._last_graph["edge"]) and np.array_equal(graph["color_offsets"], self._last_graph["color_offsets"])
        self._last_graph = {k: np.array(v, copy=True) for k, v in graph.items()}
        r, s, edge = self.rows, self.scalar, graph["edge"]
        cnt = r["count"][edge].astype(np.int64)
        slot = np.arange(4)[None, :] < cnt[:, None]
        take = lambda a: np.ascontiguousarray(a[edge][slot])
        po = np.concatenate([[0], np.cumsum(cnt)]).astype(np.uint32)
        man = api.Manifolds(color_offsets=graph["color_offsets"], body1=graph["body1"], body2=graph["body2"], normal=np.ascontiguousarray(r["normal"][edge]),
                            friction=graph["friction"], restitution=graph["restitution"], point_offsets=po, anchor1=take(r["anchor1"]), anchor2=take(r["anchor2"]),
                            penetration=take(r["penetration"]), normal_speed=take(r["normal_speed"]), warm_start_normal_impulse=take(r["ws_n_in"]),
                            warm_start_tangent_impulse=take(r["ws_t_in"]), normal_impulse=np.zeros(int(po[-1]), dtype=s)) if edge.size else None
        oracle_lib.solver_step(params, bodies, man, joints, threads=2)
        if man is not None:          # store_contact_impulses: the rows' `out` columns
            for name, col in (("ws_n_out", man.warm_start_normal_impulse), ("ws_t_out", man.warm_start_tangent_impulse), ("nimp_out", man.normal_impulse)):
                rows = r[name][edge]
                rows[slot] = col
                r[name][edge] = rows


@pytest.mark.parametrize("scene_fn,steps,substeps,kick", [
    (lambda: scenes.cubes_example(4), 120, 6, _tumble),
    (lambda: scenes.cube_stack(5, 4, 4, brick=True), 20, 4, None),
    (lambda: scenes.ragdoll_field(6, pitch=1.2, drop_height=0.5), 40, 4, None),
])
def test_device_resident_world_host_logic_with_a_cpu_contact_store(scene_fn, steps, substeps, kick):
    wa = plugins.World(scene_fn(), oracle_lib.oracle_plugins(threads=2), substeps=substeps)
    wb = plugins.DeviceResidentWorld(scene_fn(), oracle_lib.oracle_plugins(threads=2), MockContactStore(wa.scalar), substeps=substeps)
    if kick:
        kick(wa); kick(wb)
    added = removed = 0
    for i in range(steps):
        wa.broad_phase(); wb.broad_phase()
        before = dict(wb.known)
        ma, gb = wa.narrow_phase(), wb.narrow_phase()
        added += sum(1 for e, k in wb.known.items() if before.get(e) != k)
        removed += sum(1 for e in before if e not in wb.known)
        assert ma.count == gb["edge"].shape[0] and np.array_equal(ma.color_offsets, gb["color_offsets"]), f"step {i}"
        wa.solve(); wb.solve()
        assert np.array_equal(wa.bodies.position, wb.bodies.position) and np.array_equal(wa.bodies.linear_velocity, wb.bodies.linear_velocity), f"step {i}"
    assert added > 0
    if kick:
        assert removed > 0          # pairs really were dropped and ContactIds reused


@pytest.mark.parametrize("scene_fn,steps,substeps,kick", [
    (lambda: scenes.cube_stack(5, 4, 4, brick=True), 40, 4, None),          # settles: the fast path takes over
    (lambda: scenes.cubes_example(4), 90, 6, _tumble),                       # pairs come and go: both paths, ContactIds reused
])
def test_step_steady_is_the_same_simulation(scene_fn, steps, substeps, kick):
    """DeviceResidentWorld.step_steady (host AABB columns -> broad phase -> resident narrow phase -> solver from the resident rows, graphs touched
    only when something started / stopped touching) against the ordinary World, step after step, bit for bit; the fast path must actually be
    taken once the contact set has settled."""
    wa = plugins.World(scene_fn(), oracle_lib.oracle_plugins(threads=2), substeps=substeps)
    wb = plugins.DeviceResidentWorld(scene_fn(), oracle_lib.oracle_plugins(threads=2), MockContactStore(wa.scalar), substeps=substeps)
    if kick:
        kick(wa); kick(wb)
    pairs_out = api.PairList.empty(1 << 16)
    fast = 0
    for i in range(steps):
        wa.step()
        mn, mx = wb.pipeline.update_aabbs(wb.bodies, wb.params.dt)
        aabbs = wb.pipeline.intervals(wb.bodies, mn, mx)
        aabbs.joint_disabled_body_pairs = wb.scene.joint_disabled_body_pairs
        if wb._colliders is None:
            wb.prepare_steady(aabbs)
        wb._colliders["aabb_min"], wb._colliders["aabb_max"] = mn, mx     # a moving scene: this step's AABBs
        fast += bool(wb.step_steady(aabbs, pairs_out))
        for k in ("position", "rotation", "linear_velocity", "angular_velocity"):
            assert np.array_equal(getattr(wa.bodies, k), getattr(wb.bodies, k)), f"step {i}: {k}"
    assert fast > 0, "the incremental path never ran"
    if kick is None:
        assert fast > steps // 2
