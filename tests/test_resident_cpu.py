"""The resident protocol (plugins.ResidentWorld): per-edge geometry with no host state, counts to the host, graph updates, edge list back,
solver input gathered from edge-indexed arrays, impulses scattered back — stepped next to the ordinary World, the solver's input
manifolds and the bodies must stay identical bit for bit, step after step (new pairs, removed pairs, reused ContactIds, matching)."""
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
from avian_b200 import plugins, scenes  # noqa: E402
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
