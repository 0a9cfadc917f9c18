"""CPU tests that pin the oracle (oracle/) itself: the reference's own numeric tests restated, analytic invariants,
and serial == colour-parallel equality.  No GPU needed."""
import numpy as np
import pytest

from avian_b200 import api, scenes, plugins

import oracle_lib
from helpers import oracle_world, advance_to_solver_input, snapshot


def _one_body(scalar=np.float32, angvel=(0, 0, 2.0), inv_inertia=(6.0, 0, 0, 6.0, 0, 6.0)):
    s = np.dtype(scalar)
    return api.Bodies(kind=np.array([api.BODY_DYNAMIC], dtype=np.uint8), position=np.zeros((1, 3), dtype=s),
                      rotation=np.array([[0, 0, 0, 1]], dtype=s), linear_velocity=np.zeros((1, 3), dtype=s),
                      angular_velocity=np.array([angvel], dtype=s), inverse_mass=np.ones(1, dtype=s),
                      inverse_inertia_local=np.array([inv_inertia], dtype=s))


@pytest.mark.parametrize("scalar", [np.float32, np.float64])
def test_semi_implicit_euler_reference_test(scalar):
    """integrator/mod.rs:561-629: unit cube (m=1, I=1/6), w = 2z, 100 steps at 10 Hz, 1 substep."""
    b = _one_body(scalar)
    prm = api.default_step_params(dt=0.1, substeps=1)
    for _ in range(100):
        oracle_lib.solver_step(prm, b)
    assert abs(b.position[0, 1] - (-490.5)) < 10.0
    assert np.allclose(b.linear_velocity[0], [0, -98.1, 0], atol=1e-4)
    assert np.allclose(b.angular_velocity[0], [0, 0, 2.0], atol=1e-5)
    want = np.array([0, 0, np.sin(10.0), np.cos(10.0)])
    got = b.rotation[0].astype(np.float64)
    assert min(np.abs(got - want).max(), np.abs(got + want).max()) < 0.01
    # semi-implicit Euler closed form: y = -g h^2 n(n+1)/2
    assert abs(b.position[0, 1] - (-9.81 * 0.01 * 100 * 101 / 2)) < 1e-2


def test_constant_acceleration_displacement():
    """rigid_body/forces/tests.rs:54-654 family: x = 1/2 a t^2 (eps 0.05) with 20 substeps, no gravity."""
    b = _one_body(angvel=(0, 0, 0))
    b.linear_acceleration = np.array([[2.0, 0, 0]], dtype=np.float32)
    prm = api.default_step_params(dt=1.0 / 60.0, substeps=20, gravity=(0, 0, 0))
    for _ in range(60):
        oracle_lib.solver_step(prm, b)
    assert abs(b.position[0, 0] - 1.0) < 0.05


def test_damping_and_locked_axes():
    b = _one_body(angvel=(1.0, 2.0, 3.0))
    b.linear_velocity[:] = (1.0, 1.0, 1.0)
    b.linear_damping = np.array([0.5], dtype=np.float32)
    b.locked_axes = np.array([0x10 | 0x01], dtype=np.uint8)  # translation y, rotation z
    prm = api.default_step_params(dt=1.0 / 60.0, substeps=4, gravity=(0, -9.81, 0))
    v0 = b.linear_velocity.copy()
    oracle_lib.solver_step(prm, b)
    h = prm.h
    assert np.isclose(b.linear_velocity[0, 0], v0[0, 0] * (1.0 / (1.0 + h * 0.5)) ** 4, rtol=1e-5)
    assert np.isclose(b.linear_velocity[0, 1], v0[0, 1] * (1.0 / (1.0 + h * 0.5)) ** 4, rtol=1e-5)  # gravity masked on locked y


def test_gyroscopic_conserves_momentum_magnitude():
    b = _one_body(np.float64, angvel=(0.3, 5.0, 0.1), inv_inertia=(1.0, 0, 0, 0.25, 0, 0.5))
    prm = api.default_step_params(dt=1.0 / 60.0, substeps=6, gravity=(0, 0, 0))
    inertia = np.diag([1.0, 4.0, 2.0])
    def L(b):
        q = b.rotation[0]; R = _rotm(q)
        return np.linalg.norm(R @ inertia @ R.T @ b.angular_velocity[0])
    l0 = L(b)
    for _ in range(120):
        oracle_lib.solver_step(prm, b)
    assert abs(L(b) - l0) / l0 < 2e-2   # world inertia is frozen within a step, so only approximately
    assert np.linalg.norm(b.angular_velocity[0] - np.array([0.3, 5.0, 0.1])) > 1e-3  # it does precess


def _rotm(q):
    x, y, z, w = q
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                     [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                     [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


def test_resting_cube_settles():
    sc = scenes.cube_stack(1, 1, 1)
    w = oracle_world(sc, substeps=6)
    for _ in range(180):
        w.step()
    y = float(w.bodies.position[1, 1])
    assert 0.49 < y < 0.505, y                       # rests on the ground within contact tolerance
    assert np.abs(w.bodies.linear_velocity[1]).max() < 2e-2
    m = w.last_manifolds
    assert m.count == 1 and m.penetration.shape[0] == 4
    # accumulated normal impulses carry the weight: sum(lambda_n) / h ~ m g  (last substep's accumulated impulse)
    lam = float(m.warm_start_normal_impulse.sum())
    assert abs(lam / w.params.h - 9.81) / 9.81 < 0.05
    assert (m.warm_start_normal_impulse >= 0).all()


def test_friction_cone_and_nonnegative_normal_impulses():
    sc = scenes.cubes_example(3)
    w = oracle_world(sc, substeps=4)
    for _ in range(90):
        w.step()
    m = w.last_manifolds
    assert m.count > 0
    lam_n = m.warm_start_normal_impulse
    lam_t = np.linalg.norm(m.warm_start_tangent_impulse, axis=1)
    mu = np.repeat(m.friction, np.diff(m.point_offsets))
    assert (lam_n >= 0).all()
    assert (lam_t <= mu * lam_n * (1 + 1e-5) + 1e-7).all()


def test_stack_stays_up():
    sc = scenes.cube_stack(3, 4, 3, brick=True)
    w = oracle_world(sc, substeps=8)
    y0 = w.bodies.position[:, 1].copy()
    sc_xz = w.bodies.position[1:, [0, 2]].copy()
    for _ in range(120):
        w.step()
    assert np.abs(w.bodies.position[1:, 1] - y0[1:]).max() < 0.08
    assert np.abs(w.bodies.position[1:, [0, 2]] - sc_xz).max() < 0.05
    assert np.isfinite(w.bodies.position).all()


def test_colour_parallel_equals_serial():
    """the reference's parallel path (colours in chunks) must equal the serial sweep bit for bit"""
    sc = scenes.cube_stack(6, 5, 6, brick=True)
    _, (prm, b, m, j) = advance_to_solver_input(sc, steps=3, substeps=4)
    b1, m1 = b.copy(), m.copy()
    b4, m4 = b.copy(), m.copy()
    oracle_lib.solver_step(prm, b1, m1, None, threads=1)
    oracle_lib.solver_step(prm, b4, m4, None, threads=4)
    for name in ("position", "rotation", "linear_velocity", "angular_velocity"):
        assert np.array_equal(getattr(b1, name), getattr(b4, name)), name
    assert np.array_equal(m1.warm_start_normal_impulse, m4.warm_start_normal_impulse)


def test_colouring_is_conflict_free():
    sc = scenes.cube_stack(5, 6, 5, brick=True)
    w, (prm, b, m, j) = advance_to_solver_input(sc, steps=2, substeps=2)
    assert m.count > 100
    for c in range(api.COLOR_OVERFLOW):
        lo, hi = int(m.color_offsets[c]), int(m.color_offsets[c + 1])
        bodies = np.concatenate([m.body1[lo:hi], m.body2[lo:hi]])
        dyn = bodies[b.kind[bodies] != api.BODY_STATIC]
        assert len(np.unique(dyn)) == len(dyn), f"colour {c} touches a dynamic body twice"
    # static contacts sit in the high colours, dynamic-only ones in 0..19 (constraint_graph.rs:178-219)
    static_contact = (b.kind[m.body1] == api.BODY_STATIC) | (b.kind[m.body2] == api.BODY_STATIC)
    color_of = np.repeat(np.arange(api.GRAPH_COLOR_COUNT), np.diff(m.color_offsets))
    assert (color_of[static_contact] >= 1).all() and (color_of[~static_contact] < api.DYNAMIC_COLOR_COUNT).all()


def test_f32_f64_agree_on_one_step():
    """Same scene in f32 and f64: identical manifolds/colours, accumulated impulses agree to f32 accuracy.
    (Velocities are NOT compared: the relax pass is discontinuous at separation == 0 — normal_part.rs:129-156 —
    so a contact that ends a step within rounding of zero separation may or may not keep its push-out speed.)"""
    sc32 = scenes.cube_stack(3, 3, 3, brick=True, scalar=np.float32)
    sc64 = scenes.cube_stack(3, 3, 3, brick=True, scalar=np.float64)
    w32, w64 = oracle_world(sc32, substeps=4), oracle_world(sc64, substeps=4)
    w32.step(); w64.step()
    m32, m64 = w32.last_manifolds, w64.last_manifolds
    assert m32.count == m64.count and np.array_equal(m32.color_offsets, m64.color_offsets)
    assert np.abs(w32.bodies.position - w64.bodies.position).max() < 2e-3
    assert np.abs(m32.normal_impulse - m64.normal_impulse).max() < 1e-4 * m64.normal_impulse.max()


def test_pendulum_distance_joint_length_is_kept():
    s = np.float64
    b = api.Bodies(kind=np.array([api.BODY_STATIC, api.BODY_DYNAMIC], dtype=np.uint8), position=np.array([[0, 0, 0], [1.0, 0, 0]], dtype=s),
                   rotation=np.array([[0, 0, 0, 1], [0, 0, 0, 1]], dtype=s), linear_velocity=np.zeros((2, 3), dtype=s),
                   angular_velocity=np.zeros((2, 3), dtype=s), inverse_mass=np.array([0, 1.0], dtype=s),
                   inverse_inertia_local=np.array([[0] * 6, [2.5, 0, 0, 2.5, 0, 2.5]], dtype=s))
    j = api.Joints(body1=np.array([0], dtype=np.int32), body2=np.array([1], dtype=np.int32), local_anchor1=np.zeros((1, 3), dtype=s),
                   local_anchor2=np.zeros((1, 3), dtype=s), limit_min=np.array([1.0], dtype=s), limit_max=np.array([1.0], dtype=s),
                   force=np.zeros((1, 3), dtype=s), torque=np.zeros((1, 3), dtype=s))
    js = api.JointSet({api.JOINT_DISTANCE: j})
    prm = api.default_step_params(dt=1.0 / 60.0, substeps=8)
    ys = []
    for _ in range(240):
        oracle_lib.solver_step(prm, b, None, js)
        ys.append(b.position[1, 1])
        assert abs(np.linalg.norm(b.position[1]) - 1.0) < 5e-3
    assert min(ys) < -0.9                      # it swings through the bottom
    # period of a large-amplitude (90 deg) pendulum of length 1: T = 2 pi sqrt(L/g) * 1.18 = 2.37 s -> bottom at ~0.59 s
    first_min = next(i for i in range(1, len(ys) - 1) if ys[i] < ys[i - 1] and ys[i] <= ys[i + 1])
    t_bottom = (first_min + 1) / 60.0
    assert abs(t_bottom - 0.59) < 0.06
    assert np.isfinite(j.force).all() and np.linalg.norm(j.force[0]) > 1.0


def test_spherical_chain_hangs():
    sc = scenes.spherical_chain(20, scalar=np.float64)
    w = plugins.World(sc, oracle_lib.oracle_plugins(), substeps=20)
    for _ in range(60):
        w.step()
    p = w.bodies.position
    gaps = np.linalg.norm(np.diff(p, axis=0), axis=1)
    assert np.allclose(gaps, gaps[0], atol=5e-3)       # links keep their spacing
    assert np.allclose(p[0], 0.0)                      # the kinematic anchor does not move
    assert np.isfinite(p).all()


def test_body_with_velocity_moves_reference_test():
    """src/tests/mod.rs:103-147: a sphere (r = 0.5, density 1) moving at 1 m/s in x without gravity, 500 updates at 60 Hz:
    y = z = 0 exactly, x = 500/60 within 0.03."""
    b = _one_body(angvel=(0, 0, 0), inv_inertia=(19.098593, 0, 0, 19.098593, 0, 19.098593))   # 1 / (2/5 m r^2), m = 4/3 pi r^3
    b.inverse_mass[:] = 1.0 / (4.0 / 3.0 * np.pi * 0.125)
    b.linear_velocity[:] = (1.0, 0.0, 0.0)
    prm = api.default_step_params(dt=1.0 / 60.0, substeps=6, gravity=(0, 0, 0))
    for _ in range(500):
        oracle_lib.solver_step(prm, b)
    assert b.position[0, 1] == 0.0 and b.position[0, 2] == 0.0
    assert abs(b.position[0, 0] - 500.0 / 60.0) < 0.03


def test_cubes_simulation_is_locally_deterministic_reference_test():
    """src/tests/mod.rs:149-183: the 4x4x4 cubes scene run several times gives identical transforms (5 s in the reference; 1.5 s here:
    the cubes have landed and are in contact).  Restated for the oracle with a different thread count per run, which is the way
    the reference's run-to-run scheduling differs."""
    def run(threads):
        w = oracle_world(scenes.cubes_example(4), substeps=6, threads=threads)
        for _ in range(90):
            w.step()
        return w.bodies.position.copy(), w.bodies.rotation.copy()
    runs = [run(t) for t in (1, 3, 8, 3)]
    assert np.abs(runs[0][0][1:, 1]).max() < 12.0 and runs[0][0][1:, 1].min() > -1.0      # the pile is on the ground, nothing fell through
    for p, q in runs[1:]:
        assert np.array_equal(p, runs[0][0]) and np.array_equal(q, runs[0][1])


def _two_bodies(mass=(1.0, 0.5), inertia=(1.0, 0.5)):
    s = np.float32
    n = len(mass)
    inv_i = np.array([[1.0 / i, 0, 0, 1.0 / i, 0, 1.0 / i] for i in inertia], dtype=s)
    return api.Bodies(kind=np.full(n, api.BODY_DYNAMIC, dtype=np.uint8), position=np.zeros((n, 3), dtype=s), rotation=np.tile(np.array([[0, 0, 0, 1]], dtype=s), (n, 1)),
                      linear_velocity=np.zeros((n, 3), dtype=s), angular_velocity=np.zeros((n, 3), dtype=s), inverse_mass=np.array([1.0 / m for m in mass], dtype=s),
                      inverse_inertia_local=inv_i)


def test_apply_force_reference_test():
    """rigid_body/forces/tests.rs:53-96: 9.81 N upwards against gravity, 20 substeps, 5 s at 64 Hz (TIMESTEP = 1/64): the 1 kg body stays (1e-6), the 0.5 kg body
    rises by 1/2 * 9.81 * 25 (eps 0.05).  The ForcePlugin turns the force into the acceleration column F / m (forces/plugin.rs:207-241)."""
    b = _two_bodies()
    b.linear_acceleration = np.array([[0, 9.81 * 1.0, 0], [0, 9.81 * 2.0, 0]], dtype=np.float32)
    prm = api.default_step_params(dt=1.0 / 64.0, substeps=20, gravity=(0, -9.81, 0))
    for _ in range(int(5.0 * 64.0)):
        oracle_lib.solver_step(prm, b)
    assert abs(b.position[0, 1]) < 1e-6
    assert abs(b.position[1, 1] - 0.5 * 9.81 * 25.0) < 0.05


def test_apply_torque_reference_test():
    """rigid_body/forces/tests.rs:346-394: 1.5 N m about Z for 1.5 s: the body with unit inertia turns by 1/2 * 1.5 * t^2 (0.1 rad), the one
    with inertia 0.5 by 1.5 * t^2 (0.15 rad)."""
    b = _two_bodies()
    b.angular_acceleration = np.array([[0, 0, 1.5 / 1.0], [0, 0, 1.5 / 0.5]], dtype=np.float32)
    prm = api.default_step_params(dt=1.0 / 64.0, substeps=20, gravity=(0, 0, 0))
    for _ in range(int(1.5 * 64.0)):
        oracle_lib.solver_step(prm, b)
    def angle_to_z(q, theta):
        want = np.array([0, 0, np.sin(theta / 2), np.cos(theta / 2)])
        return 2.0 * np.arccos(min(1.0, abs(float(np.dot(q.astype(np.float64), want)))))
    assert angle_to_z(b.rotation[0], 0.5 * 1.5 * 1.5 ** 2) < 0.1
    assert angle_to_z(b.rotation[1], 1.5 * 1.5 ** 2) < 0.15


def test_broadphase_drops_nonfinite_aabbs():
    """update_aabb_intervals' retain (broad_phase.rs:243-245): an interval with a NaN / infinite AABB leaves the list before the sort."""
    mn = np.array([[0, 0, 0], [0.5, 0, 0], [np.nan, 0, 0], [0.2, 0, 0], [0.4, 0, 0]], dtype=np.float32)
    mx = mn + 1.0
    mx[3, 1] = np.inf
    a = api.Aabbs(collider=np.arange(5, dtype=np.uint32), body=np.arange(5, dtype=np.uint32), aabb_min=mn, aabb_max=mx,
                  flags=np.full(5, api.AABB_GENERATE_CONSTRAINTS, dtype=np.uint8), order_out=np.zeros(5, dtype=np.uint32))
    p = oracle_lib.broadphase(a)
    assert a.retained_count == 3 and list(a.order_out[:3]) == [0, 4, 1]
    assert sorted(zip(p.collider1.tolist(), p.collider2.tolist())) == [(0, 1), (0, 4), (4, 1)]
