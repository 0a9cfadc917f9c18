"""Shared helpers for the parity tests."""
from __future__ import annotations

import copy

import numpy as np

from avian_b200 import api, plugins, scenes

import oracle_lib

RTOL = 1e-5   # BASELINE.json north_star: "post-step positions/velocities within 1e-5 relative (f32)"


def rel_err(a: np.ndarray, b: np.ndarray) -> float:
    """ELEMENT-WISE relative error with an absolute floor of one unit (1 m, 1 m/s, 1 rad/s, 1 N s):
    max_i |a_i - b_i| / max(1, |b_i|).  (Round 1 divided by the largest magnitude of the whole column, which in a 50 m scene let
    every position be off by 5e-4 m)"""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    if a.size == 0:
        return 0.0
    return float((np.abs(a - b) / np.maximum(1.0, np.abs(b))).max())


def ulp_distance(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    """Distance in units in the last place between two arrays of the same float dtype (+0 and -0 are 0 apart)."""
    a = np.ascontiguousarray(a)
    b = np.ascontiguousarray(b)
    it = np.int32 if a.dtype == np.float32 else np.int64
    ia, ib = a.view(it).astype(np.int64), b.view(it).astype(np.int64)
    sign = np.int64(-2 ** 31) if a.dtype == np.float32 else np.int64(-2 ** 63)
    ia = np.where(ia < 0, sign - ia, ia)
    ib = np.where(ib < 0, sign - ib, ib)
    return np.abs(ia - ib)


def parity_report(got, want, names) -> dict:
    """Per column: element-wise relative error (floor 1), largest absolute error, share of bit-identical elements, largest ulp distance."""
    out = {}
    for name in names:
        a, b = getattr(got, name), getattr(want, name)
        if a is None or a.size == 0:
            continue
        u = ulp_distance(a, b)
        out[name] = {"max_rel_err": rel_err(a, b), "max_abs_err": float(np.abs(a.astype(np.float64) - b.astype(np.float64)).max()),
                     "bit_identical": float((u == 0).mean()), "max_ulp": int(u.max())}
    return out


BODY_OUT = ("position", "rotation", "linear_velocity", "angular_velocity")
IMPULSE_OUT = ("warm_start_normal_impulse", "warm_start_tangent_impulse", "normal_impulse")


def assert_bodies_close(got: api.Bodies, want: api.Bodies, rtol: float = RTOL, what: str = ""):
    for name in BODY_OUT:
        e = rel_err(getattr(got, name), getattr(want, name))
        assert np.isfinite(getattr(got, name)).all(), f"{what}{name} has non-finite values"
        assert e <= rtol, f"{what}{name}: relative error {e:.3e} > {rtol:.1e}"


def assert_manifolds_close(got: api.Manifolds, want: api.Manifolds, rtol: float = RTOL, what: str = ""):
    for name in IMPULSE_OUT:
        e = rel_err(getattr(got, name), getattr(want, name))
        assert e <= rtol, f"{what}{name}: relative error {e:.3e} > {rtol:.1e}"


def oracle_world(scene: scenes.Scene, substeps: int = 6, threads: int = 1, **kw) -> plugins.World:
    return plugins.World(scene, oracle_lib.oracle_plugins(threads=threads), substeps=substeps, **kw)


def snapshot(world: plugins.World):
    """(params, bodies, manifolds, joints) deep-copied: the inputs of the next solver stage."""
    m = world.last_manifolds
    return (world.params, world.bodies.copy(), None if m is None else m.copy(), None if world.joints is None else world.joints.copy())


def advance_to_solver_input(scene: scenes.Scene, steps: int, substeps: int = 6, **kw):
    """Run `steps` full oracle steps, then broad+narrow phase of the next step; return the world and the solver input."""
    w = oracle_world(scene, substeps=substeps, **kw)
    for _ in range(steps):
        w.step()
    w.broad_phase()
    w.narrow_phase()
    return w, snapshot(w)
