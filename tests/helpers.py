"""Shared helpers for the parity tests."""
from __future__ import annotations

import copy

import numpy as np

from avian_b200 import api, plugins, scenes

import oracle_lib

RTOL = 1e-5   # BASELINE.json north_star: "post-step positions/velocities within 1e-5 relative (f32)"


def rel_err(a: np.ndarray, b: np.ndarray) -> float:
    """max |a-b| / max(1, max|b|) per column block: relative to the scale of the quantity (positions ~ scene size)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    if a.size == 0:
        return 0.0
    scale = max(1.0, float(np.abs(b).max()))
    return float(np.abs(a - b).max() / scale)


def assert_bodies_close(got: api.Bodies, want: api.Bodies, rtol: float = RTOL, what: str = ""):
    for name in ("position", "rotation", "linear_velocity", "angular_velocity"):
        e = rel_err(getattr(got, name), getattr(want, name))
        assert np.isfinite(getattr(got, name)).all(), f"{what}{name} has non-finite values"
        assert e <= rtol, f"{what}{name}: relative error {e:.3e} > {rtol:.1e}"


def assert_manifolds_close(got: api.Manifolds, want: api.Manifolds, rtol: float = RTOL, what: str = ""):
    for name in ("warm_start_normal_impulse", "warm_start_tangent_impulse", "normal_impulse"):
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
