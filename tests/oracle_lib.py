"""TEST INFRASTRUCTURE: ctypes access to oracle/_build/liboracle.so (the CPU restatement of the reference) through
the same column structs as the product ABI, plus oracle-backed plugins for the headless World."""
from __future__ import annotations

import ctypes as C

import numpy as np

from avian_b200 import _build, api, plugins

_lib = None


def lib():
    global _lib
    if _lib is None:
        l = C.CDLL(str(_build.build_oracle()))
        P = C.POINTER
        l.orc_solver_step.argtypes = [C.c_uint32, P(api.AvnStepParams), P(api.AvnBodyColumns), P(api.AvnManifoldColumns), P(api.AvnJointSet), C.c_int]
        l.orc_solver_step.restype = C.c_int
        l.orc_broadphase.argtypes = [C.c_uint32, P(api.AvnAabbColumns), P(api.AvnPairList)]
        l.orc_broadphase.restype = C.c_int
        l.orc_update_aabbs.argtypes = [C.c_uint32, P(api.AvnAabbParams), P(api.AvnColliderColumns)]
        l.orc_update_aabbs.restype = C.c_int
        _lib = l
    return _lib


def _bits(dtype) -> int:
    return 32 if np.dtype(dtype) == np.float32 else 64


def solver_step(params, bodies: api.Bodies, manifolds: api.Manifolds | None = None, joints: api.JointSet | None = None, threads: int = 1) -> None:
    b = bodies.as_struct()
    m = manifolds.as_struct() if manifolds is not None and manifolds.count else None
    j = joints.as_struct() if joints is not None and joints.count else None
    st = lib().orc_solver_step(_bits(bodies.position.dtype), C.byref(params), C.byref(b), C.byref(m) if m is not None else None,
                               C.byref(j) if j is not None else None, threads)
    assert st == 0, f"oracle solver_step failed: {st}"


def broadphase(aabbs: api.Aabbs, capacity: int | None = None) -> api.PairList:
    cap = capacity or max(1024, 64 * int(aabbs.collider.shape[0]))
    a = aabbs.as_struct()
    out = api.PairList.empty(cap)
    s = out.as_struct()
    st = lib().orc_broadphase(_bits(aabbs.aabb_min.dtype), C.byref(a), C.byref(s))
    if st == api.ERR_CAPACITY:
        out = api.PairList.empty(int(s.count))
        s = out.as_struct()
        st = lib().orc_broadphase(_bits(aabbs.aabb_min.dtype), C.byref(a), C.byref(s))
    assert st == 0, f"oracle broadphase failed: {st}"
    out.count = int(s.count)
    return out.trimmed()


def update_aabbs(params, colliders: api.Colliders) -> None:
    c = colliders.as_struct()
    st = lib().orc_update_aabbs(_bits(colliders.position.dtype), C.byref(params), C.byref(c))
    assert st == 0, f"oracle update_aabbs failed: {st}"


class OracleBroadPhasePlugin(plugins.BroadPhasePlugin):
    def __init__(self):
        pass

    def collect_collision_pairs(self, aabbs):
        return broadphase(aabbs)


class OracleSolverPlugin(plugins.SolverPlugin):
    def __init__(self, config=None, threads: int = 1):
        self.config = config or plugins.SolverConfig()
        self.threads = threads

    def step(self, params, bodies, manifolds, joints):
        solver_step(params, bodies, manifolds, joints, self.threads)


def oracle_plugins(threads: int = 1, gravity=None) -> plugins.PhysicsPlugins:
    return plugins.PhysicsPlugins().add(plugins.IntegratorPlugin(gravity)).add(OracleBroadPhasePlugin()).add(OracleSolverPlugin(threads=threads))
