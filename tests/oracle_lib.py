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
    aabbs.retained_count = int(a.retained_count)
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


class OracleSlabEngine:
    """TEST INFRASTRUCTURE: the x-slab partition's per-rank engine backed by the oracle's resumable stage (orc_step_*): the same
    interface as parallel.GpuSlabEngine with numpy tables, so the orchestration and the collective can be tested without a GPU."""

    def __init__(self, threads: int = 1):
        l = lib()
        P = C.POINTER
        if not hasattr(l, "_slab_bound"):
            l.orc_step_begin.argtypes = [C.c_uint32, P(api.AvnStepParams), P(api.AvnBodyColumns), P(api.AvnManifoldColumns), P(api.AvnJointSet), C.c_int]
            l.orc_step_begin.restype = C.c_void_p
            for name, args in (("substeps", [C.c_void_p, C.c_uint32]), ("restitution", [C.c_void_p]), ("needs_restitution", [C.c_void_p]),
                               ("set_boundary", [C.c_void_p, P(api.AvnBoundary)]), ("boundary_snapshot", [C.c_void_p]),
                               ("boundary_pack", [C.c_void_p, C.c_void_p]), ("boundary_apply", [C.c_void_p, C.c_void_p]), ("finish", [C.c_void_p])):
                fn = getattr(l, f"orc_step_{name}")
                fn.argtypes, fn.restype = args, C.c_int
            l._slab_bound = True
        self.l, self.threads, self.h = l, threads, None

    def begin(self, prm, shard, rank: int, world: int):
        self.prm, self.shard = prm, shard
        self.dtype = shard.bodies.position.dtype
        self._b = shard.bodies.as_struct()
        self._m = shard.manifolds.as_struct() if shard.manifolds is not None and shard.manifolds.count else None
        self.rank, self.world = rank, world
        self.h = None

    def _ensure(self):
        if self.h is None:
            self.h = self.l.orc_step_begin(_bits(self.dtype), C.byref(self.prm), C.byref(self._b), C.byref(self._m) if self._m is not None else None, None,
                                           self.threads)
            assert self.h, "oracle step_begin failed"
            sh = self.shard
            self._bnd = tuple(np.ascontiguousarray(x, dtype=np.int32) for x in (sh.bnd_body, sh.bnd_source, sh.bnd_owner))
            b = api.AvnBoundary(int(self._bnd[0].shape[0]), int(sh.record_count), self.rank, self.world, *(x.ctypes.data for x in self._bnd))
            assert self.l.orc_step_set_boundary(self.h, C.byref(b)) == 0

    def tables(self, record_count: int, world: int):
        n = max(record_count, 1) * api.BOUNDARY_RECORD_SCALARS
        return np.zeros(n, dtype=self.dtype), np.zeros(world * n, dtype=self.dtype)

    def run(self, first: int, count: int, flags: int):
        if flags & api.RUN_PREPARE:
            self._ensure()
        if count:
            assert self.l.orc_step_substeps(self.h, count) == 0
        if flags & api.RUN_RESTITUTION:
            assert self.l.orc_step_restitution(self.h) == 0
        self._finalize = bool(flags & api.RUN_FINALIZE)

    def snapshot(self): assert self.l.orc_step_boundary_snapshot(self.h) == 0
    def pack(self, table): assert self.l.orc_step_boundary_pack(self.h, table.ctypes.data) == 0
    def apply(self, gathered): assert self.l.orc_step_boundary_apply(self.h, gathered.ctypes.data) == 0
    def needs_restitution(self) -> bool: return bool(self.l.orc_step_needs_restitution(self.h))

    def finish(self):
        assert self._finalize
        assert self.l.orc_step_finish(self.h) == 0
        self.h = None

    def all_gather(self, gathered, table):
        import torch
        import torch.distributed as dist
        dist.all_gather_into_tensor(torch.from_numpy(gathered), torch.from_numpy(table))

    def copy_table(self, gathered, r: int, table): gathered[r * table.size:(r + 1) * table.size] = table
    def sync(self): pass
