"""ctypes wrapper of libavian_host.so (avian_b200/host/host_api.cpp): the CPU-side fixture around the hot path —
swept AABBs, contact graph, narrow-phase manifolds for cuboids/spheres, constraint-graph colouring.
None of this is the hot path; it produces the INPUTS the hot path consumes (identical for oracle and GPU)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _build, api

_vp = C.c_void_p
_lib = None

SHAPE_CUBOID, SHAPE_SPHERE = 0, 1


def _load():
    global _lib
    if _lib is None:
        lib = C.CDLL(str(_build.build_host()))
        lib.avh_create.argtypes = [C.c_uint32]
        lib.avh_create.restype = _vp
        lib.avh_destroy.argtypes = [_vp]
        lib.avh_set_shapes.argtypes = [_vp, _vp, _vp, _vp, _vp]
        lib.avh_update_aabbs.argtypes = [_vp, C.c_uint32, _vp, _vp, _vp, _vp, C.c_double, _vp, _vp]
        lib.avh_get_order.argtypes = [_vp, _vp]
        lib.avh_get_order.restype = C.c_uint32
        lib.avh_set_order.argtypes = [_vp, _vp]
        lib.avh_existing_pairs.argtypes = [_vp, _vp, C.c_uint64]
        lib.avh_existing_pairs.restype = C.c_uint64
        lib.avh_add_pairs.argtypes = [_vp, _vp, _vp, _vp, _vp, _vp, C.c_uint64]
        lib.avh_narrow_phase.argtypes = [_vp, C.c_uint32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, C.c_double, C.c_uint32, C.POINTER(C.c_uint32)]
        lib.avh_narrow_phase.restype = C.c_uint32
        lib.avh_export_manifolds.argtypes = [_vp, C.c_uint32] + [_vp] * 13
        lib.avh_store_impulses.argtypes = [_vp, C.c_uint32, _vp, _vp, _vp]
        lib.avh_raw_manifolds.argtypes = [C.c_uint32, C.c_uint32] + [_vp] * 12 + [C.c_double, C.c_double] + [_vp] * 9
        lib.avh_active_edges.argtypes = [_vp] * 6
        lib.avh_active_edges.restype = C.c_uint32
        lib.avh_apply_counts.argtypes = [_vp, _vp, _vp, _vp, _vp, C.c_uint32, C.POINTER(C.c_uint32)]
        lib.avh_apply_counts.restype = C.c_uint32
        lib.avh_export_edges.argtypes = [_vp] * 7
        lib.avh_export_edges.restype = None
        lib.avh_match_raw.argtypes = [C.c_uint32, C.c_uint32, _vp, _vp, _vp, _vp, C.c_double, C.c_uint32, _vp, _vp, _vp, _vp, _vp]
        lib.avh_match_raw.restype = None
        lib.avh_rows_narrow.argtypes = [C.c_uint32, C.c_uint32] + [_vp] * 27 + [C.c_double, C.c_double, C.c_double, C.c_uint32]
        lib.avh_rows_narrow.restype = None
        lib.avh_raw_manifolds.restype = None
        lib.avh_pair_count.argtypes = [_vp]
        lib.avh_pair_count.restype = C.c_uint32
        _lib = lib
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data


def raw_manifolds(scalar, dt: float, contact_tolerance: float, pairs, colliders: dict, lin_vel: np.ndarray, ang_vel: np.ndarray,
                  f64_anchors: bool = False) -> dict:
    """The geometry stage of the fixture's narrow phase for an explicit pair list (same columns as Context.narrow_phase).
    f64_anchors adds the unrounded anchors (what match_contacts compares on the next step)."""
    lib = _load()
    dt_ = np.dtype(scalar)
    c1, c2, b1, b2 = (np.ascontiguousarray(x, dtype=np.uint32) for x in pairs)
    n = int(c1.shape[0])
    cols = {k: (None if colliders.get(k) is None else np.ascontiguousarray(colliders[k], dtype=(np.uint8 if k == "shape" else dt_)))
            for k in ("shape", "dims", "position", "rotation", "aabb_min", "aabb_max")}
    lv, av = np.ascontiguousarray(lin_vel, dtype=dt_), np.ascontiguousarray(ang_vel, dtype=dt_)
    out = {"point_count": np.zeros(n, dtype=np.uint8), "disjoint": np.zeros(n, dtype=np.uint8), "normal": np.zeros((n, 3), dtype=dt_),
           "anchor1": np.zeros((n, 4, 3), dtype=dt_), "anchor2": np.zeros((n, 4, 3), dtype=dt_), "penetration": np.zeros((n, 4), dtype=dt_),
           "normal_speed": np.zeros((n, 4), dtype=dt_)}
    a1d = np.zeros((n, 4, 3), dtype=np.float64) if f64_anchors else None
    a2d = np.zeros((n, 4, 3), dtype=np.float64) if f64_anchors else None
    lib.avh_raw_manifolds(32 if dt_ == np.float32 else 64, n, _p(c1), _p(c2), _p(b1), _p(b2), _p(cols["shape"]), _p(cols["dims"]), _p(cols["position"]),
                          _p(cols["rotation"]), _p(lv), _p(av), _p(cols["aabb_min"]), _p(cols["aabb_max"]), float(dt), float(contact_tolerance),
                          *(_p(out[k]) for k in ("point_count", "disjoint", "normal", "anchor1", "anchor2", "penetration", "normal_speed")), _p(a1d), _p(a2d))
    if f64_anchors:
        out["anchor1_f64"], out["anchor2_f64"] = a1d, a2d
    return out


class HostPipeline:
    """Contact graph + narrow-phase fixture + constraint graph for a fixed set of bodies (one collider per body,
    collider Entity::index() == body Entity::index() == row in the body columns)."""

    def __init__(self, shape_type: np.ndarray, dims: np.ndarray, friction: np.ndarray, restitution: np.ndarray, scalar=np.float32):
        self.lib = _load()
        self.n = int(shape_type.shape[0])
        self.scalar = np.dtype(scalar)
        self.bits = 32 if self.scalar == np.float32 else 64
        dims = np.asarray(dims, dtype=self.scalar).astype(np.float64)      # shapes carry the world's scalar type (Collider is f32 in an f32 build)
        self.h = self.lib.avh_create(self.n)
        st = np.ascontiguousarray(shape_type, dtype=np.int32)
        dm = np.ascontiguousarray(dims, dtype=np.float64)
        fr = np.ascontiguousarray(friction, dtype=np.float64)
        rs = np.ascontiguousarray(restitution, dtype=np.float64)
        self.lib.avh_set_shapes(self.h, _p(st), _p(dm), _p(fr), _p(rs))

    def __del__(self):
        if getattr(self, "h", None):
            self.lib.avh_destroy(self.h)
            self.h = None

    def update_aabbs(self, bodies: api.Bodies, dt: float):
        mn = np.empty((self.n, 3), dtype=self.scalar)
        mx = np.empty((self.n, 3), dtype=self.scalar)
        self.lib.avh_update_aabbs(self.h, self.bits, _p(bodies.position), _p(bodies.rotation), _p(bodies.linear_velocity),
                                  _p(bodies.angular_velocity), dt, _p(mn), _p(mx))
        return mn, mx

    def intervals(self, bodies: api.Bodies, aabb_min: np.ndarray, aabb_max: np.ndarray, with_existing: bool = True) -> api.Aabbs:
        """AabbIntervals in the persistent order + the pair set, as the broad phase's input columns."""
        order = np.empty(self.n, dtype=np.uint32)
        self.lib.avh_get_order(self.h, _p(order))
        flags = np.where(bodies.kind[order] == api.BODY_STATIC, api.AABB_IS_INACTIVE, 0).astype(np.uint8) | np.uint8(api.AABB_GENERATE_CONSTRAINTS)
        existing = None
        if with_existing:
            cnt = int(self.lib.avh_existing_pairs(self.h, None, 0))
            if cnt:
                existing = np.empty(cnt, dtype=np.uint64)
                self.lib.avh_existing_pairs(self.h, _p(existing), cnt)
        return api.Aabbs(collider=order.copy(), body=order.copy(), aabb_min=np.ascontiguousarray(aabb_min[order]),
                         aabb_max=np.ascontiguousarray(aabb_max[order]), flags=np.ascontiguousarray(flags),
                         order_out=np.empty(self.n, dtype=np.uint32), existing_pairs=existing)

    def commit_broadphase(self, aabbs: api.Aabbs, pairs: api.PairList) -> None:
        """Persist the sorted interval order and add the new pairs to the contact graph (broad_phase.rs:443-471)."""
        new_order = np.ascontiguousarray(aabbs.collider[aabbs.order_out])
        self.lib.avh_set_order(self.h, _p(new_order))
        n = int(pairs.count)
        if n:
            c1, c2, b1, b2, fl = (np.ascontiguousarray(x[:n]) for x in (pairs.collider1, pairs.collider2, pairs.body1, pairs.body2, pairs.flags))
            self.lib.avh_add_pairs(self.h, _p(c1), _p(c2), _p(b1), _p(b2), _p(fl), n)

    def narrow_phase(self, bodies: api.Bodies, aabb_min: np.ndarray, aabb_max: np.ndarray, dt: float, match_contacts: bool = True) -> api.Manifolds:
        pts = C.c_uint32(0)
        kind = np.ascontiguousarray(bodies.kind, dtype=np.uint8)
        m = int(self.lib.avh_narrow_phase(self.h, self.bits, _p(kind), _p(bodies.position), _p(bodies.rotation), _p(bodies.linear_velocity),
                                          _p(bodies.angular_velocity), _p(aabb_min), _p(aabb_max), dt, 1 if match_contacts else 0, C.byref(pts)))
        p = int(pts.value)
        s = self.scalar
        man = api.Manifolds(
            color_offsets=np.zeros(api.GRAPH_COLOR_COUNT + 1, dtype=np.uint32), body1=np.zeros(m, dtype=np.int32), body2=np.zeros(m, dtype=np.int32),
            normal=np.zeros((m, 3), dtype=s), friction=np.zeros(m, dtype=s), restitution=np.zeros(m, dtype=s),
            point_offsets=np.zeros(m + 1, dtype=np.uint32), anchor1=np.zeros((p, 3), dtype=s), anchor2=np.zeros((p, 3), dtype=s),
            penetration=np.zeros(p, dtype=s), normal_speed=np.zeros(p, dtype=s), warm_start_normal_impulse=np.zeros(p, dtype=s),
            warm_start_tangent_impulse=np.zeros((p, 2), dtype=s), normal_impulse=np.zeros(p, dtype=s))
        self.lib.avh_export_manifolds(self.h, self.bits, _p(man.color_offsets), _p(man.body1), _p(man.body2), _p(man.normal), _p(man.friction),
                                      _p(man.restitution), _p(man.point_offsets), _p(man.anchor1), _p(man.anchor2), _p(man.penetration),
                                      _p(man.normal_speed), _p(man.warm_start_normal_impulse), _p(man.warm_start_tangent_impulse))
        return man

    # ---- resident mode: the geometry runs elsewhere, the host keeps only the graphs
    def active_edges(self):
        n = self.pair_count
        ids, c1, c2, b1, b2 = (np.zeros(n, dtype=np.uint32) for _ in range(5))
        k = int(self.lib.avh_active_edges(self.h, _p(ids), _p(c1), _p(c2), _p(b1), _p(b2)))
        return ids[:k], c1[:k], c2[:k], b1[:k], b2[:k]

    def apply_counts(self, bodies: api.Bodies, ids: np.ndarray, point_count: np.ndarray, disjoint: np.ndarray):
        """Touching state machine + contact / constraint graph updates from per-edge point counts.  Returns (manifolds, points) in the graph."""
        pts = C.c_uint32(0)
        kind = np.ascontiguousarray(bodies.kind, dtype=np.uint8)
        ids, point_count, disjoint = np.ascontiguousarray(ids, dtype=np.uint32), np.ascontiguousarray(point_count, dtype=np.uint8), np.ascontiguousarray(disjoint, dtype=np.uint8)
        m = int(self.lib.avh_apply_counts(self.h, _p(kind), _p(ids), _p(point_count), _p(disjoint), int(ids.shape[0]), C.byref(pts)))
        return m, int(pts.value)

    def export_edges(self, m: int):
        """The constraint graph as a colour-major list of edge ids + per-edge bodies and material."""
        co = np.zeros(api.GRAPH_COLOR_COUNT + 1, dtype=np.uint32)
        edge, b1, b2 = np.zeros(m, dtype=np.uint32), np.zeros(m, dtype=np.int32), np.zeros(m, dtype=np.int32)
        fr, re = np.zeros(m, dtype=np.float64), np.zeros(m, dtype=np.float64)
        self.lib.avh_export_edges(self.h, _p(co), _p(edge), _p(b1), _p(b2), _p(fr), _p(re))
        return co, edge, b1, b2, fr, re

    def store_impulses(self, man: api.Manifolds) -> None:
        self.lib.avh_store_impulses(self.h, self.bits, _p(man.warm_start_normal_impulse), _p(man.warm_start_tangent_impulse), _p(man.normal_impulse))

    @property
    def pair_count(self) -> int:
        return int(self.lib.avh_pair_count(self.h))
