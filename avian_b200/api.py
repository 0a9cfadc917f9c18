"""ctypes binding of include/avian_b200.h — the same C ABI the Rust shim binds.

Nothing here computes: arrays go in as numpy buffers, the CUDA library does the work.  If the library or a
CUDA device is missing the constructors raise — there is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field

import numpy as np

from . import _build

GRAPH_COLOR_COUNT = 24
COLOR_OVERFLOW = 23
DYNAMIC_COLOR_COUNT = 20
MAX_MANIFOLD_POINTS = 4
NO_BODY = -1
BODY_DYNAMIC, BODY_KINEMATIC, BODY_STATIC = 0, 1, 2
JOINT_FIXED, JOINT_REVOLUTE, JOINT_SPHERICAL, JOINT_PRISMATIC, JOINT_DISTANCE = range(5)
JOINT_TYPE_COUNT = 5
AABB_IS_INACTIVE, AABB_CONTACT_EVENTS, AABB_GENERATE_CONSTRAINTS, AABB_CUSTOM_FILTER, AABB_MODIFY_CONTACTS = 1, 2, 4, 8, 16
PAIR_CONTACT_EVENTS, PAIR_MODIFY_CONTACTS, PAIR_GENERATE_CONSTRAINTS, PAIR_NEEDS_HOOK = 1, 2, 4, 8
CFG_FAST_TRIG = 1
OK, ERR_INVALID_ARGUMENT, ERR_CUDA, ERR_OUT_OF_MEMORY, ERR_UNSUPPORTED, ERR_CAPACITY, ERR_NCCL = 0, -1, -2, -3, -4, -5, -6

_vp = C.c_void_p


class AvnConfig(C.Structure):
    _fields_ = [("abi_version", C.c_uint32), ("device", C.c_int32), ("scalar_bits", C.c_uint32), ("flags", C.c_uint32)]


class AvnStepParams(C.Structure):
    _fields_ = [
        ("dt", C.c_double), ("h", C.c_double), ("substeps", C.c_uint32), ("restitution_iterations", C.c_uint32),
        ("gravity", C.c_double * 3), ("contact_damping_ratio", C.c_double), ("contact_frequency_factor", C.c_double),
        ("max_overlap_solve_speed", C.c_double), ("warm_start_coefficient", C.c_double), ("restitution_threshold", C.c_double),
        ("length_unit", C.c_double), ("match_contacts", C.c_uint32), ("solver_iterations", C.c_uint32),
    ]


class AvnBodyColumns(C.Structure):
    _fields_ = [("count", C.c_uint32), ("_pad", C.c_uint32)] + [
        (n, _vp) for n in (
            "kind", "position", "rotation", "linear_velocity", "angular_velocity", "inverse_mass", "inverse_inertia_local",
            "center_of_mass", "locked_axes", "dominance", "linear_damping", "angular_damping", "gravity_scale",
            "linear_acceleration", "angular_acceleration", "max_linear_speed", "max_angular_speed", "integration_flags")]


class AvnManifoldColumns(C.Structure):
    _fields_ = [("count", C.c_uint32), ("point_count", C.c_uint32), ("color_offsets", C.c_uint32 * (GRAPH_COLOR_COUNT + 1))] + [
        (n, _vp) for n in (
            "body1", "body2", "normal", "friction", "restitution", "tangent_velocity", "point_offsets", "anchor1", "anchor2",
            "penetration", "normal_speed", "warm_start_normal_impulse", "warm_start_tangent_impulse", "normal_impulse")]


class AvnJointColumns(C.Structure):
    _fields_ = [("count", C.c_uint32), ("_pad", C.c_uint32)] + [
        (n, _vp) for n in (
            "body1", "body2", "local_anchor1", "local_anchor2", "local_basis1", "local_basis2", "axis", "limit_enabled",
            "limit_min", "limit_max", "limit2_min", "limit2_max", "compliance0", "compliance1", "compliance2",
            "damping_enabled", "damping_linear", "damping_angular", "force", "torque")]


class AvnJointSet(C.Structure):
    _fields_ = [("types", AvnJointColumns * JOINT_TYPE_COUNT)]


class AvnAabbColumns(C.Structure):
    _fields_ = [("count", C.c_uint32), ("retained_count", C.c_uint32)] + [
        (n, _vp) for n in ("collider", "body", "aabb_min", "aabb_max", "memberships", "filters", "flags", "order_out")] + [
        ("existing_pairs", _vp), ("existing_pair_count", C.c_uint64), ("joint_disabled_body_pairs", _vp), ("joint_disabled_pair_count", C.c_uint64)]


class AvnPairList(C.Structure):
    _fields_ = [("capacity", C.c_uint64), ("count", C.c_uint64)] + [(n, _vp) for n in ("collider1", "collider2", "body1", "body2", "flags")]


class AvnAabbParams(C.Structure):
    _fields_ = [("dt", C.c_double), ("contact_tolerance", C.c_double), ("default_speculative_margin", C.c_double)]


class AvnColliderColumns(C.Structure):
    _fields_ = [("count", C.c_uint32), ("_pad", C.c_uint32)] + [
        (n, _vp) for n in ("shape", "dims", "position", "rotation", "linear_velocity", "angular_velocity", "collision_margin", "speculative_margin",
                           "aabb_min", "aabb_max")]


class AvnTimings(C.Structure):
    _fields_ = [(n, C.c_float) for n in ("h2d_ms", "prepare_ms", "substep_loop_ms", "finalize_ms", "d2h_ms", "broad_phase_ms", "total_ms")] + [
        (n, C.c_uint32) for n in ("kernel_launches", "contact_constraint_count", "joint_levels", "active_colors", "launch_mode")]


class AvianError(RuntimeError):
    def __init__(self, status: int, message: str):
        super().__init__(f"avian_b200 error {status}: {message}")
        self.status = status


def _ptr(a):
    if a is None:
        return None
    assert a.flags["C_CONTIGUOUS"], "column must be C-contiguous"
    return a.ctypes.data


def default_step_params(dt: float = 1.0 / 60.0, substeps: int = 6, gravity=(0.0, -9.81, 0.0), scalar=np.float32, **kw) -> AvnStepParams:
    """SolverConfig / Gravity / SubstepCount defaults of the reference (solver/plugin.rs:291-302,
    integrator/mod.rs:158-162, solver/schedule.rs:187-191).  dt, h are derived like the reference derives them:
    Duration arithmetic in integer nanoseconds (solver/schedule.rs:195-200)."""
    p = AvnStepParams()
    dt_ns = round(dt * 1e9)                       # Duration::from_secs_f64 rounds to the nearest ns
    h_ns = round(dt_ns / 1e9 / substeps * 1e9)    # Duration::div_f64 = from_secs_f64(as_secs_f64() / rhs)
    p.dt = dt_ns / 1e9
    p.h = h_ns / 1e9
    p.substeps = substeps
    p.restitution_iterations = 1
    p.gravity[0], p.gravity[1], p.gravity[2] = gravity
    p.contact_damping_ratio = 10.0
    p.contact_frequency_factor = 1.5
    p.max_overlap_solve_speed = 4.0
    p.warm_start_coefficient = 1.0
    p.restitution_threshold = 1.0
    p.length_unit = 1.0
    p.match_contacts = 1
    p.solver_iterations = 1
    for k, v in kw.items():
        setattr(p, k, v)
    return p


@dataclass
class Bodies:
    """Host columns of the rigid bodies (numpy, dtype = scalar type).  Mirrors AvnBodyColumns."""
    kind: np.ndarray
    position: np.ndarray
    rotation: np.ndarray
    linear_velocity: np.ndarray
    angular_velocity: np.ndarray
    inverse_mass: np.ndarray
    inverse_inertia_local: np.ndarray
    center_of_mass: np.ndarray | None = None
    locked_axes: np.ndarray | None = None
    dominance: np.ndarray | None = None
    linear_damping: np.ndarray | None = None
    angular_damping: np.ndarray | None = None
    gravity_scale: np.ndarray | None = None
    linear_acceleration: np.ndarray | None = None
    angular_acceleration: np.ndarray | None = None
    max_linear_speed: np.ndarray | None = None
    max_angular_speed: np.ndarray | None = None
    integration_flags: np.ndarray | None = None

    @property
    def count(self) -> int:
        return int(self.position.shape[0])

    def as_struct(self) -> AvnBodyColumns:
        s = AvnBodyColumns()
        s.count = self.count
        for name, _ in AvnBodyColumns._fields_[2:]:
            setattr(s, name, _ptr(getattr(self, name)))
        return s

    def copy(self) -> "Bodies":
        return Bodies(**{k: (None if v is None else v.copy()) for k, v in self.__dict__.items()})


@dataclass
class Manifolds:
    color_offsets: np.ndarray          # uint32[25]
    body1: np.ndarray                  # int32[M]
    body2: np.ndarray
    normal: np.ndarray                 # [M,3]
    friction: np.ndarray
    restitution: np.ndarray
    point_offsets: np.ndarray          # uint32[M+1]
    anchor1: np.ndarray                # [P,3]
    anchor2: np.ndarray
    penetration: np.ndarray
    normal_speed: np.ndarray
    warm_start_normal_impulse: np.ndarray
    warm_start_tangent_impulse: np.ndarray   # [P,2]
    normal_impulse: np.ndarray
    tangent_velocity: np.ndarray | None = None

    @property
    def count(self) -> int:
        return int(self.body1.shape[0])

    def as_struct(self) -> AvnManifoldColumns:
        s = AvnManifoldColumns()
        s.count = self.count
        s.point_count = int(self.penetration.shape[0])
        for i in range(GRAPH_COLOR_COUNT + 1):
            s.color_offsets[i] = int(self.color_offsets[i])
        for name, _ in AvnManifoldColumns._fields_[3:]:
            setattr(s, name, _ptr(getattr(self, name)))
        return s

    def copy(self) -> "Manifolds":
        return Manifolds(**{k: (None if v is None else v.copy()) for k, v in self.__dict__.items()})


@dataclass
class Joints:
    """One typed joint array (AvnJointColumns)."""
    body1: np.ndarray
    body2: np.ndarray
    local_anchor1: np.ndarray
    local_anchor2: np.ndarray
    local_basis1: np.ndarray | None = None
    local_basis2: np.ndarray | None = None
    axis: np.ndarray | None = None
    limit_enabled: np.ndarray | None = None
    limit_min: np.ndarray | None = None
    limit_max: np.ndarray | None = None
    limit2_min: np.ndarray | None = None
    limit2_max: np.ndarray | None = None
    compliance0: np.ndarray | None = None
    compliance1: np.ndarray | None = None
    compliance2: np.ndarray | None = None
    damping_enabled: np.ndarray | None = None
    damping_linear: np.ndarray | None = None
    damping_angular: np.ndarray | None = None
    force: np.ndarray | None = None
    torque: np.ndarray | None = None

    @property
    def count(self) -> int:
        return int(self.body1.shape[0])

    def fill(self, s: AvnJointColumns) -> None:
        s.count = self.count
        for name, _ in AvnJointColumns._fields_[2:]:
            setattr(s, name, _ptr(getattr(self, name)))

    def copy(self) -> "Joints":
        return Joints(**{k: (None if v is None else v.copy()) for k, v in self.__dict__.items()})


@dataclass
class JointSet:
    types: dict = field(default_factory=dict)   # AvnJointType -> Joints

    def as_struct(self) -> AvnJointSet:
        s = AvnJointSet()
        for t, j in self.types.items():
            j.fill(s.types[t])
        return s

    def copy(self) -> "JointSet":
        return JointSet({t: j.copy() for t, j in self.types.items()})

    @property
    def count(self) -> int:
        return sum(j.count for j in self.types.values())


@dataclass
class Aabbs:
    collider: np.ndarray     # uint32[C]
    body: np.ndarray         # uint32[C]
    aabb_min: np.ndarray     # [C,3]
    aabb_max: np.ndarray
    flags: np.ndarray        # uint8[C]
    memberships: np.ndarray | None = None
    filters: np.ndarray | None = None
    order_out: np.ndarray | None = None
    existing_pairs: np.ndarray | None = None          # uint64
    joint_disabled_body_pairs: np.ndarray | None = None
    retained_count: int | None = None                 # out: entries of order_out (intervals with a non-finite AABB are dropped)

    def as_struct(self) -> AvnAabbColumns:
        s = AvnAabbColumns()
        s.count = int(self.collider.shape[0])
        for name in ("collider", "body", "aabb_min", "aabb_max", "memberships", "filters", "flags", "order_out"):
            setattr(s, name, _ptr(getattr(self, name)))
        s.existing_pairs = _ptr(self.existing_pairs)
        s.existing_pair_count = 0 if self.existing_pairs is None else int(self.existing_pairs.shape[0])
        s.joint_disabled_body_pairs = _ptr(self.joint_disabled_body_pairs)
        s.joint_disabled_pair_count = 0 if self.joint_disabled_body_pairs is None else int(self.joint_disabled_body_pairs.shape[0])
        return s


@dataclass
class PairList:
    collider1: np.ndarray
    collider2: np.ndarray
    body1: np.ndarray
    body2: np.ndarray
    flags: np.ndarray
    count: int = 0

    @staticmethod
    def empty(capacity: int) -> "PairList":
        u = lambda: np.zeros(capacity, dtype=np.uint32)
        return PairList(u(), u(), u(), u(), np.zeros(capacity, dtype=np.uint8))

    def as_struct(self) -> AvnPairList:
        s = AvnPairList()
        s.capacity = int(self.collider1.shape[0])
        for name in ("collider1", "collider2", "body1", "body2", "flags"):
            setattr(s, name, _ptr(getattr(self, name)))
        return s

    def trimmed(self) -> "PairList":
        n = self.count
        return PairList(self.collider1[:n], self.collider2[:n], self.body1[:n], self.body2[:n], self.flags[:n], n)


class AvnEdgeManifolds(C.Structure):
    _fields_ = [("count", C.c_uint32), ("edge_capacity", C.c_uint32), ("color_offsets", C.c_uint32 * (GRAPH_COLOR_COUNT + 1))] + [
        (n, _vp) for n in ("edge", "body1", "body2", "friction", "restitution", "point_count", "normal", "anchor1", "anchor2", "penetration",
                           "normal_speed", "warm_start_normal_impulse", "warm_start_tangent_impulse", "normal_impulse")]


class AvnNarrowParams(C.Structure):
    _fields_ = [("dt", C.c_double), ("contact_tolerance", C.c_double)]


class AvnContactGraphConfig(C.Structure):
    _fields_ = [("body_count", C.c_uint32), ("collider_count", C.c_uint32), ("body_kind", _vp), ("friction", _vp), ("restitution", _vp)]


class AvnContactStep(C.Structure):
    _fields_ = [(n, C.c_uint32) for n in ("rows_high_water", "rows_live", "pairs_added", "pairs_removed", "started_touching", "stopped_touching",
                                         "manifold_count", "colouring_rounds", "any_restitution", "_pad")] + [("color_offsets", C.c_uint32 * (GRAPH_COLOR_COUNT + 1))]


CONTACTS_TAKE_BROADPHASE_PAIRS, CONTACTS_SHAPES_UNCHANGED = 1, 2
BODIES_STATIC_UNCHANGED = 1


class AvnIslandsConfig(C.Structure):
    _fields_ = [("body_count", C.c_uint32), ("joint_count", C.c_uint32), ("body_kind", _vp), ("sleep_threshold_linear", _vp), ("sleep_threshold_angular", _vp),
                ("sleeping_disabled", _vp), ("joint_body1", _vp), ("joint_body2", _vp), ("time_to_sleep", C.c_float), ("length_unit", C.c_float)]


class AvnIslandsStep(C.Structure):
    _fields_ = [("delta_secs", C.c_float), ("_pad", C.c_uint32), ("linear_velocity", _vp), ("angular_velocity", _vp), ("wake", _vp), ("island", _vp),
                ("sleeping", _vp), ("sleep_timer", _vp)] + [(n, C.c_uint32) for n in ("island_count", "sleeping_islands", "islands_put_to_sleep", "islands_woken",
                                                                                     "split_bodies", "merges")]


class AvnNarrowInput(C.Structure):
    _fields_ = [("pair_count", C.c_uint32), ("collider_count", C.c_uint32), ("body_count", C.c_uint32), ("_pad", C.c_uint32)] + [
        (n, _vp) for n in ("collider1", "collider2", "body1", "body2", "shape", "dims", "position", "rotation", "linear_velocity", "angular_velocity",
                           "aabb_min", "aabb_max")]


class AvnRawManifolds(C.Structure):
    _fields_ = [(n, _vp) for n in ("point_count", "disjoint", "normal", "anchor1", "anchor2", "penetration", "normal_speed")]


class AvnBoundary(C.Structure):
    _fields_ = [("count", C.c_uint32), ("record_count", C.c_uint32), ("rank", C.c_uint32), ("world", C.c_uint32),
                ("body", _vp), ("source", _vp), ("owner_rank", _vp)]


def bind_abi(lib: C.CDLL, prefix: str = "avn") -> None:
    """Declare argument/return types of every entry point of include/avian_b200.h on `lib`."""
    P = C.POINTER
    sig = {
        "create": ([P(AvnConfig), P(_vp)], C.c_int),
        "destroy": ([_vp], None),
        "last_error": ([_vp], C.c_char_p),
        "abi_version": ([], C.c_uint32),
        "alloc_pinned": ([_vp, C.c_size_t, P(_vp)], C.c_int),
        "free_pinned": ([_vp, _vp], C.c_int),
        "solver_step": ([_vp, P(AvnStepParams), P(AvnBodyColumns), P(AvnManifoldColumns), P(AvnJointSet)], C.c_int),
        "solver_upload": ([_vp, P(AvnStepParams), P(AvnBodyColumns), P(AvnManifoldColumns), P(AvnJointSet)], C.c_int),
        "solver_run": ([_vp], C.c_int),
        "solver_download": ([_vp], C.c_int),
        "broadphase": ([_vp, P(AvnAabbColumns), P(AvnPairList)], C.c_int),
        "broadphase_upload": ([_vp, P(AvnAabbColumns)], C.c_int),
        "broadphase_run": ([_vp], C.c_int),
        "broadphase_download": ([_vp, P(AvnPairList)], C.c_int),
        "get_timings": ([_vp, P(AvnTimings)], C.c_int),
        "joint_levels": ([P(AvnBodyColumns), P(AvnJointSet), _vp, P(C.c_uint32)], C.c_int),
        "update_aabbs": ([_vp, P(AvnAabbParams), P(AvnColliderColumns)], C.c_int),
        "solver_run_range": ([_vp, C.c_uint32, C.c_uint32, C.c_uint32], C.c_int),
        "solver_set_boundary": ([_vp, P(AvnBoundary)], C.c_int),
        "solver_boundary_snapshot": ([_vp], C.c_int),
        "solver_boundary_pack": ([_vp, _vp], C.c_int),
        "solver_boundary_apply": ([_vp, _vp], C.c_int),
        "solver_needs_restitution": ([_vp, P(C.c_int)], C.c_int),
        "get_stream": ([_vp, P(_vp)], C.c_int),
        "solver_step_partitioned": ([_vp], C.c_int),
        "comm_unique_id": ([_vp, _vp], C.c_int),
        "comm_init": ([_vp, C.c_uint32, C.c_uint32, _vp], C.c_int),
        "comm_destroy": ([_vp], C.c_int),
        "comm_all_gather": ([_vp, _vp, _vp, C.c_size_t], C.c_int),
        "narrow_phase": ([_vp, P(AvnNarrowParams), P(AvnNarrowInput), P(AvnRawManifolds)], C.c_int),
        "solver_upload_edges": ([_vp, P(AvnStepParams), P(AvnBodyColumns), P(AvnEdgeManifolds), P(AvnJointSet)], C.c_int),
        "solver_upload_graph": ([_vp, P(AvnStepParams), P(AvnBodyColumns), P(AvnEdgeManifolds), P(AvnJointSet)], C.c_int),
        "contacts_reserve": ([_vp, C.c_uint32], C.c_int),
        "contacts_add": ([_vp, C.c_uint32, _vp, _vp, _vp, _vp, _vp], C.c_int),
        "contacts_remove": ([_vp, C.c_uint32, _vp], C.c_int),
        "contacts_narrow_phase": ([_vp, P(AvnNarrowParams), P(AvnNarrowInput), C.c_uint32, C.c_double, _vp, _vp], C.c_int),
        "contacts_download_impulses": ([_vp, _vp, _vp, _vp], C.c_int),
        "contacts_configure": ([_vp, P(AvnContactGraphConfig)], C.c_int),
        "contacts_step": ([_vp, P(AvnNarrowParams), P(AvnNarrowInput), C.c_uint32, C.c_double, C.c_uint32, P(AvnContactStep)], C.c_int),
        "solver_upload_resident": ([_vp, P(AvnStepParams), P(AvnBodyColumns), P(AvnJointSet)], C.c_int),
        "broadphase_download_order": ([_vp, P(C.c_uint64)], C.c_int),
        "contacts_download_graph": ([_vp, C.c_uint32, _vp, _vp, _vp, _vp, _vp, _vp], C.c_int),
        "solver_prefetch_bodies": ([_vp, P(AvnBodyColumns), C.c_uint32], C.c_int),
        "islands_configure": ([_vp, P(AvnIslandsConfig)], C.c_int),
        "islands_step": ([_vp, P(AvnIslandsStep)], C.c_int),
    }
    for name, (argtypes, restype) in sig.items():
        fn = getattr(lib, f"{prefix}_{name}")
        fn.argtypes = argtypes
        fn.restype = restype


ABI_SYMBOLS = [
    "avn_create", "avn_destroy", "avn_last_error", "avn_abi_version", "avn_alloc_pinned", "avn_free_pinned", "avn_solver_step",
    "avn_solver_upload", "avn_solver_run", "avn_solver_download", "avn_broadphase", "avn_broadphase_upload", "avn_broadphase_run",
    "avn_broadphase_download", "avn_get_timings", "avn_joint_levels", "avn_update_aabbs", "avn_solver_run_range", "avn_solver_set_boundary",
    "avn_solver_boundary_snapshot", "avn_solver_boundary_pack", "avn_solver_boundary_apply", "avn_solver_needs_restitution", "avn_get_stream",
    "avn_solver_step_partitioned", "avn_comm_unique_id", "avn_comm_init", "avn_comm_destroy", "avn_comm_all_gather", "avn_narrow_phase", "avn_solver_upload_edges", "avn_solver_upload_graph",
    "avn_contacts_reserve", "avn_contacts_add", "avn_contacts_remove", "avn_contacts_narrow_phase", "avn_contacts_download_impulses",
    "avn_contacts_configure", "avn_contacts_step", "avn_solver_upload_resident", "avn_broadphase_download_order", "avn_contacts_download_graph",
    "avn_solver_prefetch_bodies", "avn_islands_configure", "avn_islands_step"]

RUN_PREPARE, RUN_RESTITUTION, RUN_FINALIZE = 1, 2, 4
COMM_ID_BYTES = 128
BOUNDARY_RECORD_SCALARS = 16


@dataclass
class Colliders:
    """Collider columns of avn_update_aabbs (AvnColliderColumns)."""
    shape: np.ndarray            # uint8[C]
    dims: np.ndarray             # [C,3]
    position: np.ndarray
    rotation: np.ndarray
    linear_velocity: np.ndarray | None = None
    angular_velocity: np.ndarray | None = None
    collision_margin: np.ndarray | None = None
    speculative_margin: np.ndarray | None = None
    aabb_min: np.ndarray | None = None
    aabb_max: np.ndarray | None = None

    def as_struct(self) -> "AvnColliderColumns":
        n = int(self.position.shape[0])
        if self.aabb_min is None:
            self.aabb_min = np.zeros((n, 3), dtype=self.position.dtype)
            self.aabb_max = np.zeros((n, 3), dtype=self.position.dtype)
        s = AvnColliderColumns()
        s.count = n
        for name, _ in AvnColliderColumns._fields_[2:]:
            setattr(s, name, _ptr(getattr(self, name)))
        return s


def aabb_params(dt: float, contact_tolerance: float = 0.005, default_speculative_margin: float = float("inf")) -> "AvnAabbParams":
    """NarrowPhaseConfig defaults (narrow_phase/mod.rs:247-255) times PhysicsLengthUnit = 1."""
    return AvnAabbParams(dt, contact_tolerance, default_speculative_margin)


def joint_levels(bodies: "Bodies", joints: "JointSet"):
    """avn_joint_levels: (level per joint in the reference's global order, number of levels).  Host-only."""
    lib = load_library()
    b, j = bodies.as_struct(), joints.as_struct()
    out = np.zeros(joints.count, dtype=np.uint32)
    n = C.c_uint32(0)
    st = lib.avn_joint_levels(C.byref(b), C.byref(j), _ptr(out), C.byref(n))
    if st != OK:
        raise AvianError(st, lib.avn_last_error(None).decode())
    return out, int(n.value)

_lib = None


def load_library() -> C.CDLL:
    """Load (building if needed) libavian_b200.so.  Raises if it cannot be had: no fallback."""
    global _lib
    if _lib is None:
        path = _build.build_cuda()
        lib = C.CDLL(str(path))
        bind_abi(lib)
        _lib = lib
    return _lib


class Context:
    """An AvnContext: one CUDA device, one stream, persistent device buffers."""

    def __init__(self, device: int = 0, scalar=np.float32, flags: int = 0):
        self.lib = load_library()
        self.device = int(device)
        self.scalar = np.dtype(scalar)
        cfg = AvnConfig(self.lib.avn_abi_version(), device, 32 if self.scalar == np.float32 else 64, flags)
        h = _vp()
        st = self.lib.avn_create(C.byref(cfg), C.byref(h))
        if st != OK:
            raise AvianError(st, self.lib.avn_last_error(None).decode())
        self.handle = h
        self._pinned: list[int] = []
        self._keep = None

    def close(self) -> None:
        if getattr(self, "handle", None):
            for p in self._pinned:
                self.lib.avn_free_pinned(self.handle, _vp(p))
            self._pinned.clear()
            self.lib.avn_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _check(self, st: int) -> None:
        if st != OK:
            raise AvianError(st, self.lib.avn_last_error(self.handle).decode())

    def pinned(self, shape, dtype) -> np.ndarray:
        """A numpy array backed by page-locked memory from avn_alloc_pinned (lives as long as the context)."""
        dtype = np.dtype(dtype)
        n = int(np.prod(shape)) * dtype.itemsize
        p = _vp()
        self._check(self.lib.avn_alloc_pinned(self.handle, max(n, 1), C.byref(p)))
        self._pinned.append(p.value)
        buf = (C.c_byte * max(n, 1)).from_address(p.value)
        return np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)

    def pin_like(self, a: np.ndarray | None) -> np.ndarray | None:
        if a is None:
            return None
        out = self.pinned(a.shape, a.dtype)
        out[...] = a
        return out

    # ---- solver stage ------------------------------------------------------------------------------------
    def _solver_args(self, params, bodies: Bodies, manifolds: Manifolds | None, joints: JointSet | None):
        b = bodies.as_struct()
        m = manifolds.as_struct() if manifolds is not None and manifolds.count else None
        j = joints.as_struct() if joints is not None and joints.count else None
        self._keep = (params, bodies, manifolds, joints, b, m, j)
        return (C.byref(params), C.byref(b), C.byref(m) if m is not None else None, C.byref(j) if j is not None else None)

    def solver_step(self, params, bodies, manifolds=None, joints=None) -> None:
        self._check(self.lib.avn_solver_step(self.handle, *self._solver_args(params, bodies, manifolds, joints)))

    def solver_upload(self, params, bodies, manifolds=None, joints=None) -> None:
        self._check(self.lib.avn_solver_upload(self.handle, *self._solver_args(params, bodies, manifolds, joints)))

    def solver_run(self) -> None:
        self._check(self.lib.avn_solver_run(self.handle))

    def solver_download(self) -> None:
        self._check(self.lib.avn_solver_download(self.handle))

    # ---- broad phase -------------------------------------------------------------------------------------
    def broadphase(self, aabbs: Aabbs, capacity: int | None = None) -> PairList:
        """Runs the sweep; grows the output list and retries once when the capacity guess was too small."""
        cap = capacity if capacity is not None else max(1024, 16 * int(aabbs.collider.shape[0]))
        a = aabbs.as_struct()
        self._keep_bp = (aabbs, a)
        self._check(self.lib.avn_broadphase_upload(self.handle, C.byref(a)))
        self._check(self.lib.avn_broadphase_run(self.handle))
        out = PairList.empty(cap)
        s = out.as_struct()
        st = self.lib.avn_broadphase_download(self.handle, C.byref(s))
        if st == ERR_CAPACITY:
            out = PairList.empty(int(s.count))
            s = out.as_struct()
            st = self.lib.avn_broadphase_download(self.handle, C.byref(s))
        self._check(st)
        out.count = int(s.count)
        aabbs.retained_count = int(a.retained_count)
        return out.trimmed()

    def broadphase_upload(self, aabbs: Aabbs) -> None:
        a = aabbs.as_struct()
        self._keep_bp = (aabbs, a)
        self._check(self.lib.avn_broadphase_upload(self.handle, C.byref(a)))

    def broadphase_run(self) -> None:
        self._check(self.lib.avn_broadphase_run(self.handle))

    def broadphase_download(self, out: PairList) -> PairList:
        s = out.as_struct()
        self._check(self.lib.avn_broadphase_download(self.handle, C.byref(s)))
        out.count = int(s.count)
        self._keep_bp[0].retained_count = int(self._keep_bp[1].retained_count)
        return out

    # ---- x-slab partition (include/avian_b200.h "one coupled scene over several GPUs")
    def solver_run_range(self, first: int, count: int, flags: int) -> None:
        self._check(self.lib.avn_solver_run_range(self.handle, first, count, flags))

    def solver_set_boundary(self, body: np.ndarray, source: np.ndarray, owner_rank: np.ndarray, record_count: int, rank: int, world: int) -> None:
        body, source, owner_rank = (np.ascontiguousarray(x, dtype=np.int32) for x in (body, source, owner_rank))
        assert source.size == body.shape[0] * world
        b = AvnBoundary(int(body.shape[0]), int(record_count), int(rank), int(world), _ptr(body), _ptr(source), _ptr(owner_rank))
        self._check(self.lib.avn_solver_set_boundary(self.handle, C.byref(b)))

    def solver_boundary_snapshot(self) -> None:
        self._check(self.lib.avn_solver_boundary_snapshot(self.handle))

    def solver_boundary_pack(self, device_ptr: int) -> None:
        self._check(self.lib.avn_solver_boundary_pack(self.handle, _vp(device_ptr)))

    def solver_boundary_apply(self, device_ptr: int) -> None:
        self._check(self.lib.avn_solver_boundary_apply(self.handle, _vp(device_ptr)))

    def solver_needs_restitution(self) -> bool:
        out = C.c_int(0)
        self._check(self.lib.avn_solver_needs_restitution(self.handle, C.byref(out)))
        return bool(out.value)

    # ---- communicator (NCCL inside the library; one process per GPU)
    def comm_unique_id(self) -> bytes:
        buf = C.create_string_buffer(COMM_ID_BYTES)
        self._check(self.lib.avn_comm_unique_id(self.handle, C.cast(buf, _vp)))
        return buf.raw

    def comm_init(self, rank: int, world: int, unique_id: bytes | None = None) -> None:
        buf = None if unique_id is None else C.create_string_buffer(bytes(unique_id), COMM_ID_BYTES)
        self._check(self.lib.avn_comm_init(self.handle, rank, world, None if buf is None else C.cast(buf, _vp)))

    def comm_destroy(self) -> None:
        self._check(self.lib.avn_comm_destroy(self.handle))

    def comm_all_gather(self, send_device_ptr: int, recv_device_ptr: int, bytes_per_rank: int) -> None:
        self._check(self.lib.avn_comm_all_gather(self.handle, _vp(send_device_ptr), _vp(recv_device_ptr), bytes_per_rank))

    def solver_step_partitioned(self) -> None:
        """The partitioned solver stage of this rank (after solver_upload + solver_set_boundary); then solver_download."""
        self._check(self.lib.avn_solver_step_partitioned(self.handle))

    def stream(self) -> int:
        """The context's cudaStream_t as an integer (torch.cuda.ExternalStream(ptr) orders a collective with the launches)."""
        out = _vp()
        self._check(self.lib.avn_get_stream(self.handle, C.byref(out)))
        return int(out.value or 0)

    def solver_step_edges(self, params, bodies: Bodies, graph: dict, edges: dict, joints: JointSet | None = None) -> None:
        """avn_solver_upload_edges + run + download.  graph = dict(color_offsets, edge, body1, body2, friction, restitution) (per manifold);
        edges = dict(point_count, normal, anchor1, anchor2, penetration, normal_speed, warm_start_normal_impulse,
        warm_start_tangent_impulse, normal_impulse) (edge-indexed, 4 slots per edge; the three impulse columns are updated in place)."""
        b = bodies.as_struct()
        j = joints.as_struct() if joints is not None and joints.count else None
        em = AvnEdgeManifolds()
        em.count = int(graph["edge"].shape[0])
        em.edge_capacity = int(edges["point_count"].shape[0])
        for i in range(GRAPH_COLOR_COUNT + 1):
            em.color_offsets[i] = int(graph["color_offsets"][i])
        keep = []
        def col(a, dtype):
            a = np.ascontiguousarray(a, dtype=dtype)
            keep.append(a)
            return a.ctypes.data
        em.edge = col(graph["edge"], np.uint32); em.body1 = col(graph["body1"], np.int32); em.body2 = col(graph["body2"], np.int32)
        em.friction = col(graph["friction"], self.scalar); em.restitution = col(graph["restitution"], self.scalar)
        em.point_count = col(edges["point_count"], np.uint8)
        for k in ("normal", "anchor1", "anchor2", "penetration", "normal_speed"):
            setattr(em, k, col(edges[k], self.scalar))
        for k in ("warm_start_normal_impulse", "warm_start_tangent_impulse", "normal_impulse"):
            assert edges[k].flags["C_CONTIGUOUS"] and edges[k].dtype == self.scalar
            setattr(em, k, edges[k].ctypes.data)
        self._keep = (params, bodies, graph, edges, joints, b, em, j, keep)
        self._check(self.lib.avn_solver_upload_edges(self.handle, C.byref(params), C.byref(b), C.byref(em) if em.count else None, C.byref(j) if j is not None else None))
        self._check(self.lib.avn_solver_run(self.handle))
        self._check(self.lib.avn_solver_download(self.handle))

    # ---- device-resident contact edges (include/avian_b200.h)
    def contacts_reserve(self, capacity: int) -> None:
        self._check(self.lib.avn_contacts_reserve(self.handle, int(capacity)))

    def contacts_add(self, ids, c1, c2, b1, b2) -> None:
        a = [np.ascontiguousarray(x, dtype=np.uint32) for x in (ids, c1, c2, b1, b2)]
        self._check(self.lib.avn_contacts_add(self.handle, int(a[0].shape[0]), *(x.ctypes.data for x in a)))

    def contacts_remove(self, ids) -> None:
        ids = np.ascontiguousarray(ids, dtype=np.uint32)
        self._check(self.lib.avn_contacts_remove(self.handle, int(ids.shape[0]), ids.ctypes.data))

    def contacts_narrow_phase(self, dt: float, contact_tolerance: float, colliders: dict, lin_vel, ang_vel, capacity: int, match_contacts: bool = True,
                              length_unit: float = 1.0):
        """Geometry + match_contacts for every live row.  Returns (point_count[capacity], disjoint[capacity])."""
        dt_ = self.scalar
        cols = {k: (None if colliders.get(k) is None else np.ascontiguousarray(colliders[k], dtype=(np.uint8 if k == "shape" else dt_)))
                for k in ("shape", "dims", "position", "rotation", "aabb_min", "aabb_max")}
        lv, av = np.ascontiguousarray(lin_vel, dtype=dt_), np.ascontiguousarray(ang_vel, dtype=dt_)
        inp = AvnNarrowInput(0, int(cols["position"].shape[0]), int(lv.shape[0]), 0, None, None, None, None, _ptr(cols["shape"]), _ptr(cols["dims"]),
                             _ptr(cols["position"]), _ptr(cols["rotation"]), _ptr(lv), _ptr(av), _ptr(cols["aabb_min"]), _ptr(cols["aabb_max"]))
        count, disjoint = np.zeros(capacity, dtype=np.uint8), np.zeros(capacity, dtype=np.uint8)
        prm = AvnNarrowParams(float(dt), float(contact_tolerance))
        self._check(self.lib.avn_contacts_narrow_phase(self.handle, C.byref(prm), C.byref(inp), 1 if match_contacts else 0, float(length_unit),
                                                       count.ctypes.data, disjoint.ctypes.data))
        return count, disjoint

    def solver_step_graph(self, params, bodies: Bodies, graph: dict, joints: JointSet | None = None, reuse_graph: bool = False) -> None:
        """avn_solver_upload_graph + run + download: the manifolds come from the resident rows, graph = dict(color_offsets, edge, body1, body2,
        friction, restitution)."""
        b = bodies.as_struct()
        j = joints.as_struct() if joints is not None and joints.count else None
        em = AvnEdgeManifolds()
        em.count = int(graph["edge"].shape[0])
        for i in range(GRAPH_COLOR_COUNT + 1):
            em.color_offsets[i] = int(graph["color_offsets"][i])
        keep = [np.ascontiguousarray(graph["edge"], dtype=np.uint32), np.ascontiguousarray(graph["body1"], dtype=np.int32),
                np.ascontiguousarray(graph["body2"], dtype=np.int32), np.ascontiguousarray(graph["friction"], dtype=self.scalar),
                np.ascontiguousarray(graph["restitution"], dtype=self.scalar)]
        if not reuse_graph:     # reuse: edge stays NULL = "the list of the previous call is still resident"
            em.edge, em.body1, em.body2, em.friction, em.restitution = (x.ctypes.data for x in keep)
        self._keep = (params, bodies, graph, joints, b, em, j, keep)
        self._check(self.lib.avn_solver_upload_graph(self.handle, C.byref(params), C.byref(b), C.byref(em) if em.count else None, C.byref(j) if j is not None else None))
        self._check(self.lib.avn_solver_run(self.handle))
        self._check(self.lib.avn_solver_download(self.handle))

    # ---- the ContactGraph + ConstraintGraph on the device (include/avian_b200.h avn_contacts_configure / _step)
    def contacts_configure(self, body_kind, collider_count: int, friction=None, restitution=None) -> None:
        kind = np.ascontiguousarray(body_kind, dtype=np.uint8)
        fr = None if friction is None else np.ascontiguousarray(friction, dtype=np.float64)
        re = None if restitution is None else np.ascontiguousarray(restitution, dtype=np.float64)
        cfg = AvnContactGraphConfig(int(kind.shape[0]), int(collider_count), _ptr(kind), _ptr(fr), _ptr(re))
        self._check(self.lib.avn_contacts_configure(self.handle, C.byref(cfg)))

    def solver_prefetch_bodies(self, bodies: Bodies, static_unchanged: bool = False) -> None:
        """avn_solver_prefetch_bodies: the body columns of the next solver upload start moving to the device now (second stream)."""
        b = bodies.as_struct()
        self._keep_prefetch = (bodies, b)
        self._check(self.lib.avn_solver_prefetch_bodies(self.handle, C.byref(b), BODIES_STATIC_UNCHANGED if static_unchanged else 0))

    def contacts_step(self, dt: float, contact_tolerance: float, colliders: dict, lin_vel, ang_vel, match_contacts: bool = True, take_pairs: bool = True,
                      length_unit: float = 1.0, shapes_unchanged: bool = False) -> dict:
        """avn_contacts_step: (the last broad phase's new pairs ->) rows, geometry + matching, status loop, graphs, colour-major list — all on the
        device.  Returns the step's counters and the colour offsets of the list."""
        dt_ = self.scalar
        cols = {k: (None if colliders.get(k) is None else np.ascontiguousarray(colliders[k], dtype=(np.uint8 if k == "shape" else dt_)))
                for k in ("shape", "dims", "position", "rotation", "aabb_min", "aabb_max")}
        lv, av = np.ascontiguousarray(lin_vel, dtype=dt_), np.ascontiguousarray(ang_vel, dtype=dt_)
        inp = AvnNarrowInput(0, int(cols["position"].shape[0]), int(lv.shape[0]), 0, None, None, None, None, _ptr(cols["shape"]), _ptr(cols["dims"]),
                             _ptr(cols["position"]), _ptr(cols["rotation"]), _ptr(lv), _ptr(av), _ptr(cols["aabb_min"]), _ptr(cols["aabb_max"]))
        prm = AvnNarrowParams(float(dt), float(contact_tolerance))
        out = AvnContactStep()
        self._check(self.lib.avn_contacts_step(self.handle, C.byref(prm), C.byref(inp), 1 if match_contacts else 0, float(length_unit),
                                               (CONTACTS_TAKE_BROADPHASE_PAIRS if take_pairs else 0) | (CONTACTS_SHAPES_UNCHANGED if shapes_unchanged else 0),
                                               C.byref(out)))
        st = {n: int(getattr(out, n)) for n, _ in AvnContactStep._fields_ if n not in ("_pad", "color_offsets")}
        st["color_offsets"] = np.array(list(out.color_offsets), dtype=np.uint32)
        return st

    def solver_step_resident(self, params, bodies: Bodies, joints: JointSet | None = None) -> None:
        """avn_solver_upload_resident + run + download: manifolds AND constraint graph come from the contact store on the device."""
        b = bodies.as_struct()
        j = joints.as_struct() if joints is not None and joints.count else None
        self._keep = (params, bodies, joints, b, j)
        self._check(self.lib.avn_solver_upload_resident(self.handle, C.byref(params), C.byref(b), C.byref(j) if j is not None else None))
        self._check(self.lib.avn_solver_run(self.handle))
        self._check(self.lib.avn_solver_download(self.handle))

    def broadphase_download_order(self) -> int:
        n = C.c_uint64(0)
        self._check(self.lib.avn_broadphase_download_order(self.handle, C.byref(n)))
        self._keep_bp[0].retained_count = int(self._keep_bp[1].retained_count)
        return int(n.value)

    def contacts_download_graph(self, capacity: int, manifold_count: int) -> dict:
        out = {"collider1": np.zeros(capacity, dtype=np.uint32), "collider2": np.zeros(capacity, dtype=np.uint32), "live": np.zeros(capacity, dtype=np.uint8),
               "touching": np.zeros(capacity, dtype=np.uint8), "colour": np.zeros(capacity, dtype=np.int8), "edge": np.zeros(manifold_count, dtype=np.uint32)}
        # edge_list receives the WHOLE colour-major list of the last step: only ask for it with a buffer of that size
        ptrs = [out[k].ctypes.data for k in ("collider1", "collider2", "live", "touching", "colour")] + [out["edge"].ctypes.data if manifold_count else None]
        self._check(self.lib.avn_contacts_download_graph(self.handle, int(capacity), *ptrs))
        return out

    # ---- persistent islands + sleeping decisions (include/avian_b200.h avn_islands_configure / _step)
    def islands_configure(self, body_kind, joints=None, thr_lin=None, thr_ang=None, disabled=None, time_to_sleep: float = 0.5, length_unit: float = 1.0) -> None:
        kind = np.ascontiguousarray(body_kind, dtype=np.uint8)
        f32 = lambda a: None if a is None else np.ascontiguousarray(a, dtype=np.float32)
        tl, ta = f32(thr_lin), f32(thr_ang)
        dis = None if disabled is None else np.ascontiguousarray(disabled, dtype=np.uint8)
        j1 = j2 = None
        nj = 0
        if joints is not None and len(joints):
            jj = np.ascontiguousarray(joints, dtype=np.uint32).reshape(-1, 2)
            j1, j2, nj = np.ascontiguousarray(jj[:, 0]), np.ascontiguousarray(jj[:, 1]), int(jj.shape[0])
        cfg = AvnIslandsConfig(int(kind.shape[0]), nj, _ptr(kind), _ptr(tl), _ptr(ta), _ptr(dis), _ptr(j1), _ptr(j2), float(time_to_sleep), float(length_unit))
        self._check(self.lib.avn_islands_configure(self.handle, C.byref(cfg)))
        self._isl_B = int(kind.shape[0])

    def islands_step(self, delta_secs: float, lin_vel, ang_vel, wake=None) -> dict:
        B = self._isl_B
        lv, av = np.ascontiguousarray(lin_vel, dtype=self.scalar), np.ascontiguousarray(ang_vel, dtype=self.scalar)
        wk = None if wake is None else np.ascontiguousarray(wake, dtype=np.uint8)
        out = {"island": np.zeros(B, dtype=np.uint32), "sleeping": np.zeros(B, dtype=np.uint8), "sleep_timer": np.zeros(B, dtype=np.float32)}
        st = AvnIslandsStep(float(delta_secs), 0, _ptr(lv), _ptr(av), _ptr(wk), _ptr(out["island"]), _ptr(out["sleeping"]), _ptr(out["sleep_timer"]))
        self._check(self.lib.avn_islands_step(self.handle, C.byref(st)))
        for n in ("island_count", "sleeping_islands", "islands_put_to_sleep", "islands_woken", "split_bodies", "merges"):
            out[n] = int(getattr(st, n))
        return out

    def contacts_download_impulses(self, capacity: int):
        wn, wt, ni = (np.zeros((capacity, 4), dtype=self.scalar), np.zeros((capacity, 4, 2), dtype=self.scalar), np.zeros((capacity, 4), dtype=self.scalar))
        self._check(self.lib.avn_contacts_download_impulses(self.handle, wn.ctypes.data, wt.ctypes.data, ni.ctypes.data))
        return wn, wt, ni

    def narrow_phase(self, dt: float, contact_tolerance: float, pairs, colliders: dict, lin_vel: np.ndarray, ang_vel: np.ndarray) -> dict:
        """avn_narrow_phase: pairs = (collider1, collider2, body1, body2) uint32 arrays; colliders = dict(shape, dims, position, rotation,
        aabb_min=None, aabb_max=None).  Returns the raw manifold columns (4 point slots per pair)."""
        c1, c2, b1, b2 = (np.ascontiguousarray(x, dtype=np.uint32) for x in pairs)
        n, dt_ = int(c1.shape[0]), self.scalar
        cols = {k: (None if colliders.get(k) is None else np.ascontiguousarray(colliders[k], dtype=(np.uint8 if k == "shape" else dt_)))
                for k in ("shape", "dims", "position", "rotation", "aabb_min", "aabb_max")}
        lv, av = np.ascontiguousarray(lin_vel, dtype=dt_), np.ascontiguousarray(ang_vel, dtype=dt_)
        inp = AvnNarrowInput(n, int(cols["position"].shape[0]), int(lv.shape[0]), 0, _ptr(c1), _ptr(c2), _ptr(b1), _ptr(b2), _ptr(cols["shape"]),
                             _ptr(cols["dims"]), _ptr(cols["position"]), _ptr(cols["rotation"]), _ptr(lv), _ptr(av), _ptr(cols["aabb_min"]),
                             _ptr(cols["aabb_max"]))
        out = {"point_count": np.zeros(n, dtype=np.uint8), "disjoint": np.zeros(n, dtype=np.uint8), "normal": np.zeros((n, 3), dtype=dt_),
               "anchor1": np.zeros((n, 4, 3), dtype=dt_), "anchor2": np.zeros((n, 4, 3), dtype=dt_), "penetration": np.zeros((n, 4), dtype=dt_),
               "normal_speed": np.zeros((n, 4), dtype=dt_)}
        raw = AvnRawManifolds(*(_ptr(out[k]) for k in ("point_count", "disjoint", "normal", "anchor1", "anchor2", "penetration", "normal_speed")))
        prm = AvnNarrowParams(float(dt), float(contact_tolerance))
        self._check(self.lib.avn_narrow_phase(self.handle, C.byref(prm), C.byref(inp), C.byref(raw)))
        return out

    def update_aabbs(self, params: "AvnAabbParams", colliders: "Colliders") -> None:
        c = colliders.as_struct()
        self._keep_aabb = (params, colliders, c)
        self._check(self.lib.avn_update_aabbs(self.handle, C.byref(params), C.byref(c)))

    def timings(self) -> dict:
        t = AvnTimings()
        self._check(self.lib.avn_get_timings(self.handle, C.byref(t)))
        return {n: getattr(t, n) for n, _ in AvnTimings._fields_ if n != "_pad"}
