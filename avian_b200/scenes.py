"""Synthetic scenes of BASELINE.json's configs as body columns (+ collider shapes for the host fixture).

All generators are deterministic (no RNG unless a seed is an argument).  Bodies use the reference's defaults:
density 1 (ColliderDensity), friction 0.5 / restitution 0 (physics_material.rs:152-160,320-327), gravity -9.81 y.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from . import api
from .fixture import SHAPE_CUBOID, SHAPE_SPHERE


@dataclass
class Scene:
    name: str
    bodies: api.Bodies
    shape_type: np.ndarray   # int32[n]
    dims: np.ndarray         # float64[n,3] half extents / radius
    friction: np.ndarray     # float64[n]
    restitution: np.ndarray  # float64[n]
    joints: api.JointSet | None = None
    joint_disabled_body_pairs: np.ndarray | None = None


def _cuboid_mass(he: np.ndarray, density: float = 1.0):
    """mass and local inverse inertia (diagonal) of solid cuboids given half extents [n,3]."""
    size = 2.0 * he
    m = density * size[:, 0] * size[:, 1] * size[:, 2]
    ix = m / 12.0 * (size[:, 1] ** 2 + size[:, 2] ** 2)
    iy = m / 12.0 * (size[:, 0] ** 2 + size[:, 2] ** 2)
    iz = m / 12.0 * (size[:, 0] ** 2 + size[:, 1] ** 2)
    return m, np.stack([ix, iy, iz], axis=1)


def _assemble(name, pos, rot, kind, he, shape_type, scalar, friction=0.5, restitution=0.0, linvel=None, angvel=None, density=1.0, **extra) -> Scene:
    n = pos.shape[0]
    s = np.dtype(scalar)
    he = np.asarray(he, dtype=np.float64)
    m, inertia = _cuboid_mass(he, density)
    sph = shape_type == SHAPE_SPHERE
    if sph.any():
        r = he[sph, 0]
        ms = density * 4.0 / 3.0 * np.pi * r ** 3
        m[sph] = ms
        inertia[sph] = (0.4 * ms * r * r)[:, None]
    dyn = kind == api.BODY_DYNAMIC
    inv_m = np.where(dyn, 1.0 / m, 0.0)
    inv_i = np.zeros((n, 6))
    inv_i[:, 0] = np.where(dyn, 1.0 / inertia[:, 0], 0.0)
    inv_i[:, 3] = np.where(dyn, 1.0 / inertia[:, 1], 0.0)
    inv_i[:, 5] = np.where(dyn, 1.0 / inertia[:, 2], 0.0)
    z3 = np.zeros((n, 3))
    bodies = api.Bodies(
        kind=np.ascontiguousarray(kind, dtype=np.uint8), position=np.ascontiguousarray(pos, dtype=s), rotation=np.ascontiguousarray(rot, dtype=s),
        linear_velocity=np.ascontiguousarray(z3 if linvel is None else linvel, dtype=s),
        angular_velocity=np.ascontiguousarray(z3 if angvel is None else angvel, dtype=s),
        inverse_mass=np.ascontiguousarray(inv_m, dtype=s), inverse_inertia_local=np.ascontiguousarray(inv_i, dtype=s),
        center_of_mass=np.zeros((n, 3), dtype=s))
    fr = np.full(n, friction, dtype=np.float64) if np.isscalar(friction) else np.asarray(friction, dtype=np.float64)
    rs = np.full(n, restitution, dtype=np.float64) if np.isscalar(restitution) else np.asarray(restitution, dtype=np.float64)
    return Scene(name, bodies, np.ascontiguousarray(shape_type, dtype=np.int32), np.ascontiguousarray(he), fr, rs, **extra)


def cube_stack(nx: int, ny: int, nz: int, size: float = 1.0, gap: float = 0.05, overlap: float = 0.01, brick: bool = True,
               scalar=np.float32, ground_half=(0.0, 0.5, 0.0), restitution: float = 0.0) -> Scene:
    """A box stack of nx*ny*nz dynamic cubes on a static ground slab (body 0).

    brick=True  — odd layers are shifted by half a pitch in x and z and hold (nx-1)*(nz-1) cubes, so every cube of an
                  odd layer rests on four cubes and every inner cube of an even layer on four: one coupled pile (one
                  island), ~4 manifolds per cube (the headline scene; 51x40x50 gives exactly 100 000 cubes).
    brick=False — aligned columns with a lateral gap, like crates/avian3d/examples/cubes.rs:42-52 (spacing 2.05 for
                  size-2 cubes): independent columns.
    Layers start `overlap` into each other like benches/src/dim3/large_pyramid.rs (y spacing 0.99 for unit cubes),
    so every contact exists on the first step.  Spawn order is x-major so the broad phase's persistent order starts
    sorted along x (the reference's insertion sort stays linear)."""
    pitch = size + gap
    xs = np.arange(nx) * pitch
    zs = np.arange(nz) * pitch
    layers = []
    for k in range(ny):
        y = size * 0.5 + k * (size - overlap) - overlap
        if brick and (k % 2):
            lx, lz = xs[:-1] + pitch * 0.5, zs[:-1] + pitch * 0.5   # (nx-1) x (nz-1) cubes, each over four cubes below
        else:
            lx, lz = xs, zs
        X, Z = np.meshgrid(lx, lz, indexing="ij")
        layers.append(np.stack([X.ravel(), np.full(X.size, y), Z.ravel()], axis=1))
    pos = np.concatenate(layers)
    pos = pos[np.lexsort((pos[:, 2], pos[:, 1], pos[:, 0]))]   # x-major spawn order
    n = pos.shape[0]
    # ground slab: top face at y = 0
    gx = ground_half[0] or (nx * pitch + 10.0)
    gz = ground_half[2] or (nz * pitch + 10.0)
    gpos = np.array([[xs.mean() if nx else 0.0, -ground_half[1], zs.mean() if nz else 0.0]])
    pos = np.concatenate([gpos, pos])
    he = np.concatenate([[[gx, ground_half[1], gz]], np.full((n, 3), size * 0.5)])
    kind = np.concatenate([[api.BODY_STATIC], np.full(n, api.BODY_DYNAMIC)])
    rot = np.tile(np.array([0.0, 0.0, 0.0, 1.0]), (n + 1, 1))
    st = np.full(n + 1, SHAPE_CUBOID)
    return _assemble(f"cube_stack_{nx}x{ny}x{nz}{'_brick' if brick else ''}", pos, rot, kind, he, st, scalar, restitution=restitution)


def cubes_example(n_side: int = 4, scalar=np.float32) -> Scene:
    """crates/avian3d/examples/cubes.rs:25-52: ground = unit cuboid scaled (100,1,100) at y=-2; n_side^3 cubes of
    side 2.0 at spacing 2.05, y offset +... (the example uses x,z in -2..2 and y in -2..2 plus 10).  n_side=3 is
    BASELINE config 1's reduced scene, n_side=4 the literal example."""
    lo = -(n_side // 2)
    idx = np.arange(lo, lo + n_side)
    X, Y, Z = np.meshgrid(idx, idx, idx, indexing="ij")
    cube_size, spacing = 2.0, 2.05
    pos = np.stack([X.ravel() * spacing, (Y.ravel() + 2) * spacing + 0.025, Z.ravel() * spacing], axis=1).astype(np.float64)
    n = pos.shape[0]
    pos = np.concatenate([[[0.0, -2.0, 0.0]], pos])
    he = np.concatenate([[[50.0, 0.5, 50.0]], np.full((n, 3), cube_size * 0.5)])
    kind = np.concatenate([[api.BODY_STATIC], np.full(n, api.BODY_DYNAMIC)])
    rot = np.tile(np.array([0.0, 0.0, 0.0, 1.0]), (n + 1, 1))
    return _assemble(f"cubes_{n_side}x{n_side}x{n_side}", pos, rot, kind, he, np.full(n + 1, SHAPE_CUBOID), scalar)


def falling_spheres(n: int, seed: int = 42, box=(200.0, 50.0, 200.0), radius: float = 0.5, scalar=np.float64) -> Scene:
    """BASELINE config 5: n spheres r=0.5 uniformly random in a box above a ground slab, zero velocity."""
    rng = np.random.default_rng(seed)
    pos = rng.uniform([0, radius + 0.01, 0], [box[0], box[1], box[2]], size=(n, 3))
    order = np.argsort(pos[:, 0], kind="stable")
    pos = pos[order]
    pos = np.concatenate([[[box[0] * 0.5, -0.5, box[2] * 0.5]], pos])
    he = np.concatenate([[[box[0] * 0.5 + 10, 0.5, box[2] * 0.5 + 10]], np.full((n, 3), radius)])
    kind = np.concatenate([[api.BODY_STATIC], np.full(n, api.BODY_DYNAMIC)])
    rot = np.tile(np.array([0.0, 0.0, 0.0, 1.0]), (n + 1, 1))
    st = np.concatenate([[SHAPE_CUBOID], np.full(n, SHAPE_SPHERE)])
    return _assemble(f"spheres_{n}", pos, rot, kind, he, st, scalar)


def _quat_axis_angle(axis, angle):
    axis = np.asarray(axis, dtype=np.float64)
    axis = axis / np.linalg.norm(axis)
    s = np.sin(angle * 0.5)
    return np.array([axis[0] * s, axis[1] * s, axis[2] * s, np.cos(angle * 0.5)])


# one ragdoll: (name, half extents, centre offset, parent, joint type, anchor in parent frame (world offset from parent centre))
_RAGDOLL = [
    ("pelvis", (0.15, 0.10, 0.10), (0.0, 1.00, 0.0), -1, None),
    ("spine", (0.15, 0.12, 0.10), (0.0, 1.24, 0.0), 0, "S"),
    ("chest", (0.17, 0.13, 0.11), (0.0, 1.51, 0.0), 1, "S"),
    ("neck", (0.05, 0.05, 0.05), (0.0, 1.70, 0.0), 2, "S"),
    ("head", (0.10, 0.11, 0.10), (0.0, 1.87, 0.0), 3, "S"),
    ("l_upper_arm", (0.14, 0.05, 0.05), (-0.33, 1.58, 0.0), 2, "S"),
    ("l_fore_arm", (0.13, 0.04, 0.04), (-0.61, 1.58, 0.0), 5, "R"),
    ("l_hand", (0.05, 0.03, 0.04), (-0.80, 1.58, 0.0), 6, "S"),
    ("r_upper_arm", (0.14, 0.05, 0.05), (0.33, 1.58, 0.0), 2, "S"),
    ("r_fore_arm", (0.13, 0.04, 0.04), (0.61, 1.58, 0.0), 8, "R"),
    ("r_hand", (0.05, 0.03, 0.04), (0.80, 1.58, 0.0), 9, "S"),
    ("l_thigh", (0.07, 0.20, 0.07), (-0.09, 0.69, 0.0), 0, "S"),
    ("l_shin", (0.06, 0.19, 0.06), (-0.09, 0.29, 0.0), 11, "R"),
    ("l_foot", (0.06, 0.04, 0.11), (-0.09, 0.05, 0.04), 12, "S"),
    ("r_thigh", (0.07, 0.20, 0.07), (0.09, 0.69, 0.0), 0, "S"),
    ("r_shin", (0.06, 0.19, 0.06), (0.09, 0.29, 0.0), 14, "R"),
    ("r_foot", (0.06, 0.04, 0.11), (0.09, 0.05, 0.04), 15, "S"),
]


def ragdoll_field(count: int, pitch: float = 8.0, drop_height: float = 2.0, seed: int = 1234, scalar=np.float32) -> Scene:
    """BASELINE config 4: `count` ragdolls on a square grid, 17 cuboid bodies and 16 joints each —
    spherical joints (swing +-60 deg, twist +-30 deg) for neck/spine/shoulders/hips/wrists/ankles, revolute joints
    (0..120 deg) for elbows and knees; every joint disables collision between its bodies (JointCollisionDisabled).
    A small deterministic pose jitter (PCG64, seed) breaks symmetry."""
    rng = np.random.default_rng(seed)
    side = int(np.ceil(np.sqrt(count)))
    nb = len(_RAGDOLL)
    pos, he, rot = [], [], []
    sj = {k: [] for k in ("b1", "b2", "a1", "a2")}
    rj = {k: [] for k in ("b1", "b2", "a1", "a2")}
    disabled = []
    for r in range(count):
        gx, gz = (r % side) * pitch, (r // side) * pitch
        base = 1 + r * nb
        yaw = _quat_axis_angle((0, 1, 0), rng.uniform(-np.pi, np.pi))
        tilt = _quat_axis_angle((rng.uniform(-1, 1), 0, rng.uniform(-1, 1) + 1e-3), np.deg2rad(rng.uniform(0, 5)))
        q = _qmul(tilt, yaw)
        for i, (_, h, c, parent, jt) in enumerate(_RAGDOLL):
            cw = _qrot(q, np.array(c) - np.array([0, 1.0, 0])) + np.array([gx, 1.0 + drop_height, gz])
            pos.append(cw)
            he.append(h)
            rot.append(q)
            if parent >= 0:
                pc = np.array(_RAGDOLL[parent][2])
                cc = np.array(c)
                # anchor: the point between parent and child along the segment, at the child's near face
                anchor_local_world = (pc + cc) * 0.5
                a1 = anchor_local_world - pc   # in the (common) body frame since all bodies share rotation q
                a2 = anchor_local_world - cc
                d = sj if jt == "S" else rj
                d["b1"].append(base + parent)
                d["b2"].append(base + i)
                d["a1"].append(a1)
                d["a2"].append(a2)
                disabled.append((base + parent, base + i))
    n = len(pos)
    pos = np.concatenate([[[side * pitch * 0.5, -0.5, side * pitch * 0.5]], np.array(pos)])
    he = np.concatenate([[[side * pitch * 0.5 + 20, 0.5, side * pitch * 0.5 + 20]], np.array(he)])
    rot = np.concatenate([[[0, 0, 0, 1.0]], np.array(rot)])
    kind = np.concatenate([[api.BODY_STATIC], np.full(n, api.BODY_DYNAMIC)])
    s = np.dtype(scalar)

    def joints(d, revolute):
        m = len(d["b1"])
        j = api.Joints(body1=np.array(d["b1"], dtype=np.int32), body2=np.array(d["b2"], dtype=np.int32),
                       local_anchor1=np.ascontiguousarray(d["a1"], dtype=s).reshape(m, 3), local_anchor2=np.ascontiguousarray(d["a2"], dtype=s).reshape(m, 3))
        j.limit_enabled = np.full(m, 1 if revolute else 3, dtype=np.uint8)
        if revolute:
            j.axis = np.tile(np.array([1.0, 0, 0], dtype=s), (m, 1))
            j.limit_min = np.zeros(m, dtype=s)
            j.limit_max = np.full(m, np.deg2rad(120.0), dtype=s)
        else:
            j.limit_min = np.full(m, -np.deg2rad(60.0), dtype=s)
            j.limit_max = np.full(m, np.deg2rad(60.0), dtype=s)
            j.limit2_min = np.full(m, -np.deg2rad(30.0), dtype=s)
            j.limit2_max = np.full(m, np.deg2rad(30.0), dtype=s)
        j.force = np.zeros((m, 3), dtype=s)
        j.torque = np.zeros((m, 3), dtype=s)
        return j

    js = api.JointSet({api.JOINT_REVOLUTE: joints(rj, True), api.JOINT_SPHERICAL: joints(sj, False)})
    dis = np.array([(min(a, b) << 32) | max(a, b) for a, b in disabled], dtype=np.uint64)
    return _assemble(f"ragdolls_{count}", pos, rot, kind, he, np.full(n + 1, SHAPE_CUBOID), scalar, joints=js, joint_disabled_body_pairs=dis)


def _qmul(a, b):
    ax, ay, az, aw = a
    bx, by, bz, bw = b
    return np.array([aw * bx + ax * bw + ay * bz - az * by, aw * by - ax * bz + ay * bw + az * bx,
                     aw * bz + ax * by - ay * bx + az * bw, aw * bw - ax * bx - ay * by - az * bz])


def _qrot(q, v):
    b = q[:3]
    return v * (q[3] * q[3] - b @ b) + b * (2.0 * (v @ b)) + np.cross(b, v) * (2.0 * q[3])


def spherical_chain(links: int = 100, scalar=np.float32) -> Scene:
    """crates/avian3d/examples/chain_3d.rs:36-72: a kinematic anchor and `links` small dynamic spheres joined by
    spherical joints (local_anchor2 = +y * 1.1 * radius*2), collision between neighbours disabled."""
    radius = 0.03 * 2.0  # particle_radius used by the example scaled to metres here
    pos = [np.array([0.0, 0.0, 0.0])]
    for i in range(links):
        pos.append(np.array([0.0, -(i + 1) * (radius * 2.0 + 0.02), 0.0]))
    n = len(pos)
    pos = np.array(pos)
    he = np.full((n, 3), radius)
    kind = np.concatenate([[api.BODY_KINEMATIC], np.full(links, api.BODY_DYNAMIC)])
    rot = np.tile(np.array([0.0, 0.0, 0.0, 1.0]), (n, 1))
    s = np.dtype(scalar)
    b1 = np.arange(0, links, dtype=np.int32)
    b2 = np.arange(1, links + 1, dtype=np.int32)
    j = api.Joints(body1=b1, body2=b2, local_anchor1=np.zeros((links, 3), dtype=s),
                   local_anchor2=np.tile(np.array([0.0, radius * 2.0 + 0.02, 0.0], dtype=s), (links, 1)))
    j.compliance0 = np.full(links, 1e-5, dtype=s)
    j.force = np.zeros((links, 3), dtype=s)
    j.torque = np.zeros((links, 3), dtype=s)
    dis = np.array([(int(a) << 32) | int(b) for a, b in zip(b1, b2)], dtype=np.uint64)
    sc = _assemble(f"chain_{links}", pos, rot, kind, he, np.full(n, SHAPE_SPHERE), scalar, joints=api.JointSet({api.JOINT_SPHERICAL: j}),
                   joint_disabled_body_pairs=dis)
    # a kinematic body keeps its collider mass (SolverBodyInertia::new keeps inv_mass, dominance 128)
    return sc
