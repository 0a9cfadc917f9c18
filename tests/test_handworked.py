"""Hand-worked known-answer vectors (tests/golden/handworked/): an INDEPENDENT pin of the oracle and of the CUDA path.

`vectors.json` is produced by `tests/golden/handworked/worked.py` — a plain numpy restatement of the path written from the Rust source, which
shares no code with oracle/ or the CUDA library (provenance: `derive.py`).  Both must reproduce it:
  * CPU (`-m "not gpu"`): the C++ oracle against the vectors;
  * GPU (`-m gpu`): the CUDA library, through the C ABI, against the same vectors — no oracle involved.
Bar: 1e-5 relative element-wise (floor 1 unit), BASELINE.json's bar.  (Not bit-exact: glam's SSE2 quaternion product may associate
differently from the scalar formulas of worked.py in the last bit.)  The broad-phase case is worked by hand and compared bit for bit.
"""
import json
from pathlib import Path

import numpy as np
import pytest

from avian_b200 import api

import oracle_lib
from helpers import rel_err

HERE = Path(__file__).resolve().parent / "golden" / "handworked"
VECTORS = json.loads((HERE / "vectors.json").read_text())
SCENES = sorted(VECTORS["scenes"])
KAT_RTOL = 1e-5


def _build(sc, dtype):
    s = np.dtype(dtype)
    p = sc["params"]
    prm = api.AvnStepParams()
    prm.dt, prm.h, prm.substeps = p["dt"], p["h"], p["substeps"]
    prm.restitution_iterations = p["restitution_iterations"]
    prm.gravity[0], prm.gravity[1], prm.gravity[2] = p["gravity"]
    for k in ("contact_damping_ratio", "contact_frequency_factor", "max_overlap_solve_speed", "warm_start_coefficient", "restitution_threshold", "length_unit"):
        setattr(prm, k, p[k])
    prm.match_contacts, prm.solver_iterations = p["match_contacts"], 1
    bs = sc["bodies"]
    col = lambda key, default, width: np.array([b.get(key, default) for b in bs], dtype=s).reshape((len(bs),) + ((width,) if width > 1 else ()))
    bodies = api.Bodies(kind=np.array([b["kind"] for b in bs], dtype=np.uint8), position=col("position", None, 3), rotation=col("rotation", None, 4),
                        linear_velocity=col("linear_velocity", None, 3), angular_velocity=col("angular_velocity", None, 3),
                        inverse_mass=col("inverse_mass", None, 1), inverse_inertia_local=col("inverse_inertia_local", None, 6),
                        center_of_mass=col("center_of_mass", [0, 0, 0], 3), dominance=np.array([b.get("dominance", 0) for b in bs], dtype=np.int8),
                        linear_damping=col("linear_damping", 0.0, 1), angular_damping=col("angular_damping", 0.0, 1), gravity_scale=col("gravity_scale", 1.0, 1))
    if any(b.get("locked_axes") for b in bs):
        bodies.locked_axes = np.array([b.get("locked_axes", 0) for b in bs], dtype=np.uint8)
    man = None
    if sc.get("manifolds"):
        ms = sc["manifolds"]
        pts = [pt for m in ms for pt in m["points"]]
        po = np.cumsum([0] + [len(m["points"]) for m in ms]).astype(np.uint32)
        man = api.Manifolds(
            color_offsets=np.array(sc["color_offsets"], dtype=np.uint32), body1=np.array([m["body1"] for m in ms], dtype=np.int32),
            body2=np.array([m["body2"] for m in ms], dtype=np.int32), normal=np.array([m["normal"] for m in ms], dtype=s),
            friction=np.array([m["friction"] for m in ms], dtype=s), restitution=np.array([m["restitution"] for m in ms], dtype=s), point_offsets=po,
            anchor1=np.array([q["anchor1"] for q in pts], dtype=s), anchor2=np.array([q["anchor2"] for q in pts], dtype=s),
            penetration=np.array([q["penetration"] for q in pts], dtype=s), normal_speed=np.array([q["normal_speed"] for q in pts], dtype=s),
            warm_start_normal_impulse=np.array([q["warm_start_normal_impulse"] for q in pts], dtype=s),
            warm_start_tangent_impulse=np.array([q["warm_start_tangent_impulse"] for q in pts], dtype=s).reshape(-1, 2),
            normal_impulse=np.zeros(len(pts), dtype=s))
    joints, where = None, []
    if sc.get("joints"):
        joints = api.JointSet()
        for t in range(api.JOINT_TYPE_COUNT):
            js = [(k, j) for k, j in enumerate(sc["joints"]) if j["type"] == t]
            if not js:
                continue
            n = len(js)
            g = lambda key, default, width: np.array([j.get(key, default) for _, j in js], dtype=s).reshape((n,) + ((width,) if width > 1 else ()))
            def lim(key):   # (enabled, min, max) columns of an optional limit
                en = np.array([j.get(key) is not None for _, j in js])
                lo = np.array([(j.get(key) or [0.0, 0.0])[0] for _, j in js], dtype=s)
                hi = np.array([(j.get(key) or [0.0, 0.0])[1] for _, j in js], dtype=s)
                return en, lo, hi
            if t == api.JOINT_DISTANCE:
                en1, lo1, hi1 = np.zeros(n, dtype=bool), g("limit_min", 0.0, 1), g("limit_max", 0.0, 1)
            else:
                en1, lo1, hi1 = lim({api.JOINT_REVOLUTE: "angle_limit", api.JOINT_PRISMATIC: "limits"}.get(t, "swing_limit"))
            en2, lo2, hi2 = lim("twist_limit")
            damp = np.array([j.get("damping") is not None for _, j in js])
            joints.types[t] = api.Joints(
                body1=np.array([j["body1"] for _, j in js], dtype=np.int32), body2=np.array([j["body2"] for _, j in js], dtype=np.int32),
                local_anchor1=g("local_anchor1", None, 3), local_anchor2=g("local_anchor2", None, 3), local_basis1=g("local_basis1", [0, 0, 0, 1], 4),
                local_basis2=g("local_basis2", [0, 0, 0, 1], 4), axis=g("axis", [[0, 0, 1], [0, 0, 1], [0, 1, 0], [1, 0, 0], [0, 0, 1]][t], 3),
                limit_enabled=(en1.astype(np.uint8) | (en2.astype(np.uint8) << 1)), limit_min=lo1, limit_max=hi1, limit2_min=lo2, limit2_max=hi2,
                compliance0=g("compliance0", 0.0, 1), compliance1=g("compliance1", 0.0, 1), compliance2=g("compliance2", 0.0, 1),
                damping_enabled=damp.astype(np.uint8), damping_linear=np.array([(j.get("damping") or [0.0, 0.0])[0] for _, j in js], dtype=s),
                damping_angular=np.array([(j.get("damping") or [0.0, 0.0])[1] for _, j in js], dtype=s),
                force=np.zeros((n, 3), dtype=s), torque=np.zeros((n, 3), dtype=s))
            where += [(k, t, i) for i, (k, _) in enumerate(js)]
    return prm, bodies, man, joints, where


def _check(name, dtype, run):
    sc = VECTORS["scenes"][name]
    want = sc["expected_f32" if np.dtype(dtype) == np.float32 else "expected_f64"]
    prm, b, m, j, where = _build(sc["input"], dtype)
    run(prm, b, m, j)
    tol = KAT_RTOL
    for key in ("position", "rotation", "linear_velocity"):
        e = rel_err(getattr(b, key), np.array(want[key]))
        assert e <= tol, f"{name} {key}: {e:.3e}"
    for i, av in enumerate(want["angular_velocity"]):
        if av is not None:
            e = rel_err(b.angular_velocity[i], np.array(av))
            assert e <= tol, f"{name} angular_velocity[{i}]: {e:.3e}"
    if m is not None:
        for key in ("warm_start_normal_impulse", "warm_start_tangent_impulse", "normal_impulse"):
            e = rel_err(getattr(m, key), np.array(want[key]))
            assert e <= tol, f"{name} {key}: {e:.3e}"
    for k, t, i in where:
        ef = rel_err(j.types[t].force[i], np.array(want["joint_force"][k]))
        et = rel_err(j.types[t].torque[i], np.array(want["joint_torque"][k]))
        # forces are O(10..100) N: relative to the value (rel_err's floor is 1)
        assert ef <= 10 * tol and et <= 10 * tol, f"{name} joint {k}: force {ef:.3e} torque {et:.3e}"


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("name", SCENES)
def test_oracle_reproduces_handworked(name, dtype):
    _check(name, dtype, lambda prm, b, m, j: oracle_lib.solver_step(prm, b, m, j))


def _sap_columns(dtype):
    iv = VECTORS["sap_six"]["intervals"]
    return api.Aabbs(collider=np.array([x["collider"] for x in iv], dtype=np.uint32), body=np.array([x["body"] for x in iv], dtype=np.uint32),
                     aabb_min=np.array([x["min"] for x in iv], dtype=dtype), aabb_max=np.array([x["max"] for x in iv], dtype=dtype),
                     flags=np.array([(api.AABB_IS_INACTIVE if x["inactive"] else 0) | api.AABB_GENERATE_CONSTRAINTS for x in iv], dtype=np.uint8),
                     memberships=np.array([x["memberships"] for x in iv], dtype=np.uint32), filters=np.array([x["filters"] for x in iv], dtype=np.uint32),
                     order_out=np.zeros(len(iv), dtype=np.uint32))


def _check_sap(pairs, aabbs):
    want = VECTORS["sap_six"]
    got = [[int(a), int(b)] for a, b in zip(pairs.collider1[:pairs.count], pairs.collider2[:pairs.count])]
    assert got == want["expected_pairs"], got
    assert [int(x) for x in aabbs.order_out] == want["expected_order"]
    assert np.array_equal(aabbs.aabb_min[:, 0] == 0, [False, True, True, False, False, False]) and np.signbit(aabbs.aabb_min[1, 0])   # the -0.0 survived the trip


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_oracle_sap_six_by_hand(dtype):
    a = _sap_columns(dtype)
    _check_sap(oracle_lib.broadphase(a), a)


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("name", SCENES)
def test_gpu_reproduces_handworked(name, dtype):
    with api.Context(device=0, scalar=dtype) as ctx:
        _check(name, dtype, lambda prm, b, m, j: ctx.solver_step(prm, b, m, j))


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_gpu_sap_six_by_hand(dtype):
    with api.Context(device=0, scalar=dtype) as ctx:
        a = _sap_columns(dtype)
        _check_sap(ctx.broadphase(a), a)


def test_handworked_module_is_independent():
    """worked.py / derive.py must not import anything of this repository (the point of the exercise)."""
    for f in ("worked.py", "derive.py"):
        src = (HERE / f).read_text()
        for banned in ("avian_b200", "oracle_lib", "oracle/", "ctypes"):
            assert banned not in src.replace("NOT from oracle/", "").replace("oracle/ or the CUDA", "").replace("NOT by oracle/", "") or banned == "oracle/", f"{f} mentions {banned}"
        assert "import avian_b200" not in src and "from avian_b200" not in src and "import oracle_lib" not in src
