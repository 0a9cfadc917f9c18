#!/usr/bin/env python
"""Generates tests/golden/*.npz — small input/output vectors of the hot path.

Provenance: the reference (Rust) cannot run in this environment, so these vectors are produced by the C++ ORACLE
(oracle/, the restatement of the reference path) and are therefore regression pins of the oracle + cross-checks for the CUDA
path, NOT outputs of Avian itself.  Re-run:  python tests/golden/make_golden.py
Each file holds the flattened inputs of one call and the outputs the oracle produced for it.
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
from avian_b200 import api, scenes  # noqa: E402
import oracle_lib  # noqa: E402
from helpers import advance_to_solver_input  # noqa: E402

OUT = Path(__file__).resolve().parent


def pack(prefix, obj):
    return {f"{prefix}.{k}": v for k, v in obj.__dict__.items() if isinstance(v, np.ndarray)}


def params_array(p):
    return np.array([p.dt, p.h, p.substeps, p.restitution_iterations, *p.gravity, p.contact_damping_ratio, p.contact_frequency_factor,
                     p.max_overlap_solve_speed, p.warm_start_coefficient, p.restitution_threshold, p.length_unit, p.match_contacts, p.solver_iterations])


def solver_case(name, scene, steps, substeps):
    _, (prm, b, m, j) = advance_to_solver_input(scene, steps=steps, substeps=substeps)
    bo, mo = b.copy(), m.copy()
    jo = None if j is None else j.copy()
    oracle_lib.solver_step(prm, bo, mo, jo)
    data = {"params": params_array(prm), **pack("in.bodies", b), **pack("in.manifolds", m), **pack("out.bodies", bo), **pack("out.manifolds", mo)}
    if j is not None:
        for t, jt in j.types.items():
            data.update(pack(f"in.joints{t}", jt))
            data.update(pack(f"out.joints{t}", jo.types[t]))
    np.savez_compressed(OUT / f"{name}.npz", **data)
    print(name, b.count, "bodies", m.count, "manifolds")


def broadphase_case(name, n, seed):
    sys.path.insert(0, str(ROOT / "tests"))
    import test_gpu_broadphase as T
    a = T.random_aabbs(n, seed)
    o = oracle_lib.broadphase(a)
    np.savez_compressed(OUT / f"{name}.npz", **pack("in.aabbs", a), **pack("out.pairs", o), count=np.array([o.count]))
    print(name, n, "aabbs", o.count, "pairs")


if __name__ == "__main__":
    solver_case("solver_cubes3_step45", scenes.cubes_example(3), 45, 1)                 # BASELINE config 1, cubes landing
    solver_case("solver_brick4_step3", scenes.cube_stack(4, 4, 4, brick=True), 3, 8)   # coupled stack, 8 substeps
    solver_case("solver_ragdolls4_step20", scenes.ragdoll_field(4, pitch=3.0, drop_height=0.1), 20, 8)  # joints + contacts
    broadphase_case("broadphase_random_1500", 1500, 9)
