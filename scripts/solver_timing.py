#!/usr/bin/env python
"""Quick device-time comparison of solver-stage variants on one frozen snapshot (GPU box).
usage: python scripts/solver_timing.py [scene] [reps] [NAME=VAL,NAME=VAL ...]   -> one line per variant (default list below,
or one variant per extra argument: a comma-separated list of environment overrides, "-" = none)"""
import os
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402
from avian_b200 import api  # noqa: E402

scene = sys.argv[1] if len(sys.argv) > 1 else "stack100k"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
variants = [("wave bps3", {}), ("wave no-l2persist", {"AVN_L2_PERSIST": "0"}), ("wave bps4", {"AVN_MEGA_BPS": "4"}),
            ("barrier bps3", {"AVN_LAUNCH_MODE": "barrier"}), ("phases", {"AVN_LAUNCH_MODE": "phases"})]
if len(sys.argv) > 3:
    variants = [(spec, dict(kv.split("=", 1) for kv in spec.split(",") if "=" in kv)) for spec in sys.argv[3:]]
with api.Context(device=0, scalar=np.float64 if scene.startswith("spheres") else np.float32) as ctx0:
    sc, prm, bodies, man, aabbs, joints = bench.build_snapshot(scene, bench.SCENES[scene][2], ctx0)
ref = None
for name, env in variants:
    os.environ.update(env)
    try:
        with api.Context(device=0, scalar=ctx0.scalar) as ctx:
            b, m = bodies.copy(), man.copy()
            ctx.solver_upload(prm, b, m, joints)
            for _ in range(3):
                ctx.solver_run()
            ms = []
            for _ in range(reps):
                ctx.solver_run(); ctx.solver_download()
                ms.append(ctx.timings()["total_ms"])
            same = "" if ref is None else f" bit-identical-to-first={np.array_equal(ref, b.position)}"
            if ref is None:
                ref = b.position.copy()
            print(f"{name:14s} solver stage {np.median(ms):8.3f} ms (min {min(ms):.3f}){same}", flush=True)
    finally:
        for k in env:
            os.environ.pop(k, None)
