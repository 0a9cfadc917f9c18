#!/usr/bin/env python
"""Solver-stage device time of an alternative build of the library (GPU box).
usage: python scripts/lib_timing.py path/to/lib.so [scene] [reps]"""
import ctypes as C
import hashlib
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from avian_b200 import api  # noqa: E402

lib = C.CDLL(str(Path(sys.argv[1]).resolve()))
api.bind_abi(lib)
api._lib = lib
import bench  # noqa: E402

scene = sys.argv[2] if len(sys.argv) > 2 else "stack100k"
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 10
with api.Context(device=0, scalar=np.float64 if scene.startswith("spheres") else np.float32) as ctx:
    sc, prm, bodies, man, aabbs, joints = bench.build_snapshot(scene, bench.SCENES[scene][2], ctx)
    b, m = bodies.copy(), man.copy()
    ctx.solver_upload(prm, b, m, joints)
    for _ in range(3):
        ctx.solver_run()
    ms = []
    for _ in range(reps):
        ctx.solver_run(); ctx.solver_download()
        ms.append(ctx.timings()["total_ms"])
    digest = hashlib.sha1(np.ascontiguousarray(b.position).tobytes()).hexdigest()[:12]
    print(f"{sys.argv[1]} {scene}: solver stage {np.median(ms):.3f} ms (min {min(ms):.3f}) positions sha1 {digest}", flush=True)
