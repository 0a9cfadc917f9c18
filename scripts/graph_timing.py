#!/usr/bin/env python
"""Where the time of one step of the device-resident pipeline goes (GPU box): wall time of every C-ABI call of DeviceGraphWorld.step_from,
the calls that wait for the device show the time of what was queued before them.  usage: python scripts/graph_timing.py [scene] [steps]"""
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402
from avian_b200 import api  # noqa: E402

scene = sys.argv[1] if len(sys.argv) > 1 else "stack100k"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 20


class A:
    pass


args = A()
args.scene, args.settle, args.solver_iterations, args.warmup, args.steps = scene, bench.SCENES[scene][2], 1, 3, steps
with api.Context(device=0, scalar=np.float64 if scene.startswith("spheres") else np.float32) as ctx:
    w, aabbs, mn, mx, first = bench._resident_world(args, ctx)
    b = w.bodies
    acc = {}

    def timed(name, fn):
        t0 = time.perf_counter()
        r = fn()
        acc[name] = acc.get(name, 0.0) + (time.perf_counter() - t0)
        return r

    colliders = {"shape": w._shape, "dims": w._dims, "position": b.position, "rotation": b.rotation, "aabb_min": mn, "aabb_max": mx}
    stats = []
    t_all = time.perf_counter()
    for _ in range(steps):
        timed("broadphase_upload", lambda: ctx.broadphase_upload(aabbs))
        timed("broadphase_run", lambda: ctx.broadphase_run())
        timed("solver_prefetch_bodies", lambda: ctx.solver_prefetch_bodies(b, static_unchanged=True))
        st = timed("contacts_step", lambda: ctx.contacts_step(w.params.dt, 0.005, colliders, b.linear_velocity, b.angular_velocity, True, take_pairs=True,
                                                             shapes_unchanged=True))
        timed("broadphase_download_order", lambda: ctx.broadphase_download_order())
        timed("host: new order", lambda: np.ascontiguousarray(aabbs.collider[aabbs.order_out]))
        timed("solver_step_resident", lambda: ctx.solver_step_resident(w.params, b, w.joints))
        stats.append(st)
    total = time.perf_counter() - t_all
    t = ctx.timings()
    print(f"scene {scene}: {steps} steps, {total / steps * 1e3:.3f} ms per step (calls serialised)")
    for k, v in acc.items():
        print(f"  {k:28s} {v / steps * 1e3:8.3f} ms")
    print("  last solver call on the device:", {k: (round(t[k], 3) if isinstance(t[k], float) else t[k]) for k in
                                                ("h2d_ms", "prepare_ms", "substep_loop_ms", "finalize_ms", "total_ms", "d2h_ms", "launch_mode", "kernel_launches")})
    # the same constraints as CSR columns from the host (the ordinary upload path), for comparison of the solver stage alone
    g = ctx.contacts_download_graph(st["rows_high_water"], st["manifold_count"])
    print("  colour sizes:", np.diff(st["color_offsets"]).tolist())
    print("  changes per step:", np.mean([s["started_touching"] + s["stopped_touching"] for s in stats]), " colouring rounds:",
          [s["colouring_rounds"] for s in stats][:10], " first frame:", first)
