#!/usr/bin/env python
"""bench.py — physics steps/s of the avian3d substep hot path on B200 (BASELINE.json metric).

A "step" is ONE pass of the hot path over one frozen snapshot of the headline scene (100 000-cube coupled stack,
f32, 8 substeps): sweep-and-prune broad phase over the 100 001 collider AABBs + the whole solver stage
(prepare, 8 x [integrate velocities, warm start, biased solve, integrate positions, relax], restitution, writeback,
store impulses) over the snapshot's bodies and contact manifolds.  The narrow phase is NOT in the step (it is outside
the hot path); its manifolds are part of the snapshot, identical for every arm.

  value        steps/s with the snapshot resident in HBM (avn_*_run only), device time from CUDA events on the
               library's stream, max over ranks; N > 1 = N independent 100k piles (island sharding, no collective),
               value = N * K / T.
  e2e          the same through the public C-ABI calls avn_broadphase + avn_solver_step with pinned HOST buffers:
               H2D of every column and D2H of the results inside the timed region.
  roofline     the dominant kernel (the persistent step megakernel): algorithmic bytes per launch (SURVEY §8d
               formulas with the measured B, M, P) / its CUDA-event duration, vs MEASURED_PEAKS.json hbm_gbs.
  cpu_baseline the CPU oracle (C++ restatement of the reference path, colour-parallel, all host cores) on the same
               snapshot, a bounded sample of full steps.
  --impl reference   times only that CPU arm (the reference itself is Rust and cannot be built in this image).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

SCENES = {
    # name: (builder, substeps, default settle steps)
    "stack100k": (lambda sc: sc.cube_stack(51, 40, 50, brick=True), 8, 2),   # BASELINE configs[2] (headline): exactly 100 000 cubes
    "stack10k": (lambda sc: sc.cube_stack(23, 20, 22, brick=True), 8, 2),    # BASELINE configs[1]-sized: ~10 000 cubes
    "stack1k": (lambda sc: sc.cube_stack(11, 10, 10, brick=True), 8, 2),
    "ragdolls5k": (lambda sc: sc.ragdoll_field(5000, pitch=3.0, drop_height=0.2), 8, 30),  # BASELINE configs[3]: 85 000 bodies, 80 000 joints
    "ragdolls500": (lambda sc: sc.ragdoll_field(500, pitch=3.0, drop_height=0.2), 8, 30),
    # BASELINE configs[4]: 1M spheres r=0.5, f64, uniform in a 200x50x200 box (seed 42); broad-phase heavy.  One GPU holds the whole scene
    # here; the same scene cut into x-slabs over N GPUs is measured by scripts/slab_bench.py.
    "spheres1m": (lambda sc: sc.falling_spheres(1_000_000, seed=42, scalar=np.float64), 8, 2),
    "spheres100k": (lambda sc: sc.falling_spheres(100_000, seed=42, box=(93.0, 50.0, 93.0), scalar=np.float64), 8, 2),
}


def metric_name(scene: str) -> str:
    label = {"stack100k": "100k-cube stack", "stack10k": "10k-cube stack", "stack1k": "1k-cube stack", "ragdolls5k": "5k-ragdoll field",
             "ragdolls500": "500-ragdoll field", "spheres1m": "1M falling spheres (f64)", "spheres100k": "100k falling spheres (f64)"}[scene]
    return f"physics steps/sec on {label} (broad phase + solver stage per step)"


def algorithmic_bytes(B: int, M: int, P: int, substeps: int, scalar_bytes: int = 4) -> dict:
    """SURVEY.md §8(d), f32 figures scaled by the scalar size; P = total contact points (P/M = mean points/manifold)."""
    k = scalar_bytes / 4.0
    pbar = P / max(M, 1)
    solve_pass = (52 + 76 * pbar) * M + 216 * B
    warm = (28 + 36 * pbar) * M + 136 * B
    a_sub = 84 * B + 80 * B + warm + 2 * solve_pass
    prepare = (100 + 80 * pbar) * M + 170 * B
    step = substeps * a_sub + prepare + 152 * B + 16 * pbar * M
    return {"solve_pass": solve_pass * k, "substep": a_sub * k, "step": step * k}


class ClockSampler(threading.Thread):
    """nvidia-smi clocks/throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], threading.Event()

    def run(self):
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            self.stop_flag.wait(0.2)

    def summary(self) -> dict:
        self.stop_flag.set()
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = sorted(float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 3 + i and r[3 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": float(self.rows[0][1]) if self.rows[0][1].replace(".", "").isdigit() else None,
                "reasons": reasons, "samples": len(self.rows)}


def build_snapshot(scene_name: str, settle: int, ctx=None):
    """Scene -> (scene, params, bodies, manifolds, aabbs) frozen after `settle` full pipeline steps.
    With ctx (GPU arm) the pipeline's hot path runs on the GPU; without it the snapshot is built from the scene's
    initial state only (settle must be 0) so that the CPU arm never needs the GPU."""
    from avian_b200 import plugins, scenes
    builder, substeps, _ = SCENES[scene_name]
    sc = builder(scenes)
    if ctx is not None:
        w = plugins.World(sc, plugins.PhysicsPlugins(ctx), substeps=substeps)
    else:
        sys.path.insert(0, str(ROOT / "tests"))
        import oracle_lib
        w = plugins.World(sc, oracle_lib.oracle_plugins(threads=os.cpu_count() or 1), substeps=substeps)
    for _ in range(settle):
        w.step()
    w.broad_phase()
    man = w.narrow_phase()
    # the steady-state broad-phase input: every current pair is already in the contact graph
    mn, mx = w.aabb_min, w.aabb_max
    aabbs = w.pipeline.intervals(w.bodies, mn, mx, with_existing=True)
    aabbs.joint_disabled_body_pairs = sc.joint_disabled_body_pairs
    return sc, w.params, w.bodies, man, aabbs, w.joints


def pin_columns(ctx, obj):
    """Move every numpy column of a Bodies/Manifolds/Aabbs dataclass into pinned host memory."""
    for k, v in list(obj.__dict__.items()):
        if isinstance(v, np.ndarray):
            setattr(obj, k, ctx.pin_like(v))
    return obj


def run_gpu(args, info):
    import torch
    from avian_b200 import api, parallel
    rank, world, local_rank = info.rank, info.world, info.local_rank
    torch.cuda.set_device(local_rank)
    scalar = np.float64 if args.scene.startswith("spheres") else np.float32
    ctx = api.Context(device=local_rank, scalar=scalar)
    sc, prm, bodies, man, aabbs, joints = build_snapshot(args.scene, args.settle, ctx)
    prm.solver_iterations = args.solver_iterations
    B, M, P = bodies.count, man.count, int(man.penetration.shape[0])
    J = 0 if joints is None else joints.count
    pin_columns(ctx, bodies); pin_columns(ctx, man); pin_columns(ctx, aabbs)
    pairs_out = api.PairList.empty(1 << 20)
    b0, m0 = bodies.copy(), man.copy()     # the frozen snapshot (the step writes results into bodies/man in place)

    def barrier():
        parallel.barrier(info)
        torch.cuda.synchronize()

    def restore():
        for name in ("position", "rotation", "linear_velocity", "angular_velocity"):
            getattr(bodies, name)[...] = getattr(b0, name)
        for name in ("warm_start_normal_impulse", "warm_start_tangent_impulse", "normal_impulse"):
            getattr(man, name)[...] = getattr(m0, name)

    # ---- resident arm: upload once, run K times ---------------------------------------------------------------------
    sampler = ClockSampler(local_rank); sampler.start()
    ctx.solver_upload(prm, bodies, man, joints)
    ctx.broadphase_upload(aabbs)
    for _ in range(args.warmup):
        ctx.broadphase_run(); ctx.solver_run()
    barrier()
    dev_ms, mega_ms, bp_ms, launches = 0.0, 0.0, 0.0, 0
    t0 = time.perf_counter()
    for _ in range(args.steps):
        ctx.broadphase_run()
        ctx.broadphase_download(pairs_out)
        tb = ctx.timings()
        ctx.solver_run()
        ctx.solver_download()
        ts = ctx.timings()
        bp_ms += tb["broad_phase_ms"]; mega_ms += ts["total_ms"]
        dev_ms += tb["broad_phase_ms"] + ts["total_ms"]
        launches += tb["kernel_launches"] + ts["kernel_launches"]
    barrier()
    wall_resident = time.perf_counter() - t0
    new_pairs = int(pairs_out.count)
    restore()

    # ---- end-to-end arm: host buffers in, host buffers out, every step ------------------------------------------------
    for _ in range(max(1, args.warmup // 2)):
        ctx.broadphase(aabbs); ctx.solver_step(prm, bodies, man, joints); restore()
    barrier()
    t0 = time.perf_counter()
    e2e_dev_ms = 0.0
    for _ in range(args.steps):
        ctx.broadphase_upload(aabbs); ctx.broadphase_run(); ctx.broadphase_download(pairs_out)
        ctx.solver_step(prm, bodies, man, joints)
    barrier()
    wall_e2e = time.perf_counter() - t0
    # keep the GPU under the same load until the sampler has a few readings (nvidia-smi takes ~100 ms per call)
    t_hold = time.perf_counter()
    while len(sampler.rows) < 3 and time.perf_counter() - t_hold < 3.0:
        ctx.broadphase_run(); ctx.solver_run(); ctx.solver_download()
    clocks = sampler.summary()
    sb = bodies.position.dtype.itemsize
    h2d = sum(v.nbytes for k, v in bodies.__dict__.items() if isinstance(v, np.ndarray)) + \
        sum(v.nbytes for k, v in man.__dict__.items() if isinstance(v, np.ndarray) and k != "normal_impulse") + \
        sum(v.nbytes for k, v in aabbs.__dict__.items() if isinstance(v, np.ndarray) and k != "order_out")
    d2h = B * (3 + 4 + 3 + 3) * sb + P * 4 * sb + B * 4 + new_pairs * 17
    restore()
    # one more end-to-end step from the frozen snapshot whose outputs are kept: the parity block compares them with the CPU arm's
    gpu_pairs = ctx.broadphase(aabbs)
    ctx.solver_step(prm, bodies, man, joints)
    gpu_out = (bodies.copy(), man.copy(), gpu_pairs, None if aabbs.order_out is None else aabbs.order_out.copy())
    restore()

    # max over ranks
    dev_ms, wall_res_ms, wall_e2e_ms, mega_ms, bp_ms = parallel.reduce_max(
        [dev_ms, wall_resident * 1e3, wall_e2e * 1e3, mega_ms, bp_ms], info, device="cuda")
    if rank != 0:
        ctx.close()
        return None

    K = args.steps
    value = world * K / (dev_ms / 1e3)
    e2e_value = world * K / (wall_e2e_ms / 1e3)
    alg = algorithmic_bytes(B, M, P, int(prm.substeps), sb)
    if J:
        alg["step"] += int(prm.substeps) * (300 * J + 100 * B) * (sb / 4.0)   # XPBD joint pass + velocity projection, SURVEY 8d
    peaks_path = ROOT / "MEASURED_PEAKS.json"
    if peaks_path.exists():
        peak, peak_src = float(json.loads(peaks_path.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    achieved = alg["step"] / (mega_ms / K / 1e3) / 1e9
    traffic = None
    tfile = ROOT / "profiles" / "traffic.json"
    if tfile.exists():
        try:
            traffic = json.loads(tfile.read_text()).get(args.scene, {}).get("dram_bytes_per_launch")
        except Exception:
            traffic = None
    result = {
        "metric": metric_name(args.scene), "value": value, "unit": "steps/s",
        "n_gpus": world, "steps": K, "warmup": args.warmup, "ms_per_step": dev_ms / K, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64" if sb == 8 else "f32", "data": "synthetic",
        "config": {"workload": f"{sc.name}: {B - 1} dynamic bodies on a ground slab, one scene per GPU, {'f64' if sb == 8 else 'f32'}, {int(prm.substeps)} substeps, "
                               f"reference solver semantics (1 warm start + {int(prm.solver_iterations)} biased + 1 relax pass per substep)",
                   "bodies": B, "manifolds": M, "contact_points": P, "joints": J, "solver_iterations": int(prm.solver_iterations), "colliders": int(aabbs.collider.shape[0]),
                   "existing_pairs": 0 if aabbs.existing_pairs is None else int(aabbs.existing_pairs.shape[0]), "new_pairs_per_step": new_pairs,
                   "parallelism": "1 pile per GPU (island sharding), no data-path collective", "settle_steps": args.settle,
                   "l2": "inputs larger than L2: constraint planes + columns > 126 MB per step", "timing": "CUDA events on the library stream, max over ranks"},
        "e2e": {"value": e2e_value, "unit": "steps/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h), "ms_per_step": wall_e2e_ms / K},
        "gpu_launches": launches,
        "clocks": clocks,
        "roofline": {"bound": "hbm", "kernel": f"step_megakernel<{'double' if sb == 8 else 'float'}> (whole solver stage, one launch per step)", "achieved": achieved, "peak": peak,
                     "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src, "algorithmic_bytes_per_launch": alg["step"],
                     "kernel_ms": mega_ms / K, "solve_pass_bytes": alg["solve_pass"]},
        "breakdown_ms": {"broad_phase": bp_ms / K, "solver_stage": mega_ms / K, "resident_wall": wall_res_ms / K},
    }
    if not args.no_cpu and world >= 1:
        keep = {}
        result["cpu_baseline"] = cpu_arm(args, prm, b0, m0, aabbs, sample_steps=args.cpu_steps, joints=joints, keep=keep)
        result["parity"] = parity_block(gpu_out, keep)
    ctx.close()
    return result


def parity_block(gpu_out, keep) -> dict:
    """The GPU step and the CPU oracle step from the SAME full-size snapshot, compared element-wise (tests/helpers.py parity_report):
    relative error with a floor of one unit, bar 1e-5 (BASELINE.json north_star); the broad phase's pair list and persistent order
    bit for bit.  The run FAILS (exit code 1) when the bar is missed."""
    sys.path.insert(0, str(ROOT / "tests"))
    from helpers import BODY_OUT, IMPULSE_OUT, parity_report
    gb, gm, gp, gorder = gpu_out
    ob, om, op, oorder = keep["bodies"], keep["manifolds"], keep["pairs"], keep["order"]
    rep = parity_report(gb, ob, BODY_OUT)
    rep.update(parity_report(gm, om, IMPULSE_OUT))
    pairs_ok = gp.count == op.count and all(np.array_equal(getattr(gp, k)[:gp.count], getattr(op, k)[:op.count]) for k in ("collider1", "collider2", "body1", "body2", "flags"))
    order_ok = gorder is None or oorder is None or bool(np.array_equal(gorder, oorder))
    worst = max(r["max_rel_err"] for r in rep.values())
    return {"max_rel_err_pos": rep["position"]["max_rel_err"], "max_rel_err_rot": rep["rotation"]["max_rel_err"],
            "max_rel_err_vel": max(rep["linear_velocity"]["max_rel_err"], rep["angular_velocity"]["max_rel_err"]),
            "max_rel_err_impulse": max(rep[k]["max_rel_err"] for k in IMPULSE_OUT if k in rep) if any(k in rep for k in IMPULSE_OUT) else 0.0,
            "pairs_bit_exact": bool(pairs_ok), "order_bit_exact": order_ok, "pairs": int(op.count),
            "bit_identical_share": min(r["bit_identical"] for r in rep.values()), "max_ulp": max(r["max_ulp"] for r in rep.values()),
            "definition": "element-wise |gpu - cpu| / max(1, |cpu|) after one full step from the same snapshot; cpu = oracle/ (restated reference)",
            "bar": 1e-5, "ok": bool(worst <= 1e-5 and pairs_ok and order_ok)}


def cpu_arm(args, prm, bodies, man, aabbs, sample_steps: int, joints=None, keep: dict | None = None) -> dict:
    """The oracle (restated reference path, colour-parallel like the reference) on the host cores."""
    sys.path.insert(0, str(ROOT / "tests"))
    import oracle_lib
    from avian_b200 import api
    threads = os.cpu_count() or 1
    t_total = 0.0
    for i in range(sample_steps):
        b, m = bodies.copy(), man.copy()
        a = api.Aabbs(**{k: (v.copy() if isinstance(v, np.ndarray) else v) for k, v in aabbs.__dict__.items()})
        t0 = time.perf_counter()
        pairs = oracle_lib.broadphase(a, capacity=1 << 20)
        oracle_lib.solver_step(prm, b, m, None if joints is None else joints.copy(), threads=threads)
        t_total += time.perf_counter() - t0
        if keep is not None and i == 0:
            keep.update(bodies=b, manifolds=m, pairs=pairs, order=a.order_out)
    return {"value": sample_steps / t_total, "unit": "steps/s", "cores": threads, "kind": "port",
            "sample": f"{sample_steps} full steps of the same snapshot (SAP single-threaded + solver stage colour-parallel on {threads} threads)",
            "ms_per_step": t_total / sample_steps * 1e3}


def run_reference(args, rank: int, world: int):
    """--impl reference: the reference's CPU implementation of the path = the oracle port (Rust cannot be built here)."""
    if rank != 0:
        return None
    sc, prm, bodies, man, aabbs, joints = build_snapshot(args.scene, args.settle if args.settle <= 4 else 0, None)
    prm.solver_iterations = args.solver_iterations
    for _ in range(min(args.warmup, 1)):
        cpu_arm(args, prm, bodies, man, aabbs, 1, joints)
    steps = max(1, min(args.steps, args.cpu_steps_max))
    cb = cpu_arm(args, prm, bodies, man, aabbs, steps, joints)
    B, M, P = bodies.count, man.count, int(man.penetration.shape[0])
    return {
        "impl": "reference", "metric": metric_name(args.scene), "value": cb["value"], "unit": "steps/s",
        "n_gpus": world, "steps": steps, "warmup": min(args.warmup, 1), "ms_per_step": cb["ms_per_step"], "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64" if bodies.position.dtype == np.float64 else "f32", "data": "synthetic",
        "config": {"workload": f"{sc.name}: {B - 1} dynamic bodies on a ground slab, {bodies.position.dtype.name}, {int(prm.substeps)} substeps", "bodies": B, "manifolds": M, "contact_points": P,
                   "joints": 0 if joints is None else joints.count, "settle_steps": args.settle,
                   "note": "restated Avian CPU path (C++ oracle), not Avian itself: no Rust toolchain in this image"},
        "cpu_baseline": cb, "e2e": {"value": cb["value"], "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--scene", default="stack100k", choices=sorted(SCENES))
    ap.add_argument("--settle", type=int, default=None, help="full pipeline steps before the snapshot is frozen (default: per scene)")
    ap.add_argument("--solver-iterations", type=int, default=1, help="EXTENSION: biased solve passes per substep (reference semantics = 1)")
    ap.add_argument("--cpu-steps", type=int, default=3, help="bounded CPU sample (full steps) for cpu_baseline")
    ap.add_argument("--cpu-steps-max", type=int, default=10)
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.settle is None:
        args.settle = SCENES[args.scene][2]

    from avian_b200 import parallel
    info = parallel.rank_info()
    if args.impl == "reference":
        res = run_reference(args, info.rank, info.world)
    else:
        if info.world > 1:
            import torch
            torch.cuda.set_device(info.local_rank)
            parallel.init(backend="nccl")
        res = run_gpu(args, info)
        if info.world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()
    if res is not None:
        print(json.dumps(res))
        if isinstance(res.get("parity"), dict) and not res["parity"]["ok"]:
            sys.exit(1)


if __name__ == "__main__":
    main()
