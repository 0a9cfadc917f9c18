#!/usr/bin/env python
"""ONE scene over N GPUs (strong scaling): the x-slab partition of the broad phase and of the solver stage, one NCCL all-gather of
the boundary tables per substep (avian_b200/parallel.py, include/avian_b200.h).  Launch:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P scripts/slab_bench.py \
        --scene spheres1m --steps 10 --warmup 3
bench.py's contract (N independent scenes, weak scaling, no collective) is unchanged; this is the other multi-GPU mode of SURVEY §8e.
Resident arm: every rank's share is uploaded once; a step = local broad phase + the partitioned solver stage (device timed, CUDA
events on the library stream, max over ranks).  End-to-end arm: every rank's host feeds its own slab, every step (H2D of the share,
launches, exchanges, D2H of the share's results; no gather of results).  Prints one JSON line on rank 0."""
import argparse
import json
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scene", default="spheres1m", choices=sorted(bench.SCENES))
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--settle", type=int, default=None)
    args = ap.parse_args()
    import torch
    from avian_b200 import api, parallel
    info = parallel.rank_info()
    torch.cuda.set_device(info.local_rank)
    if info.world > 1:
        parallel.init(backend="nccl")
    dev = f"cuda:{info.local_rank}"
    scalar = np.float64 if args.scene.startswith("spheres") else np.float32
    settle = bench.SCENES[args.scene][2] if args.settle is None else args.settle
    ctx = api.Context(device=info.local_rank, scalar=scalar)
    sc, prm, bodies, man, aabbs, joints = bench.build_snapshot(args.scene, settle, ctx)   # the same global snapshot on every rank
    assert joints is None or joints.count == 0, "the slab partition covers contact scenes (jointed scenes shard by island)"
    world, rank = info.world, info.rank
    cuts = parallel.body_slab_cuts(bodies, world)
    shard = parallel.shard_solver(bodies, man, cuts, rank, world)
    acuts = parallel.slab_cuts(aabbs.aabb_min[:, 0], world)
    ashard = parallel.shard_aabbs(aabbs, acuts, rank)
    bench.pin_columns(ctx, shard.bodies)                     # the share lives in pinned host memory, like bench.py's columns
    if shard.manifolds is not None:
        bench.pin_columns(ctx, shard.manifolds)
    bench.pin_columns(ctx, ashard.aabbs)
    engine = parallel.GpuSlabEngine(ctx)
    gather = parallel.dist_gather if world > 1 else (lambda e, t: None)
    if world == 1:
        engine.all_gather = lambda g, t: g.copy_(t)
        gather = parallel.dist_gather

    def sync():
        parallel.barrier(info)
        torch.cuda.synchronize()

    # ---- resident arm
    engine.begin(prm, shard, rank, world)
    ctx.broadphase_upload(ashard.aabbs)
    any_restitution = parallel.reduce_max([1.0 if engine.needs_restitution() else 0.0], info, dev)[0] > 0.0   # agreed once per upload
    agree = lambda flag: any_restitution
    tabs = None
    for _ in range(args.warmup):
        ctx.broadphase_run()
        tabs = parallel.run_slab_step([engine], [shard], prm, [rank], world, gather, agree, upload=False, finish=False, tabs=tabs)
    sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(engine.stream):
        e0.record()
    for _ in range(args.steps):
        ctx.broadphase_run()
        parallel.run_slab_step([engine], [shard], prm, [rank], world, gather, agree, upload=False, finish=False, tabs=tabs)
    with torch.cuda.stream(engine.stream):
        e1.record()
    sync()
    dev_ms = e0.elapsed_time(e1)
    engine.finish()

    # ---- end-to-end arm: every rank's host feeds its own slab: H2D of the share, launches + exchanges, D2H of the share's results
    #      (pairs whose first interval the slab owns; owned bodies; owned constraints' impulses).  No gather of results.
    b0 = {k: getattr(shard.bodies, k).copy() for k in parallel.BODY_OUTPUTS}
    pairs_out = api.PairList.empty(max(1 << 20, 4 * int(ashard.index.size)))
    sync()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        ctx.broadphase_upload(ashard.aabbs); ctx.broadphase_run(); ctx.broadphase_download(pairs_out)
        for k, v in b0.items():
            getattr(shard.bodies, k)[...] = v               # every step starts from the same snapshot
        parallel.slab_solver_step(engine, prm, None, None, info, device=dev, shard=shard, gather_results=None)
    sync()
    e2e_ms = (time.perf_counter() - t0) * 1e3
    dev_ms, e2e_ms = parallel.reduce_max([dev_ms, e2e_ms], info, dev)
    held = parallel.reduce_sum([float(shard.bodies.count), float(0 if shard.manifolds is None else shard.manifolds.count),
                                float(ashard.index.size)], info, dev)
    if rank == 0:
        K = args.steps
        print(json.dumps({
            "metric": bench.metric_name(args.scene) + " — ONE scene over N GPUs (x-slab partition)", "value": K / (dev_ms / 1e3), "unit": "steps/s",
            "n_gpus": world, "steps": K, "warmup": args.warmup, "ms_per_step": dev_ms / K, "higher_is_better": True, "scaling": "strong",
            "dtype": "f64" if scalar == np.float64 else "f32", "data": "synthetic",
            "config": {"workload": f"{sc.name}: one scene cut into {world} x-slabs", "bodies": bodies.count, "manifolds": man.count,
                       "boundary_bodies": shard.slot_count, "bodies_held_all_ranks": int(held[0]), "manifolds_all_ranks": int(held[1]),
                       "intervals_held_all_ranks": int(held[2]), "colliders": int(aabbs.collider.shape[0]),
                       "collective": "one all-gather of the boundary tables per substep (+ one after the restitution pass)",
                       "exchange_bytes_per_substep_per_rank": shard.record_count * api.BOUNDARY_RECORD_SCALARS * bodies.position.dtype.itemsize,
                       "parity": "solver tolerance across cuts (impulses cross a cut once per substep); broad phase bit-exact"},
            "e2e": {"value": K / (e2e_ms / 1e3), "unit": "steps/s", "ms_per_step": e2e_ms / K}}), flush=True)
    ctx.close()
    if info.world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
