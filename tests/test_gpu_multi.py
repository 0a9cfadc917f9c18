"""x-slab partition on the device.  One GPU is enough for the partition logic itself (the slabs run one after
the other on one context); the NCCL all-gather needs two and is skipped otherwise."""
import os
import sys
from pathlib import Path

import numpy as np
import pytest
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
from avian_b200 import api, parallel, scenes  # noqa: E402
import oracle_lib  # noqa: E402
from test_slab_cpu import assert_same_pairs, random_aabbs  # noqa: E402

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("scalar", [np.float32, np.float64])
@pytest.mark.parametrize("world", [2, 4])
def test_slabs_on_one_device_concatenate_to_the_single_list(world, scalar):
    a = random_aabbs(6000, seed=world + 10, scalar=scalar)
    want = oracle_lib.broadphase(a)
    want_order = a.order_out.copy()
    with api.Context(device=0, scalar=scalar) as ctx:
        single = ctx.broadphase(a)
        assert_same_pairs(single, want)
        cuts = parallel.slab_cuts(a.aabb_min[:, 0], world)
        parts = [parallel.slab_broadphase_local(ctx.broadphase, a, cuts, r) for r in range(world)]
    got, order = parallel.merge_slab_results(parts)
    assert_same_pairs(got, want)
    assert np.array_equal(order, want_order)
    assert max(p[0]["collider1"].shape[0] for p in parts) < want.count


def test_stack_scene_slabs_on_one_device(gpu_ctx):
    """The headline scene's geometry (brick stack on a ground slab that every slab's sweep must see) cut into 4 slabs."""
    from avian_b200 import plugins
    sc = scenes.cube_stack(12, 6, 10, brick=True)
    w = plugins.World(sc, plugins.PhysicsPlugins(gpu_ctx), substeps=2)
    w.step()
    a = w.pipeline.intervals(w.bodies, w.aabb_min, w.aabb_max, with_existing=False)
    a.order_out = np.zeros(a.collider.shape[0], dtype=np.uint32)
    want = oracle_lib.broadphase(a)
    want_order = a.order_out.copy()
    cuts = parallel.slab_cuts(a.aabb_min[:, 0], 4)
    parts = [parallel.slab_broadphase_local(gpu_ctx.broadphase, a, cuts, r) for r in range(4)]
    got, order = parallel.merge_slab_results(parts)
    assert_same_pairs(got, want)
    assert np.array_equal(order, want_order)


def _nccl_worker(rank, world, port, q):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    info = parallel.init(backend="nccl")
    a = random_aabbs(20000, seed=5)
    with api.Context(device=rank) as ctx:
        got, order = parallel.slab_broadphase(ctx.broadphase, a, info, device=f"cuda:{rank}")
    q.put((rank, got.count, {c: getattr(got, c).copy() for c in parallel.PAIR_COLUMNS}, order))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_gpu_nccl_slab_broadphase():
    import torch.multiprocessing as mp
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() % 1000)
    procs = [ctx.Process(target=_nccl_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = sorted((q.get(timeout=300) for _ in range(world)), key=lambda r: r[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    a = random_aabbs(20000, seed=5)
    want = oracle_lib.broadphase(a)
    for rank, count, cols, order in results:
        assert count == want.count
        for c in parallel.PAIR_COLUMNS:
            assert np.array_equal(cols[c], getattr(want, c)[:want.count]), (rank, c)
        assert np.array_equal(order, a.order_out)
