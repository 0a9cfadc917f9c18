"""One process per GPU.  The hot path shards by ISLAND: constraints only couple bodies of one connected
component of dynamic bodies, so independent piles / ragdolls / scenes are dealt round-robin to the ranks and stepped with
no data-path collective.  torch.distributed is used for the plumbing only: a barrier around the timed region and the
max-over-ranks reduction of the device timings (NCCL on the GPU box, gloo in the CPU tests)."""
from __future__ import annotations

import os
from dataclasses import dataclass


@dataclass
class RankInfo:
    rank: int
    world: int
    local_rank: int


def rank_info() -> RankInfo:
    return RankInfo(int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0")))


def shard_islands(n_islands: int, world: int, rank: int) -> list[int]:
    """Island ids owned by `rank`: round-robin, deterministic, every island owned exactly once."""
    return list(range(rank, n_islands, world))


def init(backend: str | None = None) -> RankInfo:
    info = rank_info()
    if info.world > 1:
        import torch
        import torch.distributed as dist
        if not dist.is_initialized():
            if backend is None:
                backend = "nccl" if torch.cuda.is_available() else "gloo"
            dist.init_process_group(backend=backend)
    return info


def barrier(info: RankInfo) -> None:
    if info.world > 1:
        import torch.distributed as dist
        dist.barrier()


def reduce_max(values: list[float], info: RankInfo, device: str = "cpu") -> list[float]:
    """max over ranks of each entry (timings are reported as the slowest rank's)"""
    if info.world == 1:
        return list(values)
    import torch
    import torch.distributed as dist
    t = torch.tensor(values, dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(x) for x in t.tolist()]


def reduce_sum(values: list[float], info: RankInfo, device: str = "cpu") -> list[float]:
    if info.world == 1:
        return list(values)
    import torch
    import torch.distributed as dist
    t = torch.tensor(values, dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return [float(x) for x in t.tolist()]


def aggregate_throughput(units_per_rank: float, seconds_per_rank: float, info: RankInfo, device: str = "cpu") -> float:
    """whole-job throughput = units processed by all ranks / the slowest rank's time"""
    total_units = reduce_sum([units_per_rank], info, device)[0]
    t = reduce_max([seconds_per_rank], info, device)[0]
    return total_units / t


# ---------------------------------------------------------------------------------------------------------------------------
# x-slab partition of ONE coupled scene (second row): the broad phase
# ---------------------------------------------------------------------------------------------------------------------------
# The reference's pair list is ordered by (rank i, rank j) in the x-sorted interval order (broad_phase.rs:375-439).  Cutting the
# min.x axis into `world` value ranges therefore cuts the list into `world` contiguous pieces: slab g owns the intervals with
# cuts[g-1] <= min.x < cuts[g] and emits exactly the pairs whose earlier element it owns.  The later element may live further
# right, so every slab also receives a "halo": the intervals of later slabs that start before its right-most max.x; they carry
# AVN_AABB_HALO and never start a sweep.  A stable sort of a subsequence keeps the relative order of the full stable sort, so the
# concatenation of the per-slab lists in slab order IS the single-GPU list, bit for bit (tests/test_parallel_cpu.py,
# tests/test_gpu_multi.py).  The only collective is the all-gather of the per-slab pair lists.
import numpy as np  # noqa: E402

AABB_HALO = 0x80


def slab_cuts(min_x: np.ndarray, world: int) -> np.ndarray:
    """world-1 cut values along min.x that balance the interval counts (equal values always land in the same slab)."""
    if world <= 1 or min_x.size == 0:
        return np.empty(0, dtype=min_x.dtype)
    s = np.sort(min_x)
    return np.array([s[(k * s.size) // world] for k in range(1, world)], dtype=min_x.dtype)


def slab_of(x: np.ndarray, cuts: np.ndarray) -> np.ndarray:
    """Owning slab of every value: slab g owns cuts[g-1] <= x < cuts[g]."""
    return np.searchsorted(cuts, x, side="right").astype(np.int32)


@dataclass
class AabbShard:
    aabbs: object            # api.Aabbs of the local intervals (owned + halo), in the persistent order
    index: np.ndarray        # local interval -> global interval
    owned: np.ndarray        # bool per local interval


def shard_aabbs(aabbs, cuts: np.ndarray, rank: int) -> AabbShard:
    from avian_b200 import api
    min_x = aabbs.aabb_min[:, 0]
    slab = slab_of(min_x, cuts)
    owned = slab == rank
    halo = np.zeros_like(owned)
    if owned.any():
        reach = aabbs.aabb_max[owned, 0].max()
        halo = (slab > rank) & (min_x <= reach)
    local = owned | halo
    index = np.nonzero(local)[0]
    flags = aabbs.flags[index].copy()
    flags[halo[index]] |= AABB_HALO
    existing = aabbs.existing_pairs
    if existing is not None and existing.size:
        top = int(max(int(aabbs.collider.max()), int((existing >> np.uint64(32)).max()), int((existing & np.uint64(0xFFFFFFFF)).max())))
        here = np.zeros(top + 1, dtype=bool)
        here[aabbs.collider[index]] = True
        existing = np.ascontiguousarray(existing[here[(existing >> np.uint64(32)).astype(np.int64)] & here[(existing & np.uint64(0xFFFFFFFF)).astype(np.int64)]])
    take = lambda a: None if a is None else np.ascontiguousarray(a[index])
    sub = api.Aabbs(collider=take(aabbs.collider), body=take(aabbs.body), aabb_min=take(aabbs.aabb_min), aabb_max=take(aabbs.aabb_max), flags=flags,
                    memberships=take(aabbs.memberships), filters=take(aabbs.filters), order_out=np.zeros(index.size, dtype=np.uint32),
                    existing_pairs=existing, joint_disabled_body_pairs=aabbs.joint_disabled_body_pairs)
    return AabbShard(sub, index, owned[index])


PAIR_COLUMNS = ("collider1", "collider2", "body1", "body2", "flags")


def slab_broadphase_local(broadphase, aabbs, cuts: np.ndarray, rank: int):
    """One slab's share: (pair columns, the owned part of the new persistent order as GLOBAL interval indices)."""
    sh = shard_aabbs(aabbs, cuts, rank)
    if sh.index.size == 0:
        return {c: np.zeros(0, dtype=np.uint8 if c == "flags" else np.uint32) for c in PAIR_COLUMNS}, np.zeros(0, dtype=np.uint32)
    pairs = broadphase(sh.aabbs)
    n = int(pairs.count)
    order_local = sh.aabbs.order_out
    n_owned = int(sh.owned.sum())
    # owned intervals sort before every halo interval (their min.x is strictly smaller)
    assert sh.owned[order_local[:n_owned]].all(), "slab order: owned intervals must precede the halo"
    return {c: getattr(pairs, c)[:n].copy() for c in PAIR_COLUMNS}, sh.index[order_local[:n_owned]].astype(np.uint32)


def merge_slab_results(parts):
    """Concatenate the per-slab (columns, order) results in slab order -> (api.PairList, global order)."""
    from avian_b200 import api
    cols = {c: np.concatenate([p[0][c] for p in parts]) for c in PAIR_COLUMNS}
    out = api.PairList(cols["collider1"], cols["collider2"], cols["body1"], cols["body2"], cols["flags"], count=int(cols["collider1"].shape[0]))
    return out, np.concatenate([p[1] for p in parts])


def allgather_ragged(arr: np.ndarray, info: RankInfo, device: str = "cpu") -> list[np.ndarray]:
    """All-gather of one 1-D array whose length differs per rank (counts first, then the padded payload)."""
    if info.world == 1:
        return [arr]
    import torch
    import torch.distributed as dist
    n = torch.tensor([arr.shape[0]], dtype=torch.int64, device=device)
    counts = [torch.zeros_like(n) for _ in range(info.world)]
    dist.all_gather(counts, n)
    counts = [int(c.item()) for c in counts]
    width = max(max(counts), 1)
    send = torch.zeros(width, dtype=torch.from_numpy(arr[:0]).dtype, device=device)
    send[:arr.shape[0]] = torch.from_numpy(np.ascontiguousarray(arr)).to(device)
    recv = torch.empty(info.world * width, dtype=send.dtype, device=device)
    dist.all_gather_into_tensor(recv, send)
    recv = recv.cpu().numpy()
    return [recv[r * width:r * width + counts[r]].copy() for r in range(info.world)]


def slab_broadphase(broadphase, aabbs, info: RankInfo, cuts: np.ndarray | None = None, device: str = "cpu"):
    """The broad phase of one scene cut into info.world x-slabs; every rank returns the full pair list and the full new
    persistent order, identical to the single-GPU result.  `broadphase` is the local engine (Context.broadphase)."""
    if cuts is None:
        cuts = slab_cuts(aabbs.aabb_min[:, 0], info.world)
    cols, order = slab_broadphase_local(broadphase, aabbs, cuts, info.rank)
    # uint32 columns travel as int32 bit patterns (NCCL / gloo have no unsigned 32-bit type)
    gathered = {c: allgather_ragged(cols[c].view(np.int32) if cols[c].dtype == np.uint32 else cols[c], info, device) for c in PAIR_COLUMNS}
    orders = allgather_ragged(order.view(np.int32), info, device)
    parts = [({c: (gathered[c][r].view(np.uint32) if c != "flags" else gathered[c][r]) for c in PAIR_COLUMNS}, orders[r].view(np.uint32))
             for r in range(info.world)]
    return merge_slab_results(parts)
