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
