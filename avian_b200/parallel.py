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

AABB_HALO, AABB_SPLIT_I, AABB_NOT_J = 0x80, 0x40, 0x20


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
    aabbs: object            # api.Aabbs of the local intervals (left halo + owned + right halo), in the persistent order
    index: np.ndarray        # local interval -> global interval
    owned: np.ndarray        # bool per local interval
    left: np.ndarray         # bool per local interval: a far-reaching interval owned by an earlier slab (AVN_AABB_SPLIT_I | NOT_J)


def wide_intervals(aabbs, cuts: np.ndarray, slab: np.ndarray) -> np.ndarray:
    """Intervals that reach past the middle of the slab after their own (a ground slab, a long wall).  Their sweep is split: every
    slab they reach pairs them with the intervals it owns, so no slab has to see the whole scene.  Any choice is exact; this one
    keeps ordinary bodies that merely straddle a cut out of it."""
    world = cuts.size + 1
    if world == 1:
        return np.zeros(slab.shape, dtype=bool)
    min_x, max_x = aabbs.aabb_min[:, 0], aabbs.aabb_max[:, 0]
    hi = np.concatenate([cuts, [min_x.max()]]).astype(np.float64)       # upper end of every slab's min.x range
    nxt = np.minimum(slab + 1, world - 1)
    mid_next = 0.5 * (hi[np.minimum(slab, world - 2)] + hi[nxt])
    return (slab < world - 1) & (max_x.astype(np.float64) >= mid_next)


def shard_aabbs(aabbs, cuts: np.ndarray, rank: int) -> AabbShard:
    from avian_b200 import api
    min_x, max_x = aabbs.aabb_min[:, 0], aabbs.aabb_max[:, 0]
    slab = slab_of(min_x, cuts)
    wide = wide_intervals(aabbs, cuts, slab)
    owned = slab == rank
    halo = np.zeros_like(owned)
    narrow = owned & ~wide
    if narrow.any():
        reach = max_x[narrow].max()
        halo = (slab > rank) & (min_x <= reach)
    lo = cuts[rank - 1] if rank > 0 else None
    left = wide & (slab < rank) & (max_x >= lo) if rank > 0 else np.zeros_like(owned)
    local = owned | halo | left
    index = np.nonzero(local)[0]
    flags = aabbs.flags[index].copy()
    flags[halo[index]] |= AABB_HALO
    flags[(wide & owned)[index]] |= AABB_SPLIT_I
    flags[left[index]] |= AABB_SPLIT_I | AABB_NOT_J
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
    return AabbShard(sub, index, owned[index], left[index])


PAIR_COLUMNS = ("collider1", "collider2", "body1", "body2", "flags")


def _empty_cols():
    return {c: np.zeros(0, dtype=np.uint8 if c == "flags" else np.uint32) for c in PAIR_COLUMNS}


def slab_broadphase_local(broadphase, aabbs, cuts: np.ndarray, rank: int, shard: AabbShard | None = None):
    """One slab's share: (pair columns, the owned part of the new persistent order as GLOBAL interval indices, n_foreign).  The first
    n_foreign pairs start from a far-reaching interval another slab owns (they sort first: its min.x is left of the slab)."""
    sh = shard_aabbs(aabbs, cuts, rank) if shard is None else shard
    if sh.index.size == 0:
        return _empty_cols(), np.zeros(0, dtype=np.uint32), 0
    pairs = broadphase(sh.aabbs)
    n = int(pairs.count)
    order_local = sh.aabbs.order_out
    n_left, n_owned = int(sh.left.sum()), int(sh.owned.sum())
    # left-halo intervals sort before every owned one, owned ones before the right halo (strictly smaller min.x each time)
    assert sh.left[order_local[:n_left]].all() and sh.owned[order_local[n_left:n_left + n_owned]].all(), "slab order: left halo, owned, right halo"
    cols = {c: getattr(pairs, c)[:n].copy() for c in PAIR_COLUMNS}
    n_foreign = 0
    if n_left and n:
        n_foreign = int(np.isin(cols["collider1"], sh.aabbs.collider[sh.left]).sum())
        assert np.isin(cols["collider1"][:n_foreign], sh.aabbs.collider[sh.left]).all(), "foreign pairs must lead the slab's list"
    return cols, sh.index[order_local[n_left:n_left + n_owned]].astype(np.uint32), n_foreign


def merge_slab_results(parts, collider: np.ndarray | None = None):
    """Per-slab (columns, order, n_foreign) in slab order -> (api.PairList, global order), the single-GPU list bit for bit: the slabs'
    own sections concatenated, then every foreign section (pairs starting from a far-reaching interval of an earlier slab) inserted
    right behind the pairs its interval already has.  `collider` = the global collider column (needed only when a section is foreign)."""
    from avian_b200 import api
    order = np.concatenate([p[1] for p in parts])
    own = {c: np.concatenate([p[0][c][p[2]:] for p in parts]) for c in PAIR_COLUMNS}
    foreign = {c: np.concatenate([p[0][c][:p[2]] for p in parts]) for c in PAIR_COLUMNS}
    if foreign["collider1"].size:
        assert collider is not None, "merging a split interval's pairs needs the global collider column"
        rank_of = np.zeros(int(collider.max()) + 1, dtype=np.int64)
        rank_of[collider[order]] = np.arange(order.size)                   # collider id -> position in the sorted order
        fkey = rank_of[foreign["collider1"]]
        keep = np.argsort(fkey, kind="stable")                              # by interval; slab order (= j order) kept within one interval
        at = np.searchsorted(rank_of[own["collider1"]], fkey[keep], side="right")
        own = {c: np.insert(own[c], at, foreign[c][keep]) for c in PAIR_COLUMNS}
    out = api.PairList(own["collider1"], own["collider2"], own["body1"], own["body2"], own["flags"], count=int(own["collider1"].shape[0]))
    return out, order


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
    cols, order, n_foreign = slab_broadphase_local(broadphase, aabbs, cuts, info.rank)
    # uint32 columns travel as int32 bit patterns (NCCL / gloo have no unsigned 32-bit type)
    gathered = {c: allgather_ragged(cols[c].view(np.int32) if cols[c].dtype == np.uint32 else cols[c], info, device) for c in PAIR_COLUMNS}
    orders = allgather_ragged(order.view(np.int32), info, device)
    foreign = allgather_ragged(np.array([n_foreign], dtype=np.int64), info, device)
    parts = [({c: (gathered[c][r].view(np.uint32) if c != "flags" else gathered[c][r]) for c in PAIR_COLUMNS}, orders[r].view(np.uint32), int(foreign[r][0]))
             for r in range(info.world)]
    return merge_slab_results(parts, aabbs.collider)


# ---------------------------------------------------------------------------------------------------------------------------
# x-slab partition of ONE coupled scene: the solver stage (include/avian_b200.h "one coupled scene over several GPUs")
# ---------------------------------------------------------------------------------------------------------------------------
# Bodies are owned by the slab of their position.x; a contact constraint by the owner of its first non-static body.  A rank
# holds its own bodies plus copies of the remote bodies its constraints touch; a body held by more than one rank is a boundary
# body.  Per substep: every rank runs the substep on what it holds, the boundary tables are all-gathered (the one collective),
# every holder rebuilds the same velocity (v_ref + every holder's impulses, summed in rank order) and takes the owner's position
# deltas.  Impulses cross a cut once per substep: the result equals the single-GPU step to solver tolerance (bit for bit when no
# constraint crosses a cut).
@dataclass
class SolverShard:
    bodies: object               # api.Bodies held by this rank, ascending global index
    manifolds: object            # api.Manifolds owned by this rank (colour-major, reference order kept) or None
    body_index: np.ndarray       # local body -> global body
    owned_body: np.ndarray       # bool per local body
    manifold_index: np.ndarray   # local manifold -> global manifold
    point_index: np.ndarray      # local contact point -> global contact point
    bnd_body: np.ndarray         # boundary bodies held here: local body index,
    bnd_slot: np.ndarray         # number of the body among all boundary bodies (bookkeeping, tests),
    bnd_owner: np.ndarray        # owning rank,
    bnd_source: np.ndarray       # [held, world]: record of the body in rank r's packed table, -1 when rank r does not hold it
    slot_count: int              # boundary bodies of the whole scene (0: nothing to exchange)
    record_count: int            # records per rank's table = the busiest rank's list


def body_slab_cuts(bodies, world: int) -> np.ndarray:
    from avian_b200 import api
    return slab_cuts(bodies.position[bodies.kind != api.BODY_STATIC, 0], world)


def _take(obj, rows, skip=()):
    out = {}
    for k, v in obj.__dict__.items():
        out[k] = v if (v is None or k in skip) else np.ascontiguousarray(v[rows])
    return out


def shard_solver(bodies, manifolds, cuts: np.ndarray, rank: int, world: int) -> SolverShard:
    from avian_b200 import api
    B = bodies.count
    static = bodies.kind == api.BODY_STATIC
    owner = slab_of(bodies.position[:, 0], cuts)
    owner[static] = -1
    held = np.zeros((world, B), dtype=bool)
    dyn = np.nonzero(~static)[0]
    held[owner[dyn], dyn] = True
    M = 0 if manifolds is None else manifolds.count
    referenced = np.zeros(B, dtype=bool)
    if M:
        b1, b2 = manifolds.body1.astype(np.int64), manifolds.body2.astype(np.int64)
        o1 = np.where(b1 >= 0, owner[np.maximum(b1, 0)], -1)
        o2 = np.where(b2 >= 0, owner[np.maximum(b2, 0)], -1)
        m_owner = np.where(o1 >= 0, o1, o2)
        m_owner[m_owner < 0] = 0                      # static-static: never solved, parked on rank 0
        for b in (b1, b2):
            ok = (b >= 0) & ~static[np.maximum(b, 0)]
            held[m_owner[ok], b[ok]] = True
        mine = np.nonzero(m_owner == rank)[0]
        for b in (b1, b2):
            bb = b[mine]
            referenced[bb[bb >= 0]] = True
    holders = held.sum(axis=0)
    boundary = np.nonzero(holders > 1)[0]             # ascending global index = slot order
    slot_of_body = np.full(B, -1, dtype=np.int32)
    slot_of_body[boundary] = np.arange(boundary.size, dtype=np.int32)
    local = held[rank] | (referenced & static)
    body_index = np.nonzero(local)[0]
    to_local = np.full(B, -1, dtype=np.int32)
    to_local[body_index] = np.arange(body_index.size, dtype=np.int32)
    lb = api.Bodies(**_take(bodies, body_index))
    lm, point_index, manifold_index = None, np.zeros(0, dtype=np.int64), np.zeros(0, dtype=np.int64)
    if M and mine.size:
        manifold_index = mine
        po = manifolds.point_offsets.astype(np.int64)
        n = po[mine + 1] - po[mine]
        new_po = np.concatenate([[0], np.cumsum(n)])
        point_index = np.repeat(po[mine] - new_po[:-1], n) + np.arange(new_po[-1])
        per_m = ("body1", "body2", "normal", "friction", "restitution", "tangent_velocity")
        per_p = ("anchor1", "anchor2", "penetration", "normal_speed", "warm_start_normal_impulse", "warm_start_tangent_impulse", "normal_impulse")
        cols = {k: (None if getattr(manifolds, k) is None else np.ascontiguousarray(getattr(manifolds, k)[mine])) for k in per_m}
        cols.update({k: np.ascontiguousarray(getattr(manifolds, k)[point_index]) for k in per_p})
        for k in ("body1", "body2"):
            g = cols[k]
            cols[k] = np.where(g >= 0, to_local[np.maximum(g, 0)], g).astype(np.int32)
        cols["point_offsets"] = new_po.astype(np.uint32)
        cols["color_offsets"] = np.searchsorted(mine, np.asarray(manifolds.color_offsets, dtype=np.int64), side="left").astype(np.uint32)
        lm = api.Manifolds(**cols)
    held_b = held[:, boundary]                                  # world x boundary bodies
    mine_bnd = boundary[held_b[rank]]
    # every rank packs only what it holds, in slot order: record of a body in rank r's table = its position in r's list
    pos = np.cumsum(held_b, axis=1) - 1
    source = np.where(held_b, pos, -1)[:, held_b[rank]].T.astype(np.int32)      # [held here, world]
    record_count = int(held_b.sum(axis=1).max()) if boundary.size else 0
    return SolverShard(lb, lm, body_index, (owner[body_index] == rank), manifold_index, point_index,
                       to_local[mine_bnd].astype(np.int32), slot_of_body[mine_bnd].astype(np.int32), owner[mine_bnd].astype(np.int32),
                       np.ascontiguousarray(source), int(boundary.size), record_count)


class GpuSlabEngine:
    """One rank's solver stage on its api.Context, launched substep by substep; the exchange tables are torch CUDA tensors so the
    all-gather runs on device memory, ordered with the launches on the context's own stream."""

    def __init__(self, ctx):
        import torch
        self.ctx = ctx
        self.torch = torch
        self.device = torch.device("cuda", ctx.device)
        self.stream = torch.cuda.ExternalStream(ctx.stream(), device=self.device)
        self.dtype = torch.float32 if ctx.scalar == np.float32 else torch.float64

    def begin(self, prm, shard: SolverShard, rank: int, world: int):
        self.ctx.solver_upload(prm, shard.bodies, shard.manifolds, None)
        self.ctx.solver_set_boundary(shard.bnd_body, shard.bnd_source, shard.bnd_owner, shard.record_count, rank, world)

    def tables(self, record_count: int, world: int):
        from avian_b200 import api
        n = max(record_count, 1) * api.BOUNDARY_RECORD_SCALARS
        return (self.torch.zeros(n, dtype=self.dtype, device=self.device), self.torch.zeros(world * n, dtype=self.dtype, device=self.device))

    def run(self, first: int, count: int, flags: int): self.ctx.solver_run_range(first, count, flags)
    def snapshot(self): self.ctx.solver_boundary_snapshot()
    def pack(self, table): self.ctx.solver_boundary_pack(table.data_ptr())
    def apply(self, gathered): self.ctx.solver_boundary_apply(gathered.data_ptr())
    def needs_restitution(self) -> bool: return self.ctx.solver_needs_restitution()
    def finish(self): self.ctx.solver_download()

    def all_gather(self, gathered, table):
        import torch.distributed as dist
        with self.torch.cuda.stream(self.stream):
            dist.all_gather_into_tensor(gathered, table)

    def copy_table(self, gathered, r: int, table):   # in-process stand-in for the collective (several engines on one device)
        with self.torch.cuda.stream(self.stream):
            gathered[r * table.numel():(r + 1) * table.numel()].copy_(table, non_blocking=True)

    def sync(self): self.stream.synchronize()


def run_slab_step(engines, shards, prm, ranks, world: int, gather, agree_any, upload: bool = True, finish: bool = True, tabs=None):
    """The partitioned solver stage in lockstep over the (engine, shard) pairs this process drives: all `world` of them in the
    in-process tests (gather = copies), exactly one under torch.distributed (gather = the NCCL / gloo all-gather).
    upload=False re-runs the snapshot already resident on the engines (a PREPARE launch restarts from it); finish=False leaves the
    results on the device."""
    from avian_b200 import api
    slot_count = shards[0].slot_count
    if upload:
        for e, sh, r in zip(engines, shards, ranks):
            e.begin(prm, sh, r, world)
    if tabs is None:
        tabs = [e.tables(shards[0].record_count, world) for e in engines]

    def exchange():
        if slot_count == 0:
            return
        for e, (t, g) in zip(engines, tabs):
            e.pack(t)
        gather(engines, tabs)
        for e, (t, g) in zip(engines, tabs):
            e.apply(g)

    substeps = int(prm.substeps)
    for s in range(substeps):
        for e in engines:
            e.run(s, 1, api.RUN_PREPARE if s == 0 else 0)
        exchange()
    if substeps == 0:
        for e in engines:
            e.run(0, 0, api.RUN_PREPARE)
    if agree_any(any(e.needs_restitution() for e in engines)):
        for e in engines:
            e.snapshot()
            e.run(substeps, 0, api.RUN_RESTITUTION)
        exchange()
    for e in engines:
        e.run(substeps, 0, api.RUN_FINALIZE)
        if finish:
            e.finish()
    return tabs


BODY_OUTPUTS = ("position", "rotation", "linear_velocity", "angular_velocity")
POINT_OUTPUTS = ("warm_start_normal_impulse", "warm_start_tangent_impulse", "normal_impulse")


def scatter_results(bodies, manifolds, shard: SolverShard) -> None:
    """Write one rank's results (owned bodies, owned constraints' impulses) back into the global columns."""
    rows = shard.body_index[shard.owned_body]
    for k in BODY_OUTPUTS:
        getattr(bodies, k)[rows] = getattr(shard.bodies, k)[shard.owned_body]
    if shard.manifolds is not None:
        for k in POINT_OUTPUTS:
            getattr(manifolds, k)[shard.point_index] = getattr(shard.manifolds, k)


def slab_solver_step_local(make_engine, prm, bodies, manifolds, world: int, cuts: np.ndarray | None = None):
    """All `world` slabs in this process (tests, single-GPU dry runs).  Results are written into bodies / manifolds."""
    if cuts is None:
        cuts = body_slab_cuts(bodies, world)
    shards = [shard_solver(bodies, manifolds, cuts, r, world) for r in range(world)]
    engines = [make_engine(r) for r in range(world)]

    def gather(engines_, tabs):
        for e in engines_:
            e.sync()            # every table is packed before anyone copies it (the engines have their own streams)
        for e, (t, g) in zip(engines_, tabs):
            for r, (tr, _) in enumerate(tabs):
                e.copy_table(g, r, tr)
        for e in engines_:
            e.sync()

    run_slab_step(engines, shards, prm, list(range(world)), world, gather, lambda flag: flag)
    for sh in shards:
        scatter_results(bodies, manifolds, sh)
    return shards


def dist_gather(engines_, tabs) -> None:
    engines_[0].all_gather(tabs[0][1], tabs[0][0])


def slab_solver_step(engine, prm, bodies, manifolds, info: RankInfo, cuts: np.ndarray | None = None, device: str = "cpu",
                     shard: SolverShard | None = None, gather_results: bool | None = True) -> SolverShard:
    """One rank of the partitioned stage under torch.distributed.  With gather_results every rank ends with the full result in
    bodies / manifolds (one gather of results per step); with False a rank writes back only the rows it owns; with None the results
    stay in the share's own columns (shard.bodies / shard.manifolds: each rank's host feeds its own slab of the application).  `shard` = this rank's share when the caller already holds it (a host application keeps its
    partition between steps)."""
    if shard is None:
        if cuts is None:
            cuts = body_slab_cuts(bodies, info.world)
        shard = shard_solver(bodies, manifolds, cuts, info.rank, info.world)
    sh = shard

    def agree_any(flag: bool) -> bool:
        return reduce_max([1.0 if flag else 0.0], info, device)[0] > 0.0

    run_slab_step([engine], [sh], prm, [info.rank], info.world, dist_gather, agree_any)
    if not gather_results:
        if gather_results is not None:
            scatter_results(bodies, manifolds, sh)
        return sh
    # results: every rank contributes its owned rows
    rows = sh.body_index[sh.owned_body].astype(np.int64)
    all_rows = allgather_ragged(rows, info, device)
    for k in BODY_OUTPUTS:
        col = getattr(sh.bodies, k)[sh.owned_body]
        parts = allgather_ragged(np.ascontiguousarray(col).reshape(-1), info, device)
        width = col.shape[1] if col.ndim > 1 else 1
        for r in range(info.world):
            getattr(bodies, k)[all_rows[r]] = parts[r].reshape(-1, width) if col.ndim > 1 else parts[r]
    if manifolds is not None and manifolds.count:
        all_pts = allgather_ragged(sh.point_index.astype(np.int64), info, device)
        for k in POINT_OUTPUTS:
            glob = getattr(manifolds, k)
            col = getattr(sh.manifolds, k) if sh.manifolds is not None else glob[:0]
            parts = allgather_ragged(np.ascontiguousarray(col).reshape(-1), info, device)
            width = glob.shape[1] if glob.ndim > 1 else 1
            for r in range(info.world):
                glob[all_pts[r]] = parts[r].reshape(-1, width) if glob.ndim > 1 else parts[r]
    return sh


# ---------------------------------------------------------------------------------------------------------------------------
# island sharding of ONE scene: exact, no collective in the data path
# ---------------------------------------------------------------------------------------------------------------------------
# Constraints only carry data between DYNAMIC bodies (static and kinematic bodies are never written by a constraint: dominance,
# solver/contact/mod.rs:129-154), so the connected components of the graph "dynamic bodies joined by contacts and joints" never
# exchange anything within a step.  Restricting the colour-major manifold list and the typed joint arrays to one component keeps the
# relative order of its constraints, which is all the Gauss-Seidel result of that component depends on: solving the components on
# different GPUs reproduces the single-GPU step bit for bit (tests/test_island_cpu.py, tests/test_gpu_multi.py).
@dataclass
class IslandShard:
    bodies: object
    manifolds: object            # api.Manifolds or None
    joints: object               # api.JointSet or None
    body_index: np.ndarray       # local body -> global body
    owned_body: np.ndarray       # bool per local body: results are taken from this rank
    manifold_index: np.ndarray
    point_index: np.ndarray
    joint_index: dict            # joint type -> local joint -> global joint of that type


def find_islands(bodies, manifolds=None, joints=None):
    """labels[B]: island id of every dynamic body (numbered by their smallest body index), -1 for static / kinematic bodies."""
    from scipy.sparse import coo_matrix
    from scipy.sparse.csgraph import connected_components
    from avian_b200 import api
    B = bodies.count
    dyn = bodies.kind == api.BODY_DYNAMIC
    src, dst = [], []

    def add(b1, b2):
        b1, b2 = np.asarray(b1, dtype=np.int64), np.asarray(b2, dtype=np.int64)
        ok = (b1 >= 0) & (b2 >= 0)
        ok &= dyn[np.maximum(b1, 0)] & dyn[np.maximum(b2, 0)]
        src.append(b1[ok]); dst.append(b2[ok])

    if manifolds is not None and manifolds.count:
        add(manifolds.body1, manifolds.body2)
    if joints is not None:
        for j in joints.types.values():
            if j.count:
                add(j.body1, j.body2)
    s = np.concatenate(src) if src else np.zeros(0, dtype=np.int64)
    d = np.concatenate(dst) if dst else np.zeros(0, dtype=np.int64)
    n, comp = connected_components(coo_matrix((np.ones(s.size, dtype=np.int8), (s, d)), shape=(B, B)), directed=False)
    labels = np.full(B, -1, dtype=np.int64)
    idx = np.nonzero(dyn)[0]
    # renumber by first appearance so that the numbering does not depend on the component search
    _, first = np.unique(comp[idx], return_index=True)
    order = np.argsort(first, kind="stable")
    remap = np.empty(order.size, dtype=np.int64)
    remap[order] = np.arange(order.size)
    uniq = np.unique(comp[idx])
    labels[idx] = remap[np.searchsorted(uniq, comp[idx])]
    return labels, int(order.size)


def assign_islands(labels: np.ndarray, weights: np.ndarray, world: int) -> np.ndarray:
    """rank of every island: heaviest first onto the least loaded rank (ties: lower island id, lower rank) — deterministic."""
    n = int(labels.max()) + 1 if labels.size and labels.max() >= 0 else 0
    w = np.bincount(labels[labels >= 0], weights=weights[labels >= 0], minlength=n) if n else np.zeros(0)
    rank_of = np.zeros(n, dtype=np.int64)
    load = np.zeros(world)
    for isl in np.lexsort((np.arange(n), -w)):
        r = int(np.argmin(load))
        rank_of[isl] = r
        load[r] += w[isl] + 1.0
    return rank_of


def shard_by_island(bodies, manifolds, joints, world: int, rank: int, labels: np.ndarray | None = None) -> IslandShard:
    from avian_b200 import api
    B = bodies.count
    if labels is None:
        labels, _ = find_islands(bodies, manifolds, joints)
    M = 0 if manifolds is None else manifolds.count
    # weight of a body = the constraints it carries (the solver's work), so that the ranks get similar loads
    weight = np.ones(B)
    cons = []   # (b1, b2) index arrays of every constraint list, manifolds first then joints by type
    if M:
        cons.append(("m", None, manifolds.body1.astype(np.int64), manifolds.body2.astype(np.int64)))
    if joints is not None:
        for t in sorted(joints.types):
            j = joints.types[t]
            if j.count:
                cons.append(("j", t, j.body1.astype(np.int64), j.body2.astype(np.int64)))
    for _, _, b1, b2 in cons:
        for b in (b1, b2):
            np.add.at(weight, b[b >= 0], 1.0)
    rank_of_island = assign_islands(labels, weight, world)
    body_rank = np.where(labels >= 0, rank_of_island[np.maximum(labels, 0)] if rank_of_island.size else 0, -1)

    def owner(b1, b2):   # rank of a constraint: the island of its dynamic body; none dynamic (never solved) -> rank 0
        r1 = np.where(b1 >= 0, body_rank[np.maximum(b1, 0)], -1)
        r2 = np.where(b2 >= 0, body_rank[np.maximum(b2, 0)], -1)
        r = np.where(r1 >= 0, r1, r2)
        return np.where(r >= 0, r, 0)

    local = body_rank == rank
    picks = {}
    for kind, t, b1, b2 in cons:
        mine = np.nonzero(owner(b1, b2) == rank)[0]
        picks[(kind, t)] = mine
        for b in (b1[mine], b2[mine]):
            local[b[b >= 0]] = True
    # non-dynamic bodies nobody references still have to be stepped by someone (kinematic motion): rank 0
    if rank == 0:
        ref = np.zeros(B, dtype=bool)
        for _, _, b1, b2 in cons:
            for b in (b1, b2):
                ref[b[b >= 0]] = True
        local |= (labels < 0) & ~ref
    body_index = np.nonzero(local)[0]
    to_local = np.full(B, -1, dtype=np.int32)
    to_local[body_index] = np.arange(body_index.size, dtype=np.int32)
    owned = body_rank[body_index] == rank
    if rank == 0:
        owned |= labels[body_index] < 0          # replicated non-dynamic bodies: every holder computes the same motion, rank 0 reports it
    lb = api.Bodies(**_take(bodies, body_index))
    remap = lambda g: np.where(g >= 0, to_local[np.maximum(g, 0)], g).astype(np.int32)
    lm, point_index, manifold_index = None, np.zeros(0, dtype=np.int64), np.zeros(0, dtype=np.int64)
    mine = picks.get(("m", None))
    if mine is not None and mine.size:
        manifold_index = mine
        po = manifolds.point_offsets.astype(np.int64)
        n = po[mine + 1] - po[mine]
        new_po = np.concatenate([[0], np.cumsum(n)])
        point_index = np.repeat(po[mine] - new_po[:-1], n) + np.arange(new_po[-1])
        per_m = ("normal", "friction", "restitution", "tangent_velocity")
        per_p = ("anchor1", "anchor2", "penetration", "normal_speed", "warm_start_normal_impulse", "warm_start_tangent_impulse", "normal_impulse")
        cols = {k: (None if getattr(manifolds, k) is None else np.ascontiguousarray(getattr(manifolds, k)[mine])) for k in per_m}
        cols.update({k: np.ascontiguousarray(getattr(manifolds, k)[point_index]) for k in per_p})
        cols["body1"], cols["body2"] = remap(manifolds.body1[mine]), remap(manifolds.body2[mine])
        cols["point_offsets"] = new_po.astype(np.uint32)
        cols["color_offsets"] = np.searchsorted(mine, np.asarray(manifolds.color_offsets, dtype=np.int64), side="left").astype(np.uint32)
        lm = api.Manifolds(**cols)
    lj, joint_index = None, {}
    if joints is not None:
        types = {}
        for t in sorted(joints.types):
            mine_j = picks.get(("j", t))
            if mine_j is None or mine_j.size == 0:
                continue
            j = joints.types[t]
            cols = _take(j, mine_j)
            cols["body1"], cols["body2"] = remap(j.body1[mine_j]), remap(j.body2[mine_j])
            types[t] = api.Joints(**cols)
            joint_index[t] = mine_j
        if types:
            lj = api.JointSet(types)
    return IslandShard(lb, lm, lj, body_index, owned, manifold_index, point_index, joint_index)


JOINT_OUTPUTS = ("force", "torque")


def scatter_island_results(bodies, manifolds, joints, shard: IslandShard) -> None:
    rows = shard.body_index[shard.owned_body]
    for k in BODY_OUTPUTS:
        getattr(bodies, k)[rows] = getattr(shard.bodies, k)[shard.owned_body]
    if shard.manifolds is not None:
        for k in POINT_OUTPUTS:
            getattr(manifolds, k)[shard.point_index] = getattr(shard.manifolds, k)
    if shard.joints is not None:
        for t, idx in shard.joint_index.items():
            for k in JOINT_OUTPUTS:
                dst, src = getattr(joints.types[t], k), getattr(shard.joints.types[t], k)
                if dst is not None and src is not None:
                    dst[idx] = src


def island_solver_step_local(solver_step, prm, bodies, manifolds, joints, world: int):
    """Every rank's share solved one after the other in this process (tests, single-GPU dry runs): solver_step(prm, bodies,
    manifolds, joints) is the engine (Context.solver_step).  Results are written back into the global columns."""
    labels, _ = find_islands(bodies, manifolds, joints)
    shards = [shard_by_island(bodies, manifolds, joints, world, r, labels) for r in range(world)]
    for sh in shards:
        if sh.bodies.count:
            solver_step(prm, sh.bodies, sh.manifolds, sh.joints)
    for sh in shards:
        scatter_island_results(bodies, manifolds, joints, sh)
    return shards


def island_solver_step(solver_step, prm, bodies, manifolds, joints, info: RankInfo, device: str = "cpu", gather_results: bool = True) -> IslandShard:
    """One rank of the island-sharded stage: no collective while solving; with gather_results one all-gather of the owned rows at
    the end so that every rank holds the full result."""
    labels, _ = find_islands(bodies, manifolds, joints)
    sh = shard_by_island(bodies, manifolds, joints, info.world, info.rank, labels)
    if sh.bodies.count:
        solver_step(prm, sh.bodies, sh.manifolds, sh.joints)
    scatter_island_results(bodies, manifolds, joints, sh)
    if not gather_results or info.world == 1:
        return sh
    rows = sh.body_index[sh.owned_body].astype(np.int64)
    all_rows = allgather_ragged(rows, info, device)
    for k in BODY_OUTPUTS:
        col = getattr(sh.bodies, k)[sh.owned_body]
        parts = allgather_ragged(np.ascontiguousarray(col).reshape(-1), info, device)
        for r in range(info.world):
            getattr(bodies, k)[all_rows[r]] = parts[r].reshape(-1, col.shape[1]) if col.ndim > 1 else parts[r]
    if manifolds is not None and manifolds.count:
        all_pts = allgather_ragged(sh.point_index.astype(np.int64), info, device)
        for k in POINT_OUTPUTS:
            glob = getattr(manifolds, k)
            col = getattr(sh.manifolds, k) if sh.manifolds is not None else glob[:0]
            parts = allgather_ragged(np.ascontiguousarray(col).reshape(-1), info, device)
            for r in range(info.world):
                glob[all_pts[r]] = parts[r].reshape(-1, glob.shape[1]) if glob.ndim > 1 else parts[r]
    if joints is not None:
        for t in sorted(joints.types):
            idx = sh.joint_index.get(t, np.zeros(0, dtype=np.int64)).astype(np.int64)
            all_idx = allgather_ragged(idx, info, device)
            for k in JOINT_OUTPUTS:
                glob = getattr(joints.types[t], k)
                if glob is None:
                    continue
                col = getattr(sh.joints.types[t], k) if (sh.joints is not None and t in sh.joints.types) else glob[:0]
                parts = allgather_ragged(np.ascontiguousarray(col).reshape(-1), info, device)
                for r in range(info.world):
                    glob[all_idx[r]] = parts[r].reshape(-1, glob.shape[1]) if glob.ndim > 1 else parts[r]
    return sh
