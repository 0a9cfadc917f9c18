"""TEST INFRASTRUCTURE — CPU restatement of the reference's persistent simulation islands and island sleeping.
Only tests/ may import this; the product path is csrc/contacts.cu (avn_islands_configure / avn_islands_step).  Parity unpinned by the
reference itself (no Rust toolchain here, the reference holds no golden vectors for islands): this follows the Rust source line by line in
plain Python (sequential, small scenes only).

Restated:
  PhysicsIslands::add_contact / merge_islands        dynamics/solver/islands/mod.rs:513-592, 814-993  (bigger island survives a merge)
  PhysicsIslands::remove_contact                      mod.rs:594-667                                   (constraints_removed += 1)
  PhysicsIslands::add_joint                           mod.rs:669-747
  split_island system + PhysicsIslands::split_island  mod.rs:161-179, 995-1270                         (DFS over the island's bodies)
  update_sleeping_states, wake_islands_with_sleeping_disabled, sleep_islands   islands/sleeping.rs:164-292
  WakeIslands (timers back to 0)                      islands/sleeping.rs WakeIslands::apply
  the narrow phase's calls                            collision/narrow_phase/system_param.rs:196-205, 244-258, 306-313 (events in ascending ContactId)
Schedule order of one step (schedule/mod.rs:98-105): NarrowPhase (events) -> Solver (Finalize: split_island) -> Sleeping.
"""
from __future__ import annotations

import numpy as np

STATIC = 2


class Island:
    def __init__(self, iid):
        self.id = iid
        self.bodies: list[int] = []        # linked list order (head .. tail)
        self.contacts: set[int] = set()
        self.joints: set[int] = set()
        self.removed = 0                   # constraints_removed
        self.sleeping = False


class IslandsOracle:
    def __init__(self, kind, joints=(), thr_lin=None, thr_ang=None, disabled=None, time_to_sleep=0.5, length_unit=1.0, scalar=np.float32,
                 candidate="island"):
        """candidate = "island": the reference — the split candidate is an island ID, retired when that island is the smaller side of a merge
        (remove_island, mod.rs:456-464).  candidate = "body": the device's stated deviation (include/avian_b200.h avn_islands_step) — the
        candidate is the sleepiest BODY and the island that holds it one step later is split."""
        self.candidate_mode = candidate
        self.candidate_body = None
        self.kind = np.asarray(kind)
        self.B = int(self.kind.shape[0])
        self.S = np.dtype(scalar).type
        self.thr_lin = np.full(self.B, 0.15, dtype=np.float32) if thr_lin is None else np.asarray(thr_lin, dtype=np.float32)
        self.thr_ang = np.full(self.B, 0.15, dtype=np.float32) if thr_ang is None else np.asarray(thr_ang, dtype=np.float32)
        self.disabled = np.zeros(self.B, dtype=bool) if disabled is None else np.asarray(disabled, dtype=bool)
        self.time_to_sleep = np.float32(time_to_sleep)
        self.length_unit = self.S(length_unit)
        self.timer = np.zeros(self.B, dtype=np.float32)
        self.islands: dict[int, Island] = {}
        self.free_ids: list[int] = []      # slab: the lowest vacant key first
        self.next_new = 0
        self.body_island = np.full(self.B, -1, dtype=np.int64)
        self.contact_island: dict[int, int] = {}     # linked contacts: ContactId -> island id
        self.contact_bodies: dict[int, tuple[int, int]] = {}
        self.split_candidate = None
        self.split_candidate_timer = np.float32(0.0)
        for b in range(self.B):            # every body with a SolverBody starts in its own island (BodyIslandNode::on_add, mod.rs:1327-1340)
            if self.kind[b] != STATIC:
                isl = self._create()
                isl.bodies.append(b)
                self.body_island[b] = isl.id
        self.joints = [(int(a), int(b)) for a, b in joints]
        for j, (a, b) in enumerate(self.joints):     # add_joint (mod.rs:669-747)
            if self.kind[a] == STATIC or self.kind[b] == STATIC:
                continue
            iid = self._merge(a, b)
            self.islands[iid].joints.add(j)

    # ---- slab of islands
    def _create(self) -> Island:
        if self.free_ids:
            iid = min(self.free_ids)
            self.free_ids.remove(iid)
        else:
            iid = self.next_new
            self.next_new += 1
        isl = Island(iid)
        self.islands[iid] = isl
        return isl

    def _remove(self, iid):
        if self.split_candidate == iid:          # remove_island (mod.rs:456-464)
            self.split_candidate = None
        del self.islands[iid]
        self.free_ids.append(iid)

    def _has_island(self, b) -> bool:
        return self.kind[b] != STATIC

    def _merge(self, b1, b2) -> int:              # merge_islands (mod.rs:814-993)
        if not self._has_island(b1):
            return int(self.body_island[b2])
        if not self._has_island(b2):
            return int(self.body_island[b1])
        i1, i2 = int(self.body_island[b1]), int(self.body_island[b2])
        if i1 == i2:
            return i1
        big, small = self.islands[i1], self.islands[i2]
        if len(big.bodies) < len(small.bodies):
            big, small = small, big
        for b in small.bodies:
            self.body_island[b] = big.id
        for c in small.contacts:
            self.contact_island[c] = big.id
        big.bodies.extend(small.bodies)
        big.contacts |= small.contacts
        big.joints |= small.joints
        big.removed += small.removed
        self._remove(small.id)
        return big.id

    # ---- one step -------------------------------------------------------------------------------------------------------------
    def step(self, events, lin_vel, ang_vel, delta_secs, wake=None):
        """events: list of (contact_id, kind, body1, body2) in ascending ContactId, kind 'add' (a constraint-generating pair started touching)
        or 'remove' (it stopped touching or left the contact graph).  Returns (island label per body = smallest body index of the island,
        sleeping flag per body)."""
        to_wake = []
        # -- NarrowPhase
        for cid, what, b1, b2 in sorted(events):
            if what == "remove":
                iid = self.contact_island.pop(cid, None)
                if iid is None:
                    continue
                isl = self.islands[iid]
                isl.contacts.discard(cid)
                isl.removed += 1
                self.contact_bodies.pop(cid, None)
            else:
                if not self._has_island(b1) and not self._has_island(b2):
                    continue
                iid = self._merge(b1, b2)
                self.islands[iid].contacts.add(cid)
                self.contact_island[cid] = iid
                self.contact_bodies[cid] = (b1, b2)
                if self.islands[iid].sleeping:
                    to_wake.append(iid)
        self._wake(to_wake)
        if wake is not None:                          # wake_on_changed: the application touched these bodies
            self._wake([int(self.body_island[b]) for b in np.nonzero(wake)[0] if self._has_island(b)])
        # -- Solver, SolverSystems::Finalize: split_island (mod.rs:161-179)
        if self.candidate_mode == "body":
            if self.candidate_body is not None:
                self._split(int(self.body_island[self.candidate_body]))
        elif self.split_candidate is not None and self.split_candidate in self.islands:
            self._split(self.split_candidate)
        # -- Sleeping: update_sleeping_states (sleeping.rs:185-246)
        S = self.S
        lu2 = S(self.length_unit * self.length_unit)
        awake = set()
        self.split_candidate_timer = np.float32(0.0)
        lv, av = np.asarray(lin_vel, dtype=S), np.asarray(ang_vel, dtype=S)
        for b in range(self.B):
            if not self._has_island(b):
                continue
            iid = int(self.body_island[b])
            isl = self.islands[iid]
            if isl.sleeping or self.disabled[b]:       # Without<Sleeping>, Without<SleepingDisabled>
                continue
            l2 = S(S(lv[b, 0] * lv[b, 0]) + S(lv[b, 1] * lv[b, 1])) + S(lv[b, 2] * lv[b, 2])
            a2 = S(S(av[b, 0] * av[b, 0]) + S(av[b, 1] * av[b, 1])) + S(av[b, 2] * av[b, 2])
            tl2 = np.float32(self.thr_lin[b] * np.abs(self.thr_lin[b]))
            ta2 = np.float32(self.thr_ang[b] * np.abs(self.thr_ang[b]))
            if S(l2) < S(lu2 * S(tl2)) and S(a2) < S(ta2):
                self.timer[b] = np.float32(self.timer[b] + np.float32(delta_secs))
            else:
                self.timer[b] = np.float32(0.0)
            if self.timer[b] < self.time_to_sleep:
                awake.add(iid)
            elif isl.removed > 0 and self.timer[b] > self.split_candidate_timer:
                self.split_candidate = iid
                self.candidate_body = b
                self.split_candidate_timer = self.timer[b]
        # wake_islands_with_sleeping_disabled (sleeping.rs:164-183)
        for b in np.nonzero(self.disabled)[0]:
            if self._has_island(b):
                awake.add(int(self.body_island[b]))
                self.timer[b] = np.float32(0.0)
        # sleep_islands (sleeping.rs:248-292)
        wake_buffer = []
        for iid, isl in list(self.islands.items()):
            if iid in awake:
                if isl.sleeping:
                    wake_buffer.append(iid)
            elif not isl.sleeping and isl.removed == 0:
                isl.sleeping = True
        self._wake(wake_buffer)
        return self.labels(), self.sleeping()

    def _wake(self, ids):
        for iid in ids:
            isl = self.islands.get(iid)
            if isl is None or not isl.sleeping:
                continue
            isl.sleeping = False
            for b in isl.bodies:
                self.timer[b] = np.float32(0.0)

    def _split(self, iid):                            # split_island (mod.rs:995-1270): DFS from every unvisited body of the island
        isl = self.islands[iid]
        if isl.sleeping or isl.removed == 0:
            return
        bodies = list(isl.bodies)
        contacts = {c: self.contact_bodies[c] for c in isl.contacts}
        joints = set(isl.joints)
        adj: dict[int, list[int]] = {b: [] for b in bodies}
        for c, (b1, b2) in contacts.items():
            if self._has_island(b1) and self._has_island(b2):
                adj[b1].append(b2); adj[b2].append(b1)
        for j in joints:
            a, b = self.joints[j]
            adj[a].append(b); adj[b].append(a)
        self._remove(iid)
        visited = set()
        for seed in bodies:
            if seed in visited:
                continue
            new = self._create()
            stack = [seed]
            visited.add(seed)
            while stack:
                b = stack.pop()
                new.bodies.append(b)
                self.body_island[b] = new.id
                for o in adj[b]:
                    if o not in visited:
                        visited.add(o)
                        stack.append(o)
            members = set(new.bodies)
            for c, (b1, b2) in contacts.items():
                x = b1 if self._has_island(b1) else b2
                if x in members:
                    new.contacts.add(c)
                    self.contact_island[c] = new.id
            for j in joints:
                if self.joints[j][0] in members:
                    new.joints.add(j)

    def labels(self) -> np.ndarray:
        out = np.full(self.B, 0xFFFFFFFF, dtype=np.uint32)
        for isl in self.islands.values():
            out[isl.bodies] = min(isl.bodies)
        return out

    def sleeping(self) -> np.ndarray:
        out = np.zeros(self.B, dtype=np.uint8)
        for isl in self.islands.values():
            if isl.sleeping:
                out[isl.bodies] = 1
        return out
