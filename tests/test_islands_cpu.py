"""The CPU restatement of the reference's persistent islands + sleeping (oracle/islands_oracle.py) pinned by hand-checkable scenarios:
what dynamics/solver/islands/mod.rs and islands/sleeping.rs prescribe for merges, deferred splits, sleeping and waking."""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "oracle"))
from islands_oracle import IslandsOracle  # noqa: E402

DYN, STATIC = 0, 2


def _still(n):
    return np.zeros((n, 3), dtype=np.float32), np.zeros((n, 3), dtype=np.float32)


def test_merge_deferred_split_and_sleep():
    kind = np.array([STATIC, DYN, DYN, DYN, DYN], dtype=np.uint8)
    o = IslandsOracle(kind, time_to_sleep=0.05)
    lv, av = _still(5)
    lv[:] = 1.0                                     # everything moves: nobody gets sleepy
    lab, slp = o.step([(0, "add", 1, 2), (1, "add", 2, 3), (2, "add", 0, 4)], lv, av, 1 / 60)
    assert lab.tolist() == [0xFFFFFFFF, 1, 1, 1, 4] and not slp.any()      # static bodies have no island and do not merge islands
    # the contact 2-3 goes: the island is only MARKED (constraints_removed), not split (mod.rs:594-667)
    lab, _ = o.step([(1, "remove", 2, 3)], lv, av, 1 / 60)
    assert lab.tolist() == [0xFFFFFFFF, 1, 1, 1, 4]
    # now the bodies rest: after time_to_sleep the sleepiest body's island becomes the split candidate (sleeping.rs:229-238) ...
    lv[:] = 0.0
    for _ in range(3):
        lab, slp = o.step([], lv, av, 1 / 60)
    assert lab.tolist() == [0xFFFFFFFF, 1, 1, 1, 4]
    assert slp.tolist() == [0, 0, 0, 0, 1]         # the untouched island sleeps, the marked one must be split first (sleeping.rs:262)
    # ... it is split in the next step's Finalize (mod.rs:161-179) and both halves may sleep
    lab, slp = o.step([], lv, av, 1 / 60)
    assert lab.tolist() == [0xFFFFFFFF, 1, 1, 3, 4]
    assert slp.tolist() == [0, 1, 1, 1, 1]
    # a new contact that reaches a sleeping island wakes it and resets its timers (system_param.rs:253-258, WakeIslands)
    lab, slp = o.step([(5, "add", 3, 4)], lv, av, 1 / 60)
    assert lab.tolist() == [0xFFFFFFFF, 1, 1, 3, 3]
    assert slp.tolist() == [0, 1, 1, 0, 0] and o.timer[3] == np.float32(1 / 60) and o.timer[1] > np.float32(0.05)


def test_sleeping_disabled_and_thresholds():
    kind = np.array([DYN, DYN, DYN], dtype=np.uint8)
    o = IslandsOracle(kind, disabled=[0, 0, 1], thr_lin=[0.15, -1.0, 0.15], time_to_sleep=0.03)
    lv, av = _still(3)
    for _ in range(4):
        _, slp = o.step([], lv, av, 1 / 60)
    # body 1 has a negative threshold (never sleeps: "keep signs", sleeping.rs:209-211), body 2 is SleepingDisabled
    assert slp.tolist() == [1, 0, 0] and o.timer[1] == 0 and o.timer[2] == 0


def test_joints_hold_an_island_together():
    kind = np.array([DYN, DYN, DYN], dtype=np.uint8)
    o = IslandsOracle(kind, joints=[(0, 1)], time_to_sleep=0.03)
    lv, av = _still(3)
    lab, _ = o.step([(0, "add", 1, 2)], lv, av, 1 / 60)
    assert lab.tolist() == [0, 0, 0]
    lab, _ = o.step([(0, "remove", 1, 2)], lv, av, 1 / 60)
    for _ in range(4):
        lab, slp = o.step([], lv, av, 1 / 60)
    assert lab.tolist() == [0, 0, 2] and slp.all()      # the split keeps the jointed pair together


def test_the_stated_deviation_of_the_device_candidate_rule():
    """The one place where the device's islands differ from the reference (include/avian_b200.h avn_islands_step): the reference's split
    candidate is an island ID that merge_islands retires when the candidate is the SMALLER side of a merge (remove_island, mod.rs:456-464);
    the device remembers the sleepiest BODY and splits whatever island holds it a step later.  Here: {1,2} lost a contact and rests (candidate);
    in the next step it is merged into the bigger moving island {3,4,5}.  Reference: no split that step (the candidate is gone), the merged island
    is split a step later.  Device rule: the merged island is split at once.  One step later both agree again."""
    kind = np.array([STATIC, DYN, DYN, DYN, DYN, DYN], dtype=np.uint8)
    lv, av = _still(6)
    res = {}
    for mode in ("island", "body"):
        o = IslandsOracle(kind, time_to_sleep=0.045, candidate=mode)
        fast = lv.copy(); fast[3:] = 1.0                      # bodies 3, 4, 5 keep moving: their island never gets sleepy
        o.step([(0, "add", 1, 2), (1, "add", 3, 4), (2, "add", 4, 5)], fast, av, 1 / 60)
        o.step([(0, "remove", 1, 2)], fast, av, 1 / 60)      # {1,2} is marked (constraints_removed = 1) and keeps resting
        lab, _ = o.step([], fast, av, 1 / 60)                # timers of 1 and 2 reach time_to_sleep: {1,2} becomes the split candidate
        assert lab.tolist() == [0xFFFFFFFF, 1, 1, 3, 3, 3]
        a, _ = o.step([(7, "add", 2, 3)], fast, av, 1 / 60)  # merged into the bigger island in the narrow phase of the next step
        b, _ = o.step([], fast, av, 1 / 60)
        c, _ = o.step([], fast, av, 1 / 60)
        res[mode] = (a.tolist(), b.tolist(), c.tolist())
    # reference: the candidate was retired by the merge, {1,2,3,4,5} stays whole for that step
    assert res["island"][0] == [0xFFFFFFFF, 1, 1, 1, 1, 1]
    # device rule: body 1 (the sleepiest) is still the candidate: its island is split right away — 1 is alone, 2-3-4-5 hang together
    assert res["body"][0] == [0xFFFFFFFF, 1, 2, 2, 2, 2]
    # afterwards the reference picks the merged island (it still carries constraints_removed) and splits it too
    assert res["island"][2] == res["body"][2] == [0xFFFFFFFF, 1, 2, 2, 2, 2]
