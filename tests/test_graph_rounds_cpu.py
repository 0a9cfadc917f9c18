"""The dependency-wavefront rule of the device ConstraintGraph (csrc/contacts.cu colour_rounds_kernel) as an executable model on the CPU:
processing the changed edges in ROUNDS — a push runs when it is the smallest pending edge on each of its non-static bodies, a pop when no
earlier push is pending on them — gives exactly the colours of the reference's sequential pass in ascending ContactId
(ConstraintGraph::push_manifold / pop_manifold, solver/constraint_graph.rs:163-296; the status loop of narrow_phase/system_param.rs:136-389).
The kernels themselves are checked on the GPU against the host fixture (tests/test_gpu_graph.py); this pins the RULE."""
import numpy as np

OVERFLOW, DYNAMIC_COLORS = 23, 20


def push_colour(bits, b1, b2, s1, s2):
    """the colour push_manifold picks, and the body-set update (constraint_graph.rs:163-238)"""
    if not s1 and not s2:
        for c in range(DYNAMIC_COLORS):
            if not (bits[b1] >> c) & 1 and not (bits[b2] >> c) & 1:
                bits[b1] |= 1 << c; bits[b2] |= 1 << c
                return c
        return OVERFLOW
    if s1 and s2:
        return OVERFLOW
    b = b2 if s1 else b1
    for c in range(OVERFLOW - 1, 0, -1):
        if not (bits[b] >> c) & 1:
            bits[b] |= 1 << c
            return c
    return OVERFLOW


def pop_colour(bits, b1, b2, s1, s2, c):
    if c != OVERFLOW:
        if not s1: bits[b1] &= ~(1 << c)
        if not s2: bits[b2] &= ~(1 << c)


def sequential(ops, bits, colour, static):
    for e, what, b1, b2 in ops:            # ascending ContactId
        if what == "push":
            colour[e] = push_colour(bits, b1, b2, static[b1], static[b2])
        else:
            pop_colour(bits, b1, b2, static[b1], static[b2], colour[e]); colour[e] = -1


def rounds(ops, bits, colour, static):
    pending = list(ops)
    n_rounds = 0
    while pending:
        n_rounds += 1
        min_any, min_push = {}, {}
        for e, what, b1, b2 in pending:
            for b in (b1, b2):
                if static[b]:
                    continue
                min_any[b] = min(min_any.get(b, 1 << 60), e)
                if what == "push":
                    min_push[b] = min(min_push.get(b, 1 << 60), e)
        ready = []
        for op in pending:
            e, what, b1, b2 = op
            if what == "push":
                ok = all(static[b] or min_any[b] == e for b in (b1, b2))
            else:
                ok = all(static[b] or min_push.get(b, 1 << 60) > e for b in (b1, b2))
            if ok:
                ready.append(op)
        assert ready, "the smallest pending edge can always run"
        # every edge of a round runs "at once": pushes of a round share no non-static body, pops only clear bits
        pushed, popped = set(), set()
        for e, what, b1, b2 in ready:
            for b in (b1, b2):
                if static[b]:
                    continue
                if what == "push":
                    assert b not in pushed, "two pushes of one round share a body"
                    pushed.add(b)
                else:
                    popped.add(b)
        assert not (pushed & popped), "a push and a pop of one round share a body"
        for e, what, b1, b2 in ready:
            if what == "push":
                colour[e] = push_colour(bits, b1, b2, static[b1], static[b2])
            else:
                pop_colour(bits, b1, b2, static[b1], static[b2], colour[e]); colour[e] = -1
        done = {op[0] for op in ready}
        pending = [op for op in pending if op[0] not in done]
    return n_rounds


def test_rounds_equal_the_sequential_pass():
    rng = np.random.default_rng(11)
    deepest = 0
    for trial in range(60):
        B = int(rng.integers(4, 40))
        static = rng.random(B) < 0.15
        static[0] = True
        E = int(rng.integers(5, 160))
        pairs = [(int(a), int(b)) for a, b in rng.integers(0, B, size=(E, 2)) if a != b and not (static[a] and static[b])]
        E = len(pairs)
        bits_s, bits_r = np.zeros(B, dtype=np.int64), np.zeros(B, dtype=np.int64)
        col_s, col_r = np.full(E, -1), np.full(E, -1)
        for step in range(6):            # several steps: edges flicker, the body sets persist between steps
            ops = []
            for e in range(E):
                if col_s[e] < 0 and rng.random() < 0.5:
                    ops.append((e, "push", *pairs[e]))
                elif col_s[e] >= 0 and rng.random() < 0.4:
                    ops.append((e, "pop", *pairs[e]))
            sequential(ops, bits_s, col_s, static)
            deepest = max(deepest, rounds(ops, bits_r, col_r, static))
            assert np.array_equal(col_s, col_r), (trial, step)
            assert np.array_equal(bits_s, bits_r), (trial, step)
    assert deepest > 3          # the scenes did contain dependency chains


def test_a_chain_in_ascending_id_is_sequential_and_pops_are_not():
    static = np.zeros(6, dtype=bool)
    bits, col = np.zeros(6, dtype=np.int64), np.full(5, -1)
    chain = [(e, "push", e, e + 1) for e in range(5)]            # 0-1, 1-2, 2-3, ...: every push waits for the one before
    assert rounds(chain, bits, col, static) == 5 and col.tolist() == [0, 1, 0, 1, 0]
    pops = [(e, "pop", e, e + 1) for e in range(5)]
    assert rounds(pops, bits, col, static) == 1 and not bits.any()
