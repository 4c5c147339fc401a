"""The flag-free exchange protocol (mixq_b200/csrc/mixq_kernels.cu: exchange_finish_poll_kernel / exchange_finish_rowquant_kernel,
DESIGN.md section 6) as a small state machine, run under thousands of random interleavings of the ranks.

What the kernels rely on — and what this model checks for 2, 4 and 8 ranks, one-shot and two-phase:
  * SAFETY  a store into a receive slot or a result buffer always finds the sentinel there (nobody overwrites data that has not
            been consumed yet), a reader that finds data finds the data of ITS exchange, the residual it reads is the previous
            exchange's result;
  * LIVENESS whatever the skew between the ranks, somebody can always make progress and every rank finishes every exchange.
The only ordering the model grants is what the hardware grants: program order inside a rank (GEMM pushes of exchange e, then the
finish kernel of e, then the GEMM of e + 1 — one stream) and nothing at all between ranks.  Vectors of a buffer are independent
cells (V per region) visited in random order, like the threads of the kernels."""
import random

import pytest

ARMED = None
V = 2   # independent 16-byte vectors modelled per (buffer, slice) region


class Violation(AssertionError):
    pass


def rank_program(me, world, n_exchanges, one_shot, cells, rng, nbuf=2):
    """Generator of atomic steps: yields ("wait", table, cell) when it finds the sentinel where it needs data, ("step",) after one
    store / consume."""
    slot, res = cells

    def store(cell_list, key, e, what):
        if cell_list[key] is not ARMED:
            raise Violation(f"rank {me} {what} of exchange {e} overwrites {cell_list[key]} at {key}")
        cell_list[key] = e

    for e in range(n_exchanges):
        b, ob = e % nbuf, (e + 1) % nbuf          # this exchange's buffer set, and the other one (re-armed here)
        # ---- the row-parallel GEMM: its epilogue pushes this rank's partial (slice j -> rank j; one-shot: everything -> everybody)
        pushes = [(j, v) for j in range(world) for v in range(V)]
        rng.shuffle(pushes)
        for j, v in pushes:
            store(slot, (j, b, me, v), e, "push")
            yield ("step",)
        # ---- the finish kernel
        order = [(s, v) for s in range(world) for v in range(V)]
        rng.shuffle(order)
        for s, v in order:                                   # own slice: every rank's partial, consumed and re-armed
            key = (me, b, s, v)
            while slot[key] is ARMED:
                yield ("wait", "slot", key)
            if slot[key] != e:
                raise Violation(f"rank {me} exchange {e}: slot {key} holds exchange {slot[key]}")
            slot[key] = ARMED
            yield ("step",)
        for v in range(V):                                   # residual = own slice of the previous result, then re-arm it
            key = (me, ob, me, v)
            if e > 0 and res[key] != e - 1:
                raise Violation(f"rank {me} exchange {e}: residual {key} holds {res[key]}")
            res[key] = ARMED
            yield ("step",)
        targets = [(j, v) for j in (range(world) if not one_shot else [me]) for v in range(V)]
        rng.shuffle(targets)
        for j, v in targets:                                 # broadcast of the reduced slice (one-shot: local result only)
            store(res, (j, b, me, v), e, "broadcast")
            yield ("step",)
        if one_shot:
            # every "slice" of the local result is written locally: model the other slices as written by this rank too
            for j in range(world):
                for v in range(V):
                    if j != me:
                        res[(me, ob, j, v)] = ARMED
                        store(res, (me, b, j, v), e, "local result")
                        yield ("step",)
            continue
        others = [(j, v) for j in range(world) if j != me for v in range(V)]
        rng.shuffle(others)
        for j, v in others:                                  # re-arm the dead buffer, collect the owner's slice
            res[(me, ob, j, v)] = ARMED
            key = (me, b, j, v)
            while res[key] is ARMED:
                yield ("wait", "res", key)
            if res[key] != e:
                raise Violation(f"rank {me} exchange {e}: result {key} holds exchange {res[key]}")
            yield ("step",)


def run_once(world, one_shot, seed, n_exchanges=6, nbuf=2):
    rng = random.Random(seed)
    slot = {(r, b, s, v): ARMED for r in range(world) for b in range(2) for s in range(world) for v in range(V)}
    res = {(r, b, j, v): ARMED for r in range(world) for b in range(2) for j in range(world) for v in range(V)}
    progs = [rank_program(r, world, n_exchanges, one_shot, (slot, res), random.Random(seed * 131 + r), nbuf) for r in range(world)]
    blocked = [None] * world         # the cell a rank waits for, or None
    done = [False] * world
    # a skewed scheduler: some ranks are much "faster" than others, and the skew changes over time
    weights = [rng.random() ** 3 + 0.01 for _ in range(world)]
    steps = 0
    while not all(done):
        runnable = []
        for r in range(world):
            if done[r]:
                continue
            if blocked[r] is not None:
                table = slot if blocked[r][0] == "slot" else res
                if table[blocked[r][1]] is ARMED:
                    continue
                blocked[r] = None
            runnable.append(r)
        if not runnable:
            raise Violation(f"deadlock: world {world}, one_shot {one_shot}, seed {seed}, waiting {blocked}")
        r = rng.choices(runnable, weights=[weights[x] for x in runnable])[0]
        try:
            ev = next(progs[r])
        except StopIteration:
            done[r] = True
            continue
        if ev[0] == "wait":
            blocked[r] = (ev[1], ev[2])
        steps += 1
        if steps % 97 == 0:
            weights = [rng.random() ** 3 + 0.01 for _ in range(world)]
    # everything consumed: all receive slots are armed again
    assert all(v is ARMED for v in slot.values())


@pytest.mark.parametrize("world", [2, 4, 8])
@pytest.mark.parametrize("one_shot", [False, True])
def test_flag_free_exchange_is_safe_and_live(world, one_shot):
    for seed in range(120 if world < 8 else 40):
        run_once(world, one_shot, seed)


def test_model_detects_a_broken_protocol():
    """The model is not vacuous: with ONE buffer set instead of two alternating ones a fast peer's broadcast of exchange e + 1
    can land in a result buffer that still holds exchange e — the checker must find such a schedule."""
    hits = 0
    for seed in range(200):
        try:
            run_once(4, False, seed, nbuf=1)
        except Violation:
            hits += 1
    assert hits > 0
