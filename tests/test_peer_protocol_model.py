"""Exhaustive interleaving check of the peer-memory exchange protocol (csrc/kfbo_kernels.cu: peer_exchange; the
buffers and the evaluation number come from kfbo_api.cu: eval_launch).

The model: R ranks run E evaluations each, with NO barrier between evaluations.  Per evaluation e (from 1), generation
g = e & 1, a rank
  1. stores its share, tagged e, into slab [g][me] of every rank's exchange buffer (own first, then staggered),
  2. stores e into flag [g][me] of every rank               (st.release.sys: after its data, in program order),
  3. waits until flag [g][r] of its OWN buffer is >= e for every r   (ld.acquire.sys),
  4. reads slab [g][r] of its own buffer for every r and must find the tag e there -- a tag of another evaluation
     means a total formed from the wrong data.
Every store, flag and read is one atomic step; the checker walks ALL interleavings (breadth-first over the reachable
states) and also checks that no reachable state is stuck.  With two generations the protocol is safe; with one
generation the checker finds the overwrite (a rank one evaluation ahead stores into the slab a slower rank has not read
yet), which is what the second generation is for -- and shows that the checker can see such a bug at all.
"""
from collections import deque

import pytest


def _program(me, R, E, generations):
    """The straight-line program of one rank: a list of (op, target_rank, generation, source_rank, evaluation)."""
    prog = []
    for e in range(1, E + 1):
        g = e % generations
        order = [(me + i) % R for i in range(R)]
        prog += [("store", to, g, me, e) for to in order]
        prog += [("flag", to, g, me, e) for to in range(R)]
        prog += [("wait", me, g, None, e)]
        prog += [("read", me, g, r, e) for r in range(R)]
    return prog


def _explore(R, E, generations):
    """Returns (violation or None, number of states).  State = (pcs, slabs, flags); slabs / flags are tuples indexed
    [owner rank][generation][source rank]."""
    progs = [_program(r, R, E, generations) for r in range(R)]
    zero = tuple(tuple(tuple(0 for _ in range(R)) for _ in range(generations)) for _ in range(R))
    start = (tuple(0 for _ in range(R)), zero, zero)
    seen, todo = {start}, deque([start])

    def put(mem, owner, g, src, val):
        m = [list(map(list, x)) for x in mem]
        m[owner][g][src] = val
        return tuple(tuple(tuple(y) for y in x) for x in m)

    while todo:
        pcs, slabs, flags = todo.popleft()
        moved = False
        for r in range(R):
            if pcs[r] == len(progs[r]):
                continue
            op, tgt, g, src, e = progs[r][pcs[r]]
            nslabs, nflags = slabs, flags
            if op == "store":
                nslabs = put(slabs, tgt, g, src, e)
            elif op == "flag":
                nflags = put(flags, tgt, g, src, e)
            elif op == "wait":
                if any(flags[r][g][q] < e for q in range(R)):
                    continue                                   # blocked
            elif op == "read":
                if slabs[r][g][src] != e:
                    return (f"rank {r}, evaluation {e}: slab of rank {src} holds evaluation {slabs[r][g][src]}", len(seen))
            moved = True
            nxt = (pcs[:r] + (pcs[r] + 1,) + pcs[r + 1:], nslabs, nflags)
            if nxt not in seen:
                seen.add(nxt)
                todo.append(nxt)
        if not moved and any(pcs[r] < len(progs[r]) for r in range(R)):
            return (f"stuck at program counters {pcs}", len(seen))
    return (None, len(seen))


@pytest.mark.parametrize("R,E", [(2, 5), (3, 4)])
def test_two_generations_are_safe_in_every_interleaving(R, E):
    violation, states = _explore(R, E, generations=2)
    assert violation is None, violation
    assert states > 300                                         # the walk did cover interleavings, not one schedule


def test_one_generation_is_not(capsys):
    violation, _ = _explore(2, 3, generations=1)
    assert violation is not None and "holds evaluation" in violation


def test_a_rank_is_never_two_evaluations_ahead():
    # the argument the second generation rests on: while rank a is in evaluation e, no rank has started evaluation e + 2
    R, E = 3, 4
    progs = [_program(r, R, E, 2) for r in range(R)]
    per_eval = len(progs[0]) // E
    zero = tuple(tuple(tuple(0 for _ in range(R)) for _ in range(2)) for _ in range(R))
    start = (tuple(0 for _ in range(R)), zero)                  # flags only: stores and reads do not gate progress
    seen, todo, worst = {start}, deque([start]), 0
    while todo:
        pcs, flags = todo.popleft()
        evals = [min(pc // per_eval, E - 1) for pc in pcs]
        active = [i for i in range(R) if pcs[i] < len(progs[i])]
        if active:
            worst = max(worst, max(evals) - min(evals[i] for i in active))
        for r in active:
            op, tgt, g, src, e = progs[r][pcs[r]]
            nflags = flags
            if op == "flag":
                m = [list(map(list, x)) for x in flags]
                m[tgt][g][src] = e
                nflags = tuple(tuple(tuple(y) for y in x) for x in m)
            elif op == "wait" and any(flags[r][g][q] < e for q in range(R)):
                continue
            nxt = (pcs[:r] + (pcs[r] + 1,) + pcs[r + 1:], nflags)
            if nxt not in seen:
                seen.add(nxt)
                todo.append(nxt)
    assert worst == 1
