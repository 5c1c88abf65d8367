"""Exhaustive interleaving checks of the engine's lock-free protocols (CPU, no GPU needed).

The device code orders its accesses with sequentially-consistent fences exactly where these protocols need
it (flag store / try-lock / unlock / re-check are store-buffering patterns), so a sequentially-consistent
model with one shared-memory access per step is the right abstraction.  Each protocol is written as a small
state machine per thread; the explorer visits EVERY reachable interleaving (memoised on the full state) and
checks the invariants in every state and the liveness conditions in every terminal state:

  * in-order retirement under ``retire_word = (head << 1) | locked``  (csrc/hca/engine.cuh, retire())
  * the shared-SQ doorbell hand-off: ready flags + try-lock            (csrc/hca/post.cuh, sq_submit_shared())
  * batch claims on the claim head, bounded by the doorbell            (csrc/hca/engine.cuh, try_claim())
"""
from __future__ import annotations


def explore(init_shared: tuple, init_threads: tuple, step, invariant, terminal_check, limit=2_000_000):
    """step(shared, tstate) -> (shared', tstate') for ONE thread; tstate None = thread finished."""
    seen = set()
    stack = [(init_shared, init_threads)]
    terminals = 0
    while stack:
        st = stack.pop()
        if st in seen:
            continue
        seen.add(st)
        assert len(seen) < limit, "state space larger than expected"
        shared, threads = st
        invariant(shared, threads)
        live = [i for i, t in enumerate(threads) if t is not None]
        if not live:
            terminal_check(shared)
            terminals += 1
            continue
        for i in live:
            s2, t2 = step(shared, threads[i])
            stack.append((s2, threads[:i] + (t2,) + threads[i + 1:]))
    return len(seen), terminals


# ------------------------------------------------------------------------------------------ retirement
FIN = 2
RETIRE_BATCH = 2          # the real warp takes 32; 2 exercises the "full batch, go again" loop with 3-4 WQEs


def _retire_step(shared, t):
    """shared = (states tuple, retire_word, retired tuple); t = (pc, w, h, scan_i, n)"""
    states, word, retired = shared
    pc, w, h, i, n = t
    N = len(states)
    if pc == "finish":                                  # state store (+ SC fence), then the lock attempt
        states = states[:w] + (FIN,) + states[w + 1:]
        return (states, word, retired), ("spec", w, w, 0, 0)
    if pc in ("spec", "scan"):                          # lanes read state[h + i] one by one (worst case for the model);
        nxt = "cas" if pc == "spec" else "publish"      # "spec": the loads travel WITH the lock attempt, i.e. before it is known
        if i < RETIRE_BATCH and h + i < N and n == i and states[h + i] == FIN:
            return shared, (pc, w, h, i + 1, n + 1)
        return shared, (nxt, w, h, 0, n)
    if pc == "cas":                                     # one CAS: lock AND "is it this WQE's turn"
        if word == (h << 1):
            return (states, (h << 1) | 1, retired), ("publish", w, h, 0, n)
        return shared, None                             # a holder exists or an earlier WQE is unfinished
    if pc == "publish":                                 # CQEs of the run, in order
        assert all(states[k] == FIN for k in range(h, h + n))
        retired = retired + tuple(range(h, h + n))
        if n == RETIRE_BATCH:
            return (states, word, retired), ("scan", w, h + n, 0, 0)
        return (states, word, retired), ("unlock", w, h + n, 0, 0)
    if pc == "unlock":                                  # one store publishes head and unlock (+ SC fence)
        return (states, h << 1, retired), ("recheck", w, h, 0, 0)
    if pc == "recheck":
        if h < N and states[h] == FIN:
            return shared, ("spec", w, h, 0, 0)
        return shared, None
    raise AssertionError(pc)


def _check_retire(n_wqes, order):
    def invariant(shared, threads):
        retired = shared[2]
        assert retired == tuple(range(len(retired))), f"out of order or duplicate: {retired}"
        holders = sum(1 for t in threads if t is not None and t[0] in ("scan", "publish", "unlock"))   # "spec" runs before the lock
        assert holders <= 1 and (shared[1] & 1) == (1 if holders else 0)

    def terminal(shared):
        assert shared[2] == tuple(range(n_wqes)), f"lost wake-up: retired {shared[2]} of {n_wqes}"
        assert shared[1] == n_wqes << 1

    threads = tuple(("finish", w, 0, 0, 0) for w in order)
    return explore(((0,) * n_wqes, 0, ()), threads, _retire_step, invariant, terminal)


def test_retire_word_protocol_retires_everything_once_and_in_order():
    states, terminals = _check_retire(3, (0, 1, 2))
    assert states > 100 and terminals >= 1
    _check_retire(4, (3, 1, 0, 2))          # thread identity does not matter, only the interleaving does


def test_retire_model_detects_a_missing_recheck():
    """Sanity of the checker itself: without the re-check after unlock a finished WQE can be stranded."""
    def broken_step(shared, t):
        if t[0] == "recheck":
            return shared, None
        return _retire_step(shared, t)

    def terminal(shared):
        assert shared[2] == (0, 1, 2)

    threads = tuple(("finish", w, 0, 0, 0) for w in range(3))
    try:
        explore(((0,) * 3, 0, ()), threads, broken_step, lambda s, t: None, terminal)
    except AssertionError:
        return
    raise AssertionError("the model failed to find the lost wake-up")


# ------------------------------------------------------------------------------------------ shared doorbell
def _doorbell_step(shared, t):
    """shared = (resv, flags tuple, lock, ready_head, doorbell); t = (pc, idx, h, to)"""
    resv, flags, lock, ready, db = shared
    pc, idx, h, to = t
    N = len(flags)
    if pc == "reserve":
        return (resv + 1, flags, lock, ready, db), ("flag", resv, 0, 0)
    if pc == "flag":                                   # WQE bytes + SC fence, then the generation-tagged ready flag
        flags = flags[:idx] + (idx + 1,) + flags[idx + 1:]
        return (resv, flags, lock, ready, db), ("trylock", idx, 0, 0)
    if pc == "trylock":
        if lock == 0:
            return (resv, flags, 1, ready, db), ("readhead", idx, 0, 0)
        return shared, None                            # the holder's re-check or a later poster covers this WQE
    if pc == "readhead":
        return shared, ("scan", idx, ready, ready)
    if pc == "scan":
        if to < N and flags[to] == to + 1:
            return shared, ("scan", idx, h, to + 1)
        return shared, ("ring", idx, h, to)
    if pc == "ring":
        if to > h:
            assert to > db, "doorbell must only move forward"
            return (resv, flags, lock, to, to), ("unlock", idx, h, to)
        return shared, ("unlock", idx, h, to)
    if pc == "unlock":
        return (resv, flags, 0, ready, db), ("recheck", idx, h, to)
    if pc == "recheck":
        if to < N and flags[to] == to + 1:
            return shared, ("trylock", idx, 0, 0)
        return shared, None
    raise AssertionError(pc)


def test_shared_doorbell_protocol_never_loses_a_posted_wqe():
    n = 3

    def invariant(shared, threads):
        resv, flags, lock, ready, db = shared
        assert db == ready and db <= resv
        assert all(flags[i] == i + 1 for i in range(db)), "doorbell covers a WQE whose bytes are not complete"

    def terminal(shared):
        assert shared[4] == n, f"doorbell stuck at {shared[4]} of {n}"
        assert shared[2] == 0

    states, terminals = explore((0, (0,) * n, 0, 0, 0), tuple(("reserve", 0, 0, 0) for _ in range(n)), _doorbell_step, invariant, terminal)
    assert states > 100 and terminals >= 1


# ------------------------------------------------------------------------------------------ batch claims
CLAIM_BATCH = 2


def _claim_step(shared, t):
    """shared = (doorbell, cursor, claimed tuple); poster thread t = ('post', k); engine CTA t = (pc, c, pending, tries)"""
    db, cursor, claimed = shared
    if t[0] == "post":                                  # the poster rings the doorbell twice: 2 then 2 more WQEs
        k = t[1]
        return (db + 2, cursor, claimed), (("post", k - 1) if k > 1 else None)
    pc, c, pending, tries = t
    if pc == "load":                                    # cursor and doorbell loaded one round trip before the CAS
        return shared, ("cas", cursor, db - cursor, tries)
    if pc == "cas":
        if pending <= 0:
            return shared, (("load", 0, 0, tries - 1) if tries > 1 else None)
        take = min(pending, CLAIM_BATCH)
        if cursor == c:
            assert c + take <= db, "claimed past the doorbell"
            return (db, c + take, claimed + tuple(range(c, c + take))), (("load", 0, 0, tries - 1) if tries > 1 else None)
        return shared, (("load", 0, 0, tries - 1) if tries > 1 else None)
    raise AssertionError(pc)


def test_batch_claims_hand_out_every_wqe_exactly_once_and_never_pass_the_doorbell():
    def invariant(shared, threads):
        db, cursor, claimed = shared
        assert cursor <= db
        assert claimed == tuple(range(cursor)), f"hole or duplicate in claims: {claimed}"

    states, terminals = explore((0, 0, ()), (("post", 2), ("load", 0, 0, 3), ("load", 0, 0, 3), ("load", 0, 0, 3)),
                                _claim_step, invariant, lambda shared: None)
    assert states > 50 and terminals >= 1
