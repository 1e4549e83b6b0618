"""Property tests of the C++ scheduler (hypothesis): for arbitrary set-ups the event stream must be a valid
gossip history -- conservation of messages, causality, single delivery, token accounting -- and a pure function
of the seed (replicated on every rank of a multi-GPU run) that survives a state round-trip at any round."""
import numpy as np
import pytest
from hypothesis import HealthCheck, given, settings, strategies as st

from gossipy_b200.ops.native import native_available

pytestmark = pytest.mark.skipif(not native_available(), reason="extension not built")


def _make(C, cfg):
    n, delta, proto, drop, online, samp, seed, sync, delay, token, topo = cfg
    s = C.GossipScheduler(n, delta, proto, drop, online, samp, seed)
    rng = np.random.default_rng(seed)
    if sync:
        s.set_nodes([1] * n, [int(v) for v in rng.integers(0, delta, n)], [delta] * n)
    else:
        s.set_nodes([0] * n, [int(v) for v in rng.integers(1, 2 * delta, n)], [delta] * n)
    if topo == "ring":
        s.set_topology(list(range(n + 1)), [(i + 1) % n for i in range(n)])
    elif topo == "sparse":                                  # some nodes without peers (B6: must be skipped, not abort)
        indptr, idx = [0], []
        for i in range(n):
            peers = [j for j in range(n) if j != i and (i + j) % 3 == 0]
            idx += peers
            indptr.append(len(idx))
        s.set_topology(indptr, idx)
    s.set_delay(*delay)
    s.set_message_sizes(100, 1)
    if token:
        s.set_token_account(*token)
    return s


configs = st.tuples(
    st.integers(2, 24), st.integers(2, 12), st.sampled_from([1, 2, 3]), st.sampled_from([0.0, 0.1, 0.5]),
    st.sampled_from([1.0, 0.7, 0.3]), st.sampled_from([0.0, 0.3]), st.integers(0, 2 ** 31), st.booleans(),
    st.sampled_from([(0, 0.0, 0.0), (0, 2.0, 0.0), (1, 0.0, 4.0), (2, 0.01, 1.0)]),
    st.sampled_from([None, (1, 1, 1, 1, 1), (2, 1, 1, 2, 1), (3, 2, 1, 1, 1), (4, 4, 2, 1, 1), (5, 6, 3, 1, 1)]),
    st.sampled_from(["clique", "ring", "sparse"]))


@settings(max_examples=int(__import__("os").environ.get("SCHED_EXAMPLES", "60")), deadline=None, derandomize=True,
          suppress_health_check=[HealthCheck.too_slow], database=None)
@given(cfg=configs, rounds=st.integers(1, 6), cut=st.integers(0, 5))
def test_event_stream_is_a_valid_history(cfg, rounds, cut):
    from gossipy_b200.ops.native import _try_import
    C = _try_import()
    n, delta, proto = cfg[0], cfg[1], cfg[2]
    a = _make(C, cfg)
    per_round = [a.run(1) for _ in range(rounds)]
    ev = np.concatenate(per_round) if per_round else np.zeros((0, 6), np.int32)
    # determinism: a second scheduler with the same seed, and one restored from a mid-run state, agree
    b = _make(C, cfg)
    assert np.array_equal(b.run(rounds), ev)
    c = _make(C, cfg)
    k = min(cut, rounds)
    head = [c.run(1) for _ in range(k)]
    d = _make(C, cfg)
    d.set_state(dict(c.get_state()))
    tail = [d.run(1) for _ in range(rounds - k)]
    assert np.array_equal(np.concatenate(head + tail) if head + tail else ev, ev)
    assert (d.sent, d.failed, d.total_size, d.clock, d.pending) == (a.sent, a.failed, a.total_size, a.clock, a.pending)

    kind, tick, ea, eb, slot, aux = (ev[:, i] for i in range(6))
    assert (np.diff(tick) >= 0).all() and (tick >= 0).all() and (tick < rounds * delta).all()
    born, alive, msg_type = {}, set(), {}
    replies_of = {}
    sent = failed = size = 0
    for kd, t, x, y, sl, ax in ev.tolist():
        if kd == C.EV_SEND:
            assert sl not in born and x != y and 0 <= x < n and 0 <= y < n
            assert ax == {1: 1, 2: 2, 3: 4}[proto]                     # PUSH / PULL / PUSH_PULL message types
            born[sl], msg_type[sl] = t, ax
            alive.add(sl)
            sent += 1
            size += 1 if ax == 2 else 100
        elif kd == C.EV_DROP:
            assert sl in alive
            alive.discard(sl)
            failed += 1
        elif kd == C.EV_DELIVER:
            assert sl in alive and born[sl] <= t
            alive.discard(sl)
        elif kd == C.EV_REPLY_SEND:
            assert msg_type[sl] in (2, 4) and ax not in born            # only PULL / PUSH_PULL requests are answered
            born[ax], msg_type[ax] = t, 3
            alive.add(ax)
            replies_of[sl] = ax
        elif kd == C.EV_REPLY_DELIVER:
            assert sl in alive and msg_type[sl] == 3 and born[sl] <= t
            alive.discard(sl)
            sent += 1
            size += 100
        elif kd == C.EV_EVAL:
            assert 0 <= x < n and (t + 1) % delta == 0
    assert (a.sent, a.failed, a.total_size) == (sent, failed, size)
    assert a.pending == len(alive)                                       # everything else is still on the wire
    assert all(v >= 0 for v in a.token_balances())
    if proto == 1:
        assert not replies_of
    n_eval = int((kind == C.EV_EVAL).sum())
    expect = rounds * (n if cfg[5] == 0 else max(int(n * cfg[5]), 1))
    assert n_eval == expect


def test_sparse_tick_loop_produces_the_dense_loops_events():
    """The default loop visits only the nodes that time out at a tick and evaluates availability draws on demand; the dense
    loop (every node, every tick -- what the reference does) must give the same events, counters and stream positions:
    random mixes of sync / async nodes, round lengths, topologies, faults, token accounts, broadcasts, pieces, restores."""
    import random
    from gossipy_b200.ops.native import _try_import
    C = _try_import()
    rnd = random.Random(7)
    for case in range(60):
        n = rnd.randint(2, 40)
        delta = rnd.choice([5, 10, 10, 25])
        proto = rnd.choice([1, 2, 3])
        drop, online = rnd.choice([0.0, 0.2]), rnd.choice([1.0, 1.0, 0.7, 0.2])
        seed = rnd.randint(0, 10 ** 9)
        sync = [rnd.random() < .7 for _ in range(n)]
        rl = [rnd.choice([delta, delta, 2 * delta, 3, delta + 1]) for _ in range(n)]
        off = [rnd.randint(0, rl[i] + 1) if sync[i] else rnd.randint(0, 2 * delta) for i in range(n)]
        ring = rnd.random() < .3
        tok = rnd.choice([0, 0, 3, 5, 2])
        bcast = rnd.random() < .2 and proto == 1
        delay = rnd.choice([(0, 0, 0), (1, 0, 2 * delta), (2, 0.01, 1)])

        def make(dense):
            s = C.GossipScheduler(n, delta, proto, drop, online, rnd_eval, seed)
            s.set_nodes([int(v) for v in sync], off, rl)
            if ring:
                s.set_topology(list(range(n + 1)), [(i + 1) % n for i in range(n)])
            s.set_delay(*delay)
            s.set_message_sizes(50, 1)
            if tok:
                s.set_token_account(tok, 6, 3, 1, 1)
            if bcast:
                s.set_broadcast(True)
            s.set_dense_loop(dense)
            return s
        rnd_eval = rnd.choice([0.0, 0.3])
        a, b = make(True), make(False)
        pieces = [rnd.randint(0, 3 * delta) for _ in range(6)]
        ev_a = [a.run_ticks(k) for k in pieces]
        ev_b = []
        for j, k in enumerate(pieces):
            if j == 3:                      # restore in the middle: the timeout queue is rebuilt from the clock
                st = dict(b.get_state())
                b = make(False)
                b.set_state(st)
            ev_b.append(b.run_ticks(k))
        for x, y in zip(ev_a, ev_b):
            assert np.array_equal(x, y), (case, n, delta, proto, sync, rl, off)
        sa, sb = dict(a.get_state()), dict(b.get_state())
        assert sa["streams"] == sb["streams"] and sa["order"] == sb["order"] and sa["msg_q"] == sb["msg_q"] and sa["rep_q"] == sb["rep_q"]
        assert (a.sent, a.failed, a.total_size, a.clock) == (b.sent, b.failed, b.total_size, b.clock)
        assert a.token_balances() == b.token_balances()
