"""The C++ scheduler (csrc/sched) against the Python round loop: same semantics, different RNG streams,
so deterministic set-ups must agree exactly and randomised ones statistically."""
import os

import numpy as np
import pytest
import torch

from gossipy_b200.ops.native import native_available

pytestmark = pytest.mark.skipif(not native_available(), reason="extension not built")


def _sim(engine, protocol, n=6, ring=False, drop=0., online=1., delay=None, rounds=4, tokenized=None,
         all2all=False, sampling_eval=0.):
    import gossipy_b200 as g
    from gossipy_b200.core import (AntiEntropyProtocol, ConstantDelay, CreateModelMode, StaticP2PNetwork,
                                   UniformMixing)
    from gossipy_b200.data import DataDispatcher, synthetic
    from gossipy_b200.data.handler import ClassificationDataHandler
    from gossipy_b200.model.handler import TorchModelHandler, WeightedTMH
    from gossipy_b200.model.nn import LogisticRegression
    from gossipy_b200.node import All2AllGossipNode, GossipNode
    from gossipy_b200.simul import (All2AllGossipSimulator, GossipSimulator, SimulationReport,
                                    TokenizedGossipSimulator)
    g.set_seed(3)
    (Xtr, ytr), (Xte, yte) = synthetic.spambase_like(600, 200)
    disp = DataDispatcher(ClassificationDataHandler(Xtr, ytr, Xte, yte), n=n, eval_on_user=False)
    topo = None
    if ring:   # directed ring: every node has exactly ONE peer -> peer choice is deterministic
        topo = np.zeros((n, n), dtype=int)
        for i in range(n):
            topo[i, (i + 1) % n] = 1
    net = StaticP2PNetwork(n, topo)
    cls = WeightedTMH if all2all else TorchModelHandler
    proto = cls(LogisticRegression(57, 2), torch.optim.SGD, {"lr": .5}, torch.nn.CrossEntropyLoss(),
                batch_size=16)
    node_cls = All2AllGossipNode if all2all else GossipNode
    nodes = node_cls.generate(disp, net, proto, 10, True)
    kw = dict(drop_prob=drop, online_prob=online, delay=delay or ConstantDelay(0), sampling_eval=sampling_eval)
    if tokenized is not None:
        sim = TokenizedGossipSimulator(nodes, disp, tokenized, lambda a, b, m: 1, 10, protocol, **kw)
        sim.native_utility = 1
    elif all2all:
        sim = All2AllGossipSimulator(nodes, disp, 10, protocol, **kw)
    else:
        sim = GossipSimulator(nodes, disp, 10, protocol, **kw)
    sim.progress = False
    sim.engine = engine
    rep = SimulationReport()
    sim.add_receiver(rep)
    sim.init_nodes(seed=9)
    if all2all:
        sim.start(UniformMixing(net), rounds)
    else:
        sim.start(rounds)
    rows = {i: nd.model_handler.row.clone() for i, nd in sim.nodes.items()}
    return rep, rows, sim


@pytest.mark.parametrize("protocol", ["PUSH", "PULL", "PUSH_PULL"])
def test_deterministic_ring_native_equals_python(protocol):
    """One peer per node, no faults: the event order is fully determined by the nodes' offsets except
    for the per-round shuffle, which only permutes sends of DIFFERENT ticks... so make it exact by
    comparing order-independent outcomes: counters, ages and evaluation curves."""
    from gossipy_b200.core import AntiEntropyProtocol
    import gossipy_b200 as g
    p = getattr(AntiEntropyProtocol, protocol)
    rep_p, rows_p, sim_p = _sim("python", p, ring=True)
    g.CACHE.clear()
    rep_n, rows_n, sim_n = _sim("native", p, ring=True)
    assert rep_p._sent_messages == rep_n._sent_messages
    assert rep_p._failed_messages == rep_n._failed_messages == 0
    assert rep_p._total_size == rep_n._total_size
    ages_p = sorted(int(n.model_handler.n_updates) for n in sim_p.nodes.values())
    ages_n = sorted(int(n.model_handler.n_updates) for n in sim_n.nodes.values())
    assert ages_p == ages_n
    assert len(rep_p.get_evaluation(False)) == len(rep_n.get_evaluation(False)) == 4
    assert len(g.CACHE) == 0


def test_faulty_run_statistics_and_no_leaks():
    from gossipy_b200.core import AntiEntropyProtocol, UniformDelay
    import gossipy_b200 as g
    rep, _, sim = _sim("native", AntiEntropyProtocol.PUSH_PULL, n=8, drop=.2, online=.7, delay=UniformDelay(0, 4),
                       rounds=30, sampling_eval=.5)
    # every node fires once per round; requests + delivered replies are counted as sent
    assert 8 * 30 <= rep._sent_messages <= 2 * 8 * 30
    frac_failed = rep._failed_messages / (8 * 30 * 2)
    assert .15 < frac_failed < .6
    assert len(rep.get_evaluation(False)) == 30
    # in-flight snapshots: only messages still queued may hold cache entries
    assert len(g.CACHE) <= sim._scheduler.pending + 1
    acc = [e["accuracy"] for _, e in rep.get_evaluation(False)]
    assert np.mean(acc[-5:]) > np.mean(acc[:3]) - .05


def test_tokenized_and_all2all_native():
    from gossipy_b200.core import AntiEntropyProtocol
    from gossipy_b200.flow_control import RandomizedTokenAccount, SimpleTokenAccount
    import gossipy_b200 as g
    rep, _, sim = _sim("native", AntiEntropyProtocol.PUSH, n=8, rounds=40, tokenized=RandomizedTokenAccount(C=6, A=3))
    assert rep._sent_messages > 0
    bal = sim._scheduler.token_balances()
    assert all(0 <= b <= 40 for b in bal)
    g.CACHE.clear()
    # warm-up: with C=20, A=10 nothing is sent in the first rounds (proactive() == 0 while a < A-1)
    rep2, _, _ = _sim("native", AntiEntropyProtocol.PUSH, n=8, rounds=5, tokenized=RandomizedTokenAccount(C=20, A=10))
    assert rep2._sent_messages == 0
    g.CACHE.clear()
    rep3, rows3, _ = _sim("native", AntiEntropyProtocol.PUSH, n=5, rounds=6, all2all=True)
    g.CACHE.clear()
    rep4, rows4, _ = _sim("python", AntiEntropyProtocol.PUSH, n=5, rounds=6, all2all=True)
    assert rep3._sent_messages == rep4._sent_messages == 5 * 4 * 6
    assert rep3._total_size == rep4._total_size


def test_scheduler_event_stream_invariants():
    from gossipy_b200.ops.native import _try_import
    C = _try_import()
    sch = C.GossipScheduler(50, 20, 3, 0.1, 0.8, 0.1, 1234)      # PUSH_PULL
    sch.set_delay(1, 0, 5)
    sch.set_message_sizes(117, 1)
    ev = sch.run(10)
    kinds = ev[:, 0]
    sends = ev[kinds == C.EV_SEND]
    assert len(sends) == 50 * 10                                    # sync nodes fire exactly once per round
    assert (np.diff(ev[:, 1]) >= 0).all()                           # time never runs backwards
    ids = set(sends[:, 4].tolist())
    delivered = ev[kinds == C.EV_DELIVER][:, 4].tolist()
    dropped = ev[kinds == C.EV_DROP][:, 4].tolist()
    assert set(delivered) <= ids and len(delivered) == len(set(delivered))
    # every request is delivered, dropped or still in flight -- never both
    assert not (set(delivered) & set(dropped))
    replies = ev[kinds == C.EV_REPLY_SEND]
    assert len(replies) == len(delivered)
    assert sch.sent == len(sends) + int((kinds == C.EV_REPLY_DELIVER).sum())
    assert sch.failed == len(dropped)
    assert (kinds == C.EV_EVAL).sum() == 10 * 5


# ---------------------------------------------------------------------------------------------
# batched execution of the linear learners (engine.bank.LinearBank): same schedule, same results as
# the per-event executor
# ---------------------------------------------------------------------------------------------
def _linear_sim(batched, protocol, mode, handler="pegasos", n=40, rounds=6, device="cpu", faults=True, sync=False,
                passthrough=False, cacheneigh=False):
    import gossipy_b200 as g
    from gossipy_b200.core import AntiEntropyProtocol, CreateModelMode, StaticP2PNetwork, UniformDelay
    from gossipy_b200.data import DataDispatcher, synthetic
    from gossipy_b200.data.handler import ClassificationDataHandler
    from gossipy_b200.model.handler import AdaLineHandler, PegasosHandler
    from gossipy_b200.model.nn import AdaLine
    from gossipy_b200.node import GossipNode
    from gossipy_b200.simul import GossipSimulator, SimulationReport
    g.GlobalSettings().set_device(device)
    g.CACHE.clear()
    g.set_seed(5)
    (Xtr, ytr), (Xte, yte) = synthetic.spambase_like(3 * n + 7, 150)
    disp = DataDispatcher(ClassificationDataHandler(Xtr, 2 * ytr - 1, Xte, 2 * yte - 1), n=n, eval_on_user=False)
    cls = PegasosHandler if handler == "pegasos" else AdaLineHandler
    proto = cls(AdaLine(57), .01 if handler == "pegasos" else .001, getattr(CreateModelMode, mode))
    if cacheneigh:
        from gossipy_b200.node import CacheNeighNode
        nodes = CacheNeighNode.generate(disp, StaticP2PNetwork(n), proto, 10, sync)
    elif passthrough:               # degree-aware pass-through on a topology with unequal degrees (ring + hub)
        from gossipy_b200.node import PassThroughNode
        A = np.zeros((n, n), dtype=int)
        for i in range(n):
            A[i, (i + 1) % n] = A[(i + 1) % n, i] = 1
            if i % 3 == 0 and i:
                A[i, 0] = A[0, i] = 1
        nodes = PassThroughNode.generate(disp, StaticP2PNetwork(n, A), proto, 10, sync)
    else:
        nodes = GossipNode.generate(disp, StaticP2PNetwork(n), proto, 10, sync)
    kw = dict(drop_prob=.1, online_prob=.8, delay=UniformDelay(0, 3), sampling_eval=.3) if faults else {}
    sim = GossipSimulator(nodes, disp, 10, getattr(AntiEntropyProtocol, protocol), **kw)
    sim.progress = False
    sim.engine = "native"
    sim.batched = batched
    rep = SimulationReport(); sim.add_receiver(rep)
    sim.init_nodes(seed=9)
    sim.start(rounds)
    rows = torch.stack([sim.nodes[i].model_handler.row[:57].detach().cpu().clone() for i in range(n)])
    ages = [int(sim.nodes[i].model_handler.n_updates) for i in range(n)]
    return rep, rows, ages, sim


@pytest.mark.parametrize("protocol,mode,handler", [("PUSH", "MERGE_UPDATE", "pegasos"), ("PUSH_PULL", "MERGE_UPDATE", "pegasos"),
                                                   ("PULL", "UPDATE", "pegasos"), ("PUSH", "UPDATE_MERGE", "adaline"),
                                                   ("PUSH_PULL", "UPDATE", "adaline"), ("PUSH", "PASS", "pegasos")])
def test_banked_execution_equals_per_event_execution(protocol, mode, handler):
    import gossipy_b200 as g
    rep_a, rows_a, ages_a, sim_a = _linear_sim(False, protocol, mode, handler)
    assert "_bank" not in sim_a.__dict__
    rep_b, rows_b, ages_b, sim_b = _linear_sim(True, protocol, mode, handler)
    assert "_bank" in sim_b.__dict__
    assert (rep_a._sent_messages, rep_a._failed_messages, rep_a._total_size) == \
        (rep_b._sent_messages, rep_b._failed_messages, rep_b._total_size)
    assert ages_a == ages_b
    torch.testing.assert_close(rows_a, rows_b, rtol=1e-4, atol=1e-5)
    ev_a, ev_b = rep_a.get_evaluation(False), rep_b.get_evaluation(False)
    assert len(ev_a) == len(ev_b) == 6
    for (t1, m1), (t2, m2) in zip(ev_a, ev_b):
        assert t1 == t2
        for k in m1:
            assert m1[k] == pytest.approx(m2[k], abs=1e-4), k
    g.CACHE.clear()


@pytest.mark.parametrize("protocol,mode", [("PUSH", "MERGE_UPDATE"), ("PUSH_PULL", "UPDATE"), ("PULL", "MERGE_UPDATE")])
def test_banked_pass_through_nodes_equal_per_event_execution(protocol, mode):
    """PassThroughNode (reference node.py:289-392) in the banked engine: keyed accept draws, per-message PASS / merge."""
    import gossipy_b200 as g
    rep_a, rows_a, ages_a, sim_a = _linear_sim(False, protocol, mode, passthrough=True)
    rep_b, rows_b, ages_b, sim_b = _linear_sim(True, protocol, mode, passthrough=True)
    assert "_bank" not in sim_a.__dict__ and "_bank" in sim_b.__dict__
    assert (rep_a._sent_messages, rep_a._failed_messages, rep_a._total_size) == \
        (rep_b._sent_messages, rep_b._failed_messages, rep_b._total_size)
    draws_a = [getattr(nd, "_pt_draws", 0) for nd in sim_a.nodes.values()]
    assert draws_a == [getattr(nd, "_pt_draws", 0) for nd in sim_b.nodes.values()] and sum(draws_a) > 0
    assert ages_a == ages_b
    torch.testing.assert_close(rows_a, rows_b, rtol=1e-4, atol=1e-5)
    g.CACHE.clear()


@pytest.mark.parametrize("protocol,mode", [("PUSH", "MERGE_UPDATE"), ("PUSH_PULL", "MERGE_UPDATE"), ("PUSH", "UPDATE")])
def test_banked_cache_neighbour_nodes_equal_per_event_execution(protocol, mode):
    """CacheNeighNode (reference node.py:395-496) in the banked engine: deliveries are stored per sender, one cached
    model (keyed choice) is consumed before every PUSH / PUSH_PULL send."""
    import gossipy_b200 as g
    rep_a, rows_a, ages_a, sim_a = _linear_sim(False, protocol, mode, cacheneigh=True)
    rep_b, rows_b, ages_b, sim_b = _linear_sim(True, protocol, mode, cacheneigh=True)
    assert "_bank" not in sim_a.__dict__ and "_bank" in sim_b.__dict__
    assert (rep_a._sent_messages, rep_a._failed_messages, rep_a._total_size) == \
        (rep_b._sent_messages, rep_b._failed_messages, rep_b._total_size)
    draws_a = [getattr(nd, "_cn_draws", 0) for nd in sim_a.nodes.values()]
    assert draws_a == [getattr(nd, "_cn_draws", 0) for nd in sim_b.nodes.values()] and sum(draws_a) > 0
    assert ages_a == ages_b
    torch.testing.assert_close(rows_a, rows_b, rtol=1e-4, atol=1e-5)
    ev_a, ev_b = rep_a.get_evaluation(False), rep_b.get_evaluation(False)
    for (t1, m1), (t2, m2) in zip(ev_a, ev_b):
        for k in m1:
            assert m1[k] == pytest.approx(m2[k], abs=1e-4), k
    g.CACHE.clear()


def test_bank_falls_back_when_not_bankable():
    from gossipy_b200.engine.bank import bankable
    rep, _, sim = _sim("native", __import__("gossipy_b200").core.AntiEntropyProtocol.PUSH, n=5, rounds=2)
    assert bankable(sim) is not None and "_bank" not in sim.__dict__


# ---------------------------------------------------------------------------------------------
# checkpoint / resume under the native engine: the scheduler's dynamic state and the in-flight
# snapshots are part of the checkpoint, so an interrupted run continues on the identical schedule
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("batched,variant", [(False, {}), (True, {}), (True, {"cacheneigh": True}), (True, {"passthrough": True})])
def test_native_checkpoint_resume_is_exact(batched, variant, tmp_path):
    import gossipy_b200 as g
    from gossipy_b200.simul import GossipSimulator
    rep_full, rows_full, ages_full, _ = _linear_sim(batched, "PUSH_PULL", "MERGE_UPDATE", rounds=6, **variant)
    g.CACHE.clear()
    rep_a, _, _, sim = _linear_sim(batched, "PUSH_PULL", "MERGE_UPDATE", rounds=3, **variant)
    assert sim._scheduler.pending > 0          # delays up to 3 ticks: messages are on the wire at the cut
    path = str(tmp_path / "ckpt.pkl")
    sim.save(path)
    g.CACHE.clear()
    del sim
    sim2 = GossipSimulator.load(path)
    assert "_scheduler" not in sim2.__dict__ and "_scheduler_state" in sim2.__dict__
    rep_b = [r for r in sim2._receivers if type(r).__name__ == "SimulationReport"][0]
    sim2.start(3, resume=True)
    assert ("_bank" in sim2.__dict__) == batched
    n = len(sim2.nodes)
    rows = torch.stack([sim2.nodes[i].model_handler.row[:57].detach().cpu().clone() for i in range(n)])
    ages = [int(sim2.nodes[i].model_handler.n_updates) for i in range(n)]
    assert ages == ages_full
    torch.testing.assert_close(rows, rows_full, rtol=1e-5, atol=1e-6)
    assert (rep_b._sent_messages, rep_b._failed_messages, rep_b._total_size) == \
        (rep_full._sent_messages, rep_full._failed_messages, rep_full._total_size)
    ev_full, ev_b = rep_full.get_evaluation(False), rep_b.get_evaluation(False)
    assert [t for t, _ in ev_b] == [t for t, _ in ev_full]
    for (_, m1), (_, m2) in zip(ev_full, ev_b):
        for k in m1:
            assert m1[k] == pytest.approx(m2[k], abs=1e-6), k
    g.CACHE.clear()


def test_scheduler_state_roundtrip():
    from gossipy_b200.ops.native import _try_import
    C = _try_import()

    def make():
        s = C.GossipScheduler(12, 10, 3, 0.2, 0.9, 0.5, 77)
        s.set_nodes([0] * 12, [3 + i % 4 for i in range(12)], [10] * 12)
        s.set_delay(1, 0, 5)
        s.set_token_account(5, 6, 3, 1, 1)          # randomized token account
        return s
    a = make()
    full = np.concatenate([a.run(1) for _ in range(8)])
    b = make()
    head = [b.run(1) for _ in range(4)]
    st = dict(b.get_state())
    c = make()
    c.set_state(st)
    tail = [c.run(1) for _ in range(4)]
    assert np.array_equal(np.concatenate(head + tail), full)
    assert (c.sent, c.failed, c.total_size, c.clock) == (a.sent, a.failed, a.total_size, a.clock)
    assert c.token_balances() == a.token_balances()
    with pytest.raises(Exception):
        C.GossipScheduler(5, 10, 3, 0., 1., 0., 1).set_state(st)


def test_python_checkpoint_resumed_under_native_engine_returns_pending_snapshots(tmp_path):
    """The pending queues of a Python-engine checkpoint cannot be continued by the C++ scheduler: the messages are
    discarded WITH their snapshots (arena rows must not leak)."""
    import gossipy_b200 as g
    from gossipy_b200.core import AntiEntropyProtocol, StaticP2PNetwork, UniformDelay
    from gossipy_b200.data import DataDispatcher, synthetic
    from gossipy_b200.data.handler import ClassificationDataHandler
    from gossipy_b200.model.handler import TorchModelHandler
    from gossipy_b200.model.nn import LogisticRegression
    from gossipy_b200.node import GossipNode
    from gossipy_b200.simul import GossipSimulator
    g.CACHE.clear()
    g.set_seed(3)
    (Xtr, ytr), (Xte, yte) = synthetic.spambase_like(300, 100)
    disp = DataDispatcher(ClassificationDataHandler(Xtr, ytr, Xte, yte), n=6, eval_on_user=False)
    proto = TorchModelHandler(LogisticRegression(57, 2), torch.optim.SGD, {"lr": .1}, torch.nn.CrossEntropyLoss(), batch_size=16)
    nodes = GossipNode.generate(disp, StaticP2PNetwork(6), proto, 10, True)
    sim = GossipSimulator(nodes, disp, 10, AntiEntropyProtocol.PUSH, delay=UniformDelay(5, 30))
    sim.progress = False
    sim.init_nodes(seed=1)
    sim.start(2)                                   # Python engine; long delays leave messages on the wire
    assert len(g.CACHE) > 0
    sim.engine = "native"
    sim.start(1, resume=True)
    sim.engine = "python"
    assert sum(len(q) for q in sim._msg_queues.values()) == 0
    g.CACHE.clear()


# ---------------------------------------------------------------------------------------------
# PENS under the C++ control plane: step switch between two pieces of a round, restricted peer lists
# ---------------------------------------------------------------------------------------------
def test_scheduler_run_in_pieces_and_peer_lists():
    from gossipy_b200.ops.native import _try_import
    C = _try_import()

    def make():
        s = C.GossipScheduler(9, 10, 1, 0.1, 0.9, 0.0, 5)
        s.set_nodes([1] * 9, [i % 10 for i in range(9)], [10] * 9)
        s.set_delay(1, 0, 4)
        return s
    a, b = make(), make()
    whole = np.concatenate([a.run(1) for _ in range(3)])
    pieces = np.concatenate([b.run_ticks(k) for k in (3, 7, 1, 0, 12, 7)])       # 30 ticks in odd pieces
    assert np.array_equal(whole, pieces) and a.clock == b.clock == 30
    b.set_peer_list(2, [7, 8])
    b.set_peer_list(4, [0])
    ev = np.concatenate([b.run(1) for _ in range(20)])
    sends = ev[ev[:, 0] == C.EV_SEND]
    assert set(sends[sends[:, 2] == 2][:, 3].tolist()) == {7, 8}
    assert set(sends[sends[:, 2] == 4][:, 3].tolist()) == {0}
    assert len(set(sends[sends[:, 2] == 3][:, 3].tolist())) > 2                  # the others keep their neighbourhood
    b.set_peer_list(2, [])
    ev = np.concatenate([b.run(1) for _ in range(20)])
    sends = ev[ev[:, 0] == C.EV_SEND]
    assert len(set(sends[sends[:, 2] == 2][:, 3].tolist())) > 2
    with pytest.raises(Exception):
        b.set_peer_list(0, [9])


def _pens_sim(engine, rounds, step1_rounds=4, round_len=10, executor=True, faults=False):
    import gossipy_b200 as g
    from gossipy_b200.core import AntiEntropyProtocol, CreateModelMode, StaticP2PNetwork
    from gossipy_b200.data import DataDispatcher, synthetic
    from gossipy_b200.data.handler import ClassificationDataHandler
    from gossipy_b200.model.handler import TorchModelHandler
    from gossipy_b200.model.nn import LogisticRegression
    from gossipy_b200.node import PENSNode
    from gossipy_b200.simul import GossipSimulator, SimulationEventReceiver, SimulationReport

    class Sends(SimulationEventReceiver):
        def __init__(self):
            self.log = []

        def update_message(self, failed, msg=None):
            if not failed and msg is not None:
                self.log.append((int(msg.timestamp), int(msg.sender), int(msg.receiver)))

        def update_evaluation(self, *a, **k): pass
        def update_timestep(self, t): pass
        def update_end(self): pass

    g.set_seed(11)
    (Xtr, ytr), (Xte, yte) = synthetic.spambase_like(700, 200)
    disp = DataDispatcher(ClassificationDataHandler(Xtr, ytr, Xte, yte), n=7, eval_on_user=False)
    proto = TorchModelHandler(LogisticRegression(57, 2), torch.optim.SGD, {"lr": .5}, torch.nn.CrossEntropyLoss(),
                              batch_size=16, create_model_mode=CreateModelMode.MERGE_UPDATE)
    nodes = PENSNode.generate(disp, StaticP2PNetwork(7), proto, round_len, True, n_sampled=3, m_top=1,
                              step1_rounds=step1_rounds)
    kw = {}
    if faults:              # messages stay on the wire across round boundaries, some are lost
        from gossipy_b200.core import UniformDelay
        kw = dict(drop_prob=.1, online_prob=.8, delay=UniformDelay(0, 14))
    sim = GossipSimulator(nodes, disp, 10, AntiEntropyProtocol.PUSH, **kw)
    sim.progress = False
    sim.engine = engine
    sim.native_executor = executor          # False: per-event executor also in step 2 (real messages reach the receivers)
    rep, sends = SimulationReport(), Sends()
    sim.add_receiver(rep)
    sim.add_receiver(sends)
    sim.init_nodes(seed=4)
    sim.start(rounds)
    return sim, rep, sends


@pytest.mark.parametrize("round_len", [10, 5])
def test_pens_runs_under_the_native_scheduler(round_len):
    """PENSNode overrides ``timed_out`` / ``get_peer`` (ref node.py:716-741): the C++ scheduler takes both over --
    the step switch falls between two pieces of a round, step 2 draws from the node's ``best_nodes``."""
    import gossipy_b200 as g
    sim, rep, sends = _pens_sim("native", rounds=9, step1_rounds=4, round_len=round_len, executor=False)
    assert "_scheduler" in sim.__dict__ and sim._native_supported() is None
    t_sw = 4 * round_len
    step1 = [s for s in sends.log if s[0] < t_sw]
    step2 = [s for s in sends.log if s[0] >= t_sw]
    assert step1 and step2
    for node in sim.nodes.values():
        assert node.step == 2 and node.best_nodes is not None
        mine = [r for t, s, r in step1 if s == node.idx]
        assert sum(node.selected.values()) == len(mine)               # what get_peer counts in step 1
        assert all(node.selected[p] == mine.count(p) for p in node.selected)
        if node.best_nodes:
            assert {r for t, s, r in step2 if s == node.idx} <= set(node.best_nodes)
    assert sum(sum(n.neigh_counter.values()) for n in sim.nodes.values()) > 0    # selections happened (n_sampled = 3)
    assert any(n.best_nodes for n in sim.nodes.values())
    acc = [m["accuracy"] for _, m in rep.get_evaluation(False)]
    assert len(acc) == 9 and acc[-1] > 0.6
    g.CACHE.clear()


@pytest.mark.parametrize("faults", [False, True])
def test_pens_native_checkpoint_resume_is_exact(tmp_path, faults):
    import gossipy_b200 as g
    from gossipy_b200.simul import GossipSimulator
    full, rep_full, sends_full = _pens_sim("native", rounds=8, step1_rounds=3, faults=faults)
    rows_full = {i: n.model_handler.row.clone() for i, n in full.nodes.items()}
    g.CACHE.clear()
    for cut in (2, 5):                    # one checkpoint inside step 1 (cached candidates), one inside step 2
        sim, _, _ = _pens_sim("native", rounds=cut, step1_rounds=3, faults=faults)
        path = str(tmp_path / ("pens%d.pkl" % cut))
        sim.save(path)
        g.CACHE.clear()
        del sim
        sim2 = GossipSimulator.load(path)
        sim2.start(8 - cut, resume=True)
        for i, n in sim2.nodes.items():
            torch.testing.assert_close(n.model_handler.row, rows_full[i], rtol=1e-6, atol=1e-7)
            assert n.best_nodes == full.nodes[i].best_nodes and n.selected == full.nodes[i].selected
        rep2 = [r for r in sim2._receivers if type(r).__name__ == "SimulationReport"][0]
        assert rep2._sent_messages == rep_full._sent_messages
        g.CACHE.clear()


@pytest.mark.parametrize("faults", [False, True])
def test_pens_step_two_moves_to_the_cpp_executor(faults, monkeypatch):
    """Once every PENS node has left its selection phase a delivery is a plain merge + update: the rest of the run is
    enqueued from C++ (messages on the wire become executor slots) and equals the per-event executor's run."""
    import gossipy_b200 as g
    from gossipy_b200.simul import GossipSimulator
    ref, rep_ref, sends_ref = _pens_sim("native", rounds=9, step1_rounds=3, executor=False, faults=faults)
    assert "_stream_exec" not in ref.__dict__
    rows_ref = {i: n.model_handler.row.clone() for i, n in ref.nodes.items()}
    g.CACHE.clear()
    moved = []
    orig = GossipSimulator._handover_to_executor

    def spy(self):
        n = len(self._native_msgs)
        ok = orig(self)
        moved.append((n, ok, len(self.__dict__.get("_exec_inflight", {}).get("ids", []))))
        return ok
    monkeypatch.setattr(GossipSimulator, "_handover_to_executor", spy)
    sim, rep, sends = _pens_sim("native", rounds=9, step1_rounds=3, faults=faults)
    assert moved and moved[-1][1] and (not faults or moved[-1][0] > 0 and moved[-1][2] == moved[-1][0])
    assert "_stream_exec" in sim.__dict__ and not sim._native_msgs
    for i, n in sim.nodes.items():
        torch.testing.assert_close(n.model_handler.row, rows_ref[i], rtol=1e-6, atol=1e-7)
        assert int(n.model_handler.n_updates) == int(ref.nodes[i].model_handler.n_updates)
        assert n.best_nodes == ref.nodes[i].best_nodes
    assert (rep._sent_messages, rep._failed_messages, rep._total_size) == \
        (rep_ref._sent_messages, rep_ref._failed_messages, rep_ref._total_size)
    for (t1, m1), (t2, m2) in zip(rep.get_evaluation(False), rep_ref.get_evaluation(False)):
        assert t1 == t2
        for k in m1:
            assert m1[k] == pytest.approx(m2[k], abs=1e-6)
    g.CACHE.clear()


def test_pens_random_setups_executor_and_checkpoints_equal_per_event_runs():
    """Seeded sweep over node count, n_sampled / m_top, length of the selection phase, round lengths, async nodes, faults
    and a checkpoint somewhere: the run that moves to the C++ executor (and is interrupted) equals the per-event run."""
    import random
    import gossipy_b200 as g
    from gossipy_b200.core import AntiEntropyProtocol, CreateModelMode, StaticP2PNetwork, UniformDelay
    from gossipy_b200.data import DataDispatcher, synthetic
    from gossipy_b200.data.handler import ClassificationDataHandler
    from gossipy_b200.model.handler import TorchModelHandler
    from gossipy_b200.model.nn import LogisticRegression
    from gossipy_b200.node import PENSNode
    from gossipy_b200.simul import GossipSimulator, SimulationReport

    def run(executor, n, ns, mt, s1, rounds, faults, rl, sync, seed, cut=None, path=None):
        g.CACHE.clear()
        g.set_seed(seed)
        (Xtr, ytr), (Xte, yte) = synthetic.spambase_like(60 * n, 100)
        disp = DataDispatcher(ClassificationDataHandler(Xtr, ytr, Xte, yte), n=n, eval_on_user=False)
        proto = TorchModelHandler(LogisticRegression(57, 2), torch.optim.SGD, {"lr": .3}, torch.nn.CrossEntropyLoss(),
                                  batch_size=16, create_model_mode=CreateModelMode.MERGE_UPDATE)
        nodes = PENSNode.generate(disp, StaticP2PNetwork(n), proto, rl, sync, n_sampled=ns, m_top=mt, step1_rounds=s1)
        kw = dict(drop_prob=.15, online_prob=.8, delay=UniformDelay(0, 13)) if faults else {}
        sim = GossipSimulator(nodes, disp, 10, AntiEntropyProtocol.PUSH, **kw)
        sim.progress = False
        sim.engine = "native"
        sim.native_executor = executor
        rep = SimulationReport()
        sim.add_receiver(rep)
        sim.init_nodes(seed=seed)
        if cut:
            sim.start(cut)
            sim.save(path)
            g.CACHE.clear()
            sim = GossipSimulator.load(path)
            rep = [r for r in sim._receivers if type(r).__name__ == "SimulationReport"][0]
            sim.start(rounds - cut, resume=True)
        else:
            sim.start(rounds)
        rows = torch.stack([sim.nodes[i].model_handler.row.clone() for i in range(n)])
        ages = [int(sim.nodes[i].model_handler.n_updates) for i in range(n)]
        return rows, ages, (rep._sent_messages, rep._failed_messages, rep._total_size), "_stream_exec" in sim.__dict__

    import tempfile
    rnd = random.Random(20260921)
    moved = 0
    with tempfile.TemporaryDirectory() as tmp:
        for case in range(10):
            n = rnd.randint(3, 8)
            ns = rnd.randint(1, min(4, n - 1))
            mt, s1 = rnd.randint(1, ns), rnd.randint(0, 3)
            rounds = rnd.randint(s1 + 2, s1 + 4)
            args = (n, ns, mt, s1, rounds, rnd.random() < .6, rnd.choice([10, 10, 5, 20]), rnd.random() < .7, rnd.randint(0, 10 ** 6))
            cut = rnd.choice([None, rnd.randint(1, rounds - 1)])
            a = run(False, *args)
            b = run(True, *args, cut=cut, path=os.path.join(tmp, "c%d.pkl" % case))
            moved += b[3]
            torch.testing.assert_close(a[0], b[0], rtol=1e-6, atol=1e-7, msg=lambda m: "%s cut=%s: %s" % (args, cut, m))
            assert a[1] == b[1] and a[2] == b[2], (args, cut)
    assert moved >= 4
    g.CACHE.clear()


def test_constant_utility_functions_are_recognised_from_their_byte_code():
    """A tokenized run goes native when ``utility_fun`` is nothing but ``return <int>`` (the reference scripts' lambda);
    anything that computes keeps the Python loop unless ``native_utility`` says otherwise."""
    from gossipy_b200.flow_control import SimpleTokenAccount
    from gossipy_b200.simul import TokenizedGossipSimulator

    def const_def(mh1, mh2, msg):
        return 2

    def computes(mh1, mh2, msg):
        return 1 if msg is not None else 0

    sim = TokenizedGossipSimulator.__new__(TokenizedGossipSimulator)
    for fn, want in ((lambda mh1, mh2, msg: 1, 1), (const_def, 2), (computes, None), (lambda a, b, m: True, None),
                     (lambda a, b, m: 1.5, None), (len, None)):
        sim.utility_fun = fn
        assert sim._constant_utility() == want, fn
    sim.native_utility = 3
    assert sim._constant_utility() == 3
    del sim.native_utility
    # end to end: the helper's simulation passes a lambda and no native_utility
    rep, rows, s = _sim("native", __import__("gossipy_b200.core", fromlist=["x"]).AntiEntropyProtocol.PUSH, tokenized=SimpleTokenAccount(C=2), rounds=3)
    s.native_utility = None
    assert s._constant_utility() == 1 and s._native_supported() is None


@pytest.mark.parametrize("path", ["python-loop", "per-event", "executor", "bank"])
def test_fresh_starts_do_not_leak_what_was_on_the_wire(path):
    """``start`` without ``resume`` restarts the clock and forgets the messages in flight (like the reference, whose
    queues are locals of ``start``) -- their snapshots must go back to the arenas / slot pools, run after run."""
    import gossipy_b200 as g
    from gossipy_b200.core import AntiEntropyProtocol, CreateModelMode, StaticP2PNetwork, UniformDelay
    from gossipy_b200.data import DataDispatcher, synthetic
    from gossipy_b200.data.handler import ClassificationDataHandler
    from gossipy_b200.engine import arena
    from gossipy_b200.model.handler import PegasosHandler, TorchModelHandler
    from gossipy_b200.model.nn import AdaLine, LogisticRegression
    from gossipy_b200.node import GossipNode
    from gossipy_b200.simul import GossipSimulator
    g.CACHE.clear()
    g.set_seed(5)
    n = 8
    base = sum(a.live for a in arena._ARENAS.values())       # rows of earlier tests' simulations that are still alive
    (Xtr, ytr), (Xte, yte) = synthetic.spambase_like(50 * n, 80)
    if path == "bank":
        ytr, yte = 2 * ytr - 1, 2 * yte - 1
        proto = PegasosHandler(AdaLine(57), 0.01, CreateModelMode.MERGE_UPDATE)
    else:
        proto = TorchModelHandler(LogisticRegression(57, 2), torch.optim.SGD, {"lr": .3}, torch.nn.CrossEntropyLoss(),
                                  batch_size=16, create_model_mode=CreateModelMode.MERGE_UPDATE)
    disp = DataDispatcher(ClassificationDataHandler(Xtr, ytr, Xte, yte), n=n, eval_on_user=False)
    nodes = GossipNode.generate(disp, StaticP2PNetwork(n), proto, 10, True)
    sim = GossipSimulator(nodes, disp, 10, AntiEntropyProtocol.PUSH_PULL, delay=UniformDelay(5, 25))   # always something in flight
    sim.progress = False
    sim.engine = "python" if path == "python-loop" else "native"
    sim.native_executor = path == "executor"
    sim.batched = path == "bank"
    sim.init_nodes(seed=1)

    def in_use():
        if path == "executor":
            return len(sim._stream_exec.ex.inflight())
        if path == "bank":
            return sim._bank.cap - sim._bank.n_free
        return sum(a.live for a in arena._ARENAS.values()) - base
    seen = []
    for _ in range(4):
        sim.start(3)
        seen.append(in_use())
    assert ("_stream_exec" in sim.__dict__) == (path == "executor") and ("_bank" in sim.__dict__) == (path == "bank")
    assert seen[0] > (n if path in ("python-loop", "per-event") else 0), seen      # something was on the wire at the cut
    assert max(seen[1:]) <= seen[0] + n, seen                                      # ... and it does not pile up
    sim._forget_messages_on_the_wire()
    assert in_use() == (n if path in ("python-loop", "per-event") else 0)
    g.CACHE.clear()


def test_implicit_clique_needs_no_peer_table():
    """``StaticP2PNetwork(n)`` (no topology): the scheduler's built-in clique enumerates the peers in the order of the
    network's lists, so the schedule equals the one driven by the explicit n x (n - 1) table."""
    from gossipy_b200.core import AntiEntropyProtocol
    rep, rows, sim = _sim("native", AntiEntropyProtocol.PUSH_PULL, n=7, drop=.1, online=.8, rounds=1)
    a = sim._make_scheduler()
    b = sim._make_scheduler()
    indptr, indices = sim.nodes[0].p2p_net.as_csr()
    assert len(indices) == 7 * 6
    b.set_topology(indptr.tolist(), indices.tolist())
    ea = np.concatenate([a.run(1) for _ in range(30)])
    eb = np.concatenate([b.run(1) for _ in range(30)])
    assert np.array_equal(ea, eb) and len(ea) > 400
