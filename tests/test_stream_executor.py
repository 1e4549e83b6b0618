"""The C++ executor (csrc/exec) against the per-event Python executor: same native schedule, same counter-based
shuffle keys -> the runs must be identical (weights, ages, counters, message accounting, metrics).  On CPU the
executor's two launches are callbacks into gossipy_b200.ops, so all of its bookkeeping is exercised here; the CUDA
stream / event side is covered by the gpu test at the bottom."""
import numpy as np
import pytest
import torch

from gossipy_b200.ops.native import native_available

pytestmark = pytest.mark.skipif(not native_available(), reason="extension not built")


def _sim(streamed, model="mlp", protocol="PUSH_PULL", faults=False, n=6, rounds=4, device="cpu", sync=True, start=True,
         mode="MERGE_UPDATE", limited=None, tokenized=False, partitioned=0, passthrough=False, sampled=0.0, cacheneigh=False, momentum=None):
    import gossipy_b200 as g
    from gossipy_b200.core import AntiEntropyProtocol, StaticP2PNetwork, UniformDelay
    from gossipy_b200.data import DataDispatcher, synthetic
    from gossipy_b200.data.handler import ClassificationDataHandler
    from gossipy_b200.core import CreateModelMode
    from gossipy_b200.flow_control import RandomizedTokenAccount
    from gossipy_b200.model.handler import LimitedMergeTMH, PartitionedTMH, TorchModelHandler
    from gossipy_b200.model.sampling import TorchModelPartition
    from gossipy_b200.node import PartitioningBasedNode
    from gossipy_b200.simul import TokenizedGossipSimulator
    from gossipy_b200.model.nn import LogisticRegression, TorchMLP
    from gossipy_b200.node import GossipNode
    from gossipy_b200.simul import GossipSimulator, SimulationReport
    g.GlobalSettings().set_device(device)
    g.CACHE.clear()
    g.set_seed(7)
    if model == "mlp":
        (Xtr, ytr), (Xte, yte) = synthetic.mnist_like(96 * n, 120)
        net, bs = TorchMLP(784, 10, (100,)), 32
    else:
        (Xtr, ytr), (Xte, yte) = synthetic.spambase_like(50 * n + 3, 150)
        net, bs = LogisticRegression(57, 2), 16
    disp = DataDispatcher(ClassificationDataHandler(Xtr, ytr, Xte, yte), n=n, eval_on_user=False)
    kwh = dict(batch_size=bs, create_model_mode=getattr(CreateModelMode, mode))
    node_cls = GossipNode
    if partitioned:
        proto = PartitionedTMH(net, TorchModelPartition(net, partitioned), torch.optim.SGD, {"lr": .5, "weight_decay": .001},
                               torch.nn.CrossEntropyLoss(), **kwh)
        node_cls = PartitioningBasedNode
    elif sampled:
        from gossipy_b200.model.handler import SamplingTMH
        from gossipy_b200.node import SamplingBasedNode
        proto = SamplingTMH(sampled, net, torch.optim.SGD, {"lr": .1, "weight_decay": .001}, torch.nn.CrossEntropyLoss(), **kwh)
        node_cls = SamplingBasedNode
    elif limited is not None:
        proto = LimitedMergeTMH(net, torch.optim.SGD, {"lr": .1, "weight_decay": .001}, torch.nn.CrossEntropyLoss(),
                                age_diff_threshold=limited, **kwh)
    elif momentum is not None:
        proto = TorchModelHandler(net, torch.optim.SGD, dict({"lr": .05, "weight_decay": .001}, **momentum), torch.nn.CrossEntropyLoss(), **kwh)
    else:
        proto = TorchModelHandler(net, torch.optim.SGD, {"lr": .1, "weight_decay": .001}, torch.nn.CrossEntropyLoss(), **kwh)
    topo = StaticP2PNetwork(n)
    if cacheneigh:
        from gossipy_b200.node import CacheNeighNode
        node_cls = CacheNeighNode
    if passthrough:                 # degree-aware pass-through needs unequal degrees: a ring plus a hub
        from gossipy_b200.node import PassThroughNode
        node_cls = PassThroughNode
        A = np.zeros((n, n), dtype=int)
        for i in range(n):
            A[i, (i + 1) % n] = A[(i + 1) % n, i] = 1
            if i > 1:
                A[i, 0] = A[0, i] = 1
        topo = StaticP2PNetwork(n, A)
    nodes = node_cls.generate(disp, topo, proto, 10, sync)
    kw = dict(drop_prob=.15, online_prob=.8, delay=UniformDelay(0, 4), sampling_eval=.5) if faults else {}
    if tokenized:
        sim = TokenizedGossipSimulator(nodes, disp, RandomizedTokenAccount(C=4, A=2), lambda a, b, m: 1, 10,
                                       getattr(AntiEntropyProtocol, protocol), **kw)
        sim.native_utility = 1
    else:
        sim = GossipSimulator(nodes, disp, 10, getattr(AntiEntropyProtocol, protocol), **kw)
    sim.progress = False
    sim.engine = "native"
    sim.native_executor = streamed
    rep = SimulationReport(); sim.add_receiver(rep)
    sim.init_nodes(seed=11)
    if start:
        sim.start(rounds)
    return sim, rep


def _state(sim):
    n = len(sim.nodes)
    rows = torch.stack([sim.nodes[i].model_handler.row.detach().cpu().clone() for i in range(n)])
    ages = [np.asarray(sim.nodes[i].model_handler.n_updates).tolist() for i in range(n)]
    ctr = [int(sim.nodes[i].model_handler._update_counter) for i in range(n)]
    return rows, ages, ctr


def _same(sim_a, rep_a, sim_b, rep_b, tol=0.0):
    ra, aa, ca = _state(sim_a)
    rb, ab, cb = _state(sim_b)
    assert aa == ab and ca == cb
    if tol == 0.0:
        assert torch.equal(ra, rb)
    else:
        torch.testing.assert_close(ra, rb, rtol=tol, atol=tol)
    assert (rep_a._sent_messages, rep_a._failed_messages, rep_a._total_size) == \
        (rep_b._sent_messages, rep_b._failed_messages, rep_b._total_size)
    ea, eb = rep_a.get_evaluation(False), rep_b.get_evaluation(False)
    assert [t for t, _ in ea] == [t for t, _ in eb] and len(ea) > 0
    for (_, m1), (_, m2) in zip(ea, eb):
        for k in m1:
            assert m1[k] == pytest.approx(m2[k], abs=max(tol, 1e-9)), k


@pytest.mark.parametrize("model,protocol,faults,sync", [("mlp", "PUSH_PULL", False, True), ("logreg", "PUSH", True, False),
                                                        ("logreg", "PULL", True, True), ("mlp", "PUSH", True, True)])
def test_native_executor_equals_python_executor(model, protocol, faults, sync):
    import gossipy_b200 as g
    sim_a, rep_a = _sim(False, model, protocol, faults, sync=sync)
    assert "_stream_exec" not in sim_a.__dict__
    sim_b, rep_b = _sim(True, model, protocol, faults, sync=sync)
    assert "_stream_exec" in sim_b.__dict__ and sim_b._stream_exec.ex.launches > 0
    _same(sim_a, rep_a, sim_b, rep_b)
    assert len(g.CACHE) == 0 or not faults      # the native executor never touches the message cache
    g.CACHE.clear()


@pytest.mark.parametrize("kw", [dict(mode="UPDATE", protocol="PUSH"), dict(mode="PASS", protocol="PUSH_PULL", faults=True),
                                dict(limited=20, protocol="PULL", faults=True), dict(limited=0, protocol="PUSH_PULL"),
                                dict(tokenized=True, protocol="PUSH", faults=True), dict(mode="UPDATE", model="mlp", faults=True),
                                dict(mode="UPDATE_MERGE", protocol="PUSH_PULL", faults=True), dict(mode="UPDATE_MERGE", model="mlp", protocol="PUSH"),
                                dict(mode="UPDATE_MERGE", limited=3, protocol="PULL", faults=True)])
def test_native_executor_modes_and_variants(kw):
    import gossipy_b200 as g
    kw = dict(dict(model="logreg", sync=False), **kw)
    sim_a, rep_a = _sim(False, **kw)
    sim_b, rep_b = _sim(True, **kw)
    assert "_stream_exec" in sim_b.__dict__ and "_stream_exec" not in sim_a.__dict__
    _same(sim_a, rep_a, sim_b, rep_b)
    g.CACHE.clear()


@pytest.mark.parametrize("kw", [dict(model="logreg", protocol="PUSH", partitioned=4, faults=True),
                                dict(model="mlp", protocol="PUSH_PULL", partitioned=3),
                                dict(model="logreg", protocol="PULL", partitioned=7, faults=True, sync=False),
                                # UPDATE (the mode of the reference's main_hegedus_2021.py): a private copy of the received model
                                # is trained with ITS ages, its partition merged into the (untrained) own model
                                dict(model="logreg", protocol="PUSH", partitioned=4, faults=True, mode="UPDATE", tokenized=True),
                                dict(model="mlp", protocol="PUSH_PULL", partitioned=3, mode="UPDATE"),
                                dict(model="logreg", protocol="PULL", partitioned=5, faults=True, sync=False, mode="UPDATE")])
def test_native_executor_partitioned_models(kw):
    """PartitioningBasedNode + PartitionedTMH (reference node.py:566-659, handler.py:455-525) from C++: keyed partition
    draw, per-partition ages on the wire, segment merge with age weights, 1/age-scaled local update."""
    import gossipy_b200 as g
    sim_a, rep_a = _sim(False, **kw)
    sim_b, rep_b = _sim(True, **kw)
    assert "_stream_exec" in sim_b.__dict__ and "_stream_exec" not in sim_a.__dict__
    _same(sim_a, rep_a, sim_b, rep_b)
    assert [getattr(nd, "_model_msgs", 0) for nd in sim_a.nodes.values()] == [getattr(nd, "_model_msgs", 0) for nd in sim_b.nodes.values()]
    assert sum(getattr(nd, "_model_msgs", 0) for nd in sim_b.nodes.values()) > 0
    g.CACHE.clear()


@pytest.mark.parametrize("kw", [dict(model="logreg", protocol="PUSH", passthrough=True, faults=True),
                                dict(model="mlp", protocol="PUSH_PULL", passthrough=True),
                                dict(model="logreg", protocol="PULL", passthrough=True, mode="UPDATE", sync=False)])
def test_native_executor_pass_through_nodes(kw):
    """PassThroughNode (reference node.py:289-392) from C++: keyed accept draw against the degrees, PASS or merge."""
    import gossipy_b200 as g
    sim_a, rep_a = _sim(False, n=7, **kw)
    sim_b, rep_b = _sim(True, n=7, **kw)
    assert "_stream_exec" in sim_b.__dict__ and "_stream_exec" not in sim_a.__dict__
    _same(sim_a, rep_a, sim_b, rep_b)
    draws = [getattr(nd, "_pt_draws", 0) for nd in sim_a.nodes.values()]
    assert draws == [getattr(nd, "_pt_draws", 0) for nd in sim_b.nodes.values()] and sum(draws) > 0
    g.CACHE.clear()


@pytest.mark.parametrize("kw", [dict(model="logreg", protocol="PUSH", sampled=.3, faults=True),
                                dict(model="mlp", protocol="PUSH_PULL", sampled=.1),
                                dict(model="logreg", protocol="PULL", sampled=.5, faults=True, sync=False),
                                # UPDATE: the receiver draws the sample, trains a private copy of the received model and merges
                                # the copy's sampled coordinates; its own age does not move
                                dict(model="logreg", protocol="PUSH", sampled=.3, faults=True, mode="UPDATE", tokenized=True),
                                dict(model="mlp", protocol="PUSH_PULL", sampled=.1, mode="UPDATE"),
                                dict(model="logreg", protocol="PULL", sampled=.5, faults=True, sync=False, mode="UPDATE")])
def test_native_executor_sampled_models(kw):
    """SamplingBasedNode + SamplingTMH (reference node.py:499-562, handler.py:426-452) from C++: the receiver's keyed
    coordinate sample (with replacement), indexed merge, local update."""
    import gossipy_b200 as g
    sim_a, rep_a = _sim(False, **kw)
    sim_b, rep_b = _sim(True, **kw)
    assert "_stream_exec" in sim_b.__dict__ and "_stream_exec" not in sim_a.__dict__
    _same(sim_a, rep_a, sim_b, rep_b)
    g.CACHE.clear()


@pytest.mark.parametrize("kw", [dict(model="logreg", protocol="PUSH", cacheneigh=True, faults=True),
                                dict(model="mlp", protocol="PUSH_PULL", cacheneigh=True),
                                dict(model="logreg", protocol="PUSH_PULL", cacheneigh=True, mode="UPDATE", limited=None, faults=True, sync=False)])
def test_native_executor_cache_neighbour_nodes(kw):
    """CacheNeighNode (reference node.py:395-496) from C++: deliveries are stored per sender, one cached model (keyed
    choice) is consumed before every PUSH / PUSH_PULL snapshot."""
    import gossipy_b200 as g
    sim_a, rep_a = _sim(False, n=7, **kw)
    sim_b, rep_b = _sim(True, n=7, **kw)
    assert "_stream_exec" in sim_b.__dict__ and "_stream_exec" not in sim_a.__dict__
    draws = [getattr(nd, "_cn_draws", 0) for nd in sim_a.nodes.values()]
    assert draws == [getattr(nd, "_cn_draws", 0) for nd in sim_b.nodes.values()] and sum(draws) > 0
    _same(sim_a, rep_a, sim_b, rep_b)
    g.CACHE.clear()


@pytest.mark.parametrize("kw", [dict(model="logreg", protocol="PUSH", partitioned=4, faults=True), dict(model="logreg", protocol="PUSH_PULL", sampled=.3, faults=True),
                                dict(model="logreg", protocol="PUSH", passthrough=True, faults=True, n=7),
                                dict(model="logreg", protocol="PUSH_PULL", partitioned=4, faults=True, mode="UPDATE"),
                                dict(model="logreg", protocol="PUSH_PULL", sampled=.3, faults=True, mode="UPDATE")])
def test_executor_checkpoint_of_keyed_node_classes_is_exact(kw, tmp_path):
    """Node classes with keyed draws (partition ids, samples, accept draws): building the scheduler again after a load must not
    consume a draw (it used to: the partitioned resume was off by one message counter)."""
    import gossipy_b200 as g
    from gossipy_b200.simul import GossipSimulator
    sim_full, rep_full = _sim(True, rounds=6, **kw)
    sim, rep = _sim(True, start=False, **kw)
    sim.start(3)
    path = str(tmp_path / "ck.pkl")
    sim.save(path)
    g.CACHE.clear()
    sim2 = GossipSimulator.load(path)
    rep2 = [r for r in sim2._receivers if type(r).__name__ == "SimulationReport"][0]
    sim2.start(3, resume=True)
    _same(sim_full, rep_full, sim2, rep2)
    g.CACHE.clear()


def test_cache_neighbour_executor_checkpoint_keeps_the_caches(tmp_path):
    import gossipy_b200 as g
    from gossipy_b200.simul import GossipSimulator
    kw = dict(model="logreg", protocol="PUSH", cacheneigh=True, faults=True, n=7)
    sim_full, rep_full = _sim(True, rounds=6, **kw)
    sim, rep = _sim(True, start=False, **kw)
    sim.start(3)
    assert len(sim._stream_exec.ex.caches()) > 0
    path = str(tmp_path / "ck.pkl")
    sim.save(path)
    g.CACHE.clear()
    sim2 = GossipSimulator.load(path)
    rep2 = [r for r in sim2._receivers if type(r).__name__ == "SimulationReport"][0]
    sim2.start(3, resume=True)
    assert "_stream_exec" in sim2.__dict__
    _same(sim_full, rep_full, sim2, rep2)
    g.CACHE.clear()


@pytest.mark.parametrize("kw", [dict(momentum={"momentum": .9}, protocol="PUSH_PULL"), dict(momentum={"momentum": .8, "nesterov": True}, protocol="PUSH", faults=True),
                                dict(momentum={"momentum": .9, "dampening": .1}, protocol="PULL", cacheneigh=False, passthrough=True)])
def test_native_executor_momentum_sgd(kw):
    """torch.optim.SGD with momentum (fused into the tensor-core kernel on a GPU) from C++: pair merge, then the momentum
    update with the node's own buffer row (MERGE_UPDATE keeps the optimizer state at the node)."""
    import gossipy_b200 as g
    sim_a, rep_a = _sim(False, model="mlp", **kw)
    sim_b, rep_b = _sim(True, model="mlp", **kw)
    assert "_stream_exec" in sim_b.__dict__ and "_stream_exec" not in sim_a.__dict__
    _same(sim_a, rep_a, sim_b, rep_b)
    for i in sim_a.nodes:                       # the momentum buffers carry the same state
        ba, bb = sim_a.nodes[i].model_handler._opt_rows.get("momentum"), sim_b.nodes[i].model_handler._opt_rows.get("momentum")
        if ba is not None:
            assert bb is not None and torch.equal(ba, bb)
        else:                                   # never updated: the executor's buffer exists but is marked as stateless
            assert bb is None or sim_b.nodes[i].model_handler.__dict__.get("_mom_pending")
    g.CACHE.clear()


def test_executor_race_debug_mode(monkeypatch):
    """GOSSIPY_EXEC_DEBUG=1: slot life-cycle assertions (one writer, one reader per life; no slot both free and on the
    wire; no message id twice) hold on a faulty run and fire on a forged event list."""
    import gossipy_b200 as g
    from gossipy_b200.ops.native import _try_import
    monkeypatch.setenv("GOSSIPY_EXEC_DEBUG", "1")
    sim, rep = _sim(True, "logreg", "PUSH_PULL", True, n=7, rounds=5)
    assert sim._stream_exec.ex.debug
    C = _try_import()
    ex = C.StreamExecutor(2, 1, 4, 0, 2, 2, 1, .1, 0., 5, False, 2, -1)
    for i in range(2):
        ex.set_node(i, 0, 0, 0, 6, 0, 0, 0)
    ex.set_callbacks(lambda *a: None, lambda *a: None, lambda *a: None)
    ex.set_slots(0, 4, 8, 8)
    ev = np.array([[C.EV_SEND, 0, 0, 1, 100, 1], [C.EV_SEND, 0, 0, 1, 100, 1]], dtype=np.int32)      # the same message id twice
    with pytest.raises(Exception, match="executor debug"):
        ex.run(ev, 0)
    g.CACHE.clear()


def test_slot_pool_grows_and_resume_is_exact(tmp_path):
    import gossipy_b200 as g
    from gossipy_b200.simul import GossipSimulator
    sim_full, rep_full = _sim(True, "logreg", "PUSH_PULL", True, n=12, rounds=6)
    # same run, interrupted after 3 rounds, checkpointed with messages on the wire, resumed; tiny slot pool
    sim, rep = _sim(True, "logreg", "PUSH_PULL", True, n=12, rounds=0, start=False)
    sim.start(3)
    sx = sim._stream_exec
    assert len(sx.ex.inflight()) > 0
    path = str(tmp_path / "ck.pkl")
    sim.save(path)
    g.CACHE.clear()
    sim2 = GossipSimulator.load(path)
    assert "_stream_exec" not in sim2.__dict__ and "_exec_inflight" in sim2.__dict__
    rep2 = [r for r in sim2._receivers if type(r).__name__ == "SimulationReport"][0]
    sim2.start(3, resume=True)
    _same(sim_full, rep_full, sim2, rep2)
    # growing the pool in the middle of a round keeps the in-flight snapshots
    sx2 = sim2._stream_exec
    before = int(sx2.slots.shape[0])
    sx2._grow()
    assert int(sx2.slots.shape[0]) == 2 * before and sx2.ex.free_slots >= before
    g.CACHE.clear()


def test_eligibility():
    from gossipy_b200.engine.stream_exec import eligible
    from gossipy_b200.core import CreateModelMode
    sim, _ = _sim(True, start=False)
    assert eligible(sim) is None
    sim.nodes[2].model_handler.mode = CreateModelMode.UPDATE        # nodes must agree
    assert eligible(sim) is not None
    for nd in sim.nodes.values():
        nd.model_handler.mode = CreateModelMode.UPDATE_MERGE
    assert eligible(sim) is None                   # all four modes run natively
    import gossipy_b200 as g
    g.GlobalSettings().reference_compat = True     # bug-for-bug behaviours live in the Python handlers
    try:
        assert eligible(sim) is not None
        sim.start(1)                               # falls back to the per-event executor
    finally:
        g.GlobalSettings().reference_compat = False
    assert "_stream_exec" not in sim.__dict__


def test_executor_runs_out_of_slots_gracefully(monkeypatch):
    """Direct use of the C++ class: `run` stops at the event it cannot serve and resumes after the pool grew."""
    from gossipy_b200.ops.native import _try_import
    monkeypatch.setenv("GOSSIPY_EXEC_ELIDE", "0")       # (with elision these two messages need no slot at all)
    C = _try_import()
    log = []
    ex = C.StreamExecutor(3, 1, 4, 0, 2, 2, 1, .1, 0., 5, False, 2, -1)
    for i in range(3):
        ex.set_node(i, 0, 0, 0, 6, 10 * i, 0, 0)
    ex.set_callbacks(lambda n, r, s, g, rr: log.append(("snap", n, s)), lambda n, r, s, k, ws, wp, g: log.append(("train", n, s)),
                     lambda n, r, s, g: log.append(("adopt", n, s)))
    ex.set_slots(0, 1, 8, 8)
    ev = np.array([[C.EV_SEND, 0, 0, 1, 100, 1], [C.EV_SEND, 0, 2, 1, 101, 1], [C.EV_DELIVER, 0, 0, 1, 100, 1],
                   [C.EV_DELIVER, 0, 2, 1, 101, 1], [C.EV_EVAL, 0, 1, -1, -1, 0]], dtype=np.int32)
    assert list(ex.run(ev, 0)) == [] and ex.resume_at == 1
    ex.set_slots(0, 4, 8, 8)
    assert list(ex.run(ev, 1)) == [1] and ex.resume_at == -1
    assert [e[0] for e in log] == ["snap", "snap", "train", "train"]
    assert ex.ages() == [0, 20 + 3, 20] and ex.counters() == [0, 2, 0] and ex.inflight() == []


def test_snapshot_elision_only_when_provably_safe(monkeypatch):
    """A message travels as a reference to the sender's live row iff it is delivered before the sender's next write and
    that write does not have to wait for the reader (csrc/exec/executor.cpp::can_alias)."""
    from gossipy_b200.ops.native import _try_import
    C = _try_import()

    def run(ev, n=3):
        log = []
        ex = C.StreamExecutor(n, 1, 4, 0, 2, 2, 1, .1, 0., 5, False, 2, -1)
        for i in range(n):
            ex.set_node(i, 0, 0, 0, 6, 10 * i, 0, 0)
        ex.set_callbacks(lambda nd, r, s, g, rr: log.append(("snap", nd)), lambda nd, r, s, k, ws, wp, g: log.append(("train", nd, r)),
                         lambda nd, r, s, g: log.append(("adopt", nd)))
        ex.set_slots(0, 4, 8, 8)
        ex.run(np.asarray(ev, dtype=np.int32), 0)
        return ex, log
    PUSH, PUSH_PULL, REPLY = 0, 3, 2          # message types (core.MessageType values)
    from gossipy_b200.core import MessageType
    PUSH, PUSH_PULL, REPLY = MessageType.PUSH.value, MessageType.PUSH_PULL.value, MessageType.REPLY.value
    # push-pull, no delay: the request is read live (the sender's next write is the delivery of the reply, which waits
    # for the reader anyway); the reply travels as a snapshot (node 1's next write must not wait for node 0's kernel)
    pp = [[C.EV_SEND, 0, 0, 1, 100, PUSH_PULL], [C.EV_DELIVER, 0, 0, 1, 100, PUSH_PULL], [C.EV_REPLY_SEND, 0, 0, 1, 100, 101],
          [C.EV_REPLY_DELIVER, 0, 0, 1, 101, REPLY]]
    ex, log = run(pp)
    assert ex.elided == 1 and log == [("train", 1, -1), ("snap", 1), ("train", 0, 0)]
    # the sender is written (a delivery from node 2) before its own message arrives: a real snapshot is taken
    ex, log = run([[C.EV_SEND, 0, 0, 1, 100, PUSH_PULL], [C.EV_SEND, 0, 2, 0, 101, PUSH], [C.EV_DELIVER, 0, 2, 0, 101, PUSH],
                   [C.EV_DELIVER, 0, 0, 1, 100, PUSH_PULL], [C.EV_REPLY_SEND, 0, 0, 1, 100, 102], [C.EV_REPLY_DELIVER, 0, 0, 1, 102, REPLY]])
    assert ex.elided == 0 and [e for e in log if e[0] == "snap"][0] == ("snap", 0)
    # plain PUSH: the sender's next write (if any) is unrelated to the reader -> it must not wait for the reader's kernel
    ex, log = run([[C.EV_SEND, 0, 0, 1, 100, PUSH], [C.EV_DELIVER, 0, 0, 1, 100, PUSH], [C.EV_SEND, 0, 2, 0, 101, PUSH],
                   [C.EV_DELIVER, 0, 2, 0, 101, PUSH]])
    assert ex.elided == 0 and ("snap", 0) in log and ("snap", 2) in log
    # an unrelated delivery reaches the sender between the request and its reply: snapshot
    ex, log = run(pp[:2] + [[C.EV_SEND, 0, 2, 0, 103, PUSH], [C.EV_DELIVER, 0, 2, 0, 103, PUSH]] + pp[2:])
    assert ex.elided == 0
    # not delivered in this round (delay): snapshot; dropped: nothing is ever read, no copy
    ex, log = run([[C.EV_SEND, 0, 0, 1, 100, PUSH]])
    assert log == [("snap", 0)] and ex.elided == 0 and len(ex.inflight()) == 1
    ex, log = run([[C.EV_SEND, 0, 0, 1, 100, PUSH], [C.EV_DROP, 0, 0, 1, 100, PUSH]])
    assert log == [] and ex.elided == 1 and ex.inflight() == []
    monkeypatch.setenv("GOSSIPY_EXEC_ELIDE", "0")
    ex, log = run(pp)
    assert ex.elided == 0 and log[0] == ("snap", 0)


@pytest.mark.gpu
def test_native_executor_cuda_equals_python_executor():
    import gossipy_b200 as g
    for model, protocol, faults in (("mlp", "PUSH_PULL", False), ("logreg", "PUSH", True)):
        sim_a, rep_a = _sim(False, model, protocol, faults, n=8, rounds=5, device="cuda:0")
        sim_b, rep_b = _sim(True, model, protocol, faults, n=8, rounds=5, device="cuda:0")
        torch.cuda.synchronize()
        assert "_stream_exec" in sim_b.__dict__
        _same(sim_a, rep_a, sim_b, rep_b, tol=1e-6)
        g.CACHE.clear()
    for kw in (dict(model="mlp", protocol="PUSH_PULL", mode="UPDATE_MERGE"), dict(model="logreg", protocol="PUSH", mode="UPDATE_MERGE", limited=3, faults=True),
               dict(model="mlp", protocol="PUSH", partitioned=4, faults=True), dict(model="mlp", protocol="PUSH_PULL", sampled=.2),
               dict(model="logreg", protocol="PUSH", passthrough=True, faults=True), dict(model="mlp", protocol="PUSH_PULL", cacheneigh=True),
               dict(model="mlp", protocol="PUSH_PULL", momentum={"momentum": .9})):
        sim_a, rep_a = _sim(False, n=8, rounds=4, device="cuda:0", **kw)
        sim_b, rep_b = _sim(True, n=8, rounds=4, device="cuda:0", **kw)
        torch.cuda.synchronize()
        assert "_stream_exec" in sim_b.__dict__
        _same(sim_a, rep_a, sim_b, rep_b, tol=1e-5)
        g.CACHE.clear()
    for kw in (dict(model="mlp", n=8, sync=True), dict(model="logreg", n=6, faults=True, mixing="ring")):
        sim_a, rep_a = _a2a_sim(False, device="cuda:0", **kw)
        sim_b, rep_b = _a2a_sim(True, device="cuda:0", **kw)
        torch.cuda.synchronize()
        assert "_stream_exec" in sim_b.__dict__
        _same(sim_a, rep_a, sim_b, rep_b, tol=1e-5)
        g.CACHE.clear()
    g.GlobalSettings().set_device("cpu")


def test_native_executor_equals_python_executor_random_setups():
    """Derandomised sweep over protocol x mode x faults x clocks x merge rule (hypothesis); EXEC_EXAMPLES to stress."""
    import os
    import gossipy_b200 as g
    from hypothesis import HealthCheck, given, settings, strategies as st

    @settings(max_examples=int(os.environ.get("EXEC_EXAMPLES", "12")), deadline=None, derandomize=True, database=None,
              suppress_health_check=list(HealthCheck))
    @given(protocol=st.sampled_from(["PUSH", "PULL", "PUSH_PULL"]), mode=st.sampled_from(["MERGE_UPDATE", "UPDATE", "UPDATE_MERGE", "PASS"]),
           faults=st.booleans(), sync=st.booleans(), limited=st.sampled_from([None, 0, 3, 50]), tokenized=st.booleans(),
           n=st.integers(2, 9), rounds=st.integers(1, 4), node=st.sampled_from(["gossip", "gossip", "passthrough", "cacheneigh"]))
    def check(protocol, mode, faults, sync, limited, tokenized, n, rounds, node):
        kw = dict(model="logreg", protocol=protocol, mode=mode, faults=faults, sync=sync, limited=limited,
                  tokenized=tokenized, n=max(n, 3) if node == "passthrough" else n, rounds=rounds,
                  passthrough=node == "passthrough", cacheneigh=node == "cacheneigh")
        sim_a, rep_a = _sim(False, **kw)
        sim_b, rep_b = _sim(True, **kw)
        assert "_stream_exec" in sim_b.__dict__
        ra, aa, ca = _state(sim_a)
        rb, ab, cb = _state(sim_b)
        assert aa == ab and ca == cb and torch.equal(ra, rb)
        assert (rep_a._sent_messages, rep_a._failed_messages, rep_a._total_size) == \
            (rep_b._sent_messages, rep_b._failed_messages, rep_b._total_size)
        g.CACHE.clear()
    check()


def _a2a_sim(streamed, model="logreg", n=6, rounds=4, device="cpu", faults=False, sync=False, mixing="uniform", start=True):
    import gossipy_b200 as g
    from gossipy_b200.core import AntiEntropyProtocol, CreateModelMode, StaticP2PNetwork, UniformDelay, UniformMixing
    from gossipy_b200.data import DataDispatcher, synthetic
    from gossipy_b200.data.handler import ClassificationDataHandler
    from gossipy_b200.model.handler import WeightedTMH
    from gossipy_b200.model.nn import LogisticRegression, TorchMLP
    from gossipy_b200.node import All2AllGossipNode
    from gossipy_b200.simul import All2AllGossipSimulator, SimulationReport
    g.GlobalSettings().set_device(device)
    g.CACHE.clear()
    g.set_seed(7)
    if model == "mlp":
        (Xtr, ytr), (Xte, yte) = synthetic.mnist_like(96 * n, 120)
        net, bs = TorchMLP(784, 10, (100,)), 32
    else:
        (Xtr, ytr), (Xte, yte) = synthetic.spambase_like(50 * n + 3, 150)
        net, bs = LogisticRegression(57, 2), 16
    disp = DataDispatcher(ClassificationDataHandler(Xtr, ytr, Xte, yte), n=n, eval_on_user=False)
    proto = WeightedTMH(net, torch.optim.SGD, {"lr": .1, "weight_decay": .001}, torch.nn.CrossEntropyLoss(), batch_size=bs,
                        create_model_mode=CreateModelMode.MERGE_UPDATE)
    if mixing == "ring":            # sparse topology: each node hears from two neighbours only
        A = np.zeros((n, n), dtype=int)
        for i in range(n):
            A[i, (i + 1) % n] = A[(i + 1) % n, i] = 1
        topo = StaticP2PNetwork(n, A)
    else:
        topo = StaticP2PNetwork(n)
    nodes = All2AllGossipNode.generate(disp, topo, proto, 10, sync)
    kw = dict(drop_prob=.2, online_prob=.8, delay=UniformDelay(0, 4), sampling_eval=.5) if faults else {}
    sim = All2AllGossipSimulator(nodes, disp, 10, AntiEntropyProtocol.PUSH, **kw)
    sim.progress = False
    sim.engine = "native"
    sim.native_executor = streamed
    rep = SimulationReport(); sim.add_receiver(rep)
    sim.init_nodes(seed=11)
    sim._mix = UniformMixing(topo)
    if start:
        sim.start(sim._mix, rounds)
    return sim, rep


@pytest.mark.parametrize("kw", [dict(), dict(faults=True), dict(model="mlp", sync=True), dict(mixing="ring", faults=True, n=7)])
def test_native_executor_all_to_all_nodes(kw):
    """All2AllGossipNode + WeightedTMH (reference node.py:789-870, handler.py:642-688) from C++: per-sender caches, shared
    snapshot of a timeout's pushes (reference counts), k-way merge with renormalised mixing weights on timeout, update."""
    import gossipy_b200 as g
    sim_a, rep_a = _a2a_sim(False, **kw)
    sim_b, rep_b = _a2a_sim(True, **kw)
    assert "_stream_exec" in sim_b.__dict__ and "_stream_exec" not in sim_a.__dict__
    _same(sim_a, rep_a, sim_b, rep_b)
    g.CACHE.clear()


def test_all_to_all_executor_checkpoint_keeps_the_caches(tmp_path):
    import gossipy_b200 as g
    from gossipy_b200.simul import All2AllGossipSimulator
    sim_full, rep_full = _a2a_sim(True, faults=True, rounds=6)
    sim, rep = _a2a_sim(True, faults=True, start=False)
    sim.start(sim._mix, 3)
    assert len(sim._stream_exec.ex.caches()) > 0
    path = str(tmp_path / "ck.pkl")
    sim.save(path)
    g.CACHE.clear()
    sim2 = All2AllGossipSimulator.load(path)
    rep2 = [r for r in sim2._receivers if type(r).__name__ == "SimulationReport"][0]
    sim2.start(sim._mix, 3, resume=True)
    _same(sim_full, rep_full, sim2, rep2)
    g.CACHE.clear()


def test_native_executor_variants_random_setups():
    """Derandomised sweep over the executor's other modes (all-to-all, sampled, partitioned, momentum) x protocol x faults x
    clocks x sizes; EXEC_EXAMPLES to stress (1 500 random set-ups pass, also with GOSSIPY_EXEC_DEBUG=1)."""
    import os
    import gossipy_b200 as g
    from hypothesis import HealthCheck, given, settings, strategies as st

    @settings(max_examples=int(os.environ.get("EXEC_EXAMPLES", "10")), deadline=None, derandomize=True, database=None,
              suppress_health_check=list(HealthCheck))
    @given(kind=st.sampled_from(["a2a", "sampled", "part", "momentum"]), model=st.sampled_from(["logreg", "mlp"]),
           protocol=st.sampled_from(["PUSH", "PULL", "PUSH_PULL"]), faults=st.booleans(), sync=st.booleans(), n=st.integers(2, 8),
           rounds=st.integers(1, 4), tokenized=st.booleans(), frac=st.sampled_from([.05, .3, 1.0]), parts=st.sampled_from([2, 4, 7]),
           ring=st.booleans(), nesterov=st.booleans(), update=st.booleans())
    def check(kind, model, protocol, faults, sync, n, rounds, tokenized, frac, parts, ring, nesterov, update):
        if kind == "a2a":
            kw = dict(model=model, n=max(n, 3) if ring else n, rounds=rounds, faults=faults, sync=sync, mixing="ring" if ring else "uniform")
            sim_a, rep_a = _a2a_sim(False, **kw)
            sim_b, rep_b = _a2a_sim(True, **kw)
        else:
            kw = dict(model=model, protocol=protocol, faults=faults, sync=sync, n=n, rounds=rounds, tokenized=tokenized)
            if kind == "sampled":
                kw.update(sampled=frac, mode="UPDATE" if update else "MERGE_UPDATE")
            elif kind == "part":
                kw.update(partitioned=parts, mode="UPDATE" if update else "MERGE_UPDATE")
            else:
                kw.update(model="mlp", momentum={"momentum": .9, "nesterov": nesterov})
            sim_a, rep_a = _sim(False, **kw)
            sim_b, rep_b = _sim(True, **kw)
        assert "_stream_exec" in sim_b.__dict__, (kind, kw)
        _same(sim_a, rep_a, sim_b, rep_b)
        g.CACHE.clear()
    check()
