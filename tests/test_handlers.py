"""Handler math vs the reference: merges (C1-C7), local updates (K1-K6), evaluation."""
import copy

import numpy as np
import pytest
import torch

import gossipy_b200 as g
from gossipy_b200 import CACHE
from gossipy_b200.core import CreateModelMode as M
from gossipy_b200.model import handler as H
from gossipy_b200.model.nn import AdaLine, LogisticRegression, TorchMLP
from gossipy_b200.model.sampling import TorchModelPartition, TorchModelSampling
from gossipy_b200.models import ResNet20

CE = torch.nn.CrossEntropyLoss()


def _data(n=96, d=12, c=3, seed=0):
    gen = torch.Generator().manual_seed(seed)
    X = torch.randn(n, d, generator=gen)
    return X, (X @ torch.randn(d, c, generator=gen)).argmax(1)


def _pair(ref, cls_name, net_fn, **kw):
    """Our handler and the reference's, with identical initial weights."""
    import gossipy.model.handler as RH
    import gossipy.model.nn as RN
    ours_net = net_fn(g.model.nn)
    theirs_net = net_fn(RN)
    theirs_net.load_state_dict(ours_net.state_dict())
    common = dict(optimizer=torch.optim.SGD, optimizer_params={"lr": .1, "weight_decay": .01},
                  criterion=CE, batch_size=0, local_epochs=1)
    common.update(kw)
    ours = getattr(H, cls_name)(net=ours_net, **common)
    okw = dict(common)
    if "create_model_mode" in okw:
        okw["create_model_mode"] = getattr(RH.CreateModelMode, okw["create_model_mode"].name)
    theirs = getattr(RH, cls_name)(net=theirs_net, **okw)
    return ours, theirs


def _flat(module):
    return torch.cat([p.detach().reshape(-1) for p in module.parameters()])


def _set_random(ours, theirs, seed):
    gen = torch.Generator().manual_seed(seed)
    for p in theirs.model.parameters():
        p.data = torch.randn(p.shape, generator=gen)
    ours.model.load_state_dict(theirs.model.state_dict())
    ours.model = ours.model  # re-gather into the row


def test_fullbatch_update_and_uniform_merge_match_reference(ref):
    X, y = _data()
    mk = lambda ns: ns.TorchMLP(12, 3, (8,))
    a, ra = _pair(ref, "TorchModelHandler", mk)
    b, rb = _pair(ref, "TorchModelHandler", mk)
    _set_random(a, ra, 1); _set_random(b, rb, 2)
    for h in (a, ra, b, rb):
        h._update((X, y))
    assert a.n_updates == ra.n_updates == 1
    torch.testing.assert_close(_flat(a.model), _flat(ra.model), rtol=1e-5, atol=1e-6)
    a._merge(b); ra._merge(rb)
    torch.testing.assert_close(_flat(a.model), _flat(ra.model), rtol=1e-5, atol=1e-6)
    a._merge([b, b]); ra._merge([rb, rb])
    torch.testing.assert_close(_flat(a.model), _flat(ra.model), rtol=1e-5, atol=1e-6)
    ev, rev = a.evaluate((X, y)), ra.evaluate((X, y))
    for k in rev:
        assert ev[k] == pytest.approx(float(rev[k]), abs=1e-6), k


def test_generic_path_matches_fused_path():
    """Same seed/key -> the autograd path and the explicit fused math give the same weights."""
    X, y = _data(100, 12, 3)
    net = TorchMLP(12, 3, (8,)); net.init_weights()
    fused = H.TorchModelHandler(net, torch.optim.SGD, {"lr": .05, "weight_decay": .001}, CE,
                                local_epochs=2, batch_size=32)
    generic = fused.copy()
    generic._fused = False
    assert fused._fused
    fused.owner = generic.owner = 3
    fused._update((X, y)); generic._update((X, y))
    assert fused.n_updates == generic.n_updates == 8
    torch.testing.assert_close(fused.row, generic.row, rtol=1e-4, atol=1e-6)


def test_grad_views_stay_bound():
    X, y = _data()
    h = H.TorchModelHandler(TorchMLP(12, 3, (8, 8)), torch.optim.SGD, {"lr": .1, "momentum": .9}, CE)
    assert not h._fused
    h._update((X, y))
    ptrs = [p.grad.data_ptr() for p in h.model.parameters()]
    h._update((X, y))
    assert ptrs == [p.grad.data_ptr() for p in h.model.parameters()]
    base = h._grad_row.data_ptr()
    assert all(base <= q < base + h._grad_row.numel() * 4 for q in ptrs)


@pytest.mark.parametrize("opt,params", [(torch.optim.SGD, {"lr": .1, "momentum": .9, "nesterov": True}),
                                        (torch.optim.SGD, {"lr": .1, "momentum": .5, "dampening": .1,
                                                           "weight_decay": .01}),
                                        (torch.optim.Adam, {"lr": .01, "weight_decay": .01}),
                                        (torch.optim.AdamW, {"lr": .01}),
                                        (torch.optim.RMSprop, {"lr": .01})])
def test_flat_optimizers_match_torch(opt, params):
    X, y = _data()
    net = TorchMLP(12, 3, (8,)); net.init_weights()
    h = H.TorchModelHandler(net, opt, params, CE, batch_size=0, local_epochs=3)
    twin = copy.deepcopy(net)
    topt = opt(twin.parameters(), **params)
    for _ in range(3):
        topt.zero_grad(); CE(twin(X), y).backward(); topt.step()
    h._update((X, y))
    torch.testing.assert_close(_flat(h.model), _flat(twin), rtol=2e-4, atol=1e-6)


def test_modes_match_reference(ref):
    X, y = _data()
    mk = lambda ns: ns.LogisticRegression(12, 3)
    for mode in (M.UPDATE, M.MERGE_UPDATE, M.UPDATE_MERGE, M.PASS):
        a, ra = _pair(ref, "TorchModelHandler", mk, create_model_mode=mode)
        b, rb = _pair(ref, "TorchModelHandler", mk, create_model_mode=mode)
        _set_random(a, ra, 3); _set_random(b, rb, 4)
        b.n_updates = rb.n_updates = 5
        a(b.copy(), (X, y)); ra(rb.copy(), (X, y))
        assert a.n_updates == ra.n_updates, mode
        torch.testing.assert_close(_flat(a.model), _flat(ra.model), rtol=1e-5, atol=1e-6, msg=str(mode))


def test_limited_merge_branches(ref):
    mk = lambda ns: ns.LogisticRegression(12, 3)
    for na, nb in ((10, 2), (2, 10), (4, 5), (0, 0)):
        a, ra = _pair(ref, "LimitedMergeTMH", mk, age_diff_threshold=1)
        b, rb = _pair(ref, "LimitedMergeTMH", mk, age_diff_threshold=1)
        _set_random(a, ra, 5); _set_random(b, rb, 6)
        a.n_updates = ra.n_updates = na
        b.n_updates = rb.n_updates = nb
        a._merge(b)
        if (na, nb) != (0, 0):   # the reference divides by zero for two fresh models
            ra._merge(rb)
            torch.testing.assert_close(_flat(a.model), _flat(ra.model), rtol=1e-5, atol=1e-6)
            assert a.n_updates == ra.n_updates
        else:
            assert torch.isfinite(a.row).all()


def test_weighted_kway_merge(ref):
    mk = lambda ns: ns.LogisticRegression(12, 3)
    hs = [_pair(ref, "WeightedTMH", mk) for _ in range(4)]
    for i, (o, r) in enumerate(hs):
        _set_random(o, r, 10 + i)
    w = np.array([.4, .3, .2, .1])
    hs[0][0]._merge([h[0] for h in hs[1:]], w)
    hs[0][1]._merge([h[1] for h in hs[1:]], w)
    torch.testing.assert_close(_flat(hs[0][0].model), _flat(hs[0][1].model), rtol=1e-5, atol=1e-6)


def test_partition_index_sets_and_merge_match_reference(ref):
    import gossipy.model.sampling as RS
    import gossipy.model.nn as RN
    for net_o, net_r, parts in ((LogisticRegression(5, 2), RN.LogisticRegression(5, 2), 4),
                                (TorchMLP(7, 3, (5,)), RN.TorchMLP(7, 3, (5,)), 3),
                                (TorchMLP(7, 3, (5,)), RN.TorchMLP(7, 3, (5,)), 7)):
        po, pr = TorchModelPartition(net_o, parts), RS.TorchModelPartition(net_r, parts)
        assert po.n_parts == pr.n_parts
        flat_all = []
        for p in range(po.n_parts):
            for ti, ids in pr.partitions[p].items():
                mine = po.partitions[p][ti]
                if ids is None:
                    assert mine is None
                    continue
                a = sorted(zip(*[x.tolist() for x in ids]))
                b = sorted(zip(*[x.tolist() for x in mine]))
                assert a == b, (p, ti)
            # segments cover exactly the flat index set
            cover = []
            for st, nr, rl, sd in po.segments(p).tolist():
                cover += [st + r * sd + c for r in range(nr) for c in range(rl)]
            assert sorted(cover) == po.flat_index(p).tolist()
            flat_all += cover
        assert sorted(flat_all) == list(range(po.n_params))


def test_partitioned_handler_matches_reference(ref):
    import gossipy.model.sampling as RS
    X, y = _data()
    mk = lambda ns: ns.TorchMLP(12, 3, (6,))
    import gossipy.model.nn as RN
    a, ra = _pair(ref, "PartitionedTMH", mk, tm_partition=TorchModelPartition(mk(g.model.nn), 4))
    b, rb = _pair(ref, "PartitionedTMH", mk, tm_partition=TorchModelPartition(mk(g.model.nn), 4))
    for r_ in (ra, rb):
        r_.tm_partition = RS.TorchModelPartition(mk(RN), 4)
    _set_random(a, ra, 7); _set_random(b, rb, 8)
    for h in (a, ra):
        h._update((X, y)); h._update((X, y))
    for h in (b, rb):
        h._update((X, y))
    assert a.n_updates.tolist() == ra.n_updates.tolist() == [2] * 4
    torch.testing.assert_close(_flat(a.model), _flat(ra.model), rtol=1e-5, atol=1e-6)
    a._merge(b, 2); ra._merge(rb, 2)
    a._merge(b, 5); ra._merge(rb, 1)      # ours wraps the id modulo n_parts
    torch.testing.assert_close(_flat(a.model), _flat(ra.model), rtol=1e-5, atol=1e-6)
    assert a.n_updates.tolist() == ra.n_updates.tolist()
    assert a.caching(0).key[1] == str(a.n_updates)
    CACHE.clear()


def test_sampled_merge_matches_reference(ref):
    import gossipy.model.sampling as RS
    mk = lambda ns: ns.TorchMLP(12, 3, (6,))
    a, ra = _pair(ref, "SamplingTMH", mk, sample_size=.3)
    b, rb = _pair(ref, "SamplingTMH", mk, sample_size=.3)
    _set_random(a, ra, 9); _set_random(b, rb, 10)
    sample = RS.TorchModelSampling.sample(.3, ra.model)
    a._merge(b, sample); ra._merge(rb, sample)
    torch.testing.assert_close(_flat(a.model), _flat(ra.model), rtol=1e-6, atol=1e-7)
    flat = a.draw_sample()
    assert flat.numel() == round(.3 * a.get_size()) and int(flat.max()) < a.get_size()
    d = TorchModelSampling.sample(.5, a.model)
    assert sum(v[0].numel() for v in d.values() if v is not None) == round(.5 * a.get_size())


def test_adaline_and_pegasos_match_reference(ref):
    import gossipy.model.handler as RH
    import gossipy.model.nn as RN
    gen = torch.Generator().manual_seed(0)
    X = torch.randn(40, 9, generator=gen)
    y = torch.sign(X @ torch.randn(9, generator=gen))
    for name, lr in (("AdaLineHandler", .01), ("PegasosHandler", .01)):
        o = getattr(H, name)(AdaLine(9), lr)
        r = getattr(RH, name)(RN.AdaLine(9), lr)
        o.init(); r.init()
        o._update((X, y)); r._update((X, y))
        o._update((X[:7], y[:7])); r._update((X[:7], y[:7]))
        assert o.n_updates == r.n_updates == 47
        torch.testing.assert_close(o.model.model.detach(), r.model.model.detach(), rtol=1e-4, atol=1e-5)
        o2 = o.copy(); o2._update((X[:3], y[:3]))
        r2 = copy.deepcopy(r); r2._update((X[:3], y[:3]))
        o._merge(o2); r._merge(r2)
        torch.testing.assert_close(o.model.model.detach(), r.model.model.detach(), rtol=1e-4, atol=1e-5)
        eo, er = o.evaluate((X, y)), r.evaluate((X, y))
        for k in er:
            assert eo[k] == pytest.approx(float(er[k]), abs=1e-6)


def test_kmeans_and_mf_match_reference(ref):
    import gossipy.model.handler as RH
    gen = torch.Generator().manual_seed(0)
    X = torch.randn(60, 5, generator=gen); yl = (X[:, 0] > 0).long()
    o, r = H.KMeansHandler(3, 5, alpha=.2), RH.KMeansHandler(3, 5, alpha=.2)
    o.init(); r.init()
    o.model = r.model.clone()
    o._update((X[:1], None)); r._update((X[:1], None))
    o._update((X[1:9], None)); r._update((X[1:9], None))
    torch.testing.assert_close(o.model, r.model)
    assert o.evaluate((X, yl))["nmi"] == pytest.approx(r.evaluate((X, yl))["nmi"])
    o2 = o.copy(); o2.model = o.model.flip(0)
    hung = H.KMeansHandler(3, 5, matching="hungarian"); hung.init(); hung.model = o.model.clone()
    hung._merge(o2)                       # optimal matching undoes the flip (B16 fixed)
    torch.testing.assert_close(hung.model, o.model)
    # matrix factorisation
    ratings = np.array([(1, 4.), (3, 2.), (0, 5.), (1, 1.)])
    mo, mr = H.MFModelHandler(4, 6), RH.MFModelHandler(4, 6)
    mo.init(); mr.init()
    mo.model = copy.deepcopy(mr.model)
    mo._update(ratings); mr._update(ratings)
    np.testing.assert_allclose(mo.model[1][0], mr.model[1][0], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(mo.model[0][0], mr.model[0][0], rtol=1e-5, atol=1e-6)
    assert mo.n_updates == mr.n_updates == 5
    m2, r2 = mo.copy(), copy.deepcopy(mr)
    m2._update(ratings[:2]); r2._update(ratings[:2])
    mo._merge(m2); mr._merge(r2)
    np.testing.assert_allclose(mo.model[1][0], mr.model[1][0], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(mo.model[1][1], mr.model[1][1], rtol=1e-5, atol=1e-6)
    assert mo.evaluate(ratings)["rmse"] == pytest.approx(mr.evaluate(ratings)["rmse"], rel=1e-5)
    assert mo.get_size() == mr.get_size() == 4 * 7


def test_batchnorm_model_merges():
    """B12: the reference crashes merging BN models; here float buffers average, counters max."""
    gen = torch.Generator().manual_seed(0)
    X = torch.rand(16, 3, 16, 16, generator=gen); y = torch.randint(0, 10, (16,), generator=gen)
    net = ResNet20(width=4); net.init_weights()
    a = H.TorchModelHandler(net, torch.optim.SGD, {"lr": .05}, CE, batch_size=8)
    b = a.copy()
    a._update((X, y)); a._update((X, y)); b._update((X, y))
    snap_key = b.caching(1)
    snap = CACHE.pop(snap_key)
    ra, rb = a.row.clone(), snap.row.clone()
    a._merge(snap)
    torch.testing.assert_close(a.row, (ra + rb) / 2)
    cnt = dict(a.model.named_buffers())["stem.1.num_batches_tracked"]
    assert int(cnt) == 4    # max(2 updates * 2 batches, 1 * 2)
    assert a.layout.numel > a.layout.n_params == net.get_size()


def test_snapshot_dedupe_and_release():
    X, y = _data()
    h = H.TorchModelHandler(TorchMLP(12, 3, (8,)), torch.optim.SGD, {"lr": .1}, CE)
    h.owner = 0
    h._update((X, y))
    k1, k2 = h.caching(0), h.caching(0)
    assert k1 == k2 and len(CACHE) == 1          # identical version: one snapshot, two refs
    s1 = CACHE.pop(k1); s1.release()
    assert s1._row is not None                   # still referenced by the second message
    s2 = CACHE.pop(k2); s2.release()
    assert s2._row is None and len(CACHE) == 0
    h._update((X, y))
    assert h.caching(0) != k1
    CACHE.clear()


def test_momentum_oracle_equals_torch_optim_sgd():
    """The explicit momentum update of the fused-path oracle (ops.torch_ref.mlp1_train) is torch.optim.SGD's."""
    import torch
    from gossipy_b200.ops import torch_ref as ref
    from gossipy_b200.model.nn import TorchMLP
    torch.manual_seed(3)
    dims = (20, 7, 3)
    net = TorchMLP(20, 3, (7,)).double()
    X = torch.randn(48, 20, dtype=torch.float64); y = torch.randint(0, 3, (48,))
    for kw in (dict(momentum=.9), dict(momentum=.9, nesterov=True), dict(momentum=.5, dampening=.2, weight_decay=.01)):
        m = __import__("copy").deepcopy(net)
        row = torch.cat([p.detach().reshape(-1) for p in m.parameters()]).clone()
        buf = torch.zeros_like(row)
        opt = torch.optim.SGD(m.parameters(), lr=.1, **kw)
        idx_all = list(ref._batches(48, 16, 2, 99))
        for idx in idx_all:
            loss = torch.nn.functional.cross_entropy(m(X[idx]), y[idx])
            opt.zero_grad(); loss.backward(); opt.step()
        ref.mlp1_train(row, X, y, dims, 16, 2, .1, kw.get("weight_decay", 0.), 99, None,
                       (kw["momentum"], kw.get("dampening", 0.), kw.get("nesterov", False), buf, True))
        want = torch.cat([p.detach().reshape(-1) for p in m.parameters()])
        torch.testing.assert_close(row, want, rtol=1e-10, atol=1e-12)


def test_channels_last_rows_are_a_private_layout_choice():
    """``GlobalSettings().channels_last``: conv filters live in the row as [O, kh, kw, I] and are bound as channels-last
    views, mini-batches are gathered as NHWC -- training, merging, evaluation, pickling and the model setter give the
    results of the plain layout."""
    import copy
    import pickle
    import gossipy_b200 as g
    from gossipy_b200.model.handler import PartitionedTMH, TorchModelHandler
    from gossipy_b200.model.nn import CIFAR10Net
    from gossipy_b200.model.sampling import TorchModelPartition

    def run(cl):
        g.GlobalSettings().channels_last = cl
        g.set_seed(3)
        torch.manual_seed(3)
        X, y = torch.rand(96, 3, 32, 32), torch.randint(0, 10, (96,))
        h = TorchModelHandler(CIFAR10Net(), torch.optim.SGD, {"lr": .05, "momentum": .9}, torch.nn.CrossEntropyLoss(), batch_size=32)
        h.init()
        h2 = h.copy()
        for _ in range(2):
            h._update((X, y))
        h2._update((X[:64], y[:64]))
        h._merge(h2)
        ev = h.evaluate((X, y))
        sd = {k: v.detach().clone() for k, v in h.model.state_dict().items()}
        h3 = pickle.loads(pickle.dumps(h))
        assert all(torch.equal(sd[k], v) for k, v in h3.model.state_dict().items())
        h4 = h.copy()
        h4.model = copy.deepcopy(h.model)
        assert all(torch.equal(a, b) for a, b in zip(h4.model.state_dict().values(), sd.values()))
        return h, sd, ev
    try:
        h0, sd0, ev0 = run(False)
        h1, sd1, ev1 = run(True)
        assert not h0.layout.channels_last and h1.layout.channels_last
        w = next(v for v in h1.model.parameters() if v.dim() == 4)
        assert w.is_contiguous(memory_format=torch.channels_last) and not w.is_contiguous()
        for k in sd0:
            torch.testing.assert_close(sd0[k].float(), sd1[k].float(), rtol=1e-5, atol=1e-6)
        assert ev0 == pytest.approx(ev1, abs=1e-6)
        # handlers that address parameters by flat index keep the plain order
        net = CIFAR10Net()
        hp = PartitionedTMH(net, TorchModelPartition(net, 4), torch.optim.SGD, {"lr": .1}, torch.nn.CrossEntropyLoss())
        assert not hp.layout.channels_last
    finally:
        g.GlobalSettings().channels_last = "auto"
