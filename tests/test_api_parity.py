"""Public-API parity with the reference, checked mechanically.

For every module of the reference: each name in its ``__all__`` (plus the unexported classes user scripts
import) must exist here, classes must keep their public methods, and constructors / methods must accept the
reference's parameter names in the reference's order (extra trailing keyword parameters are allowed)."""
import importlib
import inspect

import pytest

MODULES = {            # reference module -> ours, extra (unexported but used) names
    "gossipy": ("gossipy_b200", ["LOG", "CACHE", "set_seed", "CacheKey", "CacheItem", "Sizeable", "Cache",
                                 "GlobalSettings"]),
    "gossipy.core": ("gossipy_b200.core", ["Delay", "MixingMatrix", "Message", "MessageType"]),
    "gossipy.node": ("gossipy_b200.node", ["All2AllGossipNode"]),
    "gossipy.model": ("gossipy_b200.model", []),
    "gossipy.model.nn": ("gossipy_b200.model.nn", []),
    "gossipy.model.sampling": ("gossipy_b200.model.sampling", []),
    "gossipy.model.handler": ("gossipy_b200.model.handler", ["ModelHandler", "WeightedTMH", "LimitedMergeTMH",
                                                              "LimitedMergeMixin"]),
    "gossipy.flow_control": ("gossipy_b200.flow_control", []),
    "gossipy.data": ("gossipy_b200.data", ["RecSysDataDispatcher", "get_FEMNIST"]),
    "gossipy.data.handler": ("gossipy_b200.data.handler", []),
    "gossipy.simul": ("gossipy_b200.simul", ["All2AllGossipSimulator", "SimulationEventSender"]),
    "gossipy.utils": ("gossipy_b200.utils", []),
}


def _params(fn):
    try:
        sig = inspect.signature(fn)
    except (TypeError, ValueError):
        return None
    return [p for p in sig.parameters.values() if p.name != "self"]


def _check_callable(ref_fn, our_fn, where):
    rp, op = _params(ref_fn), _params(our_fn)
    if rp is None or op is None:
        return
    if any(p.kind in (p.VAR_POSITIONAL, p.VAR_KEYWORD) for p in op):
        return      # we forward *args/**kwargs (e.g. handler subclasses): accepted by construction
    ours = [p.name for p in op]
    for i, p in enumerate(rp):
        if p.kind in (p.VAR_POSITIONAL, p.VAR_KEYWORD):
            continue
        assert p.name in ours, "%s: parameter %r missing (ours: %s)" % (where, p.name, ours)
        if p.kind == p.POSITIONAL_OR_KEYWORD and p.default is p.empty:
            assert ours.index(p.name) == i, "%s: positional parameter %r moved" % (where, p.name)
    for p in op:      # anything we added must be optional
        if p.name not in [q.name for q in rp]:
            assert p.default is not p.empty or p.kind == p.KEYWORD_ONLY and p.default is not p.empty, \
                "%s: new required parameter %r" % (where, p.name)


@pytest.mark.parametrize("ref_name", sorted(MODULES))
def test_module_api_parity(ref, ref_name):
    ours_name, extra = MODULES[ref_name]
    rmod = importlib.import_module(ref_name)
    omod = importlib.import_module(ours_name)
    names = list(getattr(rmod, "__all__", [])) + [n for n in extra if hasattr(rmod, n)]
    assert names, ref_name
    for name in dict.fromkeys(names):
        assert hasattr(omod, name), "%s.%s is missing" % (ours_name, name)
        robj, oobj = getattr(rmod, name), getattr(omod, name)
        if inspect.isclass(robj):
            assert inspect.isclass(oobj), name
            if issubclass(robj, BaseException) or name in ("LOG",):
                continue
            if "__init__" in vars(robj):
                _check_callable(robj.__init__, oobj.__init__, "%s.%s.__init__" % (ours_name, name))
            for attr, member in vars(robj).items():
                if attr.startswith("_") or inspect.isclass(member) or inspect.ismodule(member):
                    continue       # (the reference imports a class inside one class body)
                assert hasattr(oobj, attr), "%s.%s.%s is missing" % (ours_name, name, attr)
                if inspect.isfunction(member):
                    _check_callable(member, getattr(oobj, attr), "%s.%s.%s" % (ours_name, name, attr))
        elif inspect.isfunction(robj):
            _check_callable(robj, oobj, "%s.%s" % (ours_name, name))


def test_enum_members_match(ref):
    import gossipy.core as rc
    import gossipy_b200.core as oc
    for name in ("CreateModelMode", "AntiEntropyProtocol", "MessageType"):
        r, o = getattr(rc, name), getattr(oc, name)
        assert [(m.name, m.value) for m in r] == [(m.name, m.value) for m in o]
