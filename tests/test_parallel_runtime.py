"""Configuration of the multi-rank runtime: placements and the transport switch."""
import pytest


def test_placement_constructors_and_queries():
    from gossipy_b200.parallel.runtime import Placement
    assert Placement.block(8, 4).ranks == [0, 0, 1, 1, 2, 2, 3, 3]
    assert Placement.block(5, 2).ranks == [0, 0, 0, 1, 1]
    assert Placement.round_robin(5, 2).ranks == [0, 1, 0, 1, 0]
    pl = Placement.by_load([5, 1, 1, 1, 4, 3], 2)
    load = [sum(w for w, r in zip([5, 1, 1, 1, 4, 3], pl.ranks) if r == q) for q in range(2)]
    assert sorted(load) == [7, 8] and len(pl) == 6
    assert pl.nodes_of(0) + pl.nodes_of(1) != [] and sorted(pl.nodes_of(0) + pl.nodes_of(1)) == list(range(6))
    assert Placement.explicit([1, 0, 1]).world == 2 and Placement.explicit([1, 0, 1]).rank_of(0) == 1
    assert Placement.block(4, 2) == Placement.explicit([0, 0, 1, 1], 2)
    with pytest.raises(ValueError):
        Placement([0, 3], world=2)


def test_set_num_nodes_keeps_an_installed_placement_and_validates():
    from gossipy_b200.parallel import runtime as prt
    st = dict(prt._state)
    try:
        prt._state.update(rank=1, world=3)
        prt.set_num_nodes(6, prt.Placement.round_robin(6, 3))
        assert [prt.rank_of(i) for i in range(6)] == [0, 1, 2, 0, 1, 2] and prt.is_mine(4) and not prt.is_mine(3)
        prt.set_num_nodes(6)                                # what init_nodes calls: the installed map survives
        assert prt.placement() == prt.Placement.round_robin(6, 3)
        prt.set_num_nodes(9)                                # another node count: back to blocks
        assert prt.placement() == prt.Placement.block(9, 3)
        with pytest.raises(ValueError):
            prt.set_num_nodes(4, [0, 1, 2])
        with pytest.raises(ValueError):
            prt.set_num_nodes(3, [0, 1, 5])
    finally:
        prt._state.clear()
        prt._state.update(st)


def test_transport_names():
    """``p2p`` | ``nccl`` (alias ``nccl-baseline``) | ``loopback`` (no inter-rank transport: this process hosts all nodes)."""
    import gossipy_b200 as g
    from gossipy_b200.parallel import runtime as prt
    st = dict(prt._state)
    try:
        prt.init(rank=0, world=1, transport="nccl-baseline")
        assert prt.transport() == "none" and not prt.active()          # one process: nothing to transport
        prt.init(rank=0, world=1, transport="loopback")
        assert prt.transport() == "loopback" and not prt.active() and prt.rank_of(5) == 0
        assert g.GlobalSettings().rank == 0 if hasattr(g.GlobalSettings(), "rank") else True
        with pytest.raises(ValueError):
            prt.init(rank=0, world=1, transport="carrier-pigeon")
    finally:
        prt._state.clear()
        prt._state.update(st)
        g.GlobalSettings().set_topology(0, 1)
