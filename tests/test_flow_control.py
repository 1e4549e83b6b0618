"""Token accounts: exhaustive table comparison with the reference."""
import numpy as np
import pytest

from gossipy_b200 import flow_control as fc


@pytest.mark.parametrize("name,args", [("PurelyReactiveTokenAccount", (2,)),
                                        ("SimpleTokenAccount", (3,)),
                                        ("GeneralizedTokenAccount", (8, 3)),
                                        ("RandomizedTokenAccount", (20, 10))])
def test_tables_match_reference(ref, name, args):
    import gossipy.flow_control as rfc
    ours, theirs = getattr(fc, name)(*args), getattr(rfc, name)(*args)
    for a in range(0, 45):
        ours.n_tokens = theirs.n_tokens = a
        assert ours.proactive() == pytest.approx(theirs.proactive())
        for u in (0, 1, 2):
            if name == "RandomizedTokenAccount":
                np.random.seed(a * 7 + u); r1 = ours.reactive(u)
                np.random.seed(a * 7 + u); r2 = theirs.reactive(u)
                assert r1 == r2
            else:
                assert ours.reactive(u) == theirs.reactive(u)


def test_add_sub_and_proactive_account():
    acc = fc.PurelyProactiveTokenAccount()
    assert acc.n_tokens == 0 and acc.proactive() == 1 and acc.reactive(5) == 0   # B21 fixed
    acc.add(3); acc.sub(5)
    assert acc.n_tokens == 0
    with pytest.raises(AssertionError):
        fc.GeneralizedTokenAccount(2, 3)
    assert fc.RandomizedTokenAccount(20, 10).spec() == (4, 20, 10, 0)
