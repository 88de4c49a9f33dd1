"""The oracle's tra_adv_fct / nonosc (C triple loops) against an independent vectorised numpy statement of the same
Fortran (tests/np_fct.py): bit for bit, closed and E-W cyclic mono-domains, 2nd and 4th order horizontal, with land."""
import numpy as np
import pytest

from oracle import oracle as O
import helpers as H
import np_fct


@pytest.mark.parametrize("jperio", [0, 1])
@pytest.mark.parametrize("h", [2, 4])
@pytest.mark.parametrize("ln_linssh", [False, True])
def test_oracle_equals_numpy_restatement(jperio, h, ln_linssh):
    G, GJ, K, kjpt = 28, 23, 9, 2
    gf = H.random_fields(O, G, GJ, K, jperio, kjpt, seed=300 + 10 * jperio + h, ln_linssh=ln_linssh)
    ref, _, _ = H.oracle_fct(O, gf, G, GJ, K, jperio, 1, 1, kjpt, h, 2, ln_linssh=ln_linssh)
    got = np_fct.tra_adv_fct(gf, kjpt, h, jperio, ln_linssh=ln_linssh)
    assert not np.array_equal(ref, gf["pta"])
    assert np.array_equal(got, ref), float(np.abs(got - ref).max())
