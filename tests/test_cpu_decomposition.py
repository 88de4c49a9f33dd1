"""mpp_init / mpp_basic_decomposition: oracle and product against the reference's own partition tables
(tests/BENCH/EXPREF/best_jpni_jpnj_*, committed as tests/golden/bench_partitions.json by golden/make_golden.py)
and against each other (mppini.F90:110-692, 695-798, 1180-1240)."""
import ctypes as C
import json
import os

import numpy as np
import pytest

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "bench_partitions.json")))
SCALARS = ("jpi jpj nimpp njmpp nlci nlcj nldi nlei nldj nlej nbondi nbondj noea nowe noso nono npolj "
           "l_Iperio l_Jperio").split()


@pytest.mark.parametrize("name", sorted(GOLD))
def test_oracle_decomposition_matches_reference_tables(O, name):
    g = GOLD[name]
    L = O.lib()
    for nb, jpni, jpnj, npts, jpimax, jpjmax in g["rows"]:
        ki, kj = C.c_int(), C.c_int()
        L.mpp_basic_decomposition(g["jpiglo"], g["jpjglo"], g["jperio"], jpni, jpnj, C.byref(ki), C.byref(kj),
                                  None, None, None, None)
        assert (ki.value, kj.value) == (jpimax, jpjmax), (name, nb, jpni, jpnj)
        assert jpni * jpnj == nb and jpimax * jpjmax == npts


@pytest.mark.parametrize("name", sorted(GOLD))
def test_product_decomposition_matches_reference_tables(N, name):
    g = GOLD[name]
    for nb, jpni, jpnj, npts, jpimax, jpjmax in g["rows"]:
        try:
            got = N.mpp_basic_decomposition(g["jpiglo"], g["jpjglo"], g["jperio"], jpni, jpnj)
        except N.NemoFctError:
            # table rows are sizes only; a grid whose last row would be thinner than the fold allows is refused
            # by mpp_basic_decomposition itself (mppini.F90:771-775) in both implementations
            continue
        assert (got[0], got[1]) == (jpimax, jpjmax), (name, nb, jpni, jpnj)
        nimppt, njmppt, nlcit, nlcjt = got[2:]
        # subdomains tile the global domain with a 2-point overlap (chap_LBC.tex:262-318)
        assert nimppt[0, -1] + nlcit[0, -1] - 1 == g["jpiglo"]
        assert njmppt[-1, 0] + nlcjt[-1, 0] - 1 == g["jpjglo"]
        assert nlcit.max() == jpimax and nlcjt.max() <= jpjmax


def test_survey_worked_example_orca025_4x2(N, O):
    """SURVEY.md App. C, hand-derived from mppini.F90:723-796, 1201-1227"""
    w = O.World(1442, 1207, 75, 4, 4, 2)
    assert [d.nimpp for d in w.doms[:4]] == [1, 361, 721, 1081]
    assert [d.nlcj for d in w.doms] == [605] * 4 + [604] * 4
    assert w.doms[4].njmpp == 604 and w.doms[4].npolj == 3 and w.doms[4].nbondj == 1
    assert w.doms[4].nono == 7 and w.doms[5].nono == 6
    assert w.doms[4].isendto == [3, 4]
    for r in range(8):
        d = N.mpp_init(1442, 1207, 75, 4, 4, 2, r + 1)
        for k in SCALARS:
            assert getattr(d, k) == getattr(w.doms[r], k), (r, k)
    w.close()


@pytest.mark.parametrize("jperio", range(8))
def test_product_mpp_init_equals_oracle(N, O, jperio):
    for (G, GJ) in [(20, 17), (21, 16), (37, 29)]:
        for (ni, nj) in [(1, 1), (2, 1), (1, 2), (2, 2), (3, 2), (4, 2), (3, 1), (5, 3)]:
            try:
                w = O.World(G, GJ, 5, jperio, ni, nj, ln_nnogather=False)
            except ValueError:
                continue
            for r, od in enumerate(w.doms):
                d = N.mpp_init(G, GJ, 5, jperio, ni, nj, r + 1)
                for k in SCALARS:
                    assert getattr(d, k) == getattr(od, k), (G, GJ, jperio, ni, nj, r, k)
                assert min(d.nsndto, 3) == od.nsndto
                assert list(d.isendto)[: min(d.nsndto, 3)] == od.isendto
            w.close()
    # without key_mpp_mpi (mppini.F90:53-102): npolj = jperio, nldi = 1, nlei = jpi
    w = O.World(20, 17, 5, jperio, 1, 1, key_mpp_mpi=False)
    d = N.mpp_init(20, 17, 5, jperio, 1, 1, 1, key_mpp_mpi=False)
    for k in SCALARS:
        assert getattr(d, k) == getattr(w.doms[0], k), k
    assert d.npolj == jperio
    w.close()


def test_invalid_layouts_are_refused(N):
    with pytest.raises(N.NemoFctError):
        N.mpp_init(10, 10, 5, 0, 9, 1, 1)          # jpi < 3
    with pytest.raises(N.NemoFctError):
        N.mpp_init(20, 12, 5, 4, 1, 4, 1)          # last row thinner than the T-pivot fold needs
    with pytest.raises(N.NemoFctError):
        N.mpp_init(20, 12, 5, 9, 1, 1, 1)          # jperio out of range
    with pytest.raises(N.NemoFctError):
        N.mpp_init(20, 12, 5, 0, 2, 1, 3)          # narea out of range
    with pytest.raises(N.NemoFctError):
        N.mpp_init(20, 12, 5, 0, 2, 1, 1, key_mpp_mpi=False)
    assert np.all(N.mpp_basic_decomposition(20, 12, 0, 2, 2)[2] >= 1)
