"""The oracle's stp_ctl (oracle/stpctl.c, restating src/OCE/stpctl.F90:115-124, 149-166, 184) against a numpy reading of the
same lines: MAXVAL / MAXLOC / MINLOC with mask = tmask == 1, first occurrence on ties, indices shifted by nimpp / njmpp."""
import numpy as np
import pytest

import helpers as H


@pytest.mark.parametrize("jperio,lay", [(0, (1, 1)), (4, (1, 1)), (4, (2, 2)), (6, (3, 1))])
def test_oracle_stp_ctl_is_maxval_maxloc(O, jperio, lay):
    G, GJ, K = 41, 33, 6
    gf = H.random_fields(O, G, GJ, K, jperio, kjpt=2, seed=70 + jperio)
    rng = np.random.default_rng(3)
    sshn = rng.standard_normal((GJ, G)); un = rng.standard_normal((K, GJ, G))
    tem = 10.0 + rng.standard_normal((K, GJ, G)); sal = 35.0 + rng.standard_normal((K, GJ, G))
    sal[2, 10, 7] = sal[4, 20, 30] = 50.0                      # a tie: the first in array order wins
    w = O.World(G, GJ, K, jperio, *lay)
    l_ssh, l_un, l_tem, l_sal, l_tm = (w.scatter(np.ascontiguousarray(a)) for a in (sshn, un, tem, sal, gf["tmask"]))
    for r, d in enumerate(w.doms):
        ts = np.ascontiguousarray(np.stack([l_tem[r], l_sal[r]]))
        got = O.stp_ctl(d, l_ssh[r], l_un[r], ts, l_tm[r])
        wet = l_tm[r] == 1.0
        if not wet.any():
            continue
        assert got["zmax"] == [np.abs(l_ssh[r]).max(), np.abs(l_un[r]).max(), -l_sal[r][wet].min(), l_sal[r][wet].max(),
                               -l_tem[r][wet].min(), l_tem[r][wet].max()]
        off = np.array([d.nimpp - 1, d.njmpp - 1, 0])

        def loc(flat, shape):
            idx = np.unravel_index(flat, shape)
            return list(np.array([idx[-1] + 1, idx[-2] + 1, (idx[0] + 1) if len(shape) == 3 else 0])[: len(shape)] + off[: len(shape)])
        assert got["ih"] == loc(np.argmax(np.abs(l_ssh[r])), l_ssh[r].shape)
        assert got["iu"] == loc(np.argmax(np.abs(l_un[r])), l_un[r].shape)
        assert got["is1"] == loc(np.argmin(np.where(wet, l_sal[r], np.inf)), l_sal[r].shape)
        assert got["is2"] == loc(np.argmax(np.where(wet, l_sal[r], -np.inf)), l_sal[r].shape)
        assert got["kindic"] == 0 and got["nan_found"] == 0
    w.close()


@pytest.mark.parametrize("what,kindic", [("ssh", -3), ("u", -3), ("s0", -3), ("s100", -3), ("nan", -3), ("edge_ok", 0)])
def test_oracle_stp_ctl_condition(O, what, kindic):
    G, GJ, K = 30, 24, 5
    gf = H.random_fields(O, G, GJ, K, 0, kjpt=2, seed=5)
    tm = gf["tmask"]
    k, j, i = np.argwhere(tm == 1.0)[17]
    sshn = np.zeros((GJ, G)); un = np.zeros((K, GJ, G)); tem = np.full((K, GJ, G), 4.0); sal = np.full((K, GJ, G), 35.0)
    if what == "ssh": sshn[j, i] = 20.5
    if what == "u": un[k, j, i] = -10.5
    if what == "s0": sal[k, j, i] = 0.0                        # zmax(3) >= 0: "negative or zero" salinity
    if what == "s100": sal[k, j, i] = 100.0
    if what == "nan": sshn[j, i] = np.nan
    if what == "edge_ok": sshn[j, i] = 20.0; un[k, j, i] = 10.0; sal[k, j, i] = 99.999
    w = O.World(G, GJ, K, 0)
    got = O.stp_ctl(w.doms[0], sshn, un, np.ascontiguousarray(np.stack([tem, sal])), tm)
    w.close()
    assert got["kindic"] == kindic, got
    assert got["nan_found"] == int(what == "nan")
