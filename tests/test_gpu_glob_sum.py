"""glob_sum on the device (csrc/glob_sum.cu) against the oracle's restatement of lib_fortran_generic.h90:32-65 / DDPDD
(lib_fortran.F90:300-332): the conservation metric of BASELINE.json ("global tracer content conserved to round-off")."""
import numpy as np
import pytest
import torch

import helpers as H

pytestmark = pytest.mark.gpu


def _tmask_i(dom, tmask):
    """dom_oce's interior mask of one subdomain (dommsk.F90:200-238): 0 on the local halo rows / columns and, under a
    T-pivot fold, on the duplicated half of the last interior row; times the surface mask"""
    jpj, jpi = tmask.shape[-2:]
    th = np.ones((jpj, jpi))
    th[:, :1] = 0.0; th[:, dom.nlci - 1:] = 0.0
    th[:1, :] = 0.0; th[dom.nlcj - 1:, :] = 0.0
    if dom.jperio in (3, 4) and dom.nlej + dom.njmpp - 1 == dom.jpjglo:
        mig = np.arange(1, jpi + 1) + dom.nimpp - 1
        tpol = np.where(mig >= dom.jpiglo // 2 + 1, 0.0, 1.0)
        th[dom.nlej - 2, 1:dom.nlci - 1] *= tpol[1:dom.nlci - 1]
    return np.ascontiguousarray(tmask.max(axis=0) * th)


def _fields(O, G, GJ, K, jperio, seed):
    gf = H.random_fields(O, G, GJ, K, jperio, kjpt=3, seed=seed)
    cvol = gf["e1e2t"][None] * gf["e3t_n"] * gf["tmask"]
    return gf, np.ascontiguousarray(cvol)


@pytest.mark.parametrize("jperio", [0, 4, 6])
def test_glob_sum_equals_the_oracle_and_ignores_the_decomposition(N, O, jperio):
    G, GJ, K = 70, 46, 13
    gf, cvol = _fields(O, G, GJ, K, jperio, 600 + jperio)
    rng = np.random.default_rng(5)
    big = gf["ptb"] * (1.0 + 1e8 * rng.standard_normal(gf["ptb"].shape))            # heavy cancellation: the pair matters
    dev = torch.device("cuda:0")
    want = None
    for (jpni, jpnj) in ((1, 1), (2, 2), (3, 2)):
        w = O.World(G, GJ, K, jperio, jpni, jpnj)
        doms = [N.mpp_init(G, GJ, K, jperio, jpni, jpnj, r + 1) for r in range(jpni * jpnj)]
        ti = [_tmask_i(d, tm) for d, tm in zip(doms, w.scatter(gf["tmask"]))]
        if jpni * jpnj == 1:
            assert np.array_equal(ti[0], gf["tmask_i"])                             # the helper restates dom_msk
        lv = w.scatter(cvol)
        lt = [w.scatter(np.ascontiguousarray(big[jn])) for jn in range(3)]
        ref = [O.glob_sum(w, [np.ascontiguousarray(a * v) for a, v in zip(lt[jn], lv)], ti)[0] for jn in range(3)]
        w.close()
        if want is None:
            want = ref
        assert ref == want                                                          # the oracle itself is decomposition-invariant
        t_i = [torch.from_numpy(a).to(dev) for a in ti]
        t_v = [torch.from_numpy(a).to(dev) for a in lv]
        t_t = [[torch.from_numpy(a).to(dev) for a in lt[jn]] for jn in range(3)]
        if jpni * jpnj == 1:
            ctx = N.FctContext(N.mpp_init(G, GJ, K, jperio, 1, 1, 1), 0)
            got = ctx.glob_sum("test", [t_t[jn][0] for jn in range(3)], t_i[0], w3d=t_v[0])
            ctx.close()
        else:
            grp = N.LocalGroup(G, GJ, K, jperio, jpni, jpnj, 0)
            got = grp.glob_sum("test", t_t, t_i, w3d=t_v)
            grp.close()
        assert got == want, (jpni, jpnj, got, want)
    plain = float(np.sum(big[0] * cvol * gf["tmask_i"][None]))
    assert plain != want[0] or abs(want[0]) == 0.0                                   # a plain fp64 sum does not get there
