"""GPU parity of tra_adv_cen and of the tra_adv_fct trend-diagnostic hooks through the C ABI, bit for bit against the
oracle (traadv_cen.F90:46-204; traadv_fct.F90:96-112, 172-176, 299-316)."""
import importlib

import numpy as np
import pytest
import torch

from oracle import oracle as O
import helpers as H

pytestmark = pytest.mark.gpu
N = importlib.import_module("nemo-fmi-devel_b200")
JPK = 11


@pytest.mark.parametrize("jperio", [0, 1, 4, 6, 7])
@pytest.mark.parametrize("h,v", [(2, 2), (2, 4), (4, 2), (4, 4)])
def test_cen_parity(jperio, h, v):
    """whole local array equal to the oracle (for h = 4 the oracle's un-poisoned work arrays: the never-assigned
    zwy(:,1,:) of the reference is 0 on both sides, see oracle/traadv_cen.c)"""
    G, GJ, kjpt = 42, 36, 2
    gf = H.random_fields(O, G, GJ, JPK, jperio, kjpt, seed=40 + jperio)
    ref, _ = H.oracle_cen(O, gf, G, GJ, JPK, jperio, 1, 1, kjpt, h, v, poison=(h == 2))
    got, loc = H.device_cen(N, gf, G, GJ, JPK, jperio, 1, 1, kjpt, h, v)
    assert np.array_equal(loc[0], ref), H.max_rel_diff(loc[0], ref)
    assert not np.array_equal(ref, gf["pta"])


@pytest.mark.parametrize("ln_linssh,ln_isfcav", [(True, False), (True, True)])
def test_cen_parity_linssh(ln_linssh, ln_isfcav):
    G, GJ, kjpt, jperio = 40, 34, 3, 4
    gf = H.random_fields(O, G, GJ, JPK, jperio, kjpt, seed=6, ln_linssh=ln_linssh, ln_isfcav=ln_isfcav)
    for (h, v) in ((2, 2), (4, 4)):
        ref, _ = H.oracle_cen(O, gf, G, GJ, JPK, jperio, 1, 1, kjpt, h, v, ln_linssh=ln_linssh, ln_isfcav=ln_isfcav)
        got, loc = H.device_cen(N, gf, G, GJ, JPK, jperio, 1, 1, kjpt, h, v, ln_linssh=ln_linssh, ln_isfcav=ln_isfcav)
        assert np.array_equal(loc[0], ref)


@pytest.mark.parametrize("jperio,jpni,jpnj", [(0, 2, 2), (4, 2, 2), (6, 3, 2)])
def test_cen_decomposed(jperio, jpni, jpnj):
    """rank-local arrays equal to the oracle's ranks (h = 4 includes the lbc_lnk on ztu / ztv); for h = 2 the assembled
    interior also equals the mono-domain result"""
    G, GJ, kjpt = 56, 48, 2
    gf = H.random_fields(O, G, GJ, JPK, jperio, kjpt, seed=17)
    for (h, v) in ((2, 2), (4, 4)):
        refg, refl = H.oracle_cen(O, gf, G, GJ, JPK, jperio, jpni, jpnj, kjpt, h, v)
        got, loc = H.device_cen(N, gf, G, GJ, JPK, jperio, jpni, jpnj, kjpt, h, v)
        for a, b in zip(loc, refl):
            assert np.array_equal(a, b)
        if h == 2:
            ref1, _ = H.oracle_cen(O, gf, G, GJ, JPK, jperio, 1, 1, kjpt, h, v)
            inner = (slice(None), slice(0, JPK - 1), slice(1, -1), slice(1, -1))
            assert np.array_equal(got[inner], ref1[inner])


@pytest.mark.parametrize("h,v", [(2, 2), (4, 4)])
@pytest.mark.parametrize("jperio", [1, 4])
def test_fct_trend_diag_parity(jperio, h, v):
    """ztrdx / ztrdy on (1:jpim1, 1:jpjm1, :), ztrdz everywhere: equal to the oracle's hooks; pta unchanged by the hooks"""
    G, GJ, kjpt = 44, 38, 2
    gf = H.random_fields(O, G, GJ, JPK, jperio, kjpt, seed=23)
    w = O.World(G, GJ, JPK, jperio, 1, 1)
    d = w.doms[0]
    d.set_fields(*[gf[k] for k in H.DOM_KEYS])
    otrd = [np.full((kjpt,) + d.shape3, np.nan) for _ in range(3)]
    d.set_diag(*otrd)
    opta = gf["pta"].copy()
    w.tra_adv_fct(gf["p2dt"], [gf["pun"]], [gf["pvn"]], [gf["pwn"]], [gf["ptb"]], [gf["ptn"]], [opta], kjpt, h, v)
    w.close()
    ctx = N.FctContext(N.mpp_init(G, GJ, JPK, jperio, 1, 1, 1), 0)
    ctx.set_domain_arrays(*[gf[k] for k in ("tmask", "umask", "vmask", "wmask", "e1e2t", "r1_e1e2t", "mikt", "mbkt")])
    ctx.set_e3t(gf["e3t_b"], gf["e3t_n"], gf["e3t_a"])
    t = {k: torch.from_numpy(gf[k]).cuda() for k in ("pun", "pvn", "pwn", "ptb", "ptn", "pta")}
    sentinel = -7.0e77
    dtrd = [torch.full((kjpt,) + d.shape3, sentinel, dtype=torch.float64, device="cuda") for _ in range(3)]
    ctx.set_trend_diag(*dtrd)
    ctx.tra_adv_fct(1, 1, "TRA", gf["p2dt"], t["pun"], t["pvn"], t["pwn"], t["ptb"], t["ptn"], t["pta"], kjpt, h, v)
    ctx.synchronize()
    assert np.array_equal(t["pta"].cpu().numpy(), opta)
    gx, gy, gz = [a.cpu().numpy() for a in dtrd]
    assert np.array_equal(gx[:, :, :-1, :-1], otrd[0][:, :, :-1, :-1])
    assert np.array_equal(gy[:, :, :-1, :-1], otrd[1][:, :, :-1, :-1])
    assert np.array_equal(gz, otrd[2])
    assert (gx[:, :, :, -1] == sentinel).all() and (gy[:, :, -1, :] == sentinel).all()      # undefined in the reference: untouched
    # hooks off again: the default (fused) schedule gives the same pta
    ctx.set_trend_diag()
    t2 = torch.from_numpy(gf["pta"]).cuda()
    ctx.tra_adv_fct(1, 1, "TRA", gf["p2dt"], t["pun"], t["pvn"], t["pwn"], t["ptb"], t["ptn"], t2, kjpt, h, v)
    ctx.synchronize()
    assert np.array_equal(t2.cpu().numpy(), opta)
    ctx.close()


def test_cen_and_diag_errors():
    dom = N.mpp_init(30, 30, JPK, 0, 1, 1, 1)
    ctx = N.FctContext(dom, 0)
    gf = H.random_fields(O, 30, 30, JPK, 0, 1, seed=1)
    ctx.set_domain_arrays(*[gf[k] for k in ("tmask", "umask", "vmask", "wmask", "e1e2t", "r1_e1e2t", "mikt", "mbkt")])
    ctx.set_e3t(gf["e3t_b"], gf["e3t_n"], gf["e3t_a"])
    z3 = torch.zeros(dom.shape3, dtype=torch.float64, device="cuda")
    z4 = torch.zeros((1,) + dom.shape3, dtype=torch.float64, device="cuda")
    with pytest.raises(N.NemoFctError, match="nn_cen_h"):
        ctx.tra_adv_cen(1, 1, "TRA", z3, z3, z3, z4, z4.clone(), 1, 3, 2)
    with pytest.raises(ValueError):
        ctx.set_trend_diag(z4, None, None)
    ctx.close()
