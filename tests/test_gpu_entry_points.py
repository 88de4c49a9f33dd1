"""The other entry points of the C ABI on the device: interp_4th_cpt (public in the reference, also used by
tra_adv_cen), the tra_adv transport build, lbc_lnk_multi on device-resident fields with a land value, and the error
behaviour (non-zero return + message, the shim's ctl_stop path)."""
import ctypes as C

import numpy as np
import pytest
import torch

import helpers as H

pytestmark = pytest.mark.gpu


def _ctx(N, O, G, GJ, K, jperio, gf, ln_isfcav=False):
    dom = N.mpp_init(G, GJ, K, jperio)
    ctx = N.FctContext(dom, 0)
    ctx.set_domain_arrays(gf["tmask"], gf["umask"], gf["vmask"], gf["wmask"], gf["e1e2t"], gf["r1_e1e2t"], gf["mikt"], gf["mbkt"],
                          False, ln_isfcav)
    ctx.set_e3t(gf["e3t_b"], gf["e3t_n"], gf["e3t_a"])
    return dom, ctx


@pytest.mark.parametrize("ln_isfcav", [False, True])
def test_interp_4th_cpt_host_and_device(N, O, ln_isfcav):
    G, GJ, K = 40, 31, 17
    gf = H.random_fields(O, G, GJ, K, 1, kjpt=1, seed=3, ln_isfcav=ln_isfcav)
    w = O.World(G, GJ, K, 1)
    w.doms[0].set_fields(*[gf[k] for k in H.DOM_KEYS], ln_isfcav=ln_isfcav)
    pt_in = np.ascontiguousarray(gf["ptn"][0])
    ref = np.full(pt_in.shape, -777.0)
    O.lib().interp_4th_cpt(w.doms[0].h, pt_in.ctypes.data, ref.ctypes.data)
    dom, ctx = _ctx(N, O, G, GJ, K, 1, gf, ln_isfcav)
    got = np.full(pt_in.shape, -777.0)
    ctx.interp_4th_cpt(pt_in, got)                                 # host pointers
    sl = (slice(1, -1), slice(1, -1), slice(1, -1))                # defined on (2:jpim1, 2:jpjm1, 2:jpkm1) only
    assert np.array_equal(got[sl], ref[sl])
    assert np.all(got[0] == -777.0) and np.all(got[-1] == -777.0) and np.all(got[:, 0] == -777.0)   # nothing else is touched
    tin = torch.from_numpy(pt_in).cuda(); tout = torch.full_like(tin, -777.0)
    ctx.interp_4th_cpt(tin, tout); ctx.synchronize()               # device pointers
    assert np.array_equal(tout.cpu().numpy()[sl], ref[sl])
    ctx.close(); w.close()


def test_tra_adv_transports(N, O):
    """zun = e2u*e3u_n*un, zvn = e1v*e3v_n*vn, zwn = e1e2t*wn, level jpk zeroed (traadv.F90:100-124)"""
    G, GJ, K = 33, 26, 9
    gf = H.random_fields(O, G, GJ, K, 0, kjpt=1, seed=8)
    rng = np.random.default_rng(0)
    e2u, e1v = rng.random((GJ, G)) + 1.0, rng.random((GJ, G)) + 1.0
    e3u, e3v = rng.random((K, GJ, G)) + 1.0, rng.random((K, GJ, G)) + 1.0
    un, vn, wn = (rng.standard_normal((K, GJ, G)) for _ in range(3))
    w = O.World(G, GJ, K, 0)
    w.doms[0].set_fields(*[gf[k] for k in H.DOM_KEYS])
    ref = [np.full((K, GJ, G), np.nan) for _ in range(3)]
    O.lib().tra_adv_transports(w.doms[0].h, *[a.ctypes.data for a in (e2u, e1v, e3u, e3v, un, vn, wn)], *[a.ctypes.data for a in ref])
    dom, ctx = _ctx(N, O, G, GJ, K, 0, gf)
    t = [torch.from_numpy(np.ascontiguousarray(a)).cuda() for a in (e2u, e1v, e3u, e3v, un, vn, wn)]
    out = [torch.full((K, GJ, G), float("nan"), dtype=torch.float64, device="cuda") for _ in range(3)]
    ctx.tra_adv_transports(*t, *out); ctx.synchronize()
    for a, b in zip(out, ref):
        assert np.array_equal(a.cpu().numpy(), b)
    ctx.close(); w.close()


@pytest.mark.parametrize("jperio", [0, 4, 6])
def test_lbc_lnk_multi_device_fields_with_pval(N, O, jperio):
    G, GJ, K = 30, 24, 5
    gf = H.random_fields(O, G, GJ, K, jperio, kjpt=1, seed=2)
    rng = np.random.default_rng(jperio)
    fields = [rng.standard_normal((K, GJ, G)) for _ in range(4)]
    w = O.World(G, GJ, K, jperio)
    ref = [f.copy() for f in fields]
    import ctypes
    # oracle with pval = 1 (e.g. masks): call lbc_lnk_multi directly on the single subdomain
    tab = (ctypes.c_void_p * 4)(*[a.ctypes.data for a in ref])
    sg = (ctypes.c_double * 4)(1.0, -1.0, -1.0, 1.0)
    L = O.lib()
    L.lbc_lnk_multi.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_char_p,
                                ctypes.POINTER(ctypes.c_double), ctypes.c_int, ctypes.c_int, ctypes.c_double]
    L.lbc_lnk_multi(w.doms[0].h, b"test", 4, tab, b"TUVF", sg, K, 1, 1.0)
    dom, ctx = _ctx(N, O, G, GJ, K, jperio, gf)
    t = [torch.from_numpy(f).cuda() for f in fields]
    ctx.lbc_lnk_multi("test", t[0], "T", 1.0, t[1], "U", -1.0, t[2], "V", -1.0, t[3], "F", 1.0, pval=1.0)
    ctx.synchronize()
    for a, b in zip(t, ref):
        assert np.array_equal(a.cpu().numpy(), b)
    ctx.close(); w.close()


def test_errors_are_reported_not_swallowed(N, O):
    G, GJ, K = 24, 22, 6
    gf = H.random_fields(O, G, GJ, K, 0, kjpt=1, seed=1)
    dom = N.mpp_init(G, GJ, K, 0)
    ctx = N.FctContext(dom, 0)
    a3 = torch.zeros((K, GJ, G), dtype=torch.float64, device="cuda"); a4 = torch.zeros((1, K, GJ, G), dtype=torch.float64, device="cuda")
    with pytest.raises(N.NemoFctError, match="set_domain_arrays"):
        ctx.tra_adv_fct(1, 1, "TRA", 1.0, a3, a3, a3, a4, a4, a4, 1, 2, 2)
    bad = gf["mbkt"].copy(); bad[3, 3] = K + 5
    with pytest.raises(N.NemoFctError, match="mikt/mbkt"):
        ctx.set_domain_arrays(gf["tmask"], gf["umask"], gf["vmask"], gf["wmask"], gf["e1e2t"], gf["r1_e1e2t"], gf["mikt"], bad)
    ctx.set_domain_arrays(gf["tmask"], gf["umask"], gf["vmask"], gf["wmask"], gf["e1e2t"], gf["r1_e1e2t"], gf["mikt"], gf["mbkt"])
    with pytest.raises(N.NemoFctError, match="set_e3t"):
        ctx.tra_adv_fct(1, 1, "TRA", 1.0, a3, a3, a3, a4, a4, a4, 1, 2, 2)
    ctx.set_e3t(gf["e3t_b"], gf["e3t_n"], gf["e3t_a"])
    with pytest.raises(N.NemoFctError, match="41"):
        ctx.tra_adv_fct(1, 1, "TRA", 1.0, a3, a3, a3, a4, a4, a4, 1, 41, 2)      # untested scheme of the reference: refused
    with pytest.raises(N.NemoFctError, match="cd_nat"):
        ctx.lbc_lnk_multi("x", a3, "Q", 1.0)
    with pytest.raises(ValueError):
        ctx.tra_adv_fct(1, 1, "TRA", 1.0, a3, a3, a3, a4, a4, a3, 1, 2, 2)       # wrong shape
    with pytest.raises(N.NemoFctError, match="schedule"):
        ctx.set_schedule(9)
    ctx.close()
    # a multi-rank subdomain without a communicator cannot exchange: loud error, not a silent local copy
    dom2 = N.mpp_init(60, 44, K, 1, 2, 1, 1)
    c2 = N.FctContext(dom2, 0)
    f = torch.zeros((K, dom2.jpj, dom2.jpi), dtype=torch.float64, device="cuda")
    with pytest.raises(N.NemoFctError, match="comm_init"):
        c2.lbc_lnk_multi("x", f, "T", 1.0)
    c2.close()


def test_tra_adv_and_trc_adv_device_resident(N, O):
    """tra_adv (traadv.F90:77-175, FCT branch) and trc_adv (trcadv.F90:70-145) on device-resident state: r2dt rule of
    :95-97 (Euler at nit000 when neuler = 0, leap-frog afterwards), transports built once and reused by the passive
    tracers.  Oracle: tra_adv_transports + tra_adv_fct with the same r2dt."""
    G, GJ, K, jperio = 52, 40, 9, 4
    gf = H.random_fields(O, G, GJ, K, jperio, kjpt=2, seed=12)
    rng = np.random.default_rng(5)
    e2u = np.full((GJ, G), 1.0e5) * (1.0 + 0.05 * rng.random((GJ, G))); e1v = np.full((GJ, G), 1.0e5) * (1.0 + 0.05 * rng.random((GJ, G)))
    e3u = gf["e3t_n"] * (1.0 + 0.01 * rng.random((K, GJ, G))); e3v = gf["e3t_n"] * (1.0 + 0.01 * rng.random((K, GJ, G)))
    un = 0.3 * 1.0e5 / 7200.0 * rng.uniform(-1, 1, (K, GJ, G)) * gf["umask"]
    vn = 0.3 * 1.0e5 / 7200.0 * rng.uniform(-1, 1, (K, GJ, G)) * gf["vmask"]
    wn = 1.0e-4 * rng.uniform(-1, 1, (K, GJ, G)) * gf["wmask"]
    w = O.World(G, GJ, K, jperio)
    w.lbc_lnk([[un], [vn], [wn]], "UVW", [-1.0, -1.0, 1.0])
    d = w.doms[0]
    d.set_fields(*[gf[k] for k in H.DOM_KEYS])
    zu, zv, zw = (np.zeros((K, GJ, G)) for _ in range(3))
    O.lib().tra_adv_transports(d.h, *[a.ctypes.data for a in (e2u, e1v, e3u, e3v, un, vn, wn)], zu.ctypes.data, zv.ctypes.data, zw.ctypes.data)
    rdt, nit000 = 3600.0, 11
    trb = np.ascontiguousarray(np.stack([gf["ptb"][0] * (1 + 0.1 * n) for n in range(5)]))      # 5 passive tracers
    trn = np.ascontiguousarray(np.stack([gf["ptn"][0] * (1 + 0.1 * n) for n in range(5)]))
    dom = N.mpp_init(G, GJ, K, jperio)
    ctx = N.FctContext(dom, 0)
    ctx.set_domain_arrays(gf["tmask"], gf["umask"], gf["vmask"], gf["wmask"], gf["e1e2t"], gf["r1_e1e2t"], gf["mikt"], gf["mbkt"])
    ctx.set_e3t(gf["e3t_b"], gf["e3t_n"], gf["e3t_a"])
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    t = [dev(a) for a in (e2u, e1v, e3u, e3v, un, vn, wn, gf["ptb"], gf["ptn"])]
    for kt, neuler, want_r2dt in ((nit000, 0, rdt), (nit000 + 1, 0, 2 * rdt), (nit000 + 5, 0, 2 * rdt)):
        ref = gf["pta"].copy()
        w.tra_adv_fct(want_r2dt, [zu], [zv], [zw], [gf["ptb"]], [gf["ptn"]], [ref], 2, 4, 4)
        tsa = dev(gf["pta"])
        ctx.tra_adv(kt, nit000, neuler, rdt, *t, tsa, 2, 4, 4)
        ctx.synchronize()
        assert np.array_equal(tsa.cpu().numpy(), ref), (kt, neuler)
    ref = np.zeros_like(trb)
    w.tra_adv_fct(2 * rdt, [zu], [zv], [zw], [trb], [trn], [ref], 5, 2, 2)
    tra = torch.zeros(trb.shape, dtype=torch.float64, device="cuda")
    ctx.trc_adv(nit000 + 5, nit000, 2 * rdt, dev(trb), dev(trn), tra, 5, 2, 2)
    ctx.synchronize()
    assert np.array_equal(tra.cpu().numpy(), ref)
    ctx.close()
    c2 = N.FctContext(dom, 0)
    c2.set_domain_arrays(gf["tmask"], gf["umask"], gf["vmask"], gf["wmask"], gf["e1e2t"], gf["r1_e1e2t"], gf["mikt"], gf["mbkt"])
    c2.set_e3t(gf["e3t_b"], gf["e3t_n"], gf["e3t_a"])
    with pytest.raises(N.NemoFctError, match="tra_adv"):
        c2.trc_adv(1, 1, rdt, dev(trb), dev(trn), tra, 5, 2, 2)       # transports not built yet
    # a host whose namelist adds eddy-induced / Stokes / mle transports (traadv.F90:103-129) is refused, not served a wrong answer
    c2.declare_transport_options(ln_ldfeiv=True)
    with pytest.raises(N.NemoFctError, match="ldf_eiv_trp"):
        c2.tra_adv(nit000, nit000, 0, rdt, *t, dev(gf["pta"]), 2, 4, 4)
    c2.declare_transport_options()
    c2.tra_adv(nit000, nit000, 0, rdt, *t, dev(gf["pta"]), 2, 4, 4)
    c2.close(); w.close()
