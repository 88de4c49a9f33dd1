"""CPU tests of the oracle's tra_adv_mus (traadv_mus.F90:55-273) and tra_nxt (tranxt.F90:65-380) restatements.

The reference holds no golden vectors for these routines either ("parity unpinned"); the pins are invariants derived
from the reference code and an independent numpy re-derivation of the closed-form parts."""
import numpy as np
import pytest

from oracle import oracle as O
import helpers as H

JPK = 9


def _setup(jpiglo, jpjglo, jperio, kjpt=2, seed=3, **kw):
    gf = H.random_fields(O, jpiglo, jpjglo, JPK, jperio, kjpt, seed=seed, **kw)
    mx = H.mus_extra_fields(O, gf, jpiglo, jpjglo, JPK, jperio, seed=seed, runoff=True)
    return gf, mx


@pytest.mark.parametrize("jperio,jpni,jpnj", [(0, 2, 2), (1, 3, 2), (4, 2, 2), (6, 2, 3), (4, 3, 1), (7, 2, 2)])
def test_mus_decomposition_invariance(jperio, jpni, jpnj):
    """SETTE's property (1x1 vs jpni x jpnj identical, bit for bit) for the MUSCL scheme and its two exchanges."""
    jpiglo, jpjglo = 30, 26
    gf, mx = _setup(jpiglo, jpjglo, jperio)
    ref, _ = H.oracle_mus(O, gf, mx, jpiglo, jpjglo, JPK, jperio, 1, 1, 2)
    got, _ = H.oracle_mus(O, gf, mx, jpiglo, jpjglo, JPK, jperio, jpni, jpnj, 2)
    inner = (slice(None), slice(0, JPK - 1), slice(1, -1), slice(1, -1))
    assert np.array_equal(ref[inner], got[inner])
    assert not np.array_equal(ref[inner], gf["pta"][inner])


def test_mus_nompi_equals_mpi_single_domain():
    gf, mx = _setup(24, 20, 4)
    a, _ = H.oracle_mus(O, gf, mx, 24, 20, JPK, 4, 1, 1, 2, key_mpp_mpi=True)
    b, _ = H.oracle_mus(O, gf, mx, 24, 20, JPK, 4, 1, 1, 2, key_mpp_mpi=False)
    assert np.array_equal(a, b)


@pytest.mark.parametrize("ln_linssh,ln_isfcav", [(False, False), (True, False), (True, True)])
def test_mus_upstream_limit_matches_numpy(ln_linssh, ln_isfcav):
    """xind = 0 switches the slopes off: the scheme must reduce to first-order upstream fluxes, which numpy can state
    independently (zalpha is 0 or 1, so the blend is exact) -- bit for bit."""
    jpiglo, jpjglo, jperio = 22, 18, 1
    gf, mx = _setup(jpiglo, jpjglo, jperio, ln_linssh=ln_linssh, ln_isfcav=ln_isfcav)
    got, _ = H.oracle_mus(O, gf, mx, jpiglo, jpjglo, JPK, jperio, 1, 1, 2, ln_linssh=ln_linssh, ln_isfcav=ln_isfcav,
                          xind_zero=True)
    ptb, pun, pvn, pwn = gf["ptb"], gf["pun"], gf["pvn"], gf["pwn"]
    exp = gf["pta"].copy()
    for jn in range(2):
        t = ptb[jn]
        fx = np.zeros_like(t); fy = np.zeros_like(t); fz = np.zeros_like(t)
        fx[:, :, :-1] = pun[:, :, :-1] * np.where(pun[:, :, :-1] >= 0, t[:, :, :-1], t[:, :, 1:])
        fy[:, :-1, :] = pvn[:, :-1, :] * np.where(pvn[:, :-1, :] >= 0, t[:, :-1, :], t[:, 1:, :])
        # vertical: face jk+1 (jk = 1..jpk-2) carries ptb(jk+1) when pwn >= 0 (upward), times wmask(jk)
        fz[1:-1] = pwn[1:-1] * np.where(pwn[1:-1] >= 0, t[1:-1], t[:-2]) * gf["wmask"][:-2]
        if ln_linssh:
            if ln_isfcav:
                k0 = gf["mikt"] - 1
                jj, ii = np.meshgrid(np.arange(t.shape[1]), np.arange(t.shape[2]), indexing="ij")
                fz[k0, jj, ii] = pwn[k0, jj, ii] * t[k0, jj, ii]
            else:
                fz[0] = pwn[0] * t[0]
        r1 = gf["r1_e1e2t"][None]
        e3 = gf["e3t_n"]
        a = exp[jn]
        hdiv = (fx[:, 1:-1, 1:-1] - fx[:, 1:-1, :-2]) + fy[:, 1:-1, 1:-1]
        hdiv = hdiv - fy[:, :-2, 1:-1]
        new = a[:-1, 1:-1, 1:-1] - (hdiv * r1[:, 1:-1, 1:-1] / e3[:, 1:-1, 1:-1])[:-1]
        new = new - ((fz[:-1] - fz[1:])[:, 1:-1, 1:-1] * r1[:, 1:-1, 1:-1] / e3[:-1, 1:-1, 1:-1])
        a[:-1, 1:-1, 1:-1] = new
    inner = (slice(None), slice(0, JPK - 1), slice(1, -1), slice(1, -1))
    assert np.array_equal(got[inner], exp[inner])


def test_mus_conservation_and_constancy():
    """Flux form: the volume-weighted sum of the trend vanishes on a cyclic domain (the fluxes telescope); a uniform
    tracer advected by a non-divergent flow gets a zero trend up to rounding."""
    jpiglo, jpjglo, jperio, kjpt = 34, 28, 1, 2
    w = O.World(jpiglo, jpjglo, JPK, jperio, 1, 1)
    gf = H.global_bench_fields(O, jpiglo, jpjglo, JPK, jperio, kjpt, cfl=0.3, ln_linssh=True)
    mx = dict(r1_e1e2u=gf["r1_e1e2t"].copy(), r1_e1e2v=gf["r1_e1e2t"].copy(), e3u_n=gf["e3t_n"].copy(),
              e3v_n=gf["e3t_n"].copy(), e3w_n=gf["e3t_n"].copy())
    w.close()
    pta0 = gf["pta"].copy()
    got, _ = H.oracle_mus(O, gf, mx, jpiglo, jpjglo, JPK, jperio, 1, 1, kjpt, ln_linssh=False)
    vol = (gf["e1e2t"][None] * gf["e3t_n"])[None]
    d = (got - pta0) * vol
    tot = d[:, :-1, 1:-1, 1:-1].sum(axis=(1, 2, 3))
    scale = np.abs(d[:, :-1, 1:-1, 1:-1]).sum(axis=(1, 2, 3))
    assert np.all(np.abs(tot) <= 1e-11 * scale), (tot, scale)
    # constancy
    gc = dict(gf)
    gc["ptb"] = np.full_like(gf["ptb"], 7.25) * gf["tmask"][None]
    gc["pta"] = np.zeros_like(gf["pta"])
    got, _ = H.oracle_mus(O, gc, mx, jpiglo, jpjglo, JPK, jperio, 1, 1, kjpt, ln_linssh=True)
    # with the surface flux pwn(1)*ptb(1) of the linear free surface the trend of a uniform tracer is -c * div(transports)
    div = np.zeros_like(gf["pun"])
    div[:-1, 1:, 1:] = ((gf["pun"][:-1, 1:, 1:] - gf["pun"][:-1, 1:, :-1]) + (gf["pvn"][:-1, 1:, 1:] - gf["pvn"][:-1, :-1, 1:])
                        + (gf["pwn"][:-1, 1:, 1:] - gf["pwn"][1:, 1:, 1:]))
    exp = -7.25 * div * gf["r1_e1e2t"][None] / gf["e3t_n"]
    err = np.abs(got[0] - exp)[:-1, 2:-2, 2:-2].max()
    assert err <= 1e-12 * np.abs(gf["pun"]).max() * 7.25 * gf["r1_e1e2t"].max() / gf["e3t_n"].min(), err


def test_mus_xind_runoff_mask():
    gf, mx = _setup(20, 18, 0)
    w = O.World(20, 18, JPK, 0, 1, 1)
    d = w.doms[0]
    d.set_fields(*[gf[k] for k in H.DOM_KEYS])
    xi = d.mus_xind(True, mx["rnfmsk"], mx["rnfmsk_z"])
    exp = np.ones_like(xi)
    exp[:-1] = 1.0 - np.maximum(mx["rnfmsk"][None] * mx["rnfmsk_z"][:-1, None, None], 0.0) * gf["tmask"][:-1]
    assert np.array_equal(xi, exp)
    assert np.array_equal(d.mus_xind(False), np.ones_like(xi))
    w.close()
    # the indicator matters: result with the runoff mask differs from the plain scheme
    a, _ = H.oracle_mus(O, gf, mx, 20, 18, JPK, 0, 1, 1, 2, ld_msc_ups=True)
    b, _ = H.oracle_mus(O, gf, mx, 20, 18, JPK, 0, 1, 1, 2, ld_msc_ups=False)
    assert not np.array_equal(a, b)


# ---- tra_nxt ---------------------------------------------------------------------------------------------------------
def _nxt_world(jpiglo, jpjglo, jperio, jpni, jpnj, gf, ln_linssh):
    w = O.World(jpiglo, jpjglo, JPK, jperio, jpni, jpnj)
    loc = {k: w.scatter(gf[k]) for k in H.DOM_KEYS + ("ptb", "ptn", "pta")}
    for r, d in enumerate(w.doms):
        d.set_fields(*[loc[k][r] for k in H.DOM_KEYS], ln_linssh=ln_linssh)
    return w, loc


def _surface_forcing(rng, w, jpj, jpi, kjpt):
    """global 2-D forcing fields, lbc-consistent"""
    arrs = {k: rng.standard_normal((1, jpj, jpi)) * 1e-4 for k in ("emp_b", "emp", "fwfisf_b", "fwfisf", "rnf_b", "rnf")}
    sbc = rng.standard_normal((kjpt, jpj, jpi)) * 1e-5
    sbc_b = rng.standard_normal((kjpt, jpj, jpi)) * 1e-5
    w.lbc_lnk([[a] for a in arrs.values()], "T" * 6, [1.0] * 6)
    w.lbc_lnk([[sbc], [sbc_b]], "TT", [1.0] * 2)                 # kjpt levels each
    return {k: np.ascontiguousarray(v[0]) for k, v in arrs.items()}, sbc, sbc_b


@pytest.mark.parametrize("ln_linssh", [True, False])
def test_nxt_matches_numpy_and_decomposition(ln_linssh):
    jpiglo, jpjglo, jperio, kjpt = 26, 22, 4, 2
    rng = np.random.default_rng(5)
    gf = H.random_fields(O, jpiglo, jpjglo, JPK, jperio, kjpt, seed=9, ln_linssh=ln_linssh)
    w1 = O.World(jpiglo, jpjglo, JPK, jperio, 1, 1)
    f2d, sbc, sbc_b = _surface_forcing(rng, w1, jpjglo, jpiglo, kjpt)
    w1.close()
    atfp, rdt, r1_rau0 = 0.1, 900.0, 1.0 / 1026.0
    res = {}
    for (jpni, jpnj) in ((1, 1), (2, 2)):
        w, loc = _nxt_world(jpiglo, jpjglo, jperio, jpni, jpnj, gf, ln_linssh)
        lf = {k: w.scatter(v) for k, v in f2d.items()}
        ls, lsb = w.scatter(sbc), w.scatter(sbc_b)
        forc = [O.NxtForcing(atfp=atfp, r1_rau0=r1_rau0, **{k: lf[k][r] for k in lf}) for r in range(w.jpnij)]
        w.tra_nxt(5, 1, False, rdt, "TRA", forc, loc["ptb"], loc["ptn"], loc["pta"], kjpt, ls, lsb)
        res[(jpni, jpnj)] = [w.gather(loc[k], gf[k].copy()) for k in ("ptb", "ptn", "pta")]
        if jpni == 1:
            full = [loc[k][0].copy() for k in ("ptb", "ptn", "pta")]
        w.close()
    for a, b in zip(res[(1, 1)], res[(2, 2)]):           # global halo columns are nobody's interior under E-W cyclicity
        assert np.array_equal(a[..., 1:-1, 1:-1], b[..., 1:-1, 1:-1])
    # independent numpy statement of the filter on the interior
    tb, tn, ta = gf["ptb"], gf["ptn"], gf["pta"]
    sl = (slice(None), slice(0, JPK - 1), slice(1, -1), slice(1, -1))
    if ln_linssh:
        exp_b = tn + atfp * (ta - 2.0 * tn + tb)
    else:
        eb, en, ea = gf["e3t_b"][None], gf["e3t_n"][None], gf["e3t_a"][None]
        tc_b, tc_n, tc_a = tb * eb, tn * en, ta * ea
        e_f = en + atfp * (ea - 2.0 * en + eb)
        tc_f = tc_n + atfp * (tc_a - 2.0 * tc_n + tc_b)
        first = (np.arange(1, JPK + 1)[:, None, None] == gf["mikt"][None])[None]
        zfact1 = atfp * rdt; zfact2 = zfact1 * r1_rau0
        e_f1 = e_f - zfact2 * ((f2d["emp_b"] - f2d["emp"]) + (f2d["fwfisf_b"] - f2d["fwfisf"]))[None, None]
        e_f1 = e_f1 - zfact2 * (-(f2d["rnf_b"] - f2d["rnf"]))[None, None]
        e_f = np.where(first, e_f1, e_f)
        tc_f = np.where(first, tc_f - zfact1 * (sbc - sbc_b)[:, None], tc_f)
        exp_b = tc_f * (1.0 / e_f)
    got_b, got_n, got_a = res[(1, 1)]
    assert np.array_equal(got_b[sl], exp_b[sl])
    assert np.array_equal(got_n[sl], ta[sl])
    # halos of all three fields hold lbc_lnk('T', +1) of the interior: applying lbc_lnk again changes nothing
    w, _ = _nxt_world(jpiglo, jpjglo, jperio, 1, 1, gf, ln_linssh)
    again = [a.copy() for a in full]
    w.lbc_lnk([[a.reshape(-1, a.shape[-2], a.shape[-1])] for a in again], "TTT", [1.0] * 3)
    w.close()
    for a, b in zip(full, again):
        assert np.array_equal(a, b)


def test_nxt_vvl_reduces_to_fix():
    jpiglo, jpjglo, jperio, kjpt = 20, 18, 1, 2
    gf = H.random_fields(O, jpiglo, jpjglo, JPK, jperio, kjpt, seed=2, ln_linssh=True)     # e3t_b = e3t_n = e3t_a
    out = {}
    for lin in (True, False):
        w, loc = _nxt_world(jpiglo, jpjglo, jperio, 1, 1, gf, lin)
        w.tra_nxt(3, 1, False, 900.0, "TRA", [O.NxtForcing(atfp=0.1)], loc["ptb"], loc["ptn"], loc["pta"], kjpt)
        out[lin] = loc["ptb"][0]
        w.close()
    assert H.max_rel_diff(out[False], out[True]) < 1e-13


@pytest.mark.parametrize("cdtype", ["TRA", "TRC"])
def test_nxt_euler_swap(cdtype):
    jpiglo, jpjglo, jperio, kjpt = 20, 18, 6, 3
    gf = H.random_fields(O, jpiglo, jpjglo, JPK, jperio, kjpt, seed=4)
    w, loc = _nxt_world(jpiglo, jpjglo, jperio, 2, 1, gf, False)
    tb0 = [a.copy() for a in loc["ptb"]]
    tn0 = [a.copy() for a in loc["ptn"]]
    w.tra_nxt(1, 1, True, 900.0, cdtype, [O.NxtForcing(), O.NxtForcing()], loc["ptb"], loc["ptn"], loc["pta"], kjpt)
    for r in range(2):
        assert np.array_equal(loc["ptn"][r][:, :-1], loc["pta"][r][:, :-1])
        assert np.array_equal(loc["ptn"][r][:, -1], tn0[r][:, -1])                # level jpk untouched
        if cdtype == "TRC":
            assert np.array_equal(loc["ptb"][r][:, :-1], loc["pta"][r][:, :-1])
        else:
            assert np.array_equal(loc["ptb"][r], tb0[r])
    w.close()
