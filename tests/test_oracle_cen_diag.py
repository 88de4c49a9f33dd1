"""CPU tests of the oracle's tra_adv_cen (traadv_cen.F90:46-204) and of the l_trd / l_hst / l_ptr hooks of tra_adv_fct
(traadv_fct.F90:96-112, 172-176, 299-316).  No reference golden vectors exist ("parity unpinned"): invariants + numpy."""
import numpy as np
import pytest

from oracle import oracle as O
import helpers as H

JPK = 9
INNER = (slice(None), slice(0, JPK - 1), slice(1, -1), slice(1, -1))


@pytest.mark.parametrize("jperio,jpni,jpnj", [(0, 2, 2), (1, 3, 2), (4, 2, 2), (6, 2, 3)])
@pytest.mark.parametrize("v", [2, 4])
def test_cen2_decomposition_invariance(jperio, jpni, jpnj, v):
    gf = H.random_fields(O, 30, 26, JPK, jperio, 2, seed=13)
    ref, _ = H.oracle_cen(O, gf, 30, 26, JPK, jperio, 1, 1, 2, 2, v, poison=True)
    got, _ = H.oracle_cen(O, gf, 30, 26, JPK, jperio, jpni, jpnj, 2, 2, v, poison=True)
    assert np.isfinite(ref[INNER]).all()
    assert np.array_equal(ref[INNER], got[INNER])
    assert not np.array_equal(ref[INNER], gf["pta"][INNER])


@pytest.mark.parametrize("ln_linssh,ln_isfcav", [(False, False), (True, False), (True, True)])
def test_cen22_matches_numpy(ln_linssh, ln_isfcav):
    """independent vectorised statement of the 2nd-order scheme, bit for bit"""
    gf = H.random_fields(O, 24, 20, JPK, 1, 2, seed=5, ln_linssh=ln_linssh, ln_isfcav=ln_isfcav)
    got, _ = H.oracle_cen(O, gf, 24, 20, JPK, 1, 1, 1, 2, 2, 2, ln_linssh=ln_linssh, ln_isfcav=ln_isfcav, poison=True)
    exp = gf["pta"].copy()
    pun, pvn, pwn = gf["pun"], gf["pvn"], gf["pwn"]
    for jn in range(2):
        t = gf["ptn"][jn]
        fx = np.zeros_like(t); fy = np.zeros_like(t); fz = np.zeros_like(t)
        fx[:, :, :-1] = 0.5 * pun[:, :, :-1] * (t[:, :, :-1] + t[:, :, 1:])
        fy[:, :-1, :] = 0.5 * pvn[:, :-1, :] * (t[:, :-1, :] + t[:, 1:, :])
        fz[1:] = 0.5 * pwn[1:] * (t[1:] + t[:-1]) * gf["wmask"][1:]
        if ln_linssh:
            if ln_isfcav:
                k0 = gf["mikt"] - 1
                jj, ii = np.meshgrid(np.arange(t.shape[1]), np.arange(t.shape[2]), indexing="ij")
                fz[k0, jj, ii] = pwn[k0, jj, ii] * t[k0, jj, ii]
            else:
                fz[0] = pwn[0] * t[0]
        div = (fx[:-1, 1:-1, 1:-1] - fx[:-1, 1:-1, :-2]) + fy[:-1, 1:-1, 1:-1]
        div = div - fy[:-1, :-2, 1:-1]
        div = div + fz[:-1, 1:-1, 1:-1]
        div = div - fz[1:, 1:-1, 1:-1]
        exp[jn][:-1, 1:-1, 1:-1] = exp[jn][:-1, 1:-1, 1:-1] - div * gf["r1_e1e2t"][None, 1:-1, 1:-1] / gf["e3t_n"][:-1, 1:-1, 1:-1]
    assert np.array_equal(got[INNER], exp[INNER])


def test_cen4_reference_defect_is_confined():
    """kn_cen_h = 4 (traadv_cen.F90:128-137) reads the never-assigned zwy(:,1,:): with NaN-poisoned automatic arrays the
    damage must be exactly the first interior row, nothing else; elsewhere the result is finite and differs from order 2."""
    gf = H.random_fields(O, 26, 22, JPK, 1, 2, seed=8)
    p4, _ = H.oracle_cen(O, gf, 26, 22, JPK, 1, 1, 1, 2, 4, 4, poison=True)
    bad = ~np.isfinite(p4[INNER])
    assert bad[:, :, 0, :].all() and not bad[:, :, 1:, :].any()
    z4, _ = H.oracle_cen(O, gf, 26, 22, JPK, 1, 1, 1, 2, 4, 4, poison=False)
    assert np.isfinite(z4).all()
    assert np.array_equal(z4[INNER][:, :, 1:, :], p4[INNER][:, :, 1:, :])
    p2, _ = H.oracle_cen(O, gf, 26, 22, JPK, 1, 1, 1, 2, 2, 4, poison=True)
    assert not np.array_equal(z4[INNER], p2[INNER])


def test_fct_equals_cen_when_limiter_is_idle():
    """SURVEY.md par. 8c invariant 5: smooth monotone field, BENCH's own tiny Courant numbers => every limiter coefficient
    is 1 away from the boundaries and tra_adv_fct must give the trend of tra_adv_cen of the same order (two independently
    restated routines; the sum upstream + (centred - upstream) differs from centred by rounding only)."""
    G, GJ, jperio, kjpt = 36, 30, 0, 2
    gf = H.global_bench_fields(O, G, GJ, JPK, jperio, kjpt)
    ii = np.arange(G)[None, None, :]; jj = np.arange(GJ)[None, :, None]; kk = np.arange(JPK)[:, None, None]
    t = (10.0 + 0.01 * ii + 0.02 * jj + 0.1 * kk) * gf["tmask"]
    gf["ptn"] = np.stack([t, 2.0 * t]); gf["ptb"] = gf["ptn"].copy()
    fct, _, _ = H.oracle_fct(O, gf, G, GJ, JPK, jperio, 1, 1, kjpt, 2, 2)
    cen, _ = H.oracle_cen(O, gf, G, GJ, JPK, jperio, 1, 1, kjpt, 2, 2)
    core = (slice(None), slice(2, JPK - 3), slice(4, -4), slice(4, -4))
    a = (fct - gf["pta"])[core]
    b = (cen - gf["pta"])[core]
    assert np.abs(b).max() > 0
    assert np.abs(a - b).max() <= 1e-11 * np.abs(b).max(), (np.abs(a - b).max(), np.abs(b).max())


@pytest.mark.parametrize("h,v", [(2, 2), (4, 4)])
def test_fct_trend_diag_hooks(h, v):
    """ztrdx/y/z = upstream + limited anti-diffusive fluxes: their divergence must reproduce the trend tra_adv_fct added to
    pta (up to the rounding of adding the two parts separately), and switching the hooks on must not change pta."""
    G, GJ, jperio, kjpt = 30, 26, 4, 2
    gf = H.random_fields(O, G, GJ, JPK, jperio, kjpt, seed=21)
    ref, _, _ = H.oracle_fct(O, gf, G, GJ, JPK, jperio, 1, 1, kjpt, h, v)
    w = O.World(G, GJ, JPK, jperio, 1, 1)
    d = w.doms[0]
    d.set_fields(*[gf[k] for k in H.DOM_KEYS])
    trd = [np.full((kjpt,) + d.shape3, np.nan) for _ in range(3)]
    d.set_diag(*trd)
    pta = gf["pta"].copy()
    w.tra_adv_fct(gf["p2dt"], [gf["pun"]], [gf["pvn"]], [gf["pwn"]], [gf["ptb"]], [gf["ptn"]], [pta], kjpt, h, v)
    w.close()
    assert np.array_equal(pta, ref)
    tx, ty, tz = trd
    assert np.isfinite(tx[:, :, :-1, :-1]).all() and np.isfinite(ty[:, :, :-1, :-1]).all() and np.isfinite(tz).all()
    div = (tx[:, :-1, 1:-1, 1:-1] - tx[:, :-1, 1:-1, :-2]) + (ty[:, :-1, 1:-1, 1:-1] - ty[:, :-1, :-2, 1:-1]) \
        + (tz[:, :-1, 1:-1, 1:-1] - tz[:, 1:, 1:-1, 1:-1])
    exp = -div * gf["r1_e1e2t"][None, None, 1:-1, 1:-1] / gf["e3t_n"][None, :-1, 1:-1, 1:-1]
    got = (ref - gf["pta"])[INNER]
    scale = np.abs(exp).max()
    assert np.abs(got - exp * gf["tmask"][None, :-1, 1:-1, 1:-1]).max() <= 1e-10 * scale
