"""Oracle self-consistency: the invariants of SURVEY.md par. 8(c) that stand in for the golden vectors the
reference does not ship.  All derived from reference code (file:line in each test)."""
import numpy as np
import pytest

import helpers as H

G, GJ, K = 34, 27, 9


@pytest.mark.parametrize("jperio", [0, 1, 4, 6])
@pytest.mark.parametrize("hv", [(2, 2), (4, 4)])
def test_decomposition_invariance_bitwise(O, jperio, hv):
    """1x1 vs jpni x jpnj give identical interiors (what SETTE checks through run.stat, stpctl.F90:134,192), with the
    no-gather and the gather fold (chap_misc.tex:314-316), with and without key_mpp_mpi."""
    h, v = hv
    gf = H.random_fields(O, G, GJ, K, jperio, kjpt=2, seed=3)
    mono, _, _ = H.oracle_fct(O, gf, G, GJ, K, jperio, 1, 1, 2, h, v)
    assert not np.isnan(mono).any(), "an undefined (NaN-poisoned) work-array value reached pta"
    nompi, _, _ = H.oracle_fct(O, gf, G, GJ, K, jperio, 1, 1, 2, h, v, key_mpp_mpi=False)
    assert np.array_equal(mono, nompi)
    for lay in [(2, 1), (1, 2), (2, 2), (3, 2), (4, 2)]:
        for nogather in (True, False):
            dec, _, _ = H.oracle_fct(O, gf, G, GJ, K, jperio, lay[0], lay[1], 2, h, v, ln_nnogather=nogather)
            assert np.array_equal(dec, mono), (lay, nogather)


def _content(O, gf, pta, jperio):
    """sum over tmask_i of e1e2t*e3t_n*pta in double-double (glob_sum, lib_fortran_generic.h90:32-65)"""
    w = O.World(G, GJ, K, jperio, 1, 1)
    out = []
    for n in range(pta.shape[0]):
        vol = np.ascontiguousarray(gf["e1e2t"][None] * gf["e3t_n"] * pta[n])
        out.append(O.glob_sum(w, [vol], [np.ascontiguousarray(gf["tmask_i"])])[0])
    w.close()
    return np.array(out)


@pytest.mark.parametrize("jperio", [0, 1, 7])
@pytest.mark.parametrize("hv", [(2, 2), (4, 4)])
def test_global_tracer_content_is_conserved(O, jperio, hv):
    """flux form (traadv_fct.F90:162-164, 291-294) telescopes: with closed / periodic boundaries and a
    non-linear free surface (no surface flux term) the volume integral of the advective trend vanishes."""
    h, v = hv
    gf = H.random_fields(O, G, GJ, K, jperio, kjpt=2, seed=11, land=0.2)
    gf["pta"][:] = 0.0
    out, _, _ = H.oracle_fct(O, gf, G, GJ, K, jperio, 1, 1, 2, h, v)
    total = _content(O, gf, out, jperio)
    scale = _content(O, gf, np.abs(out), jperio)
    assert np.all(np.abs(total) <= 1e-12 * scale), (total, scale)


def test_constancy_is_preserved(O):
    """uniform tracer + discretely non-divergent transports + steady e3t: every anti-diffusive flux is
    0.5*u*2c - u*c = 0 and the upstream divergence is c * div(U) = 0 (traadv_fct.F90:131-132, 186-187)."""
    # linear free surface: steady e3t, and the surface flux pwn(1)*ptb(1) closes the column budget (:146-156)
    gf = H.global_bench_fields(O, G, GJ, K, 1, 2, cfl=0.3, ln_linssh=True)
    c = 3.25
    gf["ptb"] = c * np.broadcast_to(gf["tmask"], gf["ptb"].shape).copy()
    gf["ptn"] = gf["ptb"].copy()
    gf["pta"][:] = 0.0
    for h, v in [(2, 2), (4, 4)]:
        out, _, _ = H.oracle_fct(O, gf, G, GJ, K, 1, 1, 1, 2, h, v, ln_linssh=True)
        trend_scale = c * np.abs(gf["pun"]).max() / (gf["e1e2t"].min() * gf["e3t_n"].min())
        assert np.abs(out).max() <= 1e-10 * trend_scale


def test_limiter_keeps_the_update_within_local_bounds(O):
    """nonosc (traadv_fct.F90:361-396): paft + dt * (limited antidiffusive divergence)/(e1e2t*e3t_n) stays within
    [zdo, zup], the extrema of (pbef, paft) over the 7-point neighbourhood."""
    jperio = 1
    gf = H.random_fields(O, G, GJ, K, jperio, kjpt=1, seed=21, land=0.1, cfl=0.3)
    cap = {"jn": 1, "names": ["zwi", "paa", "pbb", "pcc"]}
    H.oracle_fct(O, gf, G, GJ, K, jperio, 1, 1, 1, 2, 2, capture=cap)
    c = cap["out"][0]
    paft, paa, pbb, pcc = c["zwi"], c["paa"], c["pbb"], c["pcc"]
    tm, pbef = gf["tmask"], gf["ptb"][0]
    big = 1e40
    bup = np.maximum(pbef * tm - big * (1 - tm), paft * tm - big * (1 - tm))
    bdo = np.minimum(pbef * tm + big * (1 - tm), paft * tm + big * (1 - tm))
    i = slice(1, -1)
    nb_up = [bup[:-1, i, i], bup[:-1, i, :-2], bup[:-1, i, 2:], bup[:-1, :-2, i], bup[:-1, 2:, i], bup[1:, i, i],
             np.concatenate([bup[:1, i, i], bup[:-2, i, i]], 0)]
    nb_do = [bdo[:-1, i, i], bdo[:-1, i, :-2], bdo[:-1, i, 2:], bdo[:-1, :-2, i], bdo[:-1, 2:, i], bdo[1:, i, i],
             np.concatenate([bdo[:1, i, i], bdo[:-2, i, i]], 0)]
    zup = np.max(nb_up, 0); zdo = np.min(nb_do, 0)
    div = (paa[:-1, i, i] - paa[:-1, i, :-2] + pbb[:-1, i, i] - pbb[:-1, :-2, i] + pcc[:-1, i, i] - pcc[1:, i, i])
    new = paft[:-1, i, i] - gf["p2dt"] * div / (gf["e1e2t"][i, i][None] * gf["e3t_n"][:-1, i, i])
    wet = tm[:-1, i, i] == 1.0
    tol = 1e-9 * np.abs(pbef).max()
    assert np.all(new[wet] <= zup[wet] + tol) and np.all(new[wet] >= zdo[wet] - tol)
    assert (np.abs(div[wet]) > 0).mean() > 0.5, "the limiter test must exercise non-zero fluxes"


def test_interp_4th_cpt_reproduces_linear_profiles(O):
    """interp_4th_cpt (traadv_fct.F90:535-614): compact scheme + 2nd-order top/bottom rows are exact for a profile
    linear in k; output defined on (2:jpim1, 2:jpjm1, 2:jpkm1)."""
    w = O.World(12, 10, 15, 0)
    d = w.doms[0]
    k_bot = np.full((d.jpj, d.jpi), d.jpk - 1, np.int32); k_top = np.ones_like(k_bot)
    k_bot[3, 4] = 7                                             # one partial column
    m = w.dom_msk([k_top], [k_bot])[0]
    one2 = np.ones((d.jpj, d.jpi)); one3 = np.ones(d.shape3)
    d.set_fields(m["tmask"], m["umask"], m["vmask"], m["wmask"], one3, one3, one3, one2, one2, m["mikt"], m["mbkt"])
    kk = np.arange(1, d.jpk + 1, dtype=np.float64)[:, None, None]
    pt_in = np.ascontiguousarray(np.broadcast_to(2.0 + 0.5 * kk, d.shape3))
    pt_out = np.full(d.shape3, np.nan)
    O.lib().interp_4th_cpt(d.h, pt_in.ctypes.data, pt_out.ctypes.data)
    expect = 2.0 + 0.5 * (kk - 0.5)                            # w-point k sits between T-points k-1 and k
    got = pt_out[1:-1, 1:-1, 1:-1].copy()
    want = np.broadcast_to(expect, d.shape3)[1:-1, 1:-1, 1:-1].copy()
    got[:, 2, 3] = want[:, 2, 3]                                # the partial column is checked separately below
    assert np.allclose(got, want, rtol=0, atol=1e-13)
    assert np.isnan(pt_out[0]).all() and np.isnan(pt_out[-1]).all()
    # below the bottom of the partial column the rows are identity with zero RHS
    assert np.all(pt_out[8:-1, 3, 4] == 0.0) and abs(pt_out[6, 3, 4] - (2.0 + 0.5 * 6.5)) < 1e-13
    w.close()


def test_sign_follows_key_nosignedzero(O):
    """SIGN(a,b) = |a| if b >= 0 else -|a|: -0.0 counts as positive (lib_fortran.F90:339-351, nemogcm.F90:599-601)"""
    s = O.lib().sign_nosignedzero
    assert s(0.5, -0.0) == 0.5 and s(0.5, 0.0) == 0.5 and s(0.5, -1e-300) == -0.5 and s(-0.5, 3.0) == 0.5


def test_ddpdd_is_a_compensated_sum(O):
    """DDPDD (lib_fortran.F90:300-332): the double-double sum of 1e16, 1, -1e16 is exactly 1"""
    import ctypes as C
    tot = (C.c_double * 2)(0.0, 0.0)
    for x in (1e16, 1.0, -1e16):
        O.lib().ddpdd((C.c_double * 2)(x, 0.0), tot)
    assert tot[0] + tot[1] == 1.0


def test_degenerate_limiter_equals_centred_scheme(O):
    """smooth field + tiny Courant numbers: every limiter coefficient is 1 (zau = zbu = 1), so the FCT trend equals
    the centred 2nd-order flux divergence of tra_adv_cen (traadv_cen.F90:100-130, 178-190) -- cross-check between
    the FCT restatement and an independent vectorised formula."""
    jperio = 1
    gf = H.global_bench_fields(O, G, GJ, K, jperio, 1)          # native BENCH velocities: CFL ~ 1e-4
    gf["ptn"] = gf["ptb"].copy(); gf["pta"][:] = 0.0
    out, _, _ = H.oracle_fct(O, gf, G, GJ, K, jperio, 1, 1, 1, 2, 2)
    t, u, v_, w_ = gf["ptn"][0], gf["pun"], gf["pvn"], gf["pwn"]
    fx = np.zeros_like(t); fy = np.zeros_like(t); fz = np.zeros_like(t)
    fx[:, :, :-1] = 0.5 * u[:, :, :-1] * (t[:, :, :-1] + t[:, :, 1:])
    fy[:, :-1, :] = 0.5 * v_[:, :-1, :] * (t[:, :-1, :] + t[:, 1:, :])
    fz[1:] = 0.5 * w_[1:] * (t[1:] + t[:-1]) * gf["wmask"][1:]
    i = slice(1, -1)
    div = (fx[:-1, i, i] - fx[:-1, i, :-2] + fy[:-1, i, i] - fy[:-1, :-2, i] + fz[:-1, i, i] - fz[1:, i, i])
    cen = -div / (gf["e1e2t"][i, i][None] * gf["e3t_n"][:-1, i, i]) * gf["tmask"][:-1, i, i]
    got = out[0][:-1, i, i]
    assert np.abs(got - cen).max() <= 1e-9 * np.abs(cen).max()
