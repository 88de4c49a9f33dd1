"""GPU parity of the widened rows (SURVEY.md par. 8f): tra_adv_mus (MUSCL) and tra_nxt / trc_nxt through the C ABI,
bit for bit against the oracle's restatements of traadv_mus.F90:55-273 and tranxt.F90:65-380."""
import importlib

import numpy as np
import pytest
import torch

from oracle import oracle as O
import helpers as H

pytestmark = pytest.mark.gpu
N = importlib.import_module("nemo-fmi-devel_b200")
JPK = 11


def _mus_case(jpiglo, jpjglo, jperio, kjpt, seed, **kw):
    gf = H.random_fields(O, jpiglo, jpjglo, JPK, jperio, kjpt, seed=seed, **kw)
    mx = H.mus_extra_fields(O, gf, jpiglo, jpjglo, JPK, jperio, seed=seed, runoff=True)
    return gf, mx


@pytest.mark.parametrize("jperio", [0, 1, 2, 3, 4, 5, 6, 7])
@pytest.mark.parametrize("schedule", [0, 1, 2])
def test_mus_parity_all_boundaries(jperio, schedule):
    """single subdomain, every lateral boundary type, both schedules: identical to the oracle, whole array
    (interior = result, halos and level jpk = untouched input)."""
    jpiglo, jpjglo, kjpt = 44, 38, 3
    gf, mx = _mus_case(jpiglo, jpjglo, jperio, kjpt, seed=20 + jperio)
    ref, _ = H.oracle_mus(O, gf, mx, jpiglo, jpjglo, JPK, jperio, 1, 1, kjpt)
    got, loc = H.device_mus(N, gf, mx, jpiglo, jpjglo, JPK, jperio, 1, 1, kjpt, schedule=schedule)
    assert np.array_equal(loc[0], ref), H.max_rel_diff(loc[0], ref)
    assert not np.array_equal(ref, gf["pta"])


@pytest.mark.parametrize("ln_linssh,ln_isfcav,ld_msc_ups", [(True, False, False), (True, True, False), (False, False, True),
                                                            (True, True, True)])
@pytest.mark.parametrize("schedule", [0, 1, 2])
def test_mus_parity_options(ln_linssh, ln_isfcav, ld_msc_ups, schedule):
    jpiglo, jpjglo, jperio, kjpt = 40, 36, 4, 2
    gf, mx = _mus_case(jpiglo, jpjglo, jperio, kjpt, seed=7, ln_linssh=ln_linssh, ln_isfcav=ln_isfcav)
    ref, _ = H.oracle_mus(O, gf, mx, jpiglo, jpjglo, JPK, jperio, 1, 1, kjpt, ln_linssh=ln_linssh, ln_isfcav=ln_isfcav,
                          ld_msc_ups=ld_msc_ups)
    got, loc = H.device_mus(N, gf, mx, jpiglo, jpjglo, JPK, jperio, 1, 1, kjpt, ln_linssh=ln_linssh, ln_isfcav=ln_isfcav,
                            ld_msc_ups=ld_msc_ups, schedule=schedule)
    assert np.array_equal(loc[0], ref), H.max_rel_diff(loc[0], ref)


@pytest.mark.parametrize("jperio,jpni,jpnj", [(0, 2, 2), (1, 3, 1), (4, 2, 2), (6, 2, 2), (4, 1, 2)])
@pytest.mark.parametrize("schedule", [0, 1, 2])
def test_mus_decomposed_in_process(jperio, jpni, jpnj, schedule):
    """jpni x jpnj subdomains on one GPU: every rank's local array equals the oracle's rank (halo exchanges through the
    compiled plans), and the assembled interior equals the mono-domain oracle."""
    jpiglo, jpjglo, kjpt = 60, 52, 2
    gf, mx = _mus_case(jpiglo, jpjglo, jperio, kjpt, seed=31)
    ref1, _ = H.oracle_mus(O, gf, mx, jpiglo, jpjglo, JPK, jperio, 1, 1, kjpt)
    refg, refl = H.oracle_mus(O, gf, mx, jpiglo, jpjglo, JPK, jperio, jpni, jpnj, kjpt)
    got, loc = H.device_mus(N, gf, mx, jpiglo, jpjglo, JPK, jperio, jpni, jpnj, kjpt, schedule=schedule)
    for a, b in zip(loc, refl):
        assert np.array_equal(a, b)
    inner = (slice(None), slice(0, JPK - 1), slice(1, -1), slice(1, -1))
    assert np.array_equal(got[inner], ref1[inner])


def test_mus_small_domain_and_host_path():
    """below 20 x 20 the fused schedule falls back to the reference structure; the host-pointer entry stages and returns"""
    jpiglo, jpjglo, jperio, kjpt = 14, 12, 1, 2
    gf, mx = _mus_case(jpiglo, jpjglo, jperio, kjpt, seed=3)
    ref, _ = H.oracle_mus(O, gf, mx, jpiglo, jpjglo, JPK, jperio, 1, 1, kjpt)
    for host in (False, True):
        got, loc = H.device_mus(N, gf, mx, jpiglo, jpjglo, JPK, jperio, 1, 1, kjpt, schedule=2, host_path=host)
        assert np.array_equal(loc[0], ref)


def test_mus_errors():
    dom = N.mpp_init(30, 30, JPK, 0, 1, 1, 1)
    ctx = N.FctContext(dom, 0)
    z3 = torch.zeros(dom.shape3, dtype=torch.float64, device="cuda")
    z4 = torch.zeros((1,) + dom.shape3, dtype=torch.float64, device="cuda")
    with pytest.raises(N.NemoFctError, match="set_domain_arrays"):
        ctx.tra_adv_mus(1, 1, "TRA", 100.0, z3, z3, z3, z4, z4, 1)
    gf = H.random_fields(O, 30, 30, JPK, 0, 1, seed=1)
    ctx.set_domain_arrays(*[gf[k] for k in ("tmask", "umask", "vmask", "wmask", "e1e2t", "r1_e1e2t", "mikt", "mbkt")])
    ctx.set_e3t(gf["e3t_b"], gf["e3t_n"], gf["e3t_a"])
    with pytest.raises(N.NemoFctError, match="set_mus_metrics"):
        ctx.tra_adv_mus(1, 1, "TRA", 100.0, z3, z3, z3, z4, z4, 1)
    ctx.set_mus_metrics(gf["r1_e1e2t"], gf["r1_e1e2t"])
    with pytest.raises(N.NemoFctError, match="set_e3uvw"):
        ctx.tra_adv_mus(1, 1, "TRA", 100.0, z3, z3, z3, z4, z4, 1)
    ctx.close()


# ---- tra_nxt ---------------------------------------------------------------------------------------------------------
def _nxt_forcing_fields(rng, w1, jpj, jpi, kjpt, jpk, mikt, mbkt):
    f2 = {k: rng.standard_normal((1, jpj, jpi)) * 1e-4 for k in ("emp_b", "emp", "fwfisf_b", "fwfisf", "rnf_b", "rnf")}
    f2["h_rnf"] = 10.0 + 50.0 * rng.random((1, jpj, jpi))
    f2["r1_hisf_tbl"] = 1.0 / (20.0 + 10.0 * rng.random((1, jpj, jpi)))
    f2["ralpha"] = rng.random((1, jpj, jpi))
    w1.lbc_lnk([[a] for a in f2.values()], "T" * len(f2), [1.0] * len(f2))
    f2["h_rnf"][f2["h_rnf"] == 0.0] = 10.0
    out = {k: np.ascontiguousarray(v[0]) for k, v in f2.items()}
    for k in ("sbc", "sbc_b", "rnf_tsc", "rnf_tsc_b", "risf_tsc", "risf_tsc_b"):
        a = rng.standard_normal((kjpt, jpj, jpi)) * 1e-5
        w1.lbc_lnk([[a]], "T", [1.0])
        out[k] = a
    for k in ("qsr_hc", "qsr_hc_b"):
        a = rng.standard_normal((jpk, jpj, jpi)) * 1e-5
        w1.lbc_lnk([[a]], "T", [1.0])
        out[k] = a
    out["nk_rnf"] = np.minimum(mikt + 2, np.maximum(mbkt, 1)).astype(np.int32)
    out["misfkt"] = mikt.astype(np.int32).copy()
    out["misfkb"] = np.minimum(mikt + 1, np.maximum(mbkt, 1)).astype(np.int32)
    return out


NXT_2D = ("emp_b", "emp", "fwfisf_b", "fwfisf", "rnf_b", "rnf", "h_rnf", "r1_hisf_tbl", "ralpha", "nk_rnf", "misfkt", "misfkb")
NXT_ND = ("qsr_hc", "qsr_hc_b", "rnf_tsc", "rnf_tsc_b", "risf_tsc", "risf_tsc_b")


def _run_nxt_pair(jpiglo, jpjglo, jperio, jpni, jpnj, kjpt, cdtype, l_euler, ln_linssh, flags, seed=11, ln_isfcav=False):
    rng = np.random.default_rng(seed)
    gf = H.random_fields(O, jpiglo, jpjglo, JPK, jperio, kjpt, seed=seed, ln_linssh=ln_linssh, ln_isfcav=ln_isfcav)
    w1 = O.World(jpiglo, jpjglo, JPK, jperio, 1, 1)
    ff = _nxt_forcing_fields(rng, w1, jpjglo, jpiglo, kjpt, JPK, gf["mikt"], gf["mbkt"])
    w1.close()
    atfp, rdt, r1_rau0 = 0.1, 900.0, 1.0 / 1026.0
    scal = dict(atfp=atfp, r1_rau0=r1_rau0, nksr=4, **flags)
    # ---- oracle ----
    w = O.World(jpiglo, jpjglo, JPK, jperio, jpni, jpnj)
    loc = {k: w.scatter(gf[k]) for k in H.DOM_KEYS + ("ptb", "ptn", "pta")}
    lf = {k: w.scatter(ff[k]) for k in ff}
    for r, d in enumerate(w.doms):
        d.set_fields(*[loc[k][r] for k in H.DOM_KEYS], ln_linssh=ln_linssh, ln_isfcav=ln_isfcav)
    forc = [O.NxtForcing(**scal, **{k: lf[k][r] for k in NXT_2D + NXT_ND}) for r in range(w.jpnij)]
    dev_in = {k: [torch.from_numpy(a.copy()).cuda() for a in loc[k]] for k in ("ptb", "ptn", "pta")}
    w.tra_nxt(7, 1, l_euler, rdt, cdtype, forc, loc["ptb"], loc["ptn"], loc["pta"], kjpt, lf["sbc"], lf["sbc_b"])
    # ---- device ----
    n = jpni * jpnj
    if n == 1:
        ctxs = [N.FctContext(N.mpp_init(jpiglo, jpjglo, JPK, jperio, 1, 1, 1), 0)]
    else:
        grp = N.LocalGroup(jpiglo, jpjglo, JPK, jperio, jpni, jpnj, 0)
        ctxs = grp.ctx
    dforc, dsbc, dsbc_b = [], [], []
    for r, c in enumerate(ctxs):
        c.set_domain_arrays(loc["tmask"][r], loc["umask"][r], loc["vmask"][r], loc["wmask"][r], loc["e1e2t"][r],
                            loc["r1_e1e2t"][r], loc["mikt"][r], loc["mbkt"][r], ln_linssh, ln_isfcav)
        c.set_e3t(loc["e3t_b"][r], loc["e3t_n"][r], loc["e3t_a"][r])
        dforc.append(N.NxtForcing(**scal, **{k: torch.from_numpy(lf[k][r]).cuda() for k in NXT_2D + NXT_ND}))
        dsbc.append(torch.from_numpy(lf["sbc"][r]).cuda()); dsbc_b.append(torch.from_numpy(lf["sbc_b"][r]).cuda())
    if n == 1:
        ctxs[0].tra_nxt(7, 1, l_euler, rdt, cdtype, dforc[0], dev_in["ptb"][0], dev_in["ptn"][0], dev_in["pta"][0], kjpt,
                        dsbc[0], dsbc_b[0])
        ctxs[0].synchronize()
    else:
        grp.tra_nxt(7, 1, l_euler, rdt, cdtype, dforc, dev_in["ptb"], dev_in["ptn"], dev_in["pta"], kjpt, dsbc, dsbc_b)
        grp.synchronize()
    for k in ("ptb", "ptn", "pta"):
        for r in range(n):
            assert np.array_equal(dev_in[k][r].cpu().numpy(), loc[k][r]), (k, r)
    changed = not np.array_equal(loc["ptn"][0], w.scatter(gf["ptn"])[0])
    w.close()
    for c in ctxs:
        c.close()
    return changed


@pytest.mark.parametrize("cdtype", ["TRA", "TRC"])
@pytest.mark.parametrize("mode", ["euler", "fix", "vvl", "vvl_all", "vvl_rnf_depth"])
def test_nxt_parity(cdtype, mode):
    flags = {}
    if mode == "vvl_all":
        flags = dict(ln_traqsr=1, ln_rnf=1, ln_isf=1)
    if mode == "vvl_rnf_depth":
        flags = dict(ln_rnf_depth=1, ln_rnf=1)
    assert _run_nxt_pair(46, 34, 4, 1, 1, 3, cdtype, mode == "euler", mode == "fix", flags, ln_isfcav=mode == "vvl_all")


@pytest.mark.parametrize("jperio,jpni,jpnj", [(0, 2, 2), (1, 2, 1), (6, 2, 2), (4, 3, 2)])
def test_nxt_decomposed(jperio, jpni, jpnj):
    assert _run_nxt_pair(48, 40, jperio, jpni, jpnj, 2, "TRA", False, False, dict(ln_traqsr=1))
    assert _run_nxt_pair(48, 40, jperio, jpni, jpnj, 2, "TRC", True, False, {})


def test_nxt_errors():
    dom = N.mpp_init(30, 30, JPK, 0, 1, 1, 1)
    ctx = N.FctContext(dom, 0)
    gf = H.random_fields(O, 30, 30, JPK, 0, 1, seed=1)
    ctx.set_domain_arrays(*[gf[k] for k in ("tmask", "umask", "vmask", "wmask", "e1e2t", "r1_e1e2t", "mikt", "mbkt")])
    ctx.set_e3t(gf["e3t_b"], gf["e3t_n"], gf["e3t_a"])
    z4 = torch.zeros((1,) + dom.shape3, dtype=torch.float64, device="cuda")
    with pytest.raises(N.NemoFctError, match="cdtype"):
        ctx.tra_nxt(1, 1, False, 900.0, "XXX", N.NxtForcing(), z4, z4.clone(), z4.clone(), 1)
    with pytest.raises(N.NemoFctError, match="forcing"):
        ctx.tra_nxt(2, 1, False, 900.0, "TRA", None, z4, z4.clone(), z4.clone(), 1)
    with pytest.raises(N.NemoFctError, match="qsr_hc"):
        ctx.tra_nxt(2, 1, False, 900.0, "TRA", N.NxtForcing(ln_traqsr=1), z4, z4.clone(), z4.clone(), 1)
    ctx.close()
