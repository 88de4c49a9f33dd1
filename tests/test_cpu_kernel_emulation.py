"""The MUSCL and centred-scheme column kernels, compiled for the HOST from the very same source (tests/emu), against the
oracle -- bit for bit.  Covers what the GPU suite does not reach at its small jpk: the jk loop split in chunks across
blockIdx.y (chunk-start recomputation of the vertical slopes / fluxes), which production sizes (jpk = 75) use."""
import ctypes as C

import numpy as np
import pytest

from oracle import oracle as O
import helpers as H

import emu_api
from emu_api import p as _p, rect as _rect, mus as _mus

JPK = 19


@pytest.fixture(scope="module")
def emu():
    return emu_api.load()


def _lbc_uv(jpiglo, jpjglo, jperio, a, b):
    w = O.World(jpiglo, jpjglo, JPK, jperio, 1, 1)
    w.lbc_lnk([[a.reshape(-1, a.shape[-2], a.shape[-1])], [b.reshape(-1, b.shape[-2], b.shape[-1])]], "UV", [-1.0, -1.0])
    w.close()


@pytest.mark.parametrize("nk", [1, 2, 3, 5])
@pytest.mark.parametrize("jperio,lin,isf,ups", [(0, False, False, False), (4, True, True, True), (6, True, False, False)])
def test_mus_kernels_with_jk_chunks(emu, nk, jperio, lin, isf, ups):
    G, GJ, kjpt = 30, 26, 2
    gf = H.random_fields(O, G, GJ, JPK, jperio, kjpt, seed=60 + jperio, ln_linssh=lin, ln_isfcav=isf)
    mx = H.mus_extra_fields(O, gf, G, GJ, JPK, jperio, seed=60 + jperio, runoff=True)
    ref, _ = H.oracle_mus(O, gf, mx, G, GJ, JPK, jperio, 1, 1, kjpt, ln_linssh=lin, ln_isfcav=isf, ld_msc_ups=ups)
    xind = None
    if ups:
        w = O.World(G, GJ, JPK, jperio, 1, 1)
        w.doms[0].set_fields(*[gf[k] for k in H.DOM_KEYS])
        xind = w.doms[0].mus_xind(True, mx["rnfmsk"], mx["rnfmsk_z"])
        w.close()
    jpi, jpj = G, GJ
    f = 1 if jperio in (3, 4, 5, 6) else 0
    shp = gf["ptb"].shape
    # ---- reference structure: grad -> lbc -> hflux -> lbc -> trend ----
    pta = gf["pta"].copy()
    zwx, zwy, fx, fy = (np.zeros(shp) for _ in range(4))
    _mus(emu, 0, (1, jpi - 1, 1, jpj - 1), nk, gf, mx, xind, pta, zwx, zwy, fx, fy, lin, isf, kjpt)
    _lbc_uv(G, GJ, jperio, zwx, zwy)
    _mus(emu, 1, (2, jpi - 1, 2, jpj - 1), nk, gf, mx, xind, pta, zwx, zwy, fx, fy, lin, isf, kjpt)
    _lbc_uv(G, GJ, jperio, fx, fy)
    _mus(emu, 3, (2, jpi - 1, 2, jpj - 1), nk, gf, mx, xind, pta, zwx, zwy, fx, fy, lin, isf, kjpt)
    assert np.array_equal(pta, ref)
    # ---- default structure: differences in place from ptb on the inner columns, exchanged ones on the frame ----
    pta2 = gf["pta"].copy()
    fx2, fy2 = np.zeros(shp), np.zeros(shp)
    _mus(emu, 2, (3, jpi - 2, 3, jpj - 2 - f), nk, gf, mx, xind, pta2, zwx, zwy, fx2, fy2, lin, isf, kjpt)
    for rect in ((2, 2, 2, jpj - 1), (jpi - 1, jpi - 1, 2, jpj - 1), (3, jpi - 2, 2, 2), (3, jpi - 2, jpj - 1 - f, jpj - 1)):
        _mus(emu, 1, rect, nk, gf, mx, xind, pta2, zwx, zwy, fx2, fy2, lin, isf, kjpt)
    _lbc_uv(G, GJ, jperio, fx2, fy2)
    assert np.array_equal(fx2, fx) and np.array_equal(fy2, fy)
    # ---- fully fused inner kernel on the columns that need no exchanged value ----
    pta3 = gf["pta"].copy()
    _mus(emu, 4, (4, jpi - 2, 4, jpj - 2 - f), nk, gf, mx, xind, pta3, zwx, zwy, fx, fy, lin, isf, kjpt)
    sl = (slice(None), slice(None), slice(3, jpj - 2 - f), slice(3, jpi - 2))
    assert np.array_equal(pta3[sl], ref[sl])


@pytest.mark.parametrize("nk", [1, 3, 4])
@pytest.mark.parametrize("h,v", [(2, 2), (2, 4), (4, 2), (4, 4)])
@pytest.mark.parametrize("jperio,lin,isf", [(1, False, False), (4, True, True)])
def test_cen_kernel_with_jk_chunks(emu, nk, h, v, jperio, lin, isf):
    G, GJ, kjpt = 28, 24, 2
    gf = H.random_fields(O, G, GJ, JPK, jperio, kjpt, seed=80 + jperio, ln_linssh=lin, ln_isfcav=isf)
    ref, _ = H.oracle_cen(O, gf, G, GJ, JPK, jperio, 1, 1, kjpt, h, v, ln_linssh=lin, ln_isfcav=isf)
    jpi, jpj = G, GJ
    shp = gf["ptn"].shape
    ztu, ztv, ztw = np.zeros(shp), np.zeros(shp), np.zeros(shp)
    if h == 4:      # masked gradients by the MUSCL first-guess kernel on (2:jpim1, 2:jpjm1), then lbc_lnk (traadv_cen.F90:119-126)
        dummy = {k: gf["r1_e1e2t"] for k in ("r1_e1e2u", "r1_e1e2v")}
        dummy.update({k: gf["e3t_n"] for k in ("e3u_n", "e3v_n", "e3w_n")})
        _mus(emu, 0, (2, jpi - 1, 2, jpj - 1), nk, gf, dummy, None, gf["pta"].copy(), ztu, ztv, ztu, ztv, lin, isf, kjpt, ptb=gf["ptn"])
        _lbc_uv(G, GJ, jperio, ztu, ztv)
    if v == 4:
        w = O.World(G, GJ, JPK, jperio, 1, 1)
        d = w.doms[0]
        d.set_fields(*[gf[k] for k in H.DOM_KEYS], ln_linssh=lin, ln_isfcav=isf)
        for jn in range(kjpt):
            O.lib().interp_4th_cpt(d.h, gf["ptn"][jn].ctypes.data_as(C.c_void_p), ztw[jn].ctypes.data_as(C.c_void_p))
        w.close()
    pta = gf["pta"].copy()
    rc = emu.emu_cen(h, v, jpi, jpj, JPK, kjpt, _rect(2, jpi - 1, 2, jpj - 1), nk, int(lin), int(isf), _p(gf["wmask"]), _p(gf["e3t_n"]),
                     _p(gf["r1_e1e2t"]), _p(gf["mikt"]), _p(gf["pun"]), _p(gf["pvn"]), _p(gf["pwn"]), _p(gf["ptn"]), _p(pta),
                     _p(ztu), _p(ztv), _p(ztw))
    assert rc == 0
    assert np.array_equal(pta, ref)


@pytest.mark.parametrize("mode,flags", [("fix", {}), ("vvl", {}), ("vvl", dict(ln_traqsr=1, ln_rnf=1, ln_isf=1)),
                                        ("vvl", dict(ln_rnf_depth=1, ln_rnf=1)), ("euler_tra", {}), ("euler_trc", {})])
def test_nxt_kernels(emu, mode, flags):
    """k_nxt_fix / k_nxt_vvl (every forcing switch) / k_nxt_euler against the oracle's tra_nxt on a closed mono-domain, where
    the lbc_lnk of tra_nxt only zeroes the halo: compared on the interior"""
    G, GJ, kjpt, jperio = 24, 20, 3, 0
    isf = bool(flags.get("ln_isf"))
    lin = mode == "fix"
    rng = np.random.default_rng(77)
    gf = H.random_fields(O, G, GJ, JPK, jperio, kjpt, seed=91, ln_linssh=lin, ln_isfcav=isf)
    f2 = {k: np.ascontiguousarray(rng.standard_normal((GJ, G)) * 1e-4) for k in ("emp_b", "emp", "fwfisf_b", "fwfisf", "rnf_b", "rnf")}
    f2["h_rnf"] = 10.0 + 50.0 * rng.random((GJ, G)); f2["r1_hisf_tbl"] = 1.0 / (20.0 + 10.0 * rng.random((GJ, G))); f2["ralpha"] = rng.random((GJ, G))
    fn = {k: rng.standard_normal((kjpt, GJ, G)) * 1e-5 for k in ("sbc", "sbc_b", "rnf_tsc", "rnf_tsc_b", "risf_tsc", "risf_tsc_b")}
    q3 = {k: rng.standard_normal((JPK, GJ, G)) * 1e-5 for k in ("qsr_hc", "qsr_hc_b")}
    mikt, mbkt = gf["mikt"], gf["mbkt"]
    ints = {"nk_rnf": np.minimum(mikt + 2, np.maximum(mbkt, 1)).astype(np.int32), "misfkt": mikt.astype(np.int32).copy(),
            "misfkb": np.minimum(mikt + 1, np.maximum(mbkt, 1)).astype(np.int32)}
    atfp, rdt, r1_rau0, nksr = 0.1, 900.0, 1.0 / 1026.0, 5
    # ---- oracle ----
    w = O.World(G, GJ, JPK, jperio, 1, 1)
    w.doms[0].set_fields(*[gf[k] for k in H.DOM_KEYS], ln_linssh=lin, ln_isfcav=isf)
    tb, tn, ta = gf["ptb"].copy(), gf["ptn"].copy(), gf["pta"].copy()
    forc = O.NxtForcing(atfp=atfp, r1_rau0=r1_rau0, nksr=nksr, **flags, **f2, **q3, **ints,
                        rnf_tsc=fn["rnf_tsc"], rnf_tsc_b=fn["rnf_tsc_b"], risf_tsc=fn["risf_tsc"], risf_tsc_b=fn["risf_tsc_b"])
    cdtype = "TRC" if mode == "euler_trc" else "TRA"
    w.tra_nxt(4, 1, mode.startswith("euler"), rdt, cdtype, [forc], [tb], [tn], [ta], kjpt, [fn["sbc"]], [fn["sbc_b"]])
    w.close()
    # ---- emulated kernels ----
    eb, en, ea = gf["ptb"].copy(), gf["ptn"].copy(), gf["pta"].copy()
    iflags = (C.c_int * 5)(int(flags.get("ln_traqsr", 0)), int(flags.get("ln_rnf", 0)), int(flags.get("ln_isf", 0)),
                           int(flags.get("ln_rnf_depth", 0)), nksr)
    tab = (C.c_void_p * 6)(*[f2[k].ctypes.data for k in ("emp_b", "emp", "fwfisf_b", "fwfisf", "rnf_b", "rnf")])
    m = {"fix": 0, "vvl": 1, "euler_tra": 2, "euler_trc": 2}[mode]
    rc = emu.emu_nxt(m, int(mode == "euler_trc"), G, GJ, JPK, kjpt, C.c_double(atfp), C.c_double(atfp * rdt), C.c_double(atfp * rdt * r1_rau0),
                     iflags, _p(gf["e3t_b"]), _p(gf["e3t_n"]), _p(gf["e3t_a"]), _p(mikt), _p(eb), _p(en), _p(ea), _p(fn["sbc"]), _p(fn["sbc_b"]),
                     tab, _p(q3["qsr_hc"]), _p(q3["qsr_hc_b"]), _p(ints["nk_rnf"]), _p(f2["h_rnf"]), _p(fn["rnf_tsc"]), _p(fn["rnf_tsc_b"]),
                     _p(ints["misfkt"]), _p(ints["misfkb"]), _p(fn["risf_tsc"]), _p(fn["risf_tsc_b"]), _p(f2["r1_hisf_tbl"]), _p(f2["ralpha"]))
    assert rc == 0
    inner = (slice(None), slice(0, JPK - 1), slice(1, -1), slice(1, -1))
    for got, ref, name in ((eb, tb, "ptb"), (en, tn, "ptn"), (ea, ta, "pta")):
        assert np.array_equal(got[inner], ref[inner]), name
    assert not np.array_equal(en[inner], gf["ptn"][inner])


def _nonosc_case(jperio, h, seed):
    import np_fct
    G, GJ, K, kjpt = 46, 38, 9, 2
    gf = H.random_fields(O, G, GJ, K, jperio, kjpt, seed=seed)
    cap = []
    final = np_fct.tra_adv_fct(gf, kjpt, h, jperio, capture=cap)
    ref, _, _ = H.oracle_fct(O, gf, G, GJ, K, jperio, 1, 1, kjpt, h, 2)
    assert np.array_equal(final, ref)                       # the numpy statement is the oracle, bit for bit
    return G, GJ, K, kjpt, gf, cap, final


@pytest.mark.parametrize("jperio,h", [(0, 2), (1, 4)])
def test_fused_limiter_kernel_emulated(emu, jperio, h):
    """k_fct_nonosc_final (512 cooperating threads per block, shared memory, one barrier per level) compiled for the host
    from nonosc_final.cuh and run one host thread per CUDA thread: from the fields after X2 it must reproduce the final pta
    of the oracle on its output rectangle."""
    G, GJ, K, kjpt, gf, cap, final = _nonosc_case(jperio, h, 400 + jperio)
    stack = lambda k: np.ascontiguousarray(np.stack([c[k] for c in cap]))
    zwi, zwx, zwy, zwz, pta = (stack(k) for k in ("zwi", "zwx", "zwy", "zwz", "pta"))
    out = (5, G - 4, 4, GJ - 4)
    rc = emu.emu_nonosc_final(G, GJ, K, kjpt, _rect(*out), C.c_double(gf["p2dt"]), _p(gf["tmask"]), _p(gf["e3t_n"]),
                              _p(gf["e1e2t"]), _p(gf["r1_e1e2t"]), _p(gf["ptb"]), _p(zwi), _p(zwx), _p(zwy), _p(zwz), _p(pta))
    assert rc == 0
    sl = (slice(None), slice(0, K - 1), slice(out[2] - 1, out[3]), slice(out[0] - 1, out[1]))
    assert np.array_equal(pta[sl], final[sl])
    assert not np.array_equal(pta[sl], stack("pta")[sl])
    # the experimental variant of experiments/nonosc_final_v3.cuh (loads consumed behind the barrier): same bits
    pta3 = stack("pta")
    emu.emu_nonosc_final_variant.restype = C.c_int
    rc = emu.emu_nonosc_final_variant(3, G, GJ, K, kjpt, _rect(*out), C.c_double(gf["p2dt"]), _p(gf["tmask"]), _p(gf["e3t_n"]),
                                      _p(gf["e1e2t"]), _p(gf["r1_e1e2t"]), _p(gf["ptb"]), _p(zwi), _p(zwx), _p(zwy), _p(zwz), _p(pta3))
    assert rc == 0 and np.array_equal(pta3, pta)


@pytest.mark.parametrize("nk", [1, 3])
@pytest.mark.parametrize("h,v", [(2, 2), (4, 4), (4, 2), (2, 4)])
@pytest.mark.parametrize("jperio,lin,isf", [(0, False, False), (4, True, True), (6, False, False)])
def test_fct_column_kernels_reference_structure(emu, nk, h, v, jperio, lin, isf):
    """the reference-structured FCT kernels (fct_column_kernels.cuh: laplacian, P1-P5, betas, limit, final, interp_4th_cpt with
    its pivot table and simple-column classification, trend hook) in the order run_fct launches them, with the oracle's
    lbc_lnk for X1..X4, jk loop chunked or not: whole-array equality with the oracle's tra_adv_fct and its hooks"""
    G, GJ, kjpt = 30, 26, 2
    gf = H.random_fields(O, G, GJ, JPK, jperio, kjpt, seed=700 + jperio, ln_linssh=lin, ln_isfcav=isf)
    w = O.World(G, GJ, JPK, jperio, 1, 1)
    d = w.doms[0]
    d.set_fields(*[gf[k] for k in H.DOM_KEYS], ln_linssh=lin, ln_isfcav=isf)
    otrd = [np.full((kjpt,) + d.shape3, np.nan) for _ in range(3)]
    d.set_diag(*otrd)
    ref = gf["pta"].copy()
    w.tra_adv_fct(gf["p2dt"], [gf["pun"]], [gf["pvn"]], [gf["pwn"]], [gf["ptb"]], [gf["ptn"]], [ref], kjpt, h, v)

    def lbc(trip):
        w.lbc_lnk([[a.reshape(-1, GJ, G)] for a, _, _ in trip], "".join(n for _, n, _ in trip), [s for _, _, s in trip])

    pta, work = emu_api.fct_step(emu, gf, kjpt, h, v, lin, isf, nk, lbc, hooks=True)
    w.close()
    assert np.array_equal(pta, ref)
    assert np.array_equal(work["trdx"][:, :, :-1, :-1], otrd[0][:, :, :-1, :-1])
    assert np.array_equal(work["trdy"][:, :, :-1, :-1], otrd[1][:, :, :-1, :-1])
    assert np.array_equal(work["trdz"], otrd[2])


def test_interp_4th_cpt_simple_columns_and_general_path(emu):
    """the 'simple column' shortcut (pivots and masks from mbkt and a jpk-entry table) must be taken on the full-depth
    cavity-free columns only and give the same bits as the general path and as the oracle"""
    G, GJ, jperio = 26, 22, 1
    for isf in (False, True):
        gf = H.random_fields(O, G, GJ, JPK, jperio, 2, seed=800, ln_isfcav=isf)
        w = O.World(G, GJ, JPK, jperio, 1, 1)
        d = w.doms[0]
        d.set_fields(*[gf[k] for k in H.DOM_KEYS], ln_isfcav=isf)
        ref = np.zeros_like(gf["ptn"])
        for jn in range(2):
            O.lib().interp_4th_cpt(d.h, gf["ptn"][jn].ctypes.data_as(C.c_void_p), ref[jn].ctypes.data_as(C.c_void_p))
        w.close()
        fast, slow = np.zeros_like(ref), np.zeros_like(ref)
        nsimple = emu_api.interp_4th_cpt(emu, gf, gf["ptn"], fast, isf, use_simple=True)
        assert emu_api.interp_4th_cpt(emu, gf, gf["ptn"], slow, isf, use_simple=False) == 0
        ncol = (G - 2) * (GJ - 2)                                    # columns wet from level 1 down to mbkt qualify, cavities do not
        assert (0 < nsimple < ncol) if isf else (0 < nsimple <= ncol), (nsimple, ncol)
        inner = (slice(None), slice(1, JPK - 1), slice(1, -1), slice(1, -1))
        assert np.array_equal(fast[inner], ref[inner]) and np.array_equal(slow[inner], ref[inner])


@pytest.mark.parametrize("G,GJ", [(20, 20), (21, 20), (20, 23), (22, 20), (33, 21), (40, 38)])
@pytest.mark.parametrize("jperio", [0, 1, 4, 6])
def test_fct_fused_schedule_on_the_host(emu, G, GJ, jperio):
    """The WHOLE fused schedule of run_fct (schedules 1/2) on the CPU: the column sets of csrc/schedule.hpp, the fused P1-P5
    kernel (cp.async ring as plain copies), the fused limiter kernel (512 host threads per block), the reference-structured
    kernels on the frame bands and X1..X4 -- from the minimum subdomain size (20 x 20) up, closed / cyclic / T- and F-pivot
    fold, band split on and off, masks derived from tmask or read: whole-array equality with the oracle."""
    kjpt, K = 2, 7
    fold = jperio in (3, 4, 5, 6)
    gf = H.random_fields(O, G, GJ, K, jperio, kjpt, seed=900 + G + GJ + jperio)
    w = O.World(G, GJ, K, jperio, 1, 1)

    def lbc(trip):
        w.lbc_lnk([[a.reshape(-1, GJ, G)] for a, _, _ in trip], "".join(n for _, n, _ in trip), [s for _, _, s in trip])

    for (h, v, from_t, split) in ((4, 4, True, False), (2, 2, False, True), (4, 2, True, True)):
        ref, _, _ = H.oracle_fct(O, gf, G, GJ, K, jperio, 1, 1, kjpt, h, v)
        pta, plan = emu_api.fct_step_fused(emu, gf, kjpt, h, v, False, False, 1, lbc, fold, from_t, want_split=split)
        assert plan["min_size"] == 20
        assert np.array_equal(pta, ref), (h, v, from_t, split, plan["split"])
    w.close()


def test_fused_plans_cover_the_interior_exactly_once(emu):
    """schedule.hpp: inner region + frame = the interior, no column twice, for every size from the minimum up"""
    for fold in (False, True):
        for jpi in range(20, 46):
            for jpj in range(20, 44):
                p = emu_api.fct_fused_plan(emu, jpi, jpj, fold, want_split=True)
                cover = np.zeros((jpj + 1, jpi + 1), int)
                for (i0, i1, j0, j1) in p["k1"] + p["lowf"]:
                    cover[j0:j1 + 1, i0:i1 + 1] += 1
                assert (cover[2:jpj, 2:jpi] == 1).all() and cover.sum() == (jpi - 2) * (jpj - 2), (jpi, jpj, fold)
                if p["split"]:
                    c2 = np.zeros_like(cover)
                    for (i0, i1, j0, j1) in p["k1_band"] + p["k1_centre"]:
                        c2[j0:j1 + 1, i0:i1 + 1] += 1
                    c1 = np.zeros_like(cover)
                    for (i0, i1, j0, j1) in p["k1"]:
                        c1[j0:j1 + 1, i0:i1 + 1] += 1
                    assert (c1 == c2).all(), (jpi, jpj, fold)
                    assert p["k1_centre"][0][0] % 2 == 0                    # even first column: TMA box origin
                # final trend: fused output rectangle + frame `fin` = the interior, exactly once
                c3 = np.zeros_like(cover)
                i0, i1, j0, j1 = p["k2_out"]
                c3[j0:j1 + 1, i0:i1 + 1] += 1
                for (i0, i1, j0, j1) in p["fin"]:
                    c3[j0:j1 + 1, i0:i1 + 1] += 1
                assert (c3[2:jpj, 2:jpi] == 1).all() and c3.sum() == (jpi - 2) * (jpj - 2), (jpi, jpj, fold)
                for which in (0, 1):
                    m = emu_api.mus_plan(emu, which, jpi, jpj, fold)
                    c4 = np.zeros_like(cover)
                    parts = m["inner"] + (m["hflux"] if which == 0 else m["trend"])
                    for (i0, i1, j0, j1) in parts:
                        c4[j0:j1 + 1, i0:i1 + 1] += 1
                    assert (c4[2:jpj, 2:jpi] == 1).all() and c4.sum() == (jpi - 2) * (jpj - 2), (which, jpi, jpj, fold)


@pytest.mark.parametrize("G,GJ", [(20, 20), (21, 22), (23, 20), (40, 38)])
@pytest.mark.parametrize("jperio", [0, 1, 4, 6])
def test_mus_schedules_on_the_host(emu, G, GJ, jperio):
    """run_mus on the CPU with the column sets of schedule.hpp, from the minimum size up: the default structure (differences
    in place from ptb on `inner`, exchanged ones on the frame) and the fully fused one, both equal to the oracle"""
    kjpt, K = 2, 7
    fold = jperio in (3, 4, 5, 6)
    gf = H.random_fields(O, G, GJ, K, jperio, kjpt, seed=950 + G + GJ + jperio)
    mx = H.mus_extra_fields(O, gf, G, GJ, K, jperio, seed=950 + jperio)
    ref, _ = H.oracle_mus(O, gf, mx, G, GJ, K, jperio, 1, 1, kjpt)
    w = O.World(G, GJ, K, jperio, 1, 1)

    def lbc(a, b):
        w.lbc_lnk([[a.reshape(-1, GJ, G)], [b.reshape(-1, GJ, G)]], "UV", [-1.0, -1.0])

    shp = gf["ptb"].shape
    for which in (0, 1):
        plan = emu_api.mus_plan(emu, which, G, GJ, fold)
        pta = gf["pta"].copy()
        zwx, zwy, fx, fy = (np.zeros(shp) for _ in range(4))

        def on(region, kern):
            for rc in plan[region]:
                _mus(emu, kern, rc, 1, gf, mx, None, pta, zwx, zwy, fx, fy, False, False, kjpt)

        on("inner", 2 if which == 0 else 4)          # fluxes from ptb / the whole trend
        on("grad", 0)
        lbc(zwx, zwy)
        on("hflux", 1)
        lbc(fx, fy)
        on("trend", 3)
        assert np.array_equal(pta, ref), which
    w.close()


@pytest.mark.parametrize("G,GJ,jperio", [(20, 20, 0), (22, 21, 4), (34, 20, 1), (40, 29, 6), (52, 38, 4), (21, 24, 0), (66, 27, 1)])
def test_fct_fused_schedule_with_tma_tiles_on_the_host(emu, G, GJ, jperio):
    """schedule 2: the TMA-tiled P1-P5 kernel (256 host threads per block, mbarrier phases as atomic counters, bulk tensor
    copies as box copies with zero fill on the high side) in the whole fused step.  Tile origins, box coordinates, shared-memory
    indexing and ring reuse are exercised for one-tile and multi-tile rectangles with overhang, split and unsplit; the
    emulation also CHECKS the hardware rules the kernel was designed around (box origins >= 0, even innermost coordinate);
    odd jpi falls back to the cp.async kernel exactly like the launcher."""
    kjpt, K = 2, 6
    fold = jperio in (3, 4, 5, 6)
    gf = H.random_fields(O, G, GJ, K, jperio, kjpt, seed=970 + G + GJ)
    w = O.World(G, GJ, K, jperio, 1, 1)

    def lbc(trip):
        w.lbc_lnk([[a.reshape(-1, GJ, G)] for a, _, _ in trip], "".join(n for _, n, _ in trip), [s for _, _, s in trip])

    for (h, v, from_t, split) in ((4, 4, True, False), (2, 2, False, False), (4, 4, True, True)):
        ref, _, _ = H.oracle_fct(O, gf, G, GJ, K, jperio, 1, 1, kjpt, h, v)
        pta, plan = emu_api.fct_step_fused(emu, gf, kjpt, h, v, False, False, 1, lbc, fold, from_t, want_split=split, tma=True)
        assert plan["used_tma"] == (G % 2 == 0), (G, plan)
        assert np.array_equal(pta, ref), (h, v, from_t, split, plan["split"])
    w.close()


def test_fct_fused_schedule_tma_with_jk_chunks(emu):
    """the same with the jk loop of every kernel split in 3 chunks (grid z of the TMA kernel, chunk-start state from global
    memory), linear free surface with ice-shelf cavities"""
    G, GJ, K, kjpt, jperio = 44, 30, 13, 2, 4
    gf = H.random_fields(O, G, GJ, K, jperio, kjpt, seed=990, ln_linssh=True, ln_isfcav=True)
    w = O.World(G, GJ, K, jperio, 1, 1)

    def lbc(trip):
        w.lbc_lnk([[a.reshape(-1, GJ, G)] for a, _, _ in trip], "".join(n for _, n, _ in trip), [s for _, _, s in trip])

    ref, _, _ = H.oracle_fct(O, gf, G, GJ, K, jperio, 1, 1, kjpt, 4, 4, ln_linssh=True, ln_isfcav=True)
    pta, plan = emu_api.fct_step_fused(emu, gf, kjpt, 4, 4, True, True, 3, lbc, True, True, want_split=False, tma=True)
    w.close()
    assert plan["used_tma"]
    assert np.array_equal(pta, ref)


@pytest.mark.parametrize("K", [3, 4, 5])
def test_fused_and_reference_schedules_at_tiny_jpk(emu, K):
    """jpk = 3 is the smallest the ABI accepts: one level pair in the rings and pipelines"""
    G, GJ, kjpt = 26, 22, 2
    for jperio in (0, 4):
        gf = H.random_fields(O, G, GJ, K, jperio, kjpt, seed=5 + K)
        w = O.World(G, GJ, K, jperio, 1, 1)

        def lbc(trip):
            w.lbc_lnk([[a.reshape(-1, GJ, G)] for a, _, _ in trip], "".join(n for _, n, _ in trip), [s for _, _, s in trip])

        for (h, v) in ((2, 2), (4, 4)):
            ref, _, _ = H.oracle_fct(O, gf, G, GJ, K, jperio, 1, 1, kjpt, h, v)
            fused, _ = emu_api.fct_step_fused(emu, gf, kjpt, h, v, False, False, 1, lbc, jperio == 4, True, tma=True)
            plain, _ = emu_api.fct_step(emu, gf, kjpt, h, v, False, False, 1, lbc)
            assert np.array_equal(fused, ref) and np.array_equal(plain, ref), (K, jperio, h, v)
        w.close()


def test_fused_schedule_random_shapes(emu):
    """seeded sweep: random subdomain shapes, boundary types, orders, options, split / TMA choices"""
    rng = np.random.default_rng(2024)
    for it in range(10):
        G, GJ = int(rng.integers(20, 72)), int(rng.integers(20, 52))
        jperio = int(rng.choice([0, 1, 2, 3, 4, 5, 6, 7]))
        K, kjpt = int(rng.integers(3, 9)), int(rng.integers(1, 4))
        h, v = int(rng.choice([2, 4])), int(rng.choice([2, 4]))
        lin = bool(rng.integers(0, 2)); isf = lin and bool(rng.integers(0, 2))
        split, tma, from_t = bool(rng.integers(0, 2)), bool(rng.integers(0, 2)), bool(rng.integers(0, 2))
        gf = H.random_fields(O, G, GJ, K, jperio, kjpt, seed=3000 + it, ln_linssh=lin, ln_isfcav=isf)
        w = O.World(G, GJ, K, jperio, 1, 1)

        def lbc(trip):
            w.lbc_lnk([[a.reshape(-1, GJ, G)] for a, _, _ in trip], "".join(n for _, n, _ in trip), [s for _, _, s in trip])

        ref, _, _ = H.oracle_fct(O, gf, G, GJ, K, jperio, 1, 1, kjpt, h, v, ln_linssh=lin, ln_isfcav=isf)
        pta, plan = emu_api.fct_step_fused(emu, gf, kjpt, h, v, lin, isf, 1, lbc, jperio in (3, 4, 5, 6), from_t, want_split=split, tma=tma)
        w.close()
        assert np.array_equal(pta, ref), (G, GJ, jperio, K, kjpt, h, v, lin, isf, split, tma, from_t)


# ---- schedule 4: the whole step of the inner region in ONE kernel (k_fct_fused) ------------------------------------------------
@pytest.mark.parametrize("G,GJ,K,jperio,kjpt,h,v,lin,isf,nkf", [
    (40, 30, 7, 0, 2, 2, 2, False, False, 1),        # one tile row, closed
    (40, 30, 7, 1, 2, 4, 4, False, False, 1),        # E-W cyclic, 4th order + compact
    (44, 38, 9, 4, 2, 4, 4, False, False, 1),        # T-pivot fold, 2 x 3 tiles with overhang
    (66, 40, 13, 6, 1, 4, 2, True, True, 1),         # F-pivot, linear free surface + cavities, one tracer
    (66, 40, 26, 6, 3, 2, 4, True, False, 2),        # jk loop of the fused kernel split in 2 chunks, three tracers
    (36, 24, 5, 0, 2, 4, 4, False, False, 1),        # smallest size that still splits; jpk = 5: no steady-state iteration
    (52, 34, 3, 1, 2, 2, 2, False, False, 1),        # jpk = 3
])
def test_fct_one_kernel_schedule_on_the_host(emu, G, GJ, K, jperio, kjpt, h, v, lin, isf, nkf):
    """schedule 4 of run_fct on the CPU: k_fct_fused (512 host threads per block: TMA boxes as box copies with zero fill, mbarrier
    phases as counters, three software-pipelined stages per level, one barrier per level) on K2's rectangle straight from the
    inputs + the band of K1 (which leaves pta alone there) + the frame chain with X1..X4: whole-array equality with the oracle.
    The emulation also checks the TMA box rules (origins >= 0, even in ji)."""
    fold = jperio in (3, 4, 5, 6)
    gf = H.random_fields(O, G, GJ, K, jperio, kjpt, seed=1100 + G + K, ln_linssh=lin, ln_isfcav=isf)
    w = O.World(G, GJ, K, jperio, 1, 1)

    def lbc(trip):
        w.lbc_lnk([[a.reshape(-1, GJ, G)] for a, _, _ in trip], "".join(n for _, n, _ in trip), [s for _, _, s in trip])

    ref, _, _ = H.oracle_fct(O, gf, G, GJ, K, jperio, 1, 1, kjpt, h, v, ln_linssh=lin, ln_isfcav=isf)
    pta, plan = emu_api.fct_step_one_kernel(emu, gf, kjpt, h, v, lin, isf, 1, lbc, fold, nk_fused=nkf)
    w.close()
    assert pta is not None and plan["split"]
    bad = np.argwhere(pta != ref)
    assert np.array_equal(pta, ref), (bad[:5].tolist(), len(bad), plan["k2_out"])
    assert not np.array_equal(ref, gf["pta"])


def test_fct_one_kernel_falls_back_like_the_launcher(emu):
    """odd jpi (TMA strides) and subdomains too small to split keep the three-kernel schedule"""
    for (G, GJ) in ((41, 30), (22, 22)):
        gf = H.random_fields(O, G, GJ, 6, 0, 2, seed=1200)
        pta, plan = emu_api.fct_step_one_kernel(emu, gf, 2, 4, 4, False, False, 1, None, False)
        assert pta is None


@pytest.mark.parametrize("G,GJ,K,isf,jperio", [(26, 22, 19, False, 1), (26, 22, 19, True, 1), (70, 31, 75, False, 4), (34, 9, 3, False, 0),
                                               (34, 12, 10, True, 6)])
def test_interp_4th_cpt_tiled_kernel(emu, G, GJ, K, isf, jperio):
    """k_interp_4th_cpt_tiled (TMA boxes of 8 levels in a 4-deep ring, forward sweep in shared memory, tabulated elimination
    factors and once-refined reciprocal pivots on the simple columns) == interp_4th_cpt of the oracle on (2:jpim1, 2:jpjm1,
    2:jpkm1), with and without the simple-column shortcut; nothing else is written"""
    gf = H.random_fields(O, G, GJ, K, jperio, 2, seed=800 + K, ln_isfcav=isf)
    w = O.World(G, GJ, K, jperio, 1, 1)
    d = w.doms[0]
    d.set_fields(*[gf[k] for k in H.DOM_KEYS], ln_isfcav=isf)
    ref = np.zeros_like(gf["ptn"])
    for jn in range(2):
        O.lib().interp_4th_cpt(d.h, gf["ptn"][jn].ctypes.data_as(C.c_void_p), ref[jn].ctypes.data_as(C.c_void_p))
    w.close()
    inner = (slice(None), slice(1, K - 1), slice(1, -1), slice(1, -1))
    for simple in (True, False):
        out = np.full_like(ref, -7.0)
        assert emu_api.interp_4th_cpt_tiled(emu, gf, gf["ptn"], out, simple) == 0
        assert np.array_equal(out[inner], ref[inner]), simple
        out[inner] = -7.0
        assert (out == -7.0).all()
