"""ctypes access to tests/emu/libemu.so: the product's column kernels compiled for the host (test infrastructure only,
see tests/emu/emu_common.h).  Shared by test_cpu_kernel_emulation.py and test_cpu_gloo_steps.py."""
import ctypes as C
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
_L = None


def load():
    global _L
    if _L is None:
        d = os.path.join(HERE, "emu")
        subprocess.check_call(["make", "-C", d, "libemu.so"], stdout=subprocess.DEVNULL)
        L = C.CDLL(os.path.join(d, "libemu.so"))
        for f in (L.emu_mus, L.emu_cen, L.emu_nxt, L.emu_nonosc_final):
            f.restype = C.c_int
        _L = L
    return _L


def load_signcmp():
    """the fused kernel alone, built with -DNEMO_FCT_SIGN_BY_COMPARE (an experiment: sign tests by comparison, see fct_fused_kernel.cuh)"""
    d = os.path.join(HERE, "emu")
    subprocess.check_call(["make", "-C", d, "libemu_signcmp.so"], stdout=subprocess.DEVNULL)
    return C.CDLL(os.path.join(d, "libemu_signcmp.so"))


def p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def rect(*r):
    return (C.c_int * 4)(*r)


def mus(L, which, rc, nk, gf, mx, xind, pta, zwx, zwy, fx, fy, lin, isf, kjpt, ptb=None):
    """which: 0 grad, 1 hflux (exchanged differences), 2 hflux (from ptb), 3 trend, 4 inner; rc = (i0, i1, j0, j1) 1-based"""
    jpk, jpj, jpi = gf["tmask"].shape
    ret = L.emu_mus(which, jpi, jpj, jpk, kjpt, rect(*rc), nk, C.c_double(gf["p2dt"]), int(lin), int(isf),
                    p(gf["tmask"]), p(gf["umask"]), p(gf["vmask"]), p(gf["wmask"]), p(gf["e3t_n"]), p(gf["r1_e1e2t"]),
                    p(mx["r1_e1e2u"]), p(mx["r1_e1e2v"]), p(mx["e3u_n"]), p(mx["e3v_n"]), p(mx["e3w_n"]), p(xind),
                    p(gf["mikt"]), p(gf["pun"]), p(gf["pvn"]), p(gf["pwn"]), p(gf["ptb"] if ptb is None else ptb), p(pta),
                    p(zwx), p(zwy), p(fx), p(fy))
    assert ret == 0


FCT_ARRAYS = ("tmask umask vmask wmask e3t_b e3t_n e3t_a e1e2t r1_e1e2t pun pvn pwn ptb ptn pta "
              "zwi zwx zwy zwz zltu zltv ztw zbetup zbetdo trdx trdy trdz zlx zly zlz").split()


def fct(L, which, rc, nk, f, work, kjpt, h, v, lin, isf, masks_from_t=False):
    """which: 0 laplacian, 1 low_antidiff, 2 betas, 3 limit, 4 final, 5 trend hook, 6 fused inner P1-P5.  f: module arrays +
    pun.. + pta (local); work: dict of work arrays zwi.. (trdx/trdy/trdz for the hook, zlx/zly/zlz for the fused frame)"""
    jpk, jpj, jpi = f["tmask"].shape
    tab = (C.c_void_p * len(FCT_ARRAYS))()
    for n, name in enumerate(FCT_ARRAYS):
        a = work.get(name) if name in work else f.get(name)
        tab[n] = None if a is None else a.ctypes.data
    L.emu_fct.restype = C.c_int
    ret = L.emu_fct(which, jpi, jpj, jpk, kjpt, h, v, int(lin), int(isf), rect(*rc), nk, C.c_double(f["p2dt"]), tab,
                    p(f["mikt"]), p(f["mbkt"]), int(masks_from_t))
    assert ret == 0


def fct_tma(L, rc, nk, f, work, kjpt, h, v, lin, isf, masks_from_t):
    """the TMA-tiled fused P1-P5 kernel on ONE rectangle.  Returns -1 where the product falls back to the cp.async kernel (odd
    jpi or odd first column), else the number of violated hardware rules (negative / odd box origins): must be 0"""
    jpk, jpj, jpi = f["tmask"].shape
    tab = (C.c_void_p * len(FCT_ARRAYS))()
    for n, name in enumerate(FCT_ARRAYS):
        a = work.get(name) if name in work else f.get(name)
        tab[n] = None if a is None else a.ctypes.data
    L.emu_fct_low_antidiff_tma.restype = C.c_int
    return L.emu_fct_low_antidiff_tma(jpi, jpj, jpk, kjpt, h, v, int(lin), int(isf), rect(*rc), nk, C.c_double(f["p2dt"]), tab,
                                      p(f["mikt"]), p(f["mbkt"]), int(masks_from_t))


def _regions(buf, n):
    out = []
    for r in range(n):
        b = buf[17 * r:17 * r + 17]
        out.append([tuple(b[1 + 4 * q:5 + 4 * q]) for q in range(b[0])])
    return out


def fct_fused_plan(L, jpi, jpj, fold, want_split=False):
    """schedule.hpp: dict of rectangle lists k1, k1_band, k1_centre, lowf, lap, bet, lim, fin + k2_out, split, min_size"""
    buf = (C.c_int * (8 * 17 + 6))()
    L.emu_fct_fused_plan(jpi, jpj, int(fold), int(want_split), buf)
    d = dict(zip("k1 k1_band k1_centre lowf lap bet lim fin".split(), _regions(buf, 8)))
    d["k2_out"] = tuple(buf[136:140]); d["split"] = bool(buf[140]); d["min_size"] = int(buf[141])
    return d


def mus_plan(L, which, jpi, jpj, fold):
    """schedule.hpp: which = 0 default (differences in place on `inner`), 1 fully fused: dict inner, grad, hflux, trend"""
    buf = (C.c_int * (4 * 17))()
    L.emu_mus_plan(which, jpi, jpj, int(fold), buf)
    return dict(zip("inner grad hflux trend".split(), _regions(buf, 4)))


def nonosc_final(L, f, work, kjpt, out_rect):
    """the fused limiter kernel (512 cooperating threads per block) on the output rectangle, in place on work['pta']"""
    jpk, jpj, jpi = f["tmask"].shape
    ret = L.emu_nonosc_final(jpi, jpj, jpk, kjpt, rect(*out_rect), C.c_double(f["p2dt"]), p(f["tmask"]), p(f["e3t_n"]), p(f["e1e2t"]),
                             p(f["r1_e1e2t"]), p(f["ptb"]), p(work["zwi"]), p(work["zwx"]), p(work["zwy"]), p(work["zwz"]), p(work["pta"]))
    assert ret == 0


def fct_fused(L, out, nk, f, work, kjpt, h, v, lin, isf, masks_from_t):
    """the one-kernel step k_fct_fused on the output rectangle `out`, in place on work['pta'].  Returns -1 where the product
    falls back to the three-kernel schedule, else the number of violated TMA box rules: must be 0"""
    jpk, jpj, jpi = f["tmask"].shape
    tab = (C.c_void_p * len(FCT_ARRAYS))()
    for n, name in enumerate(FCT_ARRAYS):
        a = work.get(name) if name in work else f.get(name)
        tab[n] = None if a is None else a.ctypes.data
    L.emu_fct_fused.restype = C.c_int
    return L.emu_fct_fused(jpi, jpj, jpk, kjpt, h, v, int(lin), int(isf), rect(*out), nk, C.c_double(f["p2dt"]), tab,
                           p(f["mikt"]), p(f["mbkt"]), int(masks_from_t))


def fct_step_one_kernel(L, f, kjpt, h, v, lin, isf, nk, lbc, fold, nk_fused=1, fused_lib=None):
    """tra_adv_fct in schedule 4 (run_fct): k_fct_fused on K2's rectangle straight from the inputs; the band of K1 (which
    leaves pta alone inside that rectangle) and the frame chain with X1..X4 as in the other fused schedules.  Returns
    (pta, plan) or (None, plan) where the product falls back (no split possible, odd jpi)."""
    import numpy as np
    jpk, jpj, jpi = f["tmask"].shape
    shp = (kjpt, jpk, jpj, jpi)
    plan = fct_fused_plan(L, jpi, jpj, fold, True)
    if not plan["split"] or jpi % 2:
        return None, plan
    work = {k: np.zeros(shp) for k in ("zwi", "zwx", "zwy", "zwz", "zltu", "zltv", "ztw", "zbetup", "zbetdo", "zlx", "zly", "zlz")}
    work["pta"] = f["pta"].copy()
    frame = {k: v for k, v in work.items() if k not in ("zlx", "zly", "zlz")}

    def on(region, which, w, **kw):
        for rc in plan[region]:
            fct(L, which, rc, nk, f, w, kjpt, h, v, lin, isf, **kw)

    if v == 4:
        interp_4th_cpt(L, f, f["ptn"], work["ztw"], isf)
    if h == 4:
        on("lap", 0, frame)
        lbc([(work["zltu"], "T", 1.0), (work["zltv"], "T", 1.0)])                     # X1
    on("lowf", 1, frame)
    L.emu_fct_set_skip_rect(*plan["k2_out"])
    on("k1_band", 6, frame, masks_from_t=True)
    L.emu_fct_set_skip_rect(0, -1, 0, -1)
    lbc([(work["zwi"], "T", 1.0), (work["zwx"], "U", -1.0), (work["zwy"], "V", -1.0), (work["zwz"], "W", 1.0)])   # X2
    on("bet", 2, frame)
    lbc([(work["zbetup"], "T", 1.0), (work["zbetdo"], "T", 1.0)])                     # X3
    on("lim", 3, work)
    lbc([(work["zlx"], "U", -1.0), (work["zly"], "V", -1.0)])                         # X4 on the limited copies
    on("fin", 4, work)
    # the fused kernel is independent of all of the above (main stream): run it last to prove it reads inputs only
    rc = fct_fused(fused_lib or L, plan["k2_out"], nk_fused, f, frame, kjpt, h, v, lin, isf, True)
    assert rc == 0, "TMA box origin rule violated / kernel refused: %d" % rc
    return work["pta"], plan


def fct_step_fused(L, f, kjpt, h, v, lin, isf, nk, lbc, fold, masks_from_t, want_split=False, tma=False):
    """tra_adv_fct in the fused schedule (run_fct, schedules 1/2) on one subdomain: fused inner kernels on the regions of
    schedule.hpp, reference-structured kernels on the frame bands with the exchanges X1..X4 through lbc()."""
    import numpy as np
    jpk, jpj, jpi = f["tmask"].shape
    shp = (kjpt, jpk, jpj, jpi)
    plan = fct_fused_plan(L, jpi, jpj, fold, want_split)
    work = {k: np.zeros(shp) for k in ("zwi", "zwx", "zwy", "zwz", "zltu", "zltv", "ztw", "zbetup", "zbetdo", "zlx", "zly", "zlz")}
    work["pta"] = f["pta"].copy()
    frame = {k: v for k, v in work.items() if k not in ("zlx", "zly", "zlz")}        # kernels that must not see zlx..: NULL

    def on(region, which, w, **kw):
        for rc in plan[region]:
            fct(L, which, rc, nk, f, w, kjpt, h, v, lin, isf, **kw)

    if v == 4:
        interp_4th_cpt(L, f, f["ptn"], work["ztw"], isf)
    used_tma = False
    if tma and len(plan["k1_centre"]) == 1:                                          # schedule 2: TMA tiles where possible
        rc_tma = fct_tma(L, plan["k1_centre"][0], nk, f, frame, kjpt, h, v, lin, isf, masks_from_t)
        assert rc_tma <= 0, "TMA box origin rule violated %d times" % rc_tma
        used_tma = rc_tma == 0
    plan["used_tma"] = used_tma
    if not used_tma:
        on("k1_centre", 6, frame, masks_from_t=masks_from_t)                         # main stream
    if h == 4:
        on("lap", 0, frame)
        lbc([(work["zltu"], "T", 1.0), (work["zltv"], "T", 1.0)])                     # X1
    on("lowf", 1, frame)
    if plan["split"]:
        on("k1_band", 6, frame, masks_from_t=masks_from_t)
    lbc([(work["zwi"], "T", 1.0), (work["zwx"], "U", -1.0), (work["zwy"], "V", -1.0), (work["zwz"], "W", 1.0)])   # X2
    on("bet", 2, frame)
    lbc([(work["zbetup"], "T", 1.0), (work["zbetdo"], "T", 1.0)])                     # X3
    on("lim", 3, work)                                                                # limited fluxes -> zlx, zly, zlz
    lbc([(work["zlx"], "U", -1.0), (work["zly"], "V", -1.0)])                         # X4 on the limited copies
    nonosc_final(L, f, work, kjpt, plan["k2_out"])                                    # main stream; reads the unlimited fluxes
    on("fin", 4, work)
    return work["pta"], plan


def interp_4th_cpt(L, f, pt_in, pt_out, isf, use_simple=True):
    """the product's interp_4th_cpt sequence (pivots, classification, solve) on pt_in (nfld, jpk, jpj, jpi); returns the number
    of 'simple' columns found"""
    jpk, jpj, jpi = f["tmask"].shape
    L.emu_interp_4th_cpt.restype = C.c_int
    return L.emu_interp_4th_cpt(jpi, jpj, jpk, pt_in.shape[0], int(isf), p(f["wmask"]), p(f["mikt"]), p(f["mbkt"]), p(pt_in), p(pt_out),
                                int(use_simple))


def interp_4th_cpt_tiled(L, f, pt_in, pt_out, use_simple=True):
    """the TMA-tiled interp_4th_cpt kernel (k_interp_4th_cpt_tiled) on pt_in (nfld, jpk, jpj, jpi).  Returns -1 where the product
    falls back to the column kernel (odd jpi), else the number of violated TMA box rules: must be 0"""
    jpk, jpj, jpi = f["tmask"].shape
    L.emu_interp_4th_cpt_tiled.restype = C.c_int
    return L.emu_interp_4th_cpt_tiled(jpi, jpj, jpk, pt_in.shape[0], p(f["wmask"]), p(f["mikt"]), p(f["mbkt"]), p(pt_in), p(pt_out),
                                      int(use_simple))


def fct_step(L, f, kjpt, h, v, lin, isf, nk, lbc, hooks=False):
    """tra_adv_fct in the reference pass structure (run_fct, schedule 0) on one subdomain: emulated kernels + lbc(list of
    (array, cd_nat, psgn)) for the exchanges X1..X4.  Returns (pta, work arrays)."""
    import numpy as np
    jpk, jpj, jpi = f["tmask"].shape
    shp = (kjpt, jpk, jpj, jpi)
    work = {k: np.zeros(shp) for k in ("zwi", "zwx", "zwy", "zwz", "zltu", "zltv", "ztw", "zbetup", "zbetdo")}
    work["pta"] = f["pta"].copy()
    interior = (2, jpi - 1, 2, jpj - 1)
    if h == 4:
        fct(L, 0, interior, nk, f, work, kjpt, h, v, lin, isf)
        lbc([(work["zltu"], "T", 1.0), (work["zltv"], "T", 1.0)])                                   # X1
    if v == 4:
        interp_4th_cpt(L, f, f["ptn"], work["ztw"], isf)
    fct(L, 1, interior, nk, f, work, kjpt, h, v, lin, isf)
    lbc([(work["zwi"], "T", 1.0), (work["zwx"], "U", -1.0), (work["zwy"], "V", -1.0), (work["zwz"], "W", 1.0)])   # X2
    fct(L, 2, interior, nk, f, work, kjpt, h, v, lin, isf)
    lbc([(work["zbetup"], "T", 1.0), (work["zbetdo"], "T", 1.0)])                                   # X3
    fct(L, 3, interior, nk, f, work, kjpt, h, v, lin, isf)
    lbc([(work["zwx"], "U", -1.0), (work["zwy"], "V", -1.0)])                                       # X4
    fct(L, 4, interior, nk, f, work, kjpt, h, v, lin, isf)
    if hooks:
        for k in ("trdx", "trdy", "trdz"):
            work[k] = np.full(shp, -7.0e77)
        fct(L, 5, interior, 1, f, work, kjpt, h, v, lin, isf)
    return work["pta"], work


def glob_sum(L, fields, tmask_i, w3d=None, nblk=4):
    """k_glob_sum_partial + k_glob_sum_final (csrc/glob_sum.cu) on a grid of nblk blocks: [(hi, lo)] per field"""
    import numpy as np
    nfld = len(fields)
    ipk = int(np.prod(fields[0].shape[:-2])) if fields[0].ndim > 2 else 1
    jpij = fields[0].shape[-1] * fields[0].shape[-2]
    tab = (C.c_void_p * nfld)(*[a.ctypes.data for a in fields])
    out = np.zeros(2 * nfld)
    L.emu_glob_sum.restype = C.c_int
    L.emu_glob_sum(tab, nfld, p(w3d), p(tmask_i), C.c_longlong(jpij), ipk, nblk, p(out))
    return [(out[2 * f], out[2 * f + 1]) for f in range(nfld)]


def stp_ctl(L, sshn, un, tem, sal, tmask, nblk=4):
    """k_stp_ctl_partial + k_stp_ctl_final (csrc/stp_ctl.cu): (values[6], first linear indices[6], flags)"""
    import numpy as np
    vals, idx, flags = np.zeros(6), np.zeros(6, np.int64), C.c_int(0)
    L.emu_stp_ctl.restype = C.c_int
    L.emu_stp_ctl(p(sshn), p(un), p(tem), p(sal), p(tmask), C.c_longlong(sshn.size), C.c_longlong(un.size), nblk, p(vals), p(idx), C.byref(flags))
    return vals, idx, flags.value
