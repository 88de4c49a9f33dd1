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
              "zwi zwx zwy zwz zltu zltv ztw zbetup zbetdo trdx trdy trdz").split()


def fct(L, which, rc, nk, f, work, kjpt, h, v, lin, isf):
    """which: 0 laplacian, 1 low_antidiff, 2 betas, 3 limit, 4 final, 5 trend hook.  f: module arrays + pun.. + pta (local);
    work: dict of work arrays zwi.. (and trdx/trdy/trdz for the hook)"""
    jpk, jpj, jpi = f["tmask"].shape
    tab = (C.c_void_p * len(FCT_ARRAYS))()
    for n, name in enumerate(FCT_ARRAYS):
        a = work.get(name) if name in work else f.get(name)
        tab[n] = None if a is None else a.ctypes.data
    L.emu_fct.restype = C.c_int
    ret = L.emu_fct(which, jpi, jpj, jpk, kjpt, h, v, int(lin), int(isf), rect(*rc), nk, C.c_double(f["p2dt"]), tab,
                    p(f["mikt"]), p(f["mbkt"]))
    assert ret == 0


def interp_4th_cpt(L, f, pt_in, pt_out, isf, use_simple=True):
    """the product's interp_4th_cpt sequence (pivots, classification, solve) on pt_in (nfld, jpk, jpj, jpi); returns the number
    of 'simple' columns found"""
    jpk, jpj, jpi = f["tmask"].shape
    L.emu_interp_4th_cpt.restype = C.c_int
    return L.emu_interp_4th_cpt(jpi, jpj, jpk, pt_in.shape[0], int(isf), p(f["wmask"]), p(f["mikt"]), p(f["mbkt"]), p(pt_in), p(pt_out),
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
