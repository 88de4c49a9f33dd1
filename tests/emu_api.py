"""ctypes access to tests/emu/libemu.so: the product's column kernels compiled for the host (test infrastructure only,
see tests/emu/emu_common.h).  Shared by test_cpu_kernel_emulation.py and test_cpu_gloo_widened_rows.py."""
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
