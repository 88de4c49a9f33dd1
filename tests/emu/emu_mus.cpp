// host emulation of the MUSCL column kernels, see emu_common.h (test infrastructure only)
#include "emu_common.h"
#include "../../nemo-fmi-devel_b200/csrc/mus_kernels.cu"

extern "C" {

// which: 0 grad, 1 hflux (exchanged differences), 2 hflux (from ptb), 3 trend, 4 inner.  rect = i0,i1,j0,j1 (1-based).
int emu_mus(int which, int jpi, int jpj, int jpk, int kjpt, const int *rect, int nkchunk, double p2dt, int ln_linssh, int ln_isfcav,
            const double *tmask, const double *umask, const double *vmask, const double *wmask, const double *e3t_n,
            const double *r1_e1e2t, const double *r1_e1e2u, const double *r1_e1e2v, const double *e3u_n, const double *e3v_n,
            const double *e3w_n, const double *xind, const int *mikt, const double *pun, const double *pvn, const double *pwn,
            const double *ptb, double *pta, double *zwx, double *zwy, double *fx, double *fy)
{
    nemo::MusArgs a;
    std::memset(&a, 0, sizeof a);
    a.reg = nemo::Region(); a.reg.add(rect[0], rect[1], rect[2], rect[3]);
    a.jpi = jpi; a.jpj = jpj; a.jpk = jpk; a.jpij = (size_t)jpi * jpj; a.n3 = a.jpij * jpk;
    a.tmask = tmask; a.umask = umask; a.vmask = vmask; a.wmask = wmask; a.e3t_n = e3t_n; a.r1_e1e2t = r1_e1e2t;
    a.r1_e1e2u = r1_e1e2u; a.r1_e1e2v = r1_e1e2v; a.e3u_n = e3u_n; a.e3v_n = e3v_n; a.e3w_n = e3w_n; a.xind = xind; a.mikt = mikt;
    a.pun = pun; a.pvn = pvn; a.pwn = pwn; a.ptb = ptb; a.pta = pta; a.zwx = zwx; a.zwy = zwy; a.fx = fx; a.fy = fy;
    a.p2dt = p2dt; a.kjpt = kjpt; a.ln_linssh = ln_linssh; a.ln_isfcav = ln_isfcav; a.nkchunk = nkchunk;
    const int ncol = a.reg.ncol();
    using namespace nemo;
    switch (which) {
    case 0: emu_run_grid(a, k_mus_grad, ncol, nkchunk, kjpt); break;
    case 1: if (xind) emu_run_grid(a, k_mus_hflux<true, false>, ncol, nkchunk, kjpt); else emu_run_grid(a, k_mus_hflux<false, false>, ncol, nkchunk, kjpt); break;
    case 2: if (xind) emu_run_grid(a, k_mus_hflux<true, true>, ncol, nkchunk, kjpt); else emu_run_grid(a, k_mus_hflux<false, true>, ncol, nkchunk, kjpt); break;
    case 3: if (xind) emu_run_grid(a, k_mus_trend<true>, ncol, nkchunk, kjpt); else emu_run_grid(a, k_mus_trend<false>, ncol, nkchunk, kjpt); break;
    case 4: if (xind) emu_run_grid(a, k_mus_inner<true>, ncol, nkchunk, kjpt); else emu_run_grid(a, k_mus_inner<false>, ncol, nkchunk, kjpt); break;
    default: return 1;
    }
    return 0;
}

}  // extern "C"
