// host emulation of the centred-scheme column kernel, see emu_common.h (test infrastructure only)
#include "emu_common.h"
#include "../../nemo-fmi-devel_b200/csrc/cen_kernels.cu"

extern "C" {

int emu_cen(int h, int v, int jpi, int jpj, int jpk, int kjpt, const int *rect, int nkchunk, int ln_linssh, int ln_isfcav,
            const double *wmask, const double *e3t_n, const double *r1_e1e2t, const int *mikt, const double *pun, const double *pvn,
            const double *pwn, const double *ptn, double *pta, const double *ztu, const double *ztv, const double *ztw)
{
    nemo::CenArgs a;
    std::memset(&a, 0, sizeof a);
    a.reg = nemo::Region(); a.reg.add(rect[0], rect[1], rect[2], rect[3]);
    a.jpi = jpi; a.jpj = jpj; a.jpk = jpk; a.jpij = (size_t)jpi * jpj; a.n3 = a.jpij * jpk;
    a.wmask = wmask; a.e3t_n = e3t_n; a.r1_e1e2t = r1_e1e2t; a.mikt = mikt; a.pun = pun; a.pvn = pvn; a.pwn = pwn; a.ptn = ptn; a.pta = pta;
    a.ztu = ztu; a.ztv = ztv; a.ztw = ztw; a.kjpt = kjpt; a.kn_cen_h = h; a.kn_cen_v = v; a.ln_linssh = ln_linssh; a.ln_isfcav = ln_isfcav;
    a.nkchunk = nkchunk;
    const int ncol = a.reg.ncol();
    using namespace nemo;
    if (h == 2 && v == 2) emu_run_grid(a, k_cen<2, 2>, ncol, nkchunk, kjpt);
    else if (h == 2 && v == 4) emu_run_grid(a, k_cen<2, 4>, ncol, nkchunk, kjpt);
    else if (h == 4 && v == 2) emu_run_grid(a, k_cen<4, 2>, ncol, nkchunk, kjpt);
    else emu_run_grid(a, k_cen<4, 4>, ncol, nkchunk, kjpt);
    return 0;
}

}  // extern "C"
