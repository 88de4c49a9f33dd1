// host emulation of the tra_nxt kernels, see emu_common.h (test infrastructure only)
#include "emu_common.h"
#include "../../nemo-fmi-devel_b200/csrc/nxt_kernels.cu"

extern "C" {

// mode: 0 = tra_nxt_fix, 1 = tra_nxt_vvl, 2 = Euler swap (also_before = is_trc).  iflags = ll_traqsr, ll_rnf, ll_isf, ln_rnf_depth, nksr
int emu_nxt(int mode, int also_before, int jpi, int jpj, int jpk, int kjpt, double atfp, double zfact1, double zfact2, const int *iflags,
            const double *e3t_b, const double *e3t_n, const double *e3t_a, const int *mikt, double *ptb, double *ptn, double *pta,
            const double *sbc_tc, const double *sbc_tc_b, const double *const *f2d, const double *qsr_hc, const double *qsr_hc_b,
            const int *nk_rnf, const double *h_rnf, const double *rnf_tsc, const double *rnf_tsc_b, const int *misfkt, const int *misfkb,
            const double *risf_tsc, const double *risf_tsc_b, const double *r1_hisf_tbl, const double *ralpha)
{
    nemo::NxtArgs a;
    std::memset(&a, 0, sizeof a);
    a.jpi = jpi; a.jpj = jpj; a.jpk = jpk; a.kjpt = kjpt; a.jpij = (size_t)jpi * jpj; a.n3 = a.jpij * jpk;
    a.e3t_b = e3t_b; a.e3t_n = e3t_n; a.e3t_a = e3t_a; a.mikt = mikt; a.ptb = ptb; a.ptn = ptn; a.pta = pta;
    a.atfp = atfp; a.zfact1 = zfact1; a.zfact2 = zfact2;
    a.ll_traqsr = iflags[0]; a.ll_rnf = iflags[1]; a.ll_isf = iflags[2]; a.ln_rnf_depth = iflags[3]; a.nksr = iflags[4];
    a.sbc_tc = sbc_tc; a.sbc_tc_b = sbc_tc_b;
    a.emp_b = f2d[0]; a.emp = f2d[1]; a.fwfisf_b = f2d[2]; a.fwfisf = f2d[3]; a.rnf_b = f2d[4]; a.rnf = f2d[5];
    a.qsr_hc = qsr_hc; a.qsr_hc_b = qsr_hc_b; a.nk_rnf = nk_rnf; a.h_rnf = h_rnf; a.rnf_tsc = rnf_tsc; a.rnf_tsc_b = rnf_tsc_b;
    a.misfkt = misfkt; a.misfkb = misfkb; a.risf_tsc = risf_tsc; a.risf_tsc_b = risf_tsc_b; a.r1_hisf_tbl = r1_hisf_tbl; a.ralpha = ralpha;
    const int nthreads = 128;
    blockDim = {(unsigned)nthreads, 1, 1};
    using namespace nemo;
    if (mode == 2) {
        const size_t per = a.jpij * (size_t)(jpk - 1);
        for (size_t bx = 0; bx < (per + nthreads - 1) / nthreads; ++bx)
            for (int t = 0; t < nthreads; ++t) { blockIdx = {(unsigned)bx, 0, 0}; threadIdx = {(unsigned)t, 0, 0}; k_nxt_euler(a, also_before); }
        return 0;
    }
    for (int bz = 0; bz < jpk - 1; ++bz)                       // the launch grid of launch_nxt_fix / launch_nxt_vvl
        for (int by = 0; by < jpj - 2; ++by)
            for (int bx = 0; bx < (jpi - 2 + nthreads - 1) / nthreads; ++bx)
                for (int t = 0; t < nthreads; ++t) {
                    blockIdx = {(unsigned)bx, (unsigned)by, (unsigned)bz};
                    threadIdx = {(unsigned)t, 0, 0};
                    if (mode == 0) k_nxt_fix(a); else k_nxt_vvl(a);
                }
    return 0;
}

}  // extern "C"
