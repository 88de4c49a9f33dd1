// nxt_kernels.cu -- fp64 CUDA kernels (sm_100a) of the tracer time-stepping swap that follows the advection step:
// Asselin filter + field swap of tra_nxt_fix / tra_nxt_vvl (src/OCE/TRA/tranxt.F90:190-234, 237-380) and the Euler
// swap of tra_nxt / trc_nxt (tranxt.F90:148-153, src/TOP/TRP/trcnxt.F90:147-153).
//
// One thread per interior point (ji fastest => coalesced), the tracer loop inside the thread so that the three
// thicknesses and the per-column forcing terms are loaded once for all tracers.  Pure streaming: 5 array accesses per
// tracer-point + 3 per point.  Compiled with -fmad=false; every expression keeps the reference's operation order.
#include "kernels.cuh"

namespace nemo {

void note_launch();

namespace {

constexpr int kThreads = 128;

__device__ __forceinline__ double v2(const double *p, size_t i) { return p ? p[i] : 0.0; }

// tra_nxt_fix: ptb <- ptn + atfp * ( pta - 2 ptn + ptb ) ; ptn <- pta            tranxt.F90:217-232
__global__ void __launch_bounds__(kThreads) k_nxt_fix(const NxtArgs a)
{
    const int ji = 2 + (int)(blockIdx.x * blockDim.x + threadIdx.x);
    if (ji > a.jpi - 1) return;
    const int jj = 2 + (int)blockIdx.y, jk = 1 + (int)blockIdx.z;
    const size_t o = (size_t)(jk - 1) * a.jpij + (size_t)(jj - 1) * a.jpi + (size_t)(ji - 1);
    for (int jn = 0; jn < a.kjpt; ++jn) {
        const size_t q = o + (size_t)jn * a.n3;
        const double ztn = a.ptn[q];
        const double ta = a.pta[q];
        const double ztd = ta - 2.0 * ztn + a.ptb[q];
        a.ptb[q] = ztn + a.atfp * ztd;
        a.ptn[q] = ta;
    }
}

// tra_nxt_vvl: thickness-weighted Asselin filter with the surface / runoff / solar / ice-shelf corrections
//                                                                                 tranxt.F90:283-357
__global__ void __launch_bounds__(kThreads) k_nxt_vvl(const NxtArgs a)
{
    const int ji = 2 + (int)(blockIdx.x * blockDim.x + threadIdx.x);
    if (ji > a.jpi - 1) return;
    const int jj = 2 + (int)blockIdx.y, jk = 1 + (int)blockIdx.z;
    const size_t o2 = (size_t)(jj - 1) * a.jpi + (size_t)(ji - 1);
    const size_t o = (size_t)(jk - 1) * a.jpij + o2;
    const double ze3t_b = a.e3t_b[o], ze3t_n = a.e3t_n[o], ze3t_a = a.e3t_a[o];
    const double ze3t_d = ze3t_a - 2. * ze3t_n + ze3t_b;
    double ze3t_f = ze3t_n + a.atfp * ze3t_d;
    const int mik = a.mikt[o2];
    const bool first = jk == mik;
    if (first)
        ze3t_f = ze3t_f - a.zfact2 * ((v2(a.emp_b, o2) - v2(a.emp, o2)) + (v2(a.fwfisf_b, o2) - v2(a.fwfisf, o2)));
    if (a.ln_rnf_depth) {
        if (mik <= jk && jk <= a.nk_rnf[o2])
            ze3t_f = ze3t_f - a.zfact2 * (-(v2(a.rnf_b, o2) - v2(a.rnf, o2))) * (ze3t_n / a.h_rnf[o2]);
    } else if (first) {
        ze3t_f = ze3t_f - a.zfact2 * (-(v2(a.rnf_b, o2) - v2(a.rnf, o2)));
    }
    const bool l_rnf = a.ll_rnf && jk <= a.nk_rnf[o2];
    const bool l_isf_in = a.ll_isf && jk >= a.misfkt[o2] && jk < a.misfkb[o2];
    const bool l_isf_b = a.ll_isf && jk == a.misfkb[o2];
    const double r1_e3t_f = 1.e0 / ze3t_f;
    for (int jn = 0; jn < a.kjpt; ++jn) {
        const size_t q = o + (size_t)jn * a.n3, q2 = o2 + (size_t)jn * a.jpij;
        const double ta = a.pta[q];
        const double ztc_b = a.ptb[q] * ze3t_b;
        const double ztc_n = a.ptn[q] * ze3t_n;
        const double ztc_a = ta * ze3t_a;
        const double ztc_d = ztc_a - 2. * ztc_n + ztc_b;
        double ztc_f = ztc_n + a.atfp * ztc_d;
        if (first) ztc_f = ztc_f - a.zfact1 * (v2(a.sbc_tc, q2) - v2(a.sbc_tc_b, q2));
        if (a.ll_traqsr && jn == 0 && jk <= a.nksr) ztc_f = ztc_f - a.zfact1 * (a.qsr_hc[o] - a.qsr_hc_b[o]);
        if (l_rnf) ztc_f = ztc_f - a.zfact1 * (a.rnf_tsc[q2] - a.rnf_tsc_b[q2]) * ze3t_n / a.h_rnf[o2];
        if (l_isf_in) ztc_f = ztc_f - a.zfact1 * (a.risf_tsc[q2] - a.risf_tsc_b[q2]) * ze3t_n * a.r1_hisf_tbl[o2];
        if (l_isf_b)  ztc_f = ztc_f - a.zfact1 * (a.risf_tsc[q2] - a.risf_tsc_b[q2]) * ze3t_n * a.r1_hisf_tbl[o2] * a.ralpha[o2];
        a.ptb[q] = ztc_f * r1_e3t_f;
        a.ptn[q] = ta;
    }
}

// Euler step: ptn(:,:,1:jpkm1,:) = pta ; for TOP also ptb = ptn                  tranxt.F90:148-153, trcnxt.F90:147-153
__global__ void __launch_bounds__(kThreads) k_nxt_euler(const NxtArgs a, int also_before)
{
    const size_t per = a.jpij * (size_t)(a.jpk - 1);
    const size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= per) return;
    for (int jn = 0; jn < a.kjpt; ++jn) {
        const size_t q = p + (size_t)jn * a.n3;
        const double ta = a.pta[q];
        a.ptn[q] = ta;
        if (also_before) a.ptb[q] = ta;
    }
}

}  // namespace

#ifndef NEMO_EMU_KERNELS_ONLY          // launchers (CUDA launch syntax) are left out of the host emulation build of tests/emu
void launch_nxt_fix(const NxtArgs &a, cudaStream_t s)
{
    const dim3 g((unsigned)((a.jpi - 2 + kThreads - 1) / kThreads), (unsigned)(a.jpj - 2), (unsigned)(a.jpk - 1));
    k_nxt_fix<<<g, kThreads, 0, s>>>(a); note_launch();
}
void launch_nxt_vvl(const NxtArgs &a, cudaStream_t s)
{
    const dim3 g((unsigned)((a.jpi - 2 + kThreads - 1) / kThreads), (unsigned)(a.jpj - 2), (unsigned)(a.jpk - 1));
    k_nxt_vvl<<<g, kThreads, 0, s>>>(a); note_launch();
}
void launch_nxt_euler(const NxtArgs &a, int also_before, cudaStream_t s)
{
    const size_t per = a.jpij * (size_t)(a.jpk - 1);
    k_nxt_euler<<<(unsigned)((per + kThreads - 1) / kThreads), kThreads, 0, s>>>(a, also_before); note_launch();
}
#endif  // NEMO_EMU_KERNELS_ONLY

}  // namespace nemo
