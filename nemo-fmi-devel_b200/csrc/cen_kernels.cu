// cen_kernels.cu -- fp64 CUDA kernel (sm_100a) of the centred tracer-advection scheme, src/OCE/TRA/traadv_cen.F90:46-204
// (tra_adv_cen: 2nd / 4th order in the horizontal, 2nd order or 4th-order compact in the vertical).
//
// The reference's five sweeps per tracer (fluxes in i/j, flux in k, top value, divergence) are one column-marching kernel:
// each thread owns an interior (ji,jj) column (chunk), forms the six face fluxes of a cell in registers and applies the
// divergence; zwx/zwy/zwz never exist.  For kn_cen_h = 4 the masked gradients ztu, ztv come from k_mus_grad (the same
// expression as the first guess of the MUSCL slopes) and go through lbc_lnk('U',-1 / 'V',-1) as in :126; for
// kn_cen_v = 4 ztw comes from the interp_4th_cpt kernel.  Compiled with -fmad=false, reference operation order.
//
// kn_cen_h = 4 reproduces what the reference's loop bounds do at the first interior row and column (traadv_cen.F90:128-137):
// the flux at ji = 1 reads ztu(0,jj,jk) = the element before it in memory (ztu(jpi,jj-1,jk)), and zwy(:,1,:) is never
// assigned (undefined in Fortran): it is taken as 0 here, as in the oracle's un-poisoned work arrays.
#include "kernels.cuh"

namespace nemo {

void note_launch();

namespace {

constexpr int kThreads = 128;

__device__ __forceinline__ bool region_column(const Region &rg, int &ji, int &jj)
{
    const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= rg.start[rg.n]) return false;
    int q = 0;
    while (q + 1 < rg.n && p >= rg.start[q + 1]) ++q;
    const int loc = (int)p - rg.start[q];
    const int ni = rg.r[q].i1 - rg.r[q].i0 + 1;
    jj = rg.r[q].j0 + loc / ni;
    ji = rg.r[q].i0 + loc % ni;
    return true;
}

template <int H, int V>
__global__ void __launch_bounds__(kThreads) k_cen(const CenArgs a)
{
    int ji, jj;
    if (!region_column(a.reg, ji, jj)) return;
    const int kmax = a.jpk - 1;
    const int per = (kmax + a.nkchunk - 1) / a.nkchunk;
    const int ka = 1 + (int)blockIdx.y * per, kb = min(kmax, ka + per - 1);
    if (ka > kb) return;                                 // an empty trailing chunk
    const size_t toff = (size_t)blockIdx.z * a.n3;
    const double *ptn = a.ptn + toff, *ztu = a.ztu + toff, *ztv = a.ztv + toff, *ztw = a.ztw + toff;
    double *pta = a.pta + toff;
    const int jpi = a.jpi, jpk = a.jpk;
    const size_t jpij = a.jpij;
    const size_t c2 = (size_t)(jj - 1) * jpi + (ji - 1);
    const double r1 = a.r1_e1e2t[c2];
    const double r1_6 = 1.0 / 6.0;
    const int ktop = a.ln_isfcav ? a.mikt[c2] : 1;
    // vertical flux through the top face of level k (1..jpk)   :147-180
    auto fz = [&](int k) -> double {
        const size_t o = c2 + (size_t)(k - 1) * jpij;
        double v = 0.0;                                                         // zwz(:,:,1) = zwz(:,:,jpk) = 0  (:99-100)
        if (V == 2) { if (k >= 2) v = 0.5 * a.pwn[o] * (ptn[o] + ptn[o - jpij]) * a.wmask[o]; }      // jk = 2..jpk
        else        { if (k >= 2 && k <= jpk - 1) v = a.pwn[o] * ztw[o] * a.wmask[o]; }
        if (a.ln_linssh && k == ktop) v = a.pwn[o] * ptn[o];
        return v;
    };
    double fz_k = fz(ka);
    for (int k = ka; k <= kb; ++k) {
        const size_t o = c2 + (size_t)(k - 1) * jpij;
        const double t_c = ptn[o], t_w = ptn[o - 1], t_e = ptn[o + 1], t_s = ptn[o - jpi], t_n = ptn[o + jpi];
        const double u_c = a.pun[o], u_w = a.pun[o - 1], v_c = a.pvn[o], v_s = a.pvn[o - jpi];
        double fx_c, fx_w, fy_c, fy_s;
        if (H == 2) {                                                           // :106-114
            fx_c = 0.5 * u_c * (t_c + t_e); fx_w = 0.5 * u_w * (t_w + t_c);
            fy_c = 0.5 * v_c * (t_c + t_n); fy_s = 0.5 * v_s * (t_s + t_c);
        } else {                                                                // :128-137
            const double zC4t_u = (t_c + t_e) + r1_6 * (ztu[o - 1] - ztu[o + 1]);
            const double zC4t_uw = (t_w + t_c) + r1_6 * (ztu[o - 2] - ztu[o]);     // ji = 2: ztu(0,jj,jk), see the header
            const double zC4t_v = (t_c + t_n) + r1_6 * (ztv[o - jpi] - ztv[o + jpi]);
            fx_c = 0.5 * u_c * zC4t_u; fx_w = 0.5 * u_w * zC4t_uw;
            fy_c = 0.5 * v_c * zC4t_v;
            if (jj >= 3) {
                const double zC4t_vs = (t_s + t_c) + r1_6 * (ztv[o - 2 * jpi] - ztv[o]);
                fy_s = 0.5 * v_s * zC4t_vs;
            } else {
                fy_s = 0.0;                                                     // zwy(:,1,:) is never assigned in the reference
            }
        }
        const double fz_kp1 = fz(k + 1);
        pta[o] = pta[o] - (fx_c - fx_w + fy_c - fy_s + fz_k - fz_kp1) * r1 / a.e3t_n[o];   // :184-188
        fz_k = fz_kp1;
    }
}

}  // namespace

#ifndef NEMO_EMU_KERNELS_ONLY          // the launcher (CUDA launch syntax) is left out of the host emulation build of tests/emu
void launch_cen(const CenArgs &a, cudaStream_t s)
{
    if (a.reg.ncol() <= 0) return;
    const dim3 g((unsigned)((a.reg.ncol() + kThreads - 1) / kThreads), (unsigned)a.nkchunk, (unsigned)a.kjpt);
    if (a.kn_cen_h == 2 && a.kn_cen_v == 2)      k_cen<2, 2><<<g, kThreads, 0, s>>>(a);
    else if (a.kn_cen_h == 2 && a.kn_cen_v == 4) k_cen<2, 4><<<g, kThreads, 0, s>>>(a);
    else if (a.kn_cen_h == 4 && a.kn_cen_v == 2) k_cen<4, 2><<<g, kThreads, 0, s>>>(a);
    else                                         k_cen<4, 4><<<g, kThreads, 0, s>>>(a);
    note_launch();
}
#endif  // NEMO_EMU_KERNELS_ONLY

}  // namespace nemo
