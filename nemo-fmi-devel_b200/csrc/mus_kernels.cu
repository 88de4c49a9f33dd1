// mus_kernels.cu -- hand-written fp64 CUDA kernels (sm_100a) of the MUSCL tracer-advection step.
//
// What is computed follows src/OCE/TRA/traadv_mus.F90:55-273 (tra_adv_mus); how it is computed does not: the
// reference's 12 whole-array sweeps per tracer (first-guess slopes, slopes, limitation, fluxes, trend, horizontally and
// then vertically, with four automatic work arrays) become
//   * reference-structured path: k_mus_grad -> lbc_lnk(U,-1 / V,-1) -> k_mus_hflux -> lbc_lnk -> k_mus_trend,
//     three column kernels with slopes and limited slopes kept in registers, all tracers batched in one launch;
//   * default path: on the columns whose slopes need no exchanged value k_mus_hflux forms the first-guess differences
//     in place from ptb (no zwx / zwy arrays, first exchange only for the one-cell frame, on a side stream);
//   * fused inner path: k_mus_inner computes the whole trend of the columns whose 5-point-wide stencil touches no halo
//     cell straight from ptb (nothing but pta is written, no exchange is needed there); the reference-structured
//     kernels then only run on the two-cell frame around it.
// Memory-bound stencil, no tensor cores (fp64).  Bit-exactness: every expression keeps the reference's operation order,
// the file is compiled with -fmad=false, SIGN follows the key_nosignedzero override (b >= 0 -> +|a|,
// lib_fortran.F90:339-351), MIN is written as selects.
#include "kernels.cuh"

namespace nemo {

void note_launch();

namespace {

constexpr int kThreads = 128;

__device__ __forceinline__ double dmin(double a, double b) { return a < b ? a : b; }
__device__ __forceinline__ double fsign(double a, double b) { return b >= 0.0 ? fabs(a) : -fabs(a); }

// slope of the tracer from two neighbouring first-guess differences, then limited (traadv_mus.F90:149-167, 228-241):
//   zslp = ( a + b ) * ( 0.25 + SIGN( 0.25, a * b ) ) ;  zslp = SIGN( 1., zslp ) * MIN( |zslp|, 2|b|, 2|a| )
// a = difference at the point itself, b = the one before it (ji-1 / jj-1) or below it (jk+1).
__device__ __forceinline__ double mus_slope(double a, double b)
{
    const double s = (a + b) * (0.25 + fsign(0.25, a * b));
    return fsign(1., s) * dmin(dmin(fabs(s), 2. * fabs(b)), 2. * fabs(a));
}

// MUSCL flux through one face (traadv_mus.F90:175-189 and :246-251 with zalpha = 0.5 -/+ z0):
//   tr = transport, cour = 0.5 * tr * p2dt * r1_surf / e3 (already evaluated left to right by the caller),
//   t_dn / s_dn = tracer and slope on the far side (ji+1, jj+1, jk+1), t_up / s_up on the near side; xi = xind
template <bool VERT>
__device__ __forceinline__ double mus_flux(double tr, double cour, double t_dn, double s_dn, double t_up, double s_up, double xi, bool has_xi)
{
    const double z0 = fsign(0.5, tr);
    const double zalpha = VERT ? 0.5 + z0 : 0.5 - z0;
    const double zu = z0 - cour;
    const double zzwx = has_xi ? t_dn + xi * zu * s_dn : t_dn + zu * s_dn;     // xind = 1 where up-stream is not needed
    const double zzwy = has_xi ? t_up + xi * zu * s_up : t_up + zu * s_up;
    return tr * (zalpha * zzwx + (1. - zalpha) * zzwy);
}

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

__device__ __forceinline__ void k_chunk(int kmax, int nchunk, int &ka, int &kb)
{
    const int per = (kmax + nchunk - 1) / nchunk;
    ka = 1 + (int)blockIdx.y * per;
    kb = min(kmax, ka + per - 1);
}

// ------------------------------------------------------------------------------------------------------------
// first guess of the slopes on a.reg (the reference: 1:jpim1, 1:jpjm1, 1:jpkm1)   traadv_mus.F90:134-141
// ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads) k_mus_grad(const MusArgs a)
{
    int ji, jj, ka, kb;
    if (!region_column(a.reg, ji, jj)) return;
    k_chunk(a.jpk - 1, a.nkchunk, ka, kb);
    if (ka > kb) return;                                 // an empty trailing chunk (nkchunk does not divide jpk-1 evenly)
    const size_t toff = (size_t)blockIdx.z * a.n3;
    const double *ptb = a.ptb + toff;
    double *zwx = a.zwx + toff, *zwy = a.zwy + toff;
    const size_t c2 = (size_t)(jj - 1) * a.jpi + (ji - 1);
    for (int k = ka; k <= kb; ++k) {
        const size_t o = c2 + (size_t)(k - 1) * a.jpij;
        const double t = ptb[o];
        zwx[o] = a.umask[o] * (ptb[o + 1] - t);
        zwy[o] = a.vmask[o] * (ptb[o + a.jpi] - t);
    }
}

// ------------------------------------------------------------------------------------------------------------
// slopes, limitation and MUSCL horizontal fluxes on a.reg (interior columns)        traadv_mus.F90:145-191
// reads the exchanged first-guess differences zwx, zwy; writes the fluxes to fx, fy
// ------------------------------------------------------------------------------------------------------------
// FROM_T: the first-guess differences are formed in place from ptb (same expression as k_mus_grad) instead of being read
// from the exchanged zwx / zwy arrays -- valid for the columns whose ji-1..ji+1 / jj-1..jj+1 differences are not halo or
// fold-rewritten cells, i.e. (3:jpi-2, 3:jpj-2) minus one more row under a north fold.
template <bool XI, bool FROM_T>
__global__ void __launch_bounds__(kThreads) k_mus_hflux(const MusArgs a)
{
    int ji, jj, ka, kb;
    if (!region_column(a.reg, ji, jj)) return;
    k_chunk(a.jpk - 1, a.nkchunk, ka, kb);
    if (ka > kb) return;                                 // an empty trailing chunk (nkchunk does not divide jpk-1 evenly)
    const size_t toff = (size_t)blockIdx.z * a.n3;
    const double *ptb = a.ptb + toff, *zwx = a.zwx + toff, *zwy = a.zwy + toff;
    double *fx = a.fx + toff, *fy = a.fy + toff;
    const int jpi = a.jpi;
    const size_t c2 = (size_t)(jj - 1) * jpi + (ji - 1);
    const double r1u = a.r1_e1e2u[c2], r1v = a.r1_e1e2v[c2], p2dt = a.p2dt;
#pragma unroll 2
    for (int k = ka; k <= kb; ++k) {
        const size_t o = c2 + (size_t)(k - 1) * a.jpij;
        const double t_c = ptb[o], t_e = ptb[o + 1], t_n = ptb[o + jpi];
        double gx_w, gx_c, gx_e, gy_s, gy_c, gy_n;
        if (FROM_T) {
            const double t_w = ptb[o - 1], t_ee = ptb[o + 2], t_s = ptb[o - jpi], t_nn = ptb[o + 2 * jpi];
            gx_w = a.umask[o - 1] * (t_c - t_w); gx_c = a.umask[o] * (t_e - t_c); gx_e = a.umask[o + 1] * (t_ee - t_e);
            gy_s = a.vmask[o - jpi] * (t_c - t_s); gy_c = a.vmask[o] * (t_n - t_c); gy_n = a.vmask[o + jpi] * (t_nn - t_n);
        } else {
            gx_w = zwx[o - 1]; gx_c = zwx[o]; gx_e = zwx[o + 1];
            gy_s = zwy[o - jpi]; gy_c = zwy[o]; gy_n = zwy[o + jpi];
        }
        const double sx_c = mus_slope(gx_c, gx_w), sx_e = mus_slope(gx_e, gx_c);
        const double sy_c = mus_slope(gy_c, gy_s), sy_n = mus_slope(gy_n, gy_c);
        const double xi = XI ? a.xind[o] : 1.0;
        const double u = a.pun[o], v = a.pvn[o];
        fx[o] = mus_flux<false>(u, 0.5 * u * p2dt * r1u / a.e3u_n[o], t_e, sx_e, t_c, sx_c, xi, XI);
        fy[o] = mus_flux<false>(v, 0.5 * v * p2dt * r1v / a.e3v_n[o], t_n, sy_n, t_c, sy_c, xi, XI);
    }
}

// vertical first-guess difference zwx(jk) = tmask(jk) * ( ptb(jk-1) - ptb(jk) ), 0 at jk = 1 and jpk  (:219-223)
__device__ __forceinline__ double mus_gz(int k, int jpk, double tm_k, double t_km1, double t_k)
{
    return (k >= 2 && k <= jpk - 1) ? tm_k * (t_km1 - t_k) : 0.0;
}

// Marching state of one column for the vertical MUSCL flux: face f = jk+1 needs ptb(jk-1..jk+2).
// The caller walks jk = ka..kb and asks for the flux through the bottom face of each level; the flux through the top
// face of level ka is produced by the same routine one level higher.
struct MusColumn {
    const double *ptb, *tmask, *pwn, *e3w, *wmask, *xind;
    size_t c2, jpij;
    int jpk, ktop, ln_linssh;
    double r1, p2dt;
    bool has_xi;
    __device__ __forceinline__ double t(int k) const { return (k >= 1 && k <= jpk) ? ptb[c2 + (size_t)(k - 1) * jpij] : 0.0; }
    __device__ __forceinline__ double tm(int k) const { return (k >= 1 && k <= jpk) ? tmask[c2 + (size_t)(k - 1) * jpij] : 0.0; }
    // limited slope at level k (0 at k = 1 and k >= jpk)  (:225-241)
    __device__ __forceinline__ double slope(int k, double g_k, double g_kp1) const { return (k >= 2 && k <= jpk - 1) ? mus_slope(g_k, g_kp1) : 0.0; }
    // flux through the top face of level f (f = 1..jpk), given tracers and slopes of levels f-1 and f  (:243-264)
    __device__ __forceinline__ double face(int f, double t_up, double s_up, double t_dn, double s_dn) const
    {
        double v = 0.0;                                                     // zwx(:,:,1) = zwx(:,:,jpk) = 0
        if (f >= 2 && f <= jpk - 1) {
            const size_t of = c2 + (size_t)(f - 1) * jpij, ok = of - jpij;   // jk+1 and jk of the reference loop
            const double w = pwn[of];
            const double xi = has_xi ? xind[ok] : 1.0;
            v = mus_flux<true>(w, 0.5 * w * p2dt * r1 / e3w[of], t_dn, s_dn, t_up, s_up, xi, has_xi) * wmask[ok];
        }
        if (ln_linssh && f == ktop) { const size_t of = c2 + (size_t)(f - 1) * jpij; v = pwn[of] * ptb[of]; }
        return v;
    }
};

// ------------------------------------------------------------------------------------------------------------
// horizontal trend from the exchanged fluxes, then the whole vertical part, on a.reg   traadv_mus.F90:194-272
// ------------------------------------------------------------------------------------------------------------
template <bool XI>
__global__ void __launch_bounds__(kThreads) k_mus_trend(const MusArgs a)
{
    int ji, jj, ka, kb;
    if (!region_column(a.reg, ji, jj)) return;
    k_chunk(a.jpk - 1, a.nkchunk, ka, kb);
    if (ka > kb) return;                                 // an empty trailing chunk (nkchunk does not divide jpk-1 evenly)
    const size_t toff = (size_t)blockIdx.z * a.n3;
    const double *fx = a.fx + toff, *fy = a.fy + toff;
    double *pta = a.pta + toff;
    const int jpi = a.jpi, jpk = a.jpk;
    const size_t jpij = a.jpij;
    const size_t c2 = (size_t)(jj - 1) * jpi + (ji - 1);
    const double r1 = a.r1_e1e2t[c2];
    MusColumn col{a.ptb + toff, a.tmask, a.pwn, a.e3w_n, a.wmask, a.xind, c2, jpij, jpk,
                  a.ln_isfcav ? a.mikt[c2] : 1, a.ln_linssh, r1, a.p2dt, XI};
    // rolling tracers t(k-1..k+2), differences g(k..k+2), slopes s(k), s(k+1), flux through the top face of level k
    double t_m = col.t(ka - 1), t_c = col.t(ka), t_p = col.t(ka + 1);
    double g_c = mus_gz(ka, jpk, col.tm(ka), t_m, t_c), g_p = mus_gz(ka + 1, jpk, col.tm(ka + 1), t_c, t_p);
    double s_c = col.slope(ka, g_c, g_p);
    double f_top;
    {   // top face of the first level of the chunk needs the slope one level higher
        const double t_mm = col.t(ka - 2);
        const double g_m = mus_gz(ka - 1, jpk, col.tm(ka - 1), t_mm, t_m);
        const double s_m = col.slope(ka - 1, g_m, g_c);
        f_top = col.face(ka, t_m, s_m, t_c, s_c);
    }
#pragma unroll 2
    for (int k = ka; k <= kb; ++k) {
        const size_t o = c2 + (size_t)(k - 1) * jpij;
        const double t_pp = col.t(k + 2);
        const double g_pp = mus_gz(k + 2, jpk, col.tm(k + 2), t_p, t_pp);
        const double s_p = col.slope(k + 1, g_p, g_pp);
        const double f_bot = col.face(k + 1, t_c, s_c, t_p, s_p);
        const double e3 = a.e3t_n[o];
        double ta = pta[o];
        ta = ta - (fx[o] - fx[o - 1] + fy[o] - fy[o - jpi]) * r1 / e3;        // :198-200
        ta = ta - (f_top - f_bot) * r1 / e3;                                   // :269
        pta[o] = ta;
        t_m = t_c; t_c = t_p; t_p = t_pp; g_c = g_p; g_p = g_pp; s_c = s_p; f_top = f_bot;
    }
}

// ------------------------------------------------------------------------------------------------------------
// fused inner path: the complete MUSCL trend of the columns of a.reg straight from ptb.  Valid where ji-2..ji+2 and
// jj-2..jj+2 are inside the subdomain and none of those cells is rewritten by the two exchanges, i.e. on
// (3:jpi-2, 3:jpj-2) (minus one more row under a north fold); the caller restricts a.reg accordingly.
// ------------------------------------------------------------------------------------------------------------
template <bool XI>
__global__ void __launch_bounds__(kThreads) k_mus_inner(const MusArgs a)
{
    int ji, jj, ka, kb;
    if (!region_column(a.reg, ji, jj)) return;
    k_chunk(a.jpk - 1, a.nkchunk, ka, kb);
    if (ka > kb) return;                                 // an empty trailing chunk (nkchunk does not divide jpk-1 evenly)
    const size_t toff = (size_t)blockIdx.z * a.n3;
    const double *ptb = a.ptb + toff;
    double *pta = a.pta + toff;
    const int jpi = a.jpi, jpk = a.jpk;
    const size_t jpij = a.jpij;
    const size_t c2 = (size_t)(jj - 1) * jpi + (ji - 1);
    const double r1 = a.r1_e1e2t[c2], p2dt = a.p2dt;
    const double r1u_c = a.r1_e1e2u[c2], r1u_w = a.r1_e1e2u[c2 - 1], r1v_c = a.r1_e1e2v[c2], r1v_s = a.r1_e1e2v[c2 - jpi];
    MusColumn col{ptb, a.tmask, a.pwn, a.e3w_n, a.wmask, a.xind, c2, jpij, jpk,
                  a.ln_isfcav ? a.mikt[c2] : 1, a.ln_linssh, r1, p2dt, XI};
    double t_m = col.t(ka - 1), t_c = col.t(ka), t_p = col.t(ka + 1);
    double g_c = mus_gz(ka, jpk, col.tm(ka), t_m, t_c), g_p = mus_gz(ka + 1, jpk, col.tm(ka + 1), t_c, t_p);
    double s_c = col.slope(ka, g_c, g_p);
    double f_top;
    {
        const double t_mm = col.t(ka - 2);
        const double g_m = mus_gz(ka - 1, jpk, col.tm(ka - 1), t_mm, t_m);
        const double s_m = col.slope(ka - 1, g_m, g_c);
        f_top = col.face(ka, t_m, s_m, t_c, s_c);
    }
    for (int k = ka; k <= kb; ++k) {
        const size_t o = c2 + (size_t)(k - 1) * jpij;
        // ---- horizontal: tracers ji-2..ji+2 / jj-2..jj+2, first-guess differences at ji-2..ji+1 / jj-2..jj+1 ----
        const double tww = ptb[o - 2], tw = ptb[o - 1], te = ptb[o + 1], tee = ptb[o + 2];
        const double tss = ptb[o - 2 * jpi], ts = ptb[o - jpi], tn = ptb[o + jpi], tnn = ptb[o + 2 * jpi];
        const double gx_ww = a.umask[o - 2] * (tw - tww), gx_w = a.umask[o - 1] * (t_c - tw);
        const double gx_c = a.umask[o] * (te - t_c), gx_e = a.umask[o + 1] * (tee - te);
        const double gy_ss = a.vmask[o - 2 * jpi] * (ts - tss), gy_s = a.vmask[o - jpi] * (t_c - ts);
        const double gy_c = a.vmask[o] * (tn - t_c), gy_n = a.vmask[o + jpi] * (tnn - tn);
        const double sx_w = mus_slope(gx_w, gx_ww), sx_c = mus_slope(gx_c, gx_w), sx_e = mus_slope(gx_e, gx_c);
        const double sy_s = mus_slope(gy_s, gy_ss), sy_c = mus_slope(gy_c, gy_s), sy_n = mus_slope(gy_n, gy_c);
        const double xi_c = XI ? a.xind[o] : 1.0, xi_w = XI ? a.xind[o - 1] : 1.0, xi_s = XI ? a.xind[o - jpi] : 1.0;
        const double u_c = a.pun[o], u_w = a.pun[o - 1], v_c = a.pvn[o], v_s = a.pvn[o - jpi];
        const double fx_c = mus_flux<false>(u_c, 0.5 * u_c * p2dt * r1u_c / a.e3u_n[o], te, sx_e, t_c, sx_c, xi_c, XI);
        const double fx_w = mus_flux<false>(u_w, 0.5 * u_w * p2dt * r1u_w / a.e3u_n[o - 1], t_c, sx_c, tw, sx_w, xi_w, XI);
        const double fy_c = mus_flux<false>(v_c, 0.5 * v_c * p2dt * r1v_c / a.e3v_n[o], tn, sy_n, t_c, sy_c, xi_c, XI);
        const double fy_s = mus_flux<false>(v_s, 0.5 * v_s * p2dt * r1v_s / a.e3v_n[o - jpi], t_c, sy_c, ts, sy_s, xi_s, XI);
        // ---- vertical ----
        const double t_pp = col.t(k + 2);
        const double g_pp = mus_gz(k + 2, jpk, col.tm(k + 2), t_p, t_pp);
        const double s_p = col.slope(k + 1, g_p, g_pp);
        const double f_bot = col.face(k + 1, t_c, s_c, t_p, s_p);
        const double e3 = a.e3t_n[o];
        double ta = pta[o];
        ta = ta - (fx_c - fx_w + fy_c - fy_s) * r1 / e3;
        ta = ta - (f_top - f_bot) * r1 / e3;
        pta[o] = ta;
        t_m = t_c; t_c = t_p; t_p = t_pp; g_c = g_p; g_p = g_pp; s_c = s_p; f_top = f_bot;
    }
}

// Upstream / MUSCL indicator (init only)                                             traadv_mus.F90:99-113
__global__ void __launch_bounds__(kThreads) k_mus_xind(int jpi, int jpj, int jpk, const double *rnfmsk, const double *rnfmsk_z,
                                                       const double *tmask, double *xind)
{
    const size_t jpij = (size_t)jpi * jpj;
    const size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= jpij) return;
    const double upsmsk = 0.0;
    for (int k = 1; k <= jpk; ++k) {
        const size_t o = p + (size_t)(k - 1) * jpij;
        double v = 1.0;
        if (k <= jpk - 1) { const double m = rnfmsk[p] * rnfmsk_z[k - 1]; v = 1.0 - (m > upsmsk ? m : upsmsk) * tmask[o]; }
        xind[o] = v;
    }
}

#ifndef NEMO_EMU_KERNELS_ONLY          // launchers (CUDA launch syntax) are left out of the host emulation build of tests/emu
inline dim3 column_grid(const MusArgs &a)
{
    return dim3((unsigned)((a.reg.ncol() + kThreads - 1) / kThreads), (unsigned)a.nkchunk, (unsigned)a.kjpt);
}
#endif

}  // namespace

#ifndef NEMO_EMU_KERNELS_ONLY
void launch_mus_grad(const MusArgs &a, cudaStream_t s)
{
    if (a.reg.ncol() <= 0) return;
    k_mus_grad<<<column_grid(a), kThreads, 0, s>>>(a); note_launch();
}
void launch_mus_hflux(const MusArgs &a, cudaStream_t s, bool from_ptb)
{
    if (a.reg.ncol() <= 0) return;
    if (from_ptb) {
        if (a.xind) k_mus_hflux<true, true><<<column_grid(a), kThreads, 0, s>>>(a);
        else        k_mus_hflux<false, true><<<column_grid(a), kThreads, 0, s>>>(a);
    } else {
        if (a.xind) k_mus_hflux<true, false><<<column_grid(a), kThreads, 0, s>>>(a);
        else        k_mus_hflux<false, false><<<column_grid(a), kThreads, 0, s>>>(a);
    }
    note_launch();
}
void launch_mus_trend(const MusArgs &a, cudaStream_t s)
{
    if (a.reg.ncol() <= 0) return;
    if (a.xind) k_mus_trend<true><<<column_grid(a), kThreads, 0, s>>>(a);
    else        k_mus_trend<false><<<column_grid(a), kThreads, 0, s>>>(a);
    note_launch();
}
void launch_mus_inner(const MusArgs &a, cudaStream_t s)
{
    if (a.reg.ncol() <= 0) return;
    if (a.xind) k_mus_inner<true><<<column_grid(a), kThreads, 0, s>>>(a);
    else        k_mus_inner<false><<<column_grid(a), kThreads, 0, s>>>(a);
    note_launch();
}
void launch_mus_xind(int jpi, int jpj, int jpk, const double *rnfmsk, const double *rnfmsk_z, const double *tmask, double *xind,
                     cudaStream_t s)
{
    const size_t jpij = (size_t)jpi * jpj;
    k_mus_xind<<<(unsigned)((jpij + kThreads - 1) / kThreads), kThreads, 0, s>>>(jpi, jpj, jpk, rnfmsk, rnfmsk_z, tmask, xind);
    note_launch();
}
#endif  // NEMO_EMU_KERNELS_ONLY

}  // namespace nemo
