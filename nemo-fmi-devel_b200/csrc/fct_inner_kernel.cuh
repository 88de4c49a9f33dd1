// fct_inner_kernel.cuh -- P1-P5 of the FCT step on the exchange-free inner region with a per-thread asynchronous
// prefetch ring (k_fct_low_antidiff_inner), and the mask accessor it shares with the TMA-tiled variant.
//
// Included by fct_kernels.cu INSIDE namespace nemo's anonymous namespace, after fct_column_kernels.cuh and after the
// includer has defined cp_async8 / cp_async_commit / cp_async_wait<N> (inline PTX there; plain copies in the host emulation
// of tests/emu, which compiles this very source to compare it with the oracle without a GPU).  Nothing else includes it.
#pragma once

// ============================================================================================================
// Schedule 1 (fused inner region).  The inner region keeps a margin of >= 3 cells from every array edge, so none of
// the values it reads is produced by an exchange: the same arithmetic as the reference pass
// structure, without the intermediate sweeps.  The boundary frame is left to the reference-structured kernels.
// ============================================================================================================

// umask / vmask / wmask: read from the module arrays, or derived from tmask when the host arrays were verified to be
// the plain products of dommsk.F90:176-177,193 (saves three array streams)
template <bool FROM_T> struct Masks {
    const FctArgs &a;
    __device__ __forceinline__ double u(size_t o) const { return FROM_T ? a.tmask[o] * a.tmask[o + 1] : a.umask[o]; }
    __device__ __forceinline__ double v(size_t o) const { return FROM_T ? a.tmask[o] * a.tmask[o + a.jpi] : a.vmask[o]; }
    __device__ __forceinline__ double w(size_t o, int k) const {
        return FROM_T ? (k == 1 ? a.tmask[o] : a.tmask[o] * a.tmask[o - a.jpij]) : a.wmask[o];
    }
};

// ---- per-thread asynchronous prefetch ring ---------------------------------------------------------------------
// The column-marching kernels are latency bound when every level waits for its own HBM loads (ncu: long-scoreboard
// stalls, ~6 dependent load groups per level).  Each thread therefore streams the values of ITS OWN column two
// levels ahead with cp.async (LDGSTS) into private shared-memory slots: no registers are held while the loads are
// in flight and no barrier is needed, because a thread only ever reads the slots it filled itself.
constexpr int kPfStages = 3;
enum { F_PTB = 0, F_PTN, F_PTA, F_ZTW, F_PUN, F_PVN, F_PWN, F_E3B, F_E3N, F_E3A, F_TM, F_LOW_COUNT };

// P1-P5 on the inner region with the 4th-order Laplacian zltu/zltv evaluated in place (:195-208 folded into :211-221)
template <int H, int V, bool FROM_T>
__global__ void __launch_bounds__(kThreads) k_fct_low_antidiff_inner(const FctArgs a)
{
    NEMO_DYN_SMEM(double, pf_smem);                         // [kPfStages][F_LOW_COUNT][kThreads]
    int ji, jj, ka, kb;
    if (!region_column(a.reg, ji, jj)) return;
    k_chunk(a.jpk - 1, a.nkchunk, ka, kb);
    if (ka > kb) return;                                 // an empty trailing chunk (nkchunk does not divide jpk-1 evenly)
    const size_t toff = (size_t)blockIdx.z * a.n3;
    const double *__restrict__ ptb = a.ptb + toff;
    const double *__restrict__ ptn = a.ptn + toff;
    double *__restrict__ pta = a.pta + toff;
    double *__restrict__ zwi = a.zwi + toff;
    double *__restrict__ zwx = a.zwx + toff;
    double *__restrict__ zwy = a.zwy + toff;
    double *__restrict__ zwz = a.zwz + toff;
    const double *__restrict__ ztw = a.ztw + toff;
    const double *__restrict__ tmask = a.tmask;
    const Masks<FROM_T> msk{a};
    const int jpi = a.jpi, jpk = a.jpk;
    const size_t jpij = a.jpij;
    const size_t c2 = (size_t)(jj - 1) * jpi + (ji - 1);
    const double r1 = a.r1_e1e2t[c2];
    const int ktop = a.ln_linssh ? (a.ln_isfcav ? a.mikt[c2] : 1) : 0;     // level whose top flux is pwn*ptb (:146-156)
    const double p2dt = a.p2dt;
    const double r1_6 = 1.0 / 6.0;
    // schedule 4: the columns of a.out get their whole trend from k_fct_fused; this launch only feeds the frame chain there
    const bool skip_pta = ji >= a.out.i0 && ji <= a.out.i1 && jj >= a.out.j0 && jj <= a.out.j1;

    auto slot = [&](int lev, int f) -> double * { return pf_smem + ((size_t)((lev % kPfStages) * F_LOW_COUNT + f) * kThreads + threadIdx.x); };
    auto issue = [&](int lev) {                          // own-column values of level `lev`
        if (lev <= jpk) {
            const size_t o = c2 + (size_t)(lev - 1) * jpij;
            cp_async8(slot(lev, F_PTB), ptb + o); cp_async8(slot(lev, F_PTN), ptn + o); cp_async8(slot(lev, F_PTA), pta + o);
            if (V == 4) cp_async8(slot(lev, F_ZTW), ztw + o);
            cp_async8(slot(lev, F_PUN), a.pun + o); cp_async8(slot(lev, F_PVN), a.pvn + o); cp_async8(slot(lev, F_PWN), a.pwn + o);
            cp_async8(slot(lev, F_E3B), a.e3t_b + o); cp_async8(slot(lev, F_E3N), a.e3t_n + o); cp_async8(slot(lev, F_E3A), a.e3t_a + o);
            cp_async8(slot(lev, F_TM), tmask + o);
        }
        cp_async_commit();
    };
    // upstream vertical flux through the top face of level k from column values (P2 + P2b, :137-156)
    auto upw = [&](int k, double w, double tb_k, double tb_km1, double wm) -> double {
        double v = 0.0;
        if (k >= 2 && k <= jpk - 1) {
            const double zfp_wk = w + fabs(w), zfm_wk = w - fabs(w);
            v = 0.5 * (zfp_wk * tb_k + zfm_wk * tb_km1) * wm;
        }
        if (k == ktop) v = w * tb_k;
        return v;
    };

    issue(ka); issue(ka + 1);
    // values of the level above the chunk (once per chunk)
    double tb_m = 0.0, tn_m = 0.0, tm_m = 0.0;
    if (ka >= 2) { const size_t om = c2 + (size_t)(ka - 2) * jpij; tb_m = ptb[om]; tn_m = ptn[om]; tm_m = tmask[om]; }
    double upz_k = 0.0;
    bool first = true;
    for (int k = ka; k <= kb; ++k) {
        const size_t o = c2 + (size_t)(k - 1) * jpij;
        issue(k + 2);
        // neighbour columns: plain loads (L1/L2 hits on lines the neighbouring threads stream), issued before the wait
        const double tb_w = ptb[o - 1], tb_e = ptb[o + 1], tb_s = ptb[o - jpi], tb_n = ptb[o + jpi];
        const double u_w = a.pun[o - 1], v_s = a.pvn[o - jpi];
        const double tn_e = ptn[o + 1], tn_n = ptn[o + jpi];
        double tn_w = 0.0, tn_ee = 0.0, tn_s = 0.0, tn_nn = 0.0, mu_w = 0.0, mu_c = 0.0, mu_e = 0.0, mv_s = 0.0, mv_c = 0.0, mv_n = 0.0;
        if (H == 4) {
            tn_w = ptn[o - 1]; tn_ee = ptn[o + 2]; tn_s = ptn[o - jpi]; tn_nn = ptn[o + 2 * (size_t)jpi];
            mu_w = msk.u(o - 1); mu_c = msk.u(o); mu_e = msk.u(o + 1);
            mv_s = msk.v(o - jpi); mv_c = msk.v(o); mv_n = msk.v(o + jpi);
        }
        double wm_c = 0.0, wm_p = 0.0;
        if (!FROM_T) { wm_c = a.wmask[o]; wm_p = a.wmask[o + jpij]; }
        cp_async_wait<1>();                              // levels k and k+1 of this column have landed
        const double tb_c = *slot(k, F_PTB), tn_c = *slot(k, F_PTN), ta_c = *slot(k, F_PTA);
        const double u_c = *slot(k, F_PUN), v_c = *slot(k, F_PVN), w_c = *slot(k, F_PWN);
        const double e3b = *slot(k, F_E3B), e3n = *slot(k, F_E3N), e3a = *slot(k, F_E3A), tm = *slot(k, F_TM);
        const double tb_p = *slot(k + 1, F_PTB), w_p = *slot(k + 1, F_PWN), tm_p = *slot(k + 1, F_TM);
        if (FROM_T) { wm_c = (k == 1) ? tm : tm * tm_m; wm_p = tm_p * tm; }
        if (first) { upz_k = upw(k, w_c, tb_c, tb_m, wm_c); first = false; }
        double zfp, zfm;
        zfp = u_c + fabs(u_c); zfm = u_c - fabs(u_c);
        const double upx_c = 0.5 * (zfp * tb_c + zfm * tb_e);
        zfp = u_w + fabs(u_w); zfm = u_w - fabs(u_w);
        const double upx_w = 0.5 * (zfp * tb_w + zfm * tb_c);
        zfp = v_c + fabs(v_c); zfm = v_c - fabs(v_c);
        const double upy_c = 0.5 * (zfp * tb_c + zfm * tb_n);
        zfp = v_s + fabs(v_s); zfm = v_s - fabs(v_s);
        const double upy_s = 0.5 * (zfp * tb_s + zfm * tb_c);
        const double upz_kp1 = upw(k + 1, w_p, tb_p, tb_c, wm_p);
        const double ztra = -(upx_c - upx_w + upy_c - upy_s + upz_k - upz_kp1) * r1;
        if (!skip_pta) pta[o] = ta_c + ztra / e3n * tm;
        zwi[o] = (e3b * tb_c + p2dt * ztra) / e3a * tm;
        if (H == 2) {
            zwx[o] = 0.5 * u_c * (tn_c + tn_e) - upx_c;
            zwy[o] = 0.5 * v_c * (tn_c + tn_n) - upy_c;
        } else {
            // ztu(i) = (ptn(i+1)-ptn(i))*umask(i);  zltu(i) = (ztu(i) + ztu(i-1))*r1_6   (:198, :204)
            const double ztu_w = (tn_c - tn_w) * mu_w, ztu_c = (tn_e - tn_c) * mu_c, ztu_e = (tn_ee - tn_e) * mu_e;
            const double ztv_s = (tn_c - tn_s) * mv_s, ztv_c = (tn_n - tn_c) * mv_c, ztv_n = (tn_nn - tn_n) * mv_n;
            const double zltu_c = (ztu_c + ztu_w) * r1_6, zltu_e = (ztu_e + ztu_c) * r1_6;
            const double zltv_c = (ztv_c + ztv_s) * r1_6, zltv_n = (ztv_n + ztv_c) * r1_6;
            const double zC2t_u = tn_c + tn_e, zC2t_v = tn_c + tn_n;
            zwx[o] = 0.5 * u_c * (zC2t_u + zltu_c - zltu_e) - upx_c;
            zwy[o] = 0.5 * v_c * (zC2t_v + zltv_c - zltv_n) - upy_c;
        }
        double fz = 0.0;
        if (k >= 2) {
            if (V == 2) fz = (w_c * 0.5 * (tn_c + tn_m) - upz_k) * wm_c;
            else        fz = (w_c * (*slot(k, F_ZTW)) - upz_k) * wm_c;
        }
        zwz[o] = fz;
        upz_k = upz_kp1; tn_m = tn_c; tm_m = tm; tb_m = tb_c;
    }
    cp_async_wait<0>();
}
