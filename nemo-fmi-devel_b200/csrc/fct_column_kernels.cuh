// fct_column_kernels.cuh -- the column-marching fp64 kernels of the FCT step: the reference pass structure of
// src/OCE/TRA/traadv_fct.F90 (P1-P8 as k_fct_laplacian, k_fct_low_antidiff, k_fct_betas, k_fct_limit, k_fct_final), the
// compact vertical interpolation interp_4th_cpt (:517-616), the transports of tra_adv (traadv.F90:100-124) and the
// trend-diagnostic hook.  One thread owns one (ji,jj) column (chunk) and marches down jk; plain loads, no shared memory,
// no inline PTX.
//
// Included by fct_kernels.cu INSIDE namespace nemo's anonymous namespace, after kernels.cuh, kThreads and
// nonosc_final.cuh (dmax / dmin / bup_bdo).  A header of its own so that the host emulation of tests/emu can compile the
// very same kernel source and compare it with the oracle without a GPU; nothing else includes it.
#pragma once

// interior column owned by this thread: ji in 2..jpim1, jj in 2..jpjm1; returns false when out of range
__device__ __forceinline__ bool interior_column(int jpi, int jpj, int &ji, int &jj)
{
    const int ni = jpi - 2;
    const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= (long long)ni * (jpj - 2)) return false;
    jj = (int)(p / ni) + 2;
    ji = (int)(p - (long long)(jj - 2) * ni) + 2;
    return true;
}

// column owned by this thread within a Region (up to 4 rectangles)
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

// this block's chunk of the jk loop: levels ka..kb (1-based, inclusive) out of 1..kmax
__device__ __forceinline__ void k_chunk(int kmax, int nchunk, int &ka, int &kb)
{
    const int per = (kmax + nchunk - 1) / nchunk;
    ka = 1 + (int)blockIdx.y * per;
    kb = min(kmax, ka + per - 1);
}

// upstream vertical flux through the top face of level k  (P2 + P2b, traadv_fct.F90:137-156); k in 1..jpk
__device__ __forceinline__ double upstream_w(const FctArgs &a, const double *ptb, size_t col, int k, int mik)
{
    double v = 0.0;                                   // zwz(:,:,1) = zwz(:,:,jpk) = 0  (:114-115)
    if (k >= 2 && k <= a.jpk - 1) {
        const size_t o = col + (size_t)(k - 1) * a.jpij;
        const double w = a.pwn[o];
        const double zfp_wk = w + fabs(w), zfm_wk = w - fabs(w);
        v = 0.5 * (zfp_wk * ptb[o] + zfm_wk * ptb[o - a.jpij]) * a.wmask[o];
    }
    if (a.ln_linssh) {
        const int ktop = a.ln_isfcav ? mik : 1;
        if (k == ktop) { const size_t o = col + (size_t)(k - 1) * a.jpij; v = a.pwn[o] * ptb[o]; }
    }
    return v;
}

// ------------------------------------------------------------------------------------------------------------
// P4, 4th order: zltu, zltv  (traadv_fct.F90:195-208; NB the '+' of :204-205 is the reference)
// ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads) k_fct_laplacian(const FctArgs a)
{
    int ji, jj, ka, kb;
    if (!region_column(a.reg, ji, jj)) return;
    k_chunk(a.jpk - 1, a.nkchunk, ka, kb);
    if (ka > kb) return;                                 // an empty trailing chunk (nkchunk does not divide jpk-1 evenly)
    const size_t toff = (size_t)blockIdx.z * a.n3;
    const double *ptn = a.ptn + toff;
    double *zltu = a.zltu + toff, *zltv = a.zltv + toff;
    const double r1_6 = 1.0 / 6.0;
    const size_t col = (size_t)(jj - 1) * a.jpi + (ji - 1);
    const int jpi = a.jpi;
    for (int k = ka; k <= kb; ++k) {
        const size_t o = col + (size_t)(k - 1) * a.jpij;
        const double tc = ptn[o];
        const double ztu_c = (ptn[o + 1] - tc) * a.umask[o];
        const double ztu_w = (tc - ptn[o - 1]) * a.umask[o - 1];
        const double ztv_c = (ptn[o + jpi] - tc) * a.vmask[o];
        const double ztv_s = (tc - ptn[o - jpi]) * a.vmask[o - jpi];
        zltu[o] = (ztu_c + ztu_w) * r1_6;
        zltv[o] = (ztv_c + ztv_s) * r1_6;
    }
}

// ------------------------------------------------------------------------------------------------------------
// P1-P5 fused: upstream fluxes, low-order update, anti-diffusive fluxes
//   writes pta (+= low-order trend), zwi, zwx, zwy, zwz on (2:jpim1, 2:jpjm1, 1:jpkm1)
//   (halo cells of zwi/zwx/zwy/zwz are all defined by the X2 exchange that follows)
// ------------------------------------------------------------------------------------------------------------
template <int H, int V>
__global__ void __launch_bounds__(kThreads) k_fct_low_antidiff(const FctArgs a)
{
    int ji, jj, ka, kb;
    if (!region_column(a.reg, ji, jj)) return;
    k_chunk(a.jpk - 1, a.nkchunk, ka, kb);
    if (ka > kb) return;                                 // an empty trailing chunk (nkchunk does not divide jpk-1 evenly)
    const size_t toff = (size_t)blockIdx.z * a.n3;
    const double *ptb = a.ptb + toff, *ptn = a.ptn + toff;
    double *pta = a.pta + toff, *zwi = a.zwi + toff, *zwx = a.zwx + toff, *zwy = a.zwy + toff, *zwz = a.zwz + toff;
    const double *zltu = a.zltu + toff, *zltv = a.zltv + toff, *ztw = a.ztw + toff;
    const int jpi = a.jpi;
    const size_t jpij = a.jpij;
    const size_t c2 = (size_t)(jj - 1) * jpi + (ji - 1);
    const double r1 = a.r1_e1e2t[c2];
    const int mik = a.ln_isfcav ? a.mikt[c2] : 1;
    const double p2dt = a.p2dt;

    double upz_k = upstream_w(a, ptb, c2, ka, mik);                     // zwz upstream at the top face of level ka
    double tn_m = (ka >= 2) ? ptn[c2 + (size_t)(ka - 2) * jpij] : 0.0;  // ptn(jk-1)
    for (int k = ka; k <= kb; ++k) {
        const size_t o = c2 + (size_t)(k - 1) * jpij;
        const double tb_c = ptb[o], tb_w = ptb[o - 1], tb_e = ptb[o + 1], tb_s = ptb[o - jpi], tb_n = ptb[o + jpi];
        const double u_c = a.pun[o], u_w = a.pun[o - 1], v_c = a.pvn[o], v_s = a.pvn[o - jpi];
        // upstream fluxes through the 4 lateral faces (:127-132)
        double zfp, zfm;
        zfp = u_c + fabs(u_c); zfm = u_c - fabs(u_c);
        const double upx_c = 0.5 * (zfp * tb_c + zfm * tb_e);
        zfp = u_w + fabs(u_w); zfm = u_w - fabs(u_w);
        const double upx_w = 0.5 * (zfp * tb_w + zfm * tb_c);
        zfp = v_c + fabs(v_c); zfm = v_c - fabs(v_c);
        const double upy_c = 0.5 * (zfp * tb_c + zfm * tb_n);
        zfp = v_s + fabs(v_s); zfm = v_s - fabs(v_s);
        const double upy_s = 0.5 * (zfp * tb_s + zfm * tb_c);
        const double upz_kp1 = upstream_w(a, ptb, c2, k + 1, mik);
        // low-order trend and guess (:162-167)
        const double ztra = -(upx_c - upx_w + upy_c - upy_s + upz_k - upz_kp1) * r1;
        const double tm = a.tmask[o];
        pta[o] = pta[o] + ztra / a.e3t_n[o] * tm;
        zwi[o] = (a.e3t_b[o] * tb_c + p2dt * ztra) / a.e3t_a[o] * tm;
        // anti-diffusive fluxes: high order minus low order (:180-275)
        const double tn_c = ptn[o], tn_e = ptn[o + 1], tn_n = ptn[o + jpi];
        if (H == 2) {
            zwx[o] = 0.5 * u_c * (tn_c + tn_e) - upx_c;
            zwy[o] = 0.5 * v_c * (tn_c + tn_n) - upy_c;
        } else {
            const double zC2t_u = tn_c + tn_e, zC2t_v = tn_c + tn_n;
            zwx[o] = 0.5 * u_c * (zC2t_u + zltu[o] - zltu[o + 1]) - upx_c;
            zwy[o] = 0.5 * v_c * (zC2t_v + zltv[o] - zltv[o + jpi]) - upy_c;
        }
        double fz = 0.0;                                                // zwz(:,:,1): 0 (:114, :277)
        if (k >= 2) {
            if (V == 2) fz = (a.pwn[o] * 0.5 * (tn_c + tn_m) - upz_k) * a.wmask[o];
            else        fz = (a.pwn[o] * ztw[o] - upz_k) * a.wmask[o];
        }
        zwz[o] = fz;
        upz_k = upz_kp1; tn_m = tn_c;
    }
}

// ------------------------------------------------------------------------------------------------------------
// nonosc P6: local extrema and beta terms (traadv_fct.F90:356-399)
// ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads) k_fct_betas(const FctArgs a)
{
    int ji, jj, ka, kb;
    if (!region_column(a.reg, ji, jj)) return;
    k_chunk(a.jpk - 1, a.nkchunk, ka, kb);
    if (ka > kb) return;                                 // an empty trailing chunk (nkchunk does not divide jpk-1 evenly)
    const size_t toff = (size_t)blockIdx.z * a.n3;
    const double *pbef = a.ptb + toff, *paft = a.zwi + toff;
    const double *paa = a.zwx + toff, *pbb = a.zwy + toff, *pcc = a.zwz + toff;
    double *zbetup = a.zbetup + toff, *zbetdo = a.zbetdo + toff;
    const int jpi = a.jpi;
    const size_t jpij = a.jpij;
    const size_t c2 = (size_t)(jj - 1) * jpi + (ji - 1);
    const double zrtrn = 1.e-15;
    const double e12 = a.e1e2t[c2];
    const double p2dt = a.p2dt;

    double up_m, do_m, up_c, do_c, up_p, do_p;
    {
        const size_t o = c2 + (size_t)(ka - 1) * jpij;
        bup_bdo(pbef[o], paft[o], a.tmask[o], up_c, do_c);
        if (ka >= 2) bup_bdo(pbef[o - jpij], paft[o - jpij], a.tmask[o - jpij], up_m, do_m);
        else { up_m = up_c; do_m = do_c; }                              // ikm1 = MAX(jk-1,1)
    }
    double pcc_k = pcc[c2 + (size_t)(ka - 1) * jpij];
    for (int k = ka; k <= kb; ++k) {
        const size_t o = c2 + (size_t)(k - 1) * jpij;
        bup_bdo(pbef[o + jpij], paft[o + jpij], a.tmask[o + jpij], up_p, do_p);
        double up_w, do_w, up_e, do_e, up_s, do_s, up_n, do_n;
        bup_bdo(pbef[o - 1], paft[o - 1], a.tmask[o - 1], up_w, do_w);
        bup_bdo(pbef[o + 1], paft[o + 1], a.tmask[o + 1], up_e, do_e);
        bup_bdo(pbef[o - jpi], paft[o - jpi], a.tmask[o - jpi], up_s, do_s);
        bup_bdo(pbef[o + jpi], paft[o + jpi], a.tmask[o + jpi], up_n, do_n);
        const double zup = dmax(dmax(dmax(dmax(dmax(dmax(up_c, up_w), up_e), up_s), up_n), up_m), up_p);
        const double zdo = dmin(dmin(dmin(dmin(dmin(dmin(do_c, do_w), do_e), do_s), do_n), do_m), do_p);
        const double paa_c = paa[o], paa_w = paa[o - 1], pbb_c = pbb[o], pbb_s = pbb[o - jpi];
        const double pcc_kp1 = pcc[o + jpij];
        const double zpos = dmax(0., paa_w) - dmin(0., paa_c) + dmax(0., pbb_s) - dmin(0., pbb_c)
                          + dmax(0., pcc_kp1) - dmin(0., pcc_k);
        const double zneg = dmax(0., paa_c) - dmin(0., paa_w) + dmax(0., pbb_c) - dmin(0., pbb_s)
                          + dmax(0., pcc_k) - dmin(0., pcc_kp1);
        const double zbt = e12 * a.e3t_n[o] / p2dt;
        const double aft = paft[o];
        zbetup[o] = (zup - aft) / (zpos + zrtrn) * zbt;
        zbetdo[o] = (aft - zdo) / (zneg + zrtrn) * zbt;
        up_m = up_c; do_m = do_c; up_c = up_p; do_c = do_p; pcc_k = pcc_kp1;
    }
}

// limiter coefficient of one flux: zcu*zau + (1-zcu)*zbu with zcu = 0.5 + SIGN(0.5, flux)   (:407-410)
__device__ __forceinline__ double limit_coef(double flux, double bdo_here, double bup_next, double bup_here, double bdo_next)
{
    const double zau = dmin(dmin(1.0, bdo_here), bup_next);
    const double zbu = dmin(dmin(1.0, bup_here), bdo_next);
    const double zcu = 0.5 + ((flux >= 0.0) ? 0.5 : -0.5);
    return zcu * zau + (1.0 - zcu) * zbu;
}

// ------------------------------------------------------------------------------------------------------------
// nonosc P7: monotonic fluxes, in place (traadv_fct.F90:404-425)
// ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads) k_fct_limit(const FctArgs a)
{
    int ji, jj, ka, kb;
    if (!region_column(a.reg, ji, jj)) return;
    k_chunk(a.jpk - 1, a.nkchunk, ka, kb);
    if (ka > kb) return;                                 // an empty trailing chunk (nkchunk does not divide jpk-1 evenly)
    const size_t toff = (size_t)blockIdx.z * a.n3;
    const double *paa = a.zwx + toff, *pbb = a.zwy + toff, *pcc = a.zwz + toff;
    // in place (reference structure) or into separate arrays (schedule 1: the inner kernels still read zwx/zwy/zwz)
    double *oaa = (a.zlx ? a.zlx : a.zwx) + toff, *obb = (a.zly ? a.zly : a.zwy) + toff, *occ = (a.zlz ? a.zlz : a.zwz) + toff;
    const double *zbetup = a.zbetup + toff, *zbetdo = a.zbetdo + toff;
    const int jpi = a.jpi;
    const size_t jpij = a.jpij;
    const size_t c2 = (size_t)(jj - 1) * jpi + (ji - 1);
    double bup_c = zbetup[c2 + (size_t)(ka - 1) * jpij], bdo_c = zbetdo[c2 + (size_t)(ka - 1) * jpij];
    if (ka == 1 && a.zlz) occ[c2] = pcc[c2];                            // pcc(:,:,1) is never limited
    for (int k = ka; k <= kb; ++k) {
        const size_t o = c2 + (size_t)(k - 1) * jpij;
        const double bup_e = zbetup[o + 1], bdo_e = zbetdo[o + 1], bup_n = zbetup[o + jpi], bdo_n = zbetdo[o + jpi];
        const double bup_p = zbetup[o + jpij], bdo_p = zbetdo[o + jpij];
        const double fa = paa[o], fb = pbb[o], fc = pcc[o + jpij];
        oaa[o] = fa * limit_coef(fa, bdo_c, bup_e, bup_c, bdo_e);
        obb[o] = fb * limit_coef(fb, bdo_c, bup_n, bup_c, bdo_n);
        // k direction (:419-422): za = MIN(1, zbetdo(jk+1), zbetup(jk)), zb = MIN(1, zbetup(jk+1), zbetdo(jk))
        occ[o + jpij] = fc * limit_coef(fc, bdo_p, bup_c, bup_p, bdo_c);
        bup_c = bup_p; bdo_c = bdo_p;
    }
}

// ------------------------------------------------------------------------------------------------------------
// P8: final trend with corrected fluxes (traadv_fct.F90:288-297)
// ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads) k_fct_final(const FctArgs a)
{
    int ji, jj, ka, kb;
    if (!region_column(a.reg, ji, jj)) return;
    k_chunk(a.jpk - 1, a.nkchunk, ka, kb);
    if (ka > kb) return;                                 // an empty trailing chunk (nkchunk does not divide jpk-1 evenly)
    const size_t toff = (size_t)blockIdx.z * a.n3;
    const double *zwx = (a.zlx ? a.zlx : a.zwx) + toff, *zwy = (a.zly ? a.zly : a.zwy) + toff, *zwz = (a.zlz ? a.zlz : a.zwz) + toff;
    double *pta = a.pta + toff;
    const int jpi = a.jpi;
    const size_t jpij = a.jpij;
    const size_t c2 = (size_t)(jj - 1) * jpi + (ji - 1);
    const double r1 = a.r1_e1e2t[c2];
    double fz_k = zwz[c2 + (size_t)(ka - 1) * jpij];
    for (int k = ka; k <= kb; ++k) {
        const size_t o = c2 + (size_t)(k - 1) * jpij;
        const double fz_kp1 = zwz[o + jpij];
        pta[o] = pta[o] - (zwx[o] - zwx[o - 1] + zwy[o] - zwy[o - jpi] + fz_k - fz_kp1) * r1 / a.e3t_n[o];
        fz_k = fz_kp1;
    }
}

// ------------------------------------------------------------------------------------------------------------
// interp_4th_cpt (traadv_fct.F90:517-616).  The matrix depends on masks only: pivots zwt are computed once.
// Row coefficients on the fly: rows ikt = mikt+1 and ikb = mbkt are (d,i,s) = (1,0,0) with 2nd-order RHS;
// rows 3..jpkm1 otherwise (3*wmask+1, wmask, wmask); row 2 otherwise the ln_isfcav preset (1,0,0), RHS 0.
// ------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void cpt_row(int k, int ikt, int ikb, double wm, double &d, double &s)
{
    if (k == ikt || k == ikb || k == 2) { d = 1.0; s = 0.0; }
    else { d = 3.0 * wm + 1.0; s = wm; }
}

__global__ void __launch_bounds__(kThreads) k_cpt_pivots(int jpi, int jpj, int jpk, const double *wmask,
                                                         const int *mikt, const int *mbkt, double *zwt)
{
    int ji, jj;
    if (!interior_column(jpi, jpj, ji, jj)) return;
    const size_t jpij = (size_t)jpi * jpj, c2 = (size_t)(jj - 1) * jpi + (ji - 1);
    const int ikt = mikt[c2] + 1, ikb = mbkt[c2];
    double d, s, t_m, s_m;
    cpt_row(2, ikt, ikb, wmask[c2 + jpij], d, s);
    t_m = d; s_m = s;
    zwt[c2 + jpij] = t_m;                                               // zwt(2) = zwd(2)  (:579)
    for (int k = 3; k <= jpk - 1; ++k) {
        const size_t o = c2 + (size_t)(k - 1) * jpij;
        cpt_row(k, ikt, ikb, wmask[o], d, s);                           // zwi(k) == zws(k) on every row
        const double t = d - s * s_m / t_m;                             // zwd - zwi*zws(k-1)/zwt(k-1)  (:585)
        zwt[o] = t; t_m = t; s_m = s;
    }
}

// "Simple" columns (no ice-shelf cavity, wet from level 1 to mbkt: every column of tests/BENCH and most of a real ocean)
// have pivots and masks that follow from mbkt alone: zwt(k) = U(k) for k < mbkt and 1 below, with U the pivot sequence
// of a full-depth column; wmask(k) = 1 down to mbkt.  k_cpt_classify verifies this per column ONCE against the arrays
// (bitwise) and the solver then streams only ptn in and ztw out for them.
__global__ void __launch_bounds__(kThreads) k_cpt_classify(int jpi, int jpj, int jpk, const double *wmask, const int *mikt, const int *mbkt,
                                                           const double *zwt, const double *utab, unsigned char *simple)
{
    int ji, jj;
    if (!interior_column(jpi, jpj, ji, jj)) return;
    const size_t jpij = (size_t)jpi * jpj, c2 = (size_t)(jj - 1) * jpi + (ji - 1);
    const int ikb = mbkt[c2];
    bool ok = mikt[c2] == 1;
    for (int k = 2; k <= jpk - 1 && ok; ++k) {
        const size_t o = c2 + (size_t)(k - 1) * jpij;
        const double zw = (k < ikb) ? utab[k] : 1.0, wm = (k <= ikb) ? 1.0 : 0.0;
        ok = (zwt[o] == zw) && (wmask[o] == wm);
    }
    simple[c2] = ok ? 1 : 0;
}

// Thomas solve, one thread per column.  Each sweep is strictly sequential in jk; the loads of the next two levels are
// issued before the current level's dependent arithmetic (division chain) so that they overlap it.
__global__ void __launch_bounds__(kThreads) k_interp_4th_cpt(int jpi, int jpj, int jpk, const double *__restrict__ wmask,
                                                             const int *__restrict__ mikt, const int *__restrict__ mbkt,
                                                             const double *__restrict__ zwt, const unsigned char *__restrict__ simple,
                                                             const double *__restrict__ utab, const double *__restrict__ pt_in_all,
                                                             double *pt_out_all, const Region reg, double *scratch_all)
{
    int ji, jj;
    if (reg.n > 0 ? !region_column(reg, ji, jj) : !interior_column(jpi, jpj, ji, jj)) return;   // reg: only these columns (frame bands)
    const size_t jpij = (size_t)jpi * jpj, n3 = jpij * jpk, c2 = (size_t)(jj - 1) * jpi + (ji - 1);
    const double *__restrict__ pt_in = pt_in_all + (size_t)blockIdx.z * n3;
    double *pt_out = pt_out_all + (size_t)blockIdx.z * n3;
    // the forward sweep is parked in `fw`: pt_out itself, or a scratch array when another stream may read pt_out meanwhile
    double *fw = scratch_all ? scratch_all + (size_t)blockIdx.z * n3 : pt_out;
    const int ikt = mikt[c2] + 1, ikb = mbkt[c2];
    const int jpkm1 = jpk - 1;
    const bool smp = simple && simple[c2] != 0;
    auto pivot = [&](int k) -> double { return smp ? ((k < ikb) ? utab[k] : 1.0) : zwt[c2 + (size_t)(k - 1) * jpij]; };
    auto wmsk = [&](int k) -> double { return smp ? ((k <= ikb) ? 1.0 : 0.0) : wmask[c2 + (size_t)(k - 1) * jpij]; };
    auto clampk = [&](int k) -> int { return k < 1 ? 1 : (k > jpk ? jpk : k); };
    // forward sweep: pt_out(k) = zwrm(k) - zwi(k)/zwt(k-1)*pt_out(k-1)   (:590-601)
    double t_km1 = pt_in[c2], z_m = 0.0, zwt_m = 1.0;
    double t_a = pt_in[c2 + (size_t)(clampk(2) - 1) * jpij], t_b = pt_in[c2 + (size_t)(clampk(3) - 1) * jpij];
    double p_a = pivot(2), p_b = pivot(clampk(3)), w_a = wmsk(2), w_b = wmsk(clampk(3));
    for (int k = 2; k <= jpkm1; ++k) {
        const double t_k = t_a, zw = p_a, wm = w_a;
        t_a = t_b; p_a = p_b; w_a = w_b;
        { const int kn = clampk(k + 2); t_b = pt_in[c2 + (size_t)(kn - 1) * jpij]; p_b = pivot(kn); w_b = wmsk(kn); }
        double rhs, wi;
        if (k == ikt || k == ikb) { rhs = 0.5 * (t_km1 + t_k); wi = 0.0; }        // (:563-571)
        else if (k == 2)          { rhs = 0.0; wi = 0.0; }                         // ln_isfcav preset (:554-556)
        else                      { rhs = 3.0 * wm * (t_k + t_km1); wi = wm; }     // (:538-542)
        double z = rhs;
        if (k >= 3) z = rhs - wi / zwt_m * z_m;
        fw[c2 + (size_t)(k - 1) * jpij] = z;
        z_m = z; zwt_m = zw; t_km1 = t_k;
    }
    // back substitution (:603-614).  NB pt_out is read back here: the loads below are of levels this thread wrote.
    {
        double x = z_m / zwt_m;                                                    // level jpkm1: still in registers
        pt_out[c2 + (size_t)(jpkm1 - 1) * jpij] = x;
        const double *rd = fw;
        double z_a = rd[c2 + (size_t)(clampk(jpk - 2) - 1) * jpij], z_b = rd[c2 + (size_t)(clampk(jpk - 3) - 1) * jpij];
        p_a = pivot(clampk(jpk - 2)); p_b = pivot(clampk(jpk - 3)); w_a = wmsk(clampk(jpk - 2)); w_b = wmsk(clampk(jpk - 3));
        for (int k = jpk - 2; k >= 2; --k) {
            const double z = z_a, zw = p_a, wm = w_a;
            z_a = z_b; p_a = p_b; w_a = w_b;
            { const int kn = clampk(k - 2); z_b = rd[c2 + (size_t)(kn - 1) * jpij]; p_b = pivot(kn); w_b = wmsk(kn); }
            double d, s;
            cpt_row(k, ikt, ikb, wm, d, s);
            x = (z - s * x) / zw;
            pt_out[c2 + (size_t)(k - 1) * jpij] = x;
        }
    }
}

// ------------------------------------------------------------------------------------------------------------
// tra_adv: effective transports (traadv.F90:100-124), Eulerian branch
// ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_transports(int jpk, size_t jpij, const double *e2u, const double *e1v,
                                                    const double *e1e2t, const double *e3u_n, const double *e3v_n,
                                                    const double *un, const double *vn, const double *wn,
                                                    double *zun, double *zvn, double *zwn)
{
    const size_t n3 = jpij * jpk;
    for (size_t m = (size_t)blockIdx.x * blockDim.x + threadIdx.x; m < n3; m += (size_t)gridDim.x * blockDim.x) {
        const size_t k = m / jpij, c = m - k * jpij;
        if ((int)k == jpk - 1) { zun[m] = 0.0; zvn[m] = 0.0; zwn[m] = 0.0; }       // no transport through the bottom
        else {
            zun[m] = e2u[c] * e3u_n[m] * un[m];
            zvn[m] = e1v[c] * e3v_n[m] * vn[m];
            zwn[m] = e1e2t[c] * wn[m];
        }
    }
}

// trend-diagnostic hook: ztrd = upstream flux + limited anti-diffusive flux (traadv_fct.F90:172-176, 299-303)
__global__ void __launch_bounds__(kThreads) k_fct_diag(const FctArgs a, double *trdx, double *trdy, double *trdz)
{
    const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= (long long)a.jpij) return;
    const int jj = (int)(p / a.jpi) + 1, ji = (int)(p % a.jpi) + 1;
    const size_t toff = (size_t)blockIdx.z * a.n3;
    const double *ptb = a.ptb + toff, *zwx = a.zwx + toff, *zwy = a.zwy + toff, *zwz = a.zwz + toff;
    const size_t c2 = (size_t)p;
    const int mik = a.ln_isfcav ? a.mikt[c2] : 1;
    const bool hm = ji <= a.jpi - 1 && jj <= a.jpj - 1;                  // the h- range of P1 (:123-135)
    for (int k = 1; k <= a.jpk; ++k) {
        const size_t o = c2 + (size_t)(k - 1) * a.jpij;
        trdz[toff + o] = upstream_w(a, ptb, c2, k, mik) + zwz[o];
        if (hm) {
            double ux = 0.0, uy = 0.0;                                   // zwx(:,:,jpk) = zwy(:,:,jpk) = 0 (:115)
            if (k <= a.jpk - 1) {
                const double u = a.pun[o], v = a.pvn[o], t = ptb[o];
                ux = 0.5 * ((u + fabs(u)) * t + (u - fabs(u)) * ptb[o + 1]);
                uy = 0.5 * ((v + fabs(v)) * t + (v - fabs(v)) * ptb[o + a.jpi]);
            }
            trdx[toff + o] = ux + zwx[o];
            trdy[toff + o] = uy + zwy[o];
        }
    }
}
