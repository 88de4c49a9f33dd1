// fct_fused_kernel.cuh -- the WHOLE FCT step of the exchange-free inner region in one kernel (k_fct_fused):
// upstream fluxes, low-order update, anti-diffusive fluxes (P1-P5, traadv_fct.F90:123-278), the Zalesak limiter
// (nonosc P6 + P7, :356-425) and the final trend (P8, :288-297).  The intermediate arrays zwi / zwx / zwy / zwz / zbetup /
// zbetdo of the reference never exist: per tracer-point the kernel reads ptb, ptn, pta [, ztw] and writes pta.
//
// Included by fct_kernels.cu INSIDE namespace nemo's anonymous namespace, after nonosc_final.cuh (dmax / dmin / bup_bdo /
// limit_coef_sel) and after the includer has defined mbar_init, mbar_init_fence, mbar_expect_tx, mbar_wait and tma_load_3d
// (inline PTX there; an atomic phase counter and a box copy with zero fill in the host emulation of tests/emu, which
// compiles this very source to compare it with the oracle without a GPU).  Nothing else includes it.
//
// Geometry.  A block owns an extended tile of FX x FY columns = its (FX-4) x (FY-4) output columns + a halo of 2 (the
// dependency radius of one step is 3: the third cell is only READ, from the TMA boxes).  One thread per column marches down
// jk.  Iteration `it` runs three independent stages on three different levels, software-pipelined through registers:
//   A  level a = it   (all threads)   P1-P5: zwi, fx, fy, fz and zbup / zbdo; publishes zbup, zbdo, fx, fy in shared memory
//   B  level b = it-1 (halo-1 ring)   betas from the neighbours' zbup / zbdo / fluxes published one iteration earlier
//   C  level c = it-2 (output cols)   limits the six fluxes of the cell with the betas published one iteration earlier,
//                                     applies the final divergence and stores pta
// so ONE barrier per level orders every shared-memory reuse (zbup/zbdo and the betas are double-buffered by level parity,
// fx/fy triple-buffered).  Inputs arrive as TMA boxes (cp.async.bulk.tensor.3d, one elected thread) in a 3-stage ring.
// Box origins must be >= 0 and even in ji (fp64: 16-byte aligned start, measured on B200): the extended tile starts at an
// even 0-based column (out.i0 odd), the halo boxes two columns to its west.
#pragma once

constexpr int FX = 32, FY = 16, FHALO = 2, FOX = FX - 2 * FHALO, FOY = FY - 2 * FHALO;
constexpr int FBW = FX + 4, FBH = FY + 3;             // halo box: (pad, 1) west, 2 east; 1 south, 2 north
constexpr int FSTAGES = 3;
enum { FH_PTB = 0, FH_PTN, FH_TM, FH_PUN, FH_PVN, FH_COUNT };                    // boxes with halo
enum { FP_PTA = 0, FP_ZTW, FP_PWN, FP_E3B, FP_E3N, FP_E3A, FP_COUNT };           // boxes of the extended tile itself
constexpr int kFHaloBytes = ((FBW * FBH * 8 + 127) / 128) * 128, kFPlainBytes = FX * FY * 8;
constexpr int kFStageBytes = FH_COUNT * kFHaloBytes + FP_COUNT * kFPlainBytes;
constexpr int kFPlane = FX * FY;                      // one published plane (doubles)
constexpr int kFPlanes = 2 + 2 + 3 + 3 + 2 + 2;       // zbup, zbdo (x2), fx, fy (x3), zbetup, zbetdo (x2)
constexpr size_t kFusedSmemBytes = (size_t)FSTAGES * kFStageBytes + (size_t)kFPlanes * kFPlane * 8 + 64;

struct FusedMaps { CUtensorMap h[FH_COUNT]; CUtensorMap p[FP_COUNT]; };

template <int H, int V>
__global__ void __launch_bounds__(FX * FY, 1) k_fct_fused(const FctArgs a, const __grid_constant__ FusedMaps maps)
{
    NEMO_DYN_SMEM_ALIGNED(unsigned char, fu_smem, 128);
    double *planes = reinterpret_cast<double *>(fu_smem + (size_t)FSTAGES * kFStageBytes);
    double *sU = planes, *sD = sU + 2 * kFPlane, *sFx = sD + 2 * kFPlane, *sFy = sFx + 3 * kFPlane;
    double *sBu = sFy + 3 * kFPlane, *sBd = sBu + 2 * kFPlane;
    unsigned long long *full = reinterpret_cast<unsigned long long *>(sBd + 2 * kFPlane);
    const int tx = threadIdx.x % FX, ty = threadIdx.x / FX;
    const int jn = (int)blockIdx.x % a.kjpt, chunk = (int)blockIdx.z;   // tracer index fastest: the shared boxes of a tile hit L2
    const int X0 = a.out.i0 - 1 - FHALO + ((int)blockIdx.x / a.kjpt) * FOX;      // 0-based column / row of thread (0,0)
    const int Y0 = a.out.j0 - 1 - FHALO + (int)blockIdx.y * FOY;
    const int jpi = a.jpi, jpk = a.jpk;
    const size_t jpij = a.jpij;
    int ka, kb;
    { const int per = (jpk - 1 + a.nkchunk - 1) / a.nkchunk; ka = 1 + chunk * per; kb = min(jpk - 1, ka + per - 1); }
    if (ka > kb) return;
    const int a_lo = max(1, ka - 2), a_hi = min(jpk, kb + 2);          // levels of stage A
    const int b_lo = max(1, ka - 1), b_hi = min(jpk - 1, kb + 1);      // levels of stage B
    const int lastlev = min(jpk, a_hi + 1);                             // last level whose boxes are loaded
    const int gi = X0 + tx + 1, gj = Y0 + ty + 1;                       // 1-based column of this thread
    const bool is_out = tx >= FHALO && tx < FX - FHALO && ty >= FHALO && ty < FY - FHALO && gi <= a.out.i1 && gj <= a.out.j1;
    const bool is_beta = tx >= 1 && tx < FX - 1 && ty >= 1 && ty < FY - 1;
    const int ci = min(gi, jpi), cj = min(gj, a.jpj);                   // tiles overhang the rectangle: keep addresses legal
    const size_t toff = (size_t)jn * a.n3;
    const size_t c2 = (size_t)(cj - 1) * jpi + (ci - 1);
    double *__restrict__ pta = a.pta + toff;
    const double r1 = a.r1_e1e2t[c2], e12 = a.e1e2t[c2];
    const int ktop = a.ln_linssh ? (a.ln_isfcav ? a.mikt[c2] : 1) : 0; // level whose top flux is pwn*ptb (:146-156)
    const double p2dt = a.p2dt;
    const double r1_6 = 1.0 / 6.0, zrtrn = 1.e-15;

    const CUtensorMap *mh = maps.h, *mp = maps.p;                       // descriptor addresses stay in the parameter space
    auto issue = [=](int lev) {                                         // one thread: all boxes of level `lev`
        unsigned char *st = fu_smem + (size_t)(lev % FSTAGES) * kFStageBytes;
        unsigned long long *bar = &full[lev % FSTAGES];
        mbar_expect_tx(bar, FH_COUNT * (FBW * FBH * 8) + (FP_COUNT - (V == 4 ? 0 : 1)) * kFPlainBytes);
        const int z3 = lev - 1, z4 = jn * jpk + lev - 1;
        tma_load_3d(st + FH_PTB * kFHaloBytes, &mh[FH_PTB], bar, X0 - 2, Y0 - 1, z4);
        tma_load_3d(st + FH_PTN * kFHaloBytes, &mh[FH_PTN], bar, X0 - 2, Y0 - 1, z4);
        tma_load_3d(st + FH_TM * kFHaloBytes, &mh[FH_TM], bar, X0 - 2, Y0 - 1, z3);
        tma_load_3d(st + FH_PUN * kFHaloBytes, &mh[FH_PUN], bar, X0 - 2, Y0 - 1, z3);
        tma_load_3d(st + FH_PVN * kFHaloBytes, &mh[FH_PVN], bar, X0 - 2, Y0 - 1, z3);
        unsigned char *pl = st + FH_COUNT * kFHaloBytes;
        tma_load_3d(pl + FP_PTA * kFPlainBytes, &mp[FP_PTA], bar, X0, Y0, z4);
        if (V == 4) tma_load_3d(pl + FP_ZTW * kFPlainBytes, &mp[FP_ZTW], bar, X0, Y0, z4);
        tma_load_3d(pl + FP_PWN * kFPlainBytes, &mp[FP_PWN], bar, X0, Y0, z3);
        tma_load_3d(pl + FP_E3B * kFPlainBytes, &mp[FP_E3B], bar, X0, Y0, z3);
        tma_load_3d(pl + FP_E3N * kFPlainBytes, &mp[FP_E3N], bar, X0, Y0, z3);
        tma_load_3d(pl + FP_E3A * kFPlainBytes, &mp[FP_E3A], bar, X0, Y0, z3);
    };
    auto upw = [&](int k, double w, double tb_k, double tb_km1, double wm) -> double {      // P2 + P2b (:137-156)
        double v = 0.0;
        if (k >= 2 && k <= jpk - 1) {
            const double zfp_wk = w + fabs(w), zfm_wk = w - fabs(w);
            v = 0.5 * (zfp_wk * tb_k + zfm_wk * tb_km1) * wm;
        }
        if (k == ktop) v = w * tb_k;
        return v;
    };

    if (threadIdx.x == 0) {
        for (int s = 0; s < FSTAGES; ++s) mbar_init(&full[s], 1);
        mbar_init_fence();
        issue(a_lo);
        if (a_lo + 1 <= lastlev) issue(a_lo + 1);
    }
    // column registers carried from level to level
    double tb_m = 0.0, tn_m = 0.0, tm_m = 0.0;          // ptb, ptn, tmask of level a-1
    if (a_lo >= 2) { const size_t om = c2 + (size_t)(a_lo - 2) * jpij; tb_m = a.ptb[toff + om]; tn_m = a.ptn[toff + om]; tm_m = a.tmask[om]; }
    double upz_k = 0.0;                                  // upstream vertical flux through the top of level a
    bool first = true;
    double up_b = 0.0, do_b = 0.0, up_bm = 0.0, do_bm = 0.0;            // zbup / zbdo at levels b, b-1
    double aft_b = 0.0, e3n_b = 1.0, e3n_c = 1.0, pta_b = 0.0, pta_c = 0.0;
    double fx_b = 0.0, fy_b = 0.0, fz_b = 0.0, fx_c = 0.0, fy_c = 0.0, fz_c = 0.0;   // anti-diffusive fluxes at levels b, c
    double bup_c = 0.0, bdo_c = 0.0, bup_cm = 0.0, bdo_cm = 0.0;       // betas at levels c, c-1
    const int hc = (ty + 1) * FBW + (tx + 2), pc = ty * FX + tx;
    const int cell = ty * FX + tx;

    for (int it = a_lo; it <= kb + 2; ++it) {
        __syncthreads();                                 // every thread is done with iteration it-1: its stage and planes are free
        if (threadIdx.x == 0 && it + 2 <= lastlev) issue(it + 2);
        const int lev_a = it, lev_b = it - 1, lev_c = it - 2;

        // ---- stage A: level a ----------------------------------------------------------------------------------
        double up_a = 0.0, do_a = 0.0, aft_a = 0.0, e3n_a = 1.0, pta_a = 0.0, fx_a = 0.0, fy_a = 0.0, fz_a = 0.0;
        if (lev_a <= a_hi) {
            mbar_wait(&full[lev_a % FSTAGES], ((lev_a - a_lo) / FSTAGES) & 1);
            const double *sh = reinterpret_cast<const double *>(fu_smem + (size_t)(lev_a % FSTAGES) * kFStageBytes);
            const double *s_tb = sh + FH_PTB * (kFHaloBytes / 8), *s_tm = sh + FH_TM * (kFHaloBytes / 8);
            const double tb_c = s_tb[hc], tm = s_tm[hc];
            if (lev_a <= jpk - 1) {
                mbar_wait(&full[(lev_a + 1) % FSTAGES], ((lev_a + 1 - a_lo) / FSTAGES) & 1);
                const double *sh1 = reinterpret_cast<const double *>(fu_smem + (size_t)((lev_a + 1) % FSTAGES) * kFStageBytes);
                const double *sp = sh + FH_COUNT * (kFHaloBytes / 8), *sp1 = sh1 + FH_COUNT * (kFHaloBytes / 8);
                const double *s_tn = sh + FH_PTN * (kFHaloBytes / 8), *s_u = sh + FH_PUN * (kFHaloBytes / 8), *s_v = sh + FH_PVN * (kFHaloBytes / 8);
                const double tb_w = s_tb[hc - 1], tb_e = s_tb[hc + 1], tb_s = s_tb[hc - FBW], tb_n = s_tb[hc + FBW];
                const double u_c = s_u[hc], u_w = s_u[hc - 1], v_c = s_v[hc], v_s = s_v[hc - FBW];
                const double tn_c = s_tn[hc], tn_e = s_tn[hc + 1], tn_n = s_tn[hc + FBW];
                const double w_c = sp[FP_PWN * kFPlane + pc];
                const double e3b = sp[FP_E3B * kFPlane + pc], e3n = sp[FP_E3N * kFPlane + pc], e3a = sp[FP_E3A * kFPlane + pc];
                const double tb_p = sh1[FH_PTB * (kFHaloBytes / 8) + hc], tm_p = sh1[FH_TM * (kFHaloBytes / 8) + hc];
                const double w_p = sp1[FP_PWN * kFPlane + pc];
                const double wm_c = (lev_a == 1) ? tm : tm * tm_m, wm_p = tm_p * tm;       // wmask (dommsk.F90:193)
                if (first) upz_k = upw(lev_a, w_c, tb_c, tb_m, wm_c);
                double zfp, zfm;
                zfp = u_c + fabs(u_c); zfm = u_c - fabs(u_c);
                const double upx_c = 0.5 * (zfp * tb_c + zfm * tb_e);
                zfp = u_w + fabs(u_w); zfm = u_w - fabs(u_w);
                const double upx_w = 0.5 * (zfp * tb_w + zfm * tb_c);
                zfp = v_c + fabs(v_c); zfm = v_c - fabs(v_c);
                const double upy_c = 0.5 * (zfp * tb_c + zfm * tb_n);
                zfp = v_s + fabs(v_s); zfm = v_s - fabs(v_s);
                const double upy_s = 0.5 * (zfp * tb_s + zfm * tb_c);
                const double upz_kp1 = upw(lev_a + 1, w_p, tb_p, tb_c, wm_p);
                const double ztra = -(upx_c - upx_w + upy_c - upy_s + upz_k - upz_kp1) * r1;
                if (is_out) pta_a = sp[FP_PTA * kFPlane + pc] + ztra / e3n * tm;           // first half of the trend (:165)
                aft_a = (e3b * tb_c + p2dt * ztra) / e3a * tm;                            // zwi (:167)
                if (H == 2) {
                    fx_a = 0.5 * u_c * (tn_c + tn_e) - upx_c;
                    fy_a = 0.5 * v_c * (tn_c + tn_n) - upy_c;
                } else {
                    const double tn_w = s_tn[hc - 1], tn_ee = s_tn[hc + 2], tn_s = s_tn[hc - FBW], tn_nn = s_tn[hc + 2 * FBW];
                    const double tm_w = s_tm[hc - 1], tm_e = s_tm[hc + 1], tm_ee = s_tm[hc + 2];
                    const double tm_s = s_tm[hc - FBW], tm_n = s_tm[hc + FBW], tm_nn = s_tm[hc + 2 * FBW];
                    const double mu_w = tm_w * tm, mu_c = tm * tm_e, mu_e = tm_e * tm_ee;   // umask, vmask (dommsk.F90:176-177)
                    const double mv_s = tm_s * tm, mv_c = tm * tm_n, mv_n = tm_n * tm_nn;
                    const double ztu_w = (tn_c - tn_w) * mu_w, ztu_c = (tn_e - tn_c) * mu_c, ztu_e = (tn_ee - tn_e) * mu_e;
                    const double ztv_s = (tn_c - tn_s) * mv_s, ztv_c = (tn_n - tn_c) * mv_c, ztv_n = (tn_nn - tn_n) * mv_n;
                    const double zltu_c = (ztu_c + ztu_w) * r1_6, zltu_e = (ztu_e + ztu_c) * r1_6;
                    const double zltv_c = (ztv_c + ztv_s) * r1_6, zltv_n = (ztv_n + ztv_c) * r1_6;
                    const double zC2t_u = tn_c + tn_e, zC2t_v = tn_c + tn_n;
                    fx_a = 0.5 * u_c * (zC2t_u + zltu_c - zltu_e) - upx_c;
                    fy_a = 0.5 * v_c * (zC2t_v + zltv_c - zltv_n) - upy_c;
                }
                if (lev_a >= 2) {
                    if (V == 2) fz_a = (w_c * 0.5 * (tn_c + tn_m) - upz_k) * wm_c;
                    else        fz_a = (w_c * sp[FP_ZTW * kFPlane + pc] - upz_k) * wm_c;
                }
                e3n_a = e3n;
                upz_k = upz_kp1; tn_m = tn_c;
            }
            first = false;
            bup_bdo(tb_c, aft_a, tm, up_a, do_a);                                          // zbup, zbdo (:361-364); zwi(jpk) = 0
            sU[(lev_a & 1) * kFPlane + cell] = up_a; sD[(lev_a & 1) * kFPlane + cell] = do_a;
            sFx[(lev_a % 3) * kFPlane + cell] = fx_a; sFy[(lev_a % 3) * kFPlane + cell] = fy_a;
            tb_m = tb_c; tm_m = tm;
        }

        // ---- stage B: betas of level b (:366-399) ----------------------------------------------------------------
        double bup_b = 0.0, bdo_b = 0.0;                                                  // zbetup(jpk) = zbetdo(jpk) = 0 (:356)
        if (lev_b >= b_lo && lev_b <= b_hi && is_beta) {
            const double *U = sU + (lev_b & 1) * kFPlane + cell, *D = sD + (lev_b & 1) * kFPlane + cell;
            const double up_m1 = (lev_b == 1) ? up_b : up_bm, do_m1 = (lev_b == 1) ? do_b : do_bm;   // ikm1 = MAX(jk-1,1)
            const double zup = dmax(dmax(dmax(dmax(dmax(dmax(up_b, U[-1]), U[1]), U[-FX]), U[FX]), up_m1), up_a);
            const double zdo = dmin(dmin(dmin(dmin(dmin(dmin(do_b, D[-1]), D[1]), D[-FX]), D[FX]), do_m1), do_a);
            const double paa_w = sFx[(lev_b % 3) * kFPlane + cell - 1], pbb_s = sFy[(lev_b % 3) * kFPlane + cell - FX];
            const double zpos = dmax(0., paa_w) - dmin(0., fx_b) + dmax(0., pbb_s) - dmin(0., fy_b)
                              + dmax(0., fz_a) - dmin(0., fz_b);
            const double zneg = dmax(0., fx_b) - dmin(0., paa_w) + dmax(0., fy_b) - dmin(0., pbb_s)
                              + dmax(0., fz_b) - dmin(0., fz_a);
            const double zbt = e12 * e3n_b / p2dt;
            bup_b = (zup - aft_b) / (zpos + zrtrn) * zbt;
            bdo_b = (aft_b - zdo) / (zneg + zrtrn) * zbt;
        }
        if (lev_b >= 1) { sBu[(lev_b & 1) * kFPlane + cell] = bup_b; sBd[(lev_b & 1) * kFPlane + cell] = bdo_b; }

        // ---- stage C: limited fluxes and final trend of level c (:404-425, :288-297) -------------------------------
        if (lev_c >= ka && is_out) {
            const double *Bu = sBu + (lev_c & 1) * kFPlane + cell, *Bd = sBd + (lev_c & 1) * kFPlane + cell;
            const double bup_e = Bu[1], bdo_e = Bd[1], bup_w = Bu[-1], bdo_w = Bd[-1];
            const double bup_n = Bu[FX], bdo_n = Bd[FX], bup_s = Bu[-FX], bdo_s = Bd[-FX];
            const double paa_w = sFx[(lev_c % 3) * kFPlane + cell - 1], pbb_s = sFy[(lev_c % 3) * kFPlane + cell - FX];
            const double lx_e = fx_c * limit_coef_sel(fx_c, bdo_c, bup_e, bup_c, bdo_e);
            const double lx_w = paa_w * limit_coef_sel(paa_w, bdo_w, bup_c, bup_w, bdo_c);
            const double ly_n = fy_c * limit_coef_sel(fy_c, bdo_c, bup_n, bup_c, bdo_n);
            const double ly_s = pbb_s * limit_coef_sel(pbb_s, bdo_s, bup_c, bup_s, bdo_c);
            // pcc(jk+1) is limited with betas(jk), betas(jk+1) (:419-422); pcc(:,:,1) is never limited
            const double lz_t = (lev_c == 1) ? fz_c : fz_c * limit_coef_sel(fz_c, bdo_c, bup_cm, bup_c, bdo_cm);
            const double lz_b = fz_b * limit_coef_sel(fz_b, bdo_b, bup_c, bup_b, bdo_c);
            pta[c2 + (size_t)(lev_c - 1) * jpij] = pta_c - (lx_e - lx_w + ly_n - ly_s + lz_t - lz_b) * r1 / e3n_c;
        }

        // rotate the column registers: b -> c, a -> b
        bup_cm = bup_c; bdo_cm = bdo_c; bup_c = bup_b; bdo_c = bdo_b;
        fx_c = fx_b; fy_c = fy_b; fz_c = fz_b; fx_b = fx_a; fy_b = fy_a; fz_b = fz_a;
        e3n_c = e3n_b; e3n_b = e3n_a; pta_c = pta_b; pta_b = pta_a;
        up_bm = up_b; do_bm = do_b; up_b = up_a; do_b = do_a; aft_b = aft_a;
    }
}
