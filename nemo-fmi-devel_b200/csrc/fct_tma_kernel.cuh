// fct_tma_kernel.cuh -- P1-P5 of the FCT step on the exchange-free inner region with TMA-staged tiles
// (k_fct_low_antidiff_tma): the default data movement of the step's DRAM-bound kernel.
//
// Included by fct_kernels.cu INSIDE namespace nemo's anonymous namespace, after fct_inner_kernel.cuh (Masks) and after the
// includer has defined smem_u32, mbar_init, mbar_init_fence, mbar_expect_tx, mbar_wait and tma_load_3d (inline PTX there:
// mbarrier.*, cp.async.bulk.tensor.3d; in the host emulation of tests/emu an atomic phase counter and a box copy with
// zero fill, so that the very same kernel source -- tile origins, box coordinates, shared-memory indexing, ring reuse -- is
// compared with the oracle at any subdomain size without a GPU).  Nothing else includes it.
#pragma once

#ifndef NEMO_DYN_SMEM_ALIGNED
#define NEMO_DYN_SMEM_ALIGNED(type, name, alignment) extern __shared__ __align__(alignment) type name[]
#endif

// ---- TMA-tiled variant of the inner P1-P5 kernel ----------------------------------------------------------------
// Same arithmetic as k_fct_low_antidiff_inner, different data movement.  The cp.async variant still issues ~35
// 8-byte memory instructions per level and thread (ncu: LSU/MIO-throttle bound, L1 hit rate 20 %).  Here one elected
// thread issues ONE bulk tensor copy (TMA, cp.async.bulk.tensor.3d -> SASS UTMALDG) per array and level for the
// whole 32x8 tile -- with a 2-column/2-row halo for the arrays that are read at neighbours -- into a 3-stage
// shared-memory ring guarded by mbarriers; all 256 threads then read own and neighbour values from shared memory.
// Requires jpi even (TMA global strides are multiples of 16 bytes); otherwise the cp.async variant is used.
// Box origins are never negative (measured on B200: a negative tile coordinate raises an illegal-instruction error;
// overhang on the high side is zero-filled): the stencil needs 1 halo cell to the west/south and 2 to the east/north,
// so the 36x12 box starts at (tile origin - 1) >= 0.  The innermost box coordinate must also be EVEN for fp64 (16-byte
// aligned start address, measured the same way): tiles start at odd 0-based columns (i0 = 2), so every box -- also
// the ones without halo -- starts one column west of the tile.
constexpr int TTX = 32, TTY = 8, TTH = 2, TBW = TTX + 2 * TTH, TBH = TTY + 2 * TTH, TSTAGES = 3;
enum { TH_PTB = 0, TH_PTN, TH_TM, TH_PUN, TH_PVN, TH_COUNT };                    // boxes with halo
enum { TP_PTA = 0, TP_ZTW, TP_PWN, TP_E3B, TP_E3N, TP_E3A, TP_COUNT };           // plain boxes
constexpr int TPW = TTX + 2;                        // plain boxes carry one pad column west (+1 east to stay even)
constexpr int kTileHaloBytes = TBW * TBH * 8, kTilePlainBytes = TPW * TTY * 8;
constexpr int kTileStageBytes = TH_COUNT * kTileHaloBytes + TP_COUNT * kTilePlainBytes;

struct TileMaps { CUtensorMap h[TH_COUNT]; CUtensorMap p[TP_COUNT]; };

template <int H, int V, bool FROM_T>
__global__ void __launch_bounds__(TTX * TTY) k_fct_low_antidiff_tma(const FctArgs a, const __grid_constant__ TileMaps maps, Rect rc)
{
    NEMO_DYN_SMEM_ALIGNED(unsigned char, tile_smem, 128);
    unsigned long long *full = reinterpret_cast<unsigned long long *>(tile_smem + TSTAGES * kTileStageBytes);
    const int lx = threadIdx.x % TTX, ly = threadIdx.x / TTX;
    // the tracer index varies fastest over the grid: the blocks that share a tile's pun/pvn/pwn/e3t/tmask boxes are
    // scheduled together and the second tracer's copies hit L2 (ncu: 26 GB read vs 16 GB ideal when tracers were apart)
    const int jn = (int)blockIdx.x % a.kjpt, chunk = (int)blockIdx.z;
    const int ox = rc.i0 - 1 + ((int)blockIdx.x / a.kjpt) * TTX, oy = rc.j0 - 1 + (int)blockIdx.y * TTY;   // 0-based tile origin
    const int jpi = a.jpi, jpk = a.jpk;
    const size_t jpij = a.jpij;
    int ka, kb;
    { const int per = (jpk - 1 + a.nkchunk - 1) / a.nkchunk; ka = 1 + chunk * per; kb = min(jpk - 1, ka + per - 1); }
    if (ka > kb) return;                                 // an empty trailing chunk
    const int gi = ox + lx + 1, gj = oy + ly + 1;                                 // 1-based column of this thread
    const bool active = gi <= rc.i1 && gj <= rc.j1;
    const int ci = min(gi, jpi - 2), cj = min(gj, a.jpj - 2);                     // clamped: addresses of inactive threads stay legal
    const size_t toff = (size_t)jn * a.n3;
    const size_t c2 = (size_t)(cj - 1) * jpi + (ci - 1);
    double *__restrict__ pta = a.pta + toff;
    double *__restrict__ zwi = a.zwi + toff;
    double *__restrict__ zwx = a.zwx + toff;
    double *__restrict__ zwy = a.zwy + toff;
    double *__restrict__ zwz = a.zwz + toff;
    const double r1 = a.r1_e1e2t[c2];
    const int ktop = a.ln_linssh ? (a.ln_isfcav ? a.mikt[c2] : 1) : 0;
    const double p2dt = a.p2dt;
    const double r1_6 = 1.0 / 6.0;

    // descriptor addresses must stay in the kernel-parameter space: take them here, not through a by-reference closure
    const CUtensorMap *mh = maps.h, *mp = maps.p;
    auto stage = [&](int lev) -> unsigned char * { return tile_smem + (size_t)(lev % TSTAGES) * kTileStageBytes; };
    auto issue = [=](int lev) {                          // one thread: all boxes of level `lev`
        unsigned char *st = tile_smem + (size_t)(lev % TSTAGES) * kTileStageBytes;
        unsigned long long *bar = &full[lev % TSTAGES];
        const unsigned bytes = TH_COUNT * kTileHaloBytes + (TP_COUNT - (V == 4 ? 0 : 1)) * kTilePlainBytes;
        mbar_expect_tx(bar, bytes);
        const int z3 = lev - 1, z4 = jn * jpk + lev - 1;
        tma_load_3d(st + TH_PTB * kTileHaloBytes, &mh[TH_PTB], bar, ox - 1, oy - 1, z4);
        tma_load_3d(st + TH_PTN * kTileHaloBytes, &mh[TH_PTN], bar, ox - 1, oy - 1, z4);
        tma_load_3d(st + TH_TM * kTileHaloBytes, &mh[TH_TM], bar, ox - 1, oy - 1, z3);
        tma_load_3d(st + TH_PUN * kTileHaloBytes, &mh[TH_PUN], bar, ox - 1, oy - 1, z3);
        tma_load_3d(st + TH_PVN * kTileHaloBytes, &mh[TH_PVN], bar, ox - 1, oy - 1, z3);
        unsigned char *pl = st + TH_COUNT * kTileHaloBytes;
        tma_load_3d(pl + TP_PTA * kTilePlainBytes, &mp[TP_PTA], bar, ox - 1, oy, z4);
        if (V == 4) tma_load_3d(pl + TP_ZTW * kTilePlainBytes, &mp[TP_ZTW], bar, ox - 1, oy, z4);
        tma_load_3d(pl + TP_PWN * kTilePlainBytes, &mp[TP_PWN], bar, ox - 1, oy, z3);
        tma_load_3d(pl + TP_E3B * kTilePlainBytes, &mp[TP_E3B], bar, ox - 1, oy, z3);
        tma_load_3d(pl + TP_E3N * kTilePlainBytes, &mp[TP_E3N], bar, ox - 1, oy, z3);
        tma_load_3d(pl + TP_E3A * kTilePlainBytes, &mp[TP_E3A], bar, ox - 1, oy, z3);
    };
    auto upw = [&](int k, double w, double tb_k, double tb_km1, double wm) -> double {
        double v = 0.0;
        if (k >= 2 && k <= jpk - 1) {
            const double zfp_wk = w + fabs(w), zfm_wk = w - fabs(w);
            v = 0.5 * (zfp_wk * tb_k + zfm_wk * tb_km1) * wm;
        }
        if (k == ktop) v = w * tb_k;
        return v;
    };

    if (threadIdx.x == 0) {
        for (int s = 0; s < TSTAGES; ++s) mbar_init(&full[s], 1);
        mbar_init_fence();
        issue(ka);
        issue(ka + 1);                                   // ka + 1 <= jpk always
    }
    double tb_m = 0.0, tn_m = 0.0, tm_m = 0.0;
    if (ka >= 2) { const size_t om = c2 + (size_t)(ka - 2) * jpij; tb_m = a.ptb[toff + om]; tn_m = a.ptn[toff + om]; tm_m = a.tmask[om]; }
    double upz_k = 0.0;
    bool first = true;
    const int hc = (ly + 1) * TBW + (lx + 1), pc = ly * TPW + lx + 1;   // every box starts one cell west (halo boxes also one cell south) of the tile
    for (int k = ka; k <= kb; ++k) {
        __syncthreads();                                 // every thread is done with level k-1: its stage is free
        if (threadIdx.x == 0 && k + 2 <= kb + 1) issue(k + 2);
        mbar_wait(&full[k % TSTAGES], ((k - ka) / TSTAGES) & 1);
        mbar_wait(&full[(k + 1) % TSTAGES], ((k + 1 - ka) / TSTAGES) & 1);
        const double *sh = reinterpret_cast<const double *>(stage(k));
        const double *sp = sh + TH_COUNT * TBW * TBH;
        const double *sh1 = reinterpret_cast<const double *>(stage(k + 1));
        const double *sp1 = sh1 + TH_COUNT * TBW * TBH;
        const double *s_tb = sh + TH_PTB * TBW * TBH, *s_tn = sh + TH_PTN * TBW * TBH, *s_tm = sh + TH_TM * TBW * TBH;
        const double *s_u = sh + TH_PUN * TBW * TBH, *s_v = sh + TH_PVN * TBW * TBH;
        const double tb_c = s_tb[hc], tb_w = s_tb[hc - 1], tb_e = s_tb[hc + 1], tb_s = s_tb[hc - TBW], tb_n = s_tb[hc + TBW];
        const double u_c = s_u[hc], u_w = s_u[hc - 1], v_c = s_v[hc], v_s = s_v[hc - TBW];
        const double tn_c = s_tn[hc], tn_e = s_tn[hc + 1], tn_n = s_tn[hc + TBW];
        const double tm = s_tm[hc];
        const double ta_c = sp[TP_PTA * TPW * TTY + pc], w_c = sp[TP_PWN * TPW * TTY + pc];
        const double e3b = sp[TP_E3B * TPW * TTY + pc], e3n = sp[TP_E3N * TPW * TTY + pc], e3a = sp[TP_E3A * TPW * TTY + pc];
        const double tb_p = sh1[TH_PTB * TBW * TBH + hc], tm_p = sh1[TH_TM * TBW * TBH + hc], w_p = sp1[TP_PWN * TPW * TTY + pc];
        const size_t o = c2 + (size_t)(k - 1) * jpij;
        double wm_c, wm_p;
        if (FROM_T) { wm_c = (k == 1) ? tm : tm * tm_m; wm_p = tm_p * tm; }
        else { wm_c = a.wmask[o]; wm_p = a.wmask[o + jpij]; }
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
        const double new_ta = ta_c + ztra / e3n * tm;
        const double new_wi = (e3b * tb_c + p2dt * ztra) / e3a * tm;
        double fx, fy;
        if (H == 2) {
            fx = 0.5 * u_c * (tn_c + tn_e) - upx_c;
            fy = 0.5 * v_c * (tn_c + tn_n) - upy_c;
        } else {
            const double tn_w = s_tn[hc - 1], tn_ee = s_tn[hc + 2], tn_s = s_tn[hc - TBW], tn_nn = s_tn[hc + 2 * TBW];
            double mu_w, mu_c, mu_e, mv_s, mv_c, mv_n;
            if (FROM_T) {
                const double tm_w = s_tm[hc - 1], tm_e = s_tm[hc + 1], tm_ee = s_tm[hc + 2];
                const double tm_s = s_tm[hc - TBW], tm_n = s_tm[hc + TBW], tm_nn = s_tm[hc + 2 * TBW];
                mu_w = tm_w * tm; mu_c = tm * tm_e; mu_e = tm_e * tm_ee;
                mv_s = tm_s * tm; mv_c = tm * tm_n; mv_n = tm_n * tm_nn;
            } else {
                mu_w = a.umask[o - 1]; mu_c = a.umask[o]; mu_e = a.umask[o + 1];
                mv_s = a.vmask[o - jpi]; mv_c = a.vmask[o]; mv_n = a.vmask[o + jpi];
            }
            const double ztu_w = (tn_c - tn_w) * mu_w, ztu_c = (tn_e - tn_c) * mu_c, ztu_e = (tn_ee - tn_e) * mu_e;
            const double ztv_s = (tn_c - tn_s) * mv_s, ztv_c = (tn_n - tn_c) * mv_c, ztv_n = (tn_nn - tn_n) * mv_n;
            const double zltu_c = (ztu_c + ztu_w) * r1_6, zltu_e = (ztu_e + ztu_c) * r1_6;
            const double zltv_c = (ztv_c + ztv_s) * r1_6, zltv_n = (ztv_n + ztv_c) * r1_6;
            const double zC2t_u = tn_c + tn_e, zC2t_v = tn_c + tn_n;
            fx = 0.5 * u_c * (zC2t_u + zltu_c - zltu_e) - upx_c;
            fy = 0.5 * v_c * (zC2t_v + zltv_c - zltv_n) - upy_c;
        }
        double fz = 0.0;
        if (k >= 2) {
            if (V == 2) fz = (w_c * 0.5 * (tn_c + tn_m) - upz_k) * wm_c;
            else        fz = (w_c * sp[TP_ZTW * TPW * TTY + pc] - upz_k) * wm_c;
        }
        if (active) { pta[o] = new_ta; zwi[o] = new_wi; zwx[o] = fx; zwy[o] = fy; zwz[o] = fz; }
        upz_k = upz_kp1; tn_m = tn_c; tm_m = tm; tb_m = tb_c;
    }
}
