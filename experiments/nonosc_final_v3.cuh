// dev/nonosc_final_v3.cuh -- EXPERIMENTAL variant of k_fct_nonosc_final (nonosc_final.cuh).  NOT part of libnemo_fct.so.
//
// Status: bit-identical to the oracle in the host emulation (tests/test_cpu_kernel_emulation.py); written after the round's
// GPU budget was spent, so it has never run on a GPU and its speed is unknown.  To try it: include it next to
// nonosc_final.cuh in fct_kernels.cu, launch it from launch_fct_nonosc_final (same shared memory, same grid) and compare
// `bench.py --no-e2e --no-cpu-baseline` (the kernel is `fct_nonosc_final` in the roofline table).
//
// Why: the source-level ncu profile (profiles/r1b_ncu_source_nonosc_final.txt) puts 21 % of the stall samples on the
// per-level barrier and 17 % on the first consumers of the level-(k+1) loads in bup_bdo, which the product kernel runs
// BEFORE the barrier.  Nothing published before the barrier needs those loads: zbup / zbdo of level k are already in
// registers, and paa(k), pbb(k) can be loaded one level ahead (two more doubles per thread).  Here the loads are issued
// before the barrier and consumed behind it, so the barrier wait and the HBM latency overlap.  Same arithmetic, same bits.
// pta and e3t_n of level k-1 are read where the final trend uses them instead of being carried in registers, which keeps
// the kernel at the product kernel's static size (sm_100a, -O3: 888 instructions per two levels, 124 / 148 bytes of spill
// stores / loads vs 128 / 144); five of the seven loads per level now cross the barrier in flight, the two look-ahead loads
// of paa / pbb are still spilled (hence waited for) in front of it at the 64-register cap.
__global__ void __launch_bounds__(NX * NY, 2) k_fct_nonosc_final_v3(const FctArgs a)
{
    NEMO_DYN_SMEM(double, fct_smem);
    double(*sA)[4][NY][NX] = reinterpret_cast<double(*)[4][NY][NX]>(fct_smem);                          // [level % 3][zbup, zbdo, paa, pbb]
    double(*sB)[2][NY][NX] = reinterpret_cast<double(*)[2][NY][NX]>(fct_smem + 3 * 4 * NY * NX);        // [level % 2][zbetup, zbetdo]
    const int tx = threadIdx.x % NX, ty = threadIdx.x / NX;
    const int ox = NX - 2 * NHALO, oy = NY - 2 * NHALO;
    const int gi = a.out.i0 + ((int)blockIdx.x / a.kjpt) * ox + tx - NHALO;      // tracer index fastest over the grid (shared tmask/e3t_n lines hit L2)
    const int gj = a.out.j0 + (int)blockIdx.y * oy + ty - NHALO;
    // tiles overhang the rectangle at its east/north end: clamp the address, never the role
    const int ji = min(gi, a.out.i1 + NHALO), jj = min(gj, a.out.j1 + NHALO);
    const bool is_out = tx >= NHALO && tx < NX - NHALO && ty >= NHALO && ty < NY - NHALO && gi <= a.out.i1 && gj <= a.out.j1;
    const bool is_beta = tx >= 1 && tx < NX - 1 && ty >= 1 && ty < NY - 1;
    const size_t toff = (size_t)((int)blockIdx.x % a.kjpt) * a.n3;
    const double *pbef = a.ptb + toff, *paft = a.zwi + toff;
    const double *paa = a.zwx + toff, *pbb = a.zwy + toff, *pcc = a.zwz + toff;
    double *pta = a.pta + toff;
    const int jpi = a.jpi, jpk = a.jpk;
    const size_t jpij = a.jpij;
    const size_t c2 = (size_t)(jj - 1) * jpi + (ji - 1);
    const double zrtrn = 1.e-15;
    const double e12 = a.e1e2t[c2], r1 = a.r1_e1e2t[c2];
    const double p2dt = a.p2dt;

    double up_m, do_m, up_c, do_c, up_p, do_p;          // zbup/zbdo of this column at jk-1, jk, jk+1
    bup_bdo(pbef[c2], paft[c2], a.tmask[c2], up_c, do_c);
    up_m = up_c; do_m = do_c;                           // ikm1 = MAX(jk-1,1)
    double aft_c = paft[c2];
    double pcc_k = pcc[c2];                             // anti-diffusive pcc(jk), pcc(jk+1) rolling
    double bup_mm = 0.0, bdo_mm = 0.0, bup_m = 0.0, bdo_m = 0.0;   // betas of this column at jk-2, jk-1
    double paa_m = 0.0, pbb_m = 0.0, pcc_m = 0.0;       // own fluxes of level jk-1 (pcc_m = pcc(jk-1))
    double paa_n = paa[c2], pbb_n = pbb[c2];            // lateral fluxes of the NEXT level to publish, loaded one level ahead (jpk >= 3)

#pragma unroll 2
    for (int k = 1; k <= jpk; ++k) {
        const size_t o = c2 + (size_t)(k - 1) * jpij;
        const bool lev = k <= jpk - 1;                  // betas are computed for jk = 1..jpkm1, zbetup/do(jpk) = 0
        double paa_c = 0.0, pbb_c = 0.0, pcc_p = 0.0, aft_p = 0.0, e3n_c = 1.0;
        double bef_p = 0.0, tm_p = 0.0, paa_nn = 0.0, pbb_nn = 0.0;
        if (lev) {
            // everything published before the barrier comes from registers; the loads issued here are consumed behind it
            paa_c = paa_n; pbb_c = pbb_n;
            double *A = &sA[k % 3][0][0][0];
            A[0 * NX * NY + ty * NX + tx] = up_c; A[1 * NX * NY + ty * NX + tx] = do_c;
            A[2 * NX * NY + ty * NX + tx] = paa_c; A[3 * NX * NY + ty * NX + tx] = pbb_c;
            aft_p = paft[o + jpij]; bef_p = pbef[o + jpij]; tm_p = a.tmask[o + jpij];
            pcc_p = pcc[o + jpij];
            e3n_c = a.e3t_n[o];
            if (k + 1 <= jpk - 1) { paa_nn = paa[o + jpij]; pbb_nn = pbb[o + jpij]; }
        }
        __syncthreads();
        if (lev) bup_bdo(bef_p, aft_p, tm_p, up_p, do_p);
        double bup_c = 0.0, bdo_c = 0.0;
        if (lev && is_beta) {
            const double(*A)[NY][NX] = sA[k % 3];
            const double zup = dmax(dmax(dmax(dmax(dmax(dmax(up_c, A[0][ty][tx - 1]), A[0][ty][tx + 1]), A[0][ty - 1][tx]), A[0][ty + 1][tx]), up_m), up_p);
            const double zdo = dmin(dmin(dmin(dmin(dmin(dmin(do_c, A[1][ty][tx - 1]), A[1][ty][tx + 1]), A[1][ty - 1][tx]), A[1][ty + 1][tx]), do_m), do_p);
            const double paa_w = A[2][ty][tx - 1], pbb_s = A[3][ty - 1][tx];
            const double zpos = dmax(0., paa_w) - dmin(0., paa_c) + dmax(0., pbb_s) - dmin(0., pbb_c)
                              + dmax(0., pcc_p) - dmin(0., pcc_k);
            const double zneg = dmax(0., paa_c) - dmin(0., paa_w) + dmax(0., pbb_c) - dmin(0., pbb_s)
                              + dmax(0., pcc_k) - dmin(0., pcc_p);
            const double zbt = e12 * e3n_c / p2dt;
            bup_c = (zup - aft_c) / (zpos + zrtrn) * zbt;
            bdo_c = (aft_c - zdo) / (zneg + zrtrn) * zbt;
        }
        sB[k % 2][0][ty][tx] = bup_c; sB[k % 2][1][ty][tx] = bdo_c;
        // No second barrier: the betas of level k-1 read below were published before this iteration's barrier, sB[k % 2]
        // was last read (as level k-2) before it too, and sA[(k+1) % 3] is not rewritten until after the next one.
        if (k >= 2 && is_out) {
            // final trend of level kk = k-1: betas(kk) of the neighbours from sB, own betas at kk-1, kk, kk+1
            const int kk = k - 1;
            const double(*Bm)[NY][NX] = sB[kk % 2];
            const double(*Am)[NY][NX] = sA[kk % 3];
            const double bup_e = Bm[0][ty][tx + 1], bdo_e = Bm[1][ty][tx + 1], bup_w = Bm[0][ty][tx - 1], bdo_w = Bm[1][ty][tx - 1];
            const double bup_n = Bm[0][ty + 1][tx], bdo_n = Bm[1][ty + 1][tx], bup_s = Bm[0][ty - 1][tx], bdo_s = Bm[1][ty - 1][tx];
            const double paa_w = Am[2][ty][tx - 1], pbb_s = Am[3][ty - 1][tx];
            const double lx_e = paa_m * limit_coef_sel(paa_m, bdo_m, bup_e, bup_m, bdo_e);
            const double lx_w = paa_w * limit_coef_sel(paa_w, bdo_w, bup_m, bup_w, bdo_m);
            const double ly_n = pbb_m * limit_coef_sel(pbb_m, bdo_m, bup_n, bup_m, bdo_n);
            const double ly_s = pbb_s * limit_coef_sel(pbb_s, bdo_s, bup_m, bup_s, bdo_m);
            // pcc(jk+1) is limited with betas(jk), betas(jk+1) (:419-422); pcc(:,:,1) is never limited
            const double lz_t = (kk == 1) ? pcc_m : pcc_m * limit_coef_sel(pcc_m, bdo_m, bup_mm, bup_m, bdo_mm);
            const double lz_b = pcc_k * limit_coef_sel(pcc_k, bdo_c, bup_m, bup_c, bdo_m);
            const size_t om = o - jpij;
            pta[om] = pta[om] - (lx_e - lx_w + ly_n - ly_s + lz_t - lz_b) * r1 / a.e3t_n[om];
        }
        // rotate the column registers
        up_m = up_c; do_m = do_c; up_c = up_p; do_c = do_p; aft_c = aft_p;
        bup_mm = bup_m; bdo_mm = bdo_m; bup_m = bup_c; bdo_m = bdo_c;
        paa_m = paa_c; pbb_m = pbb_c; pcc_m = pcc_k; pcc_k = pcc_p;
        paa_n = paa_nn; pbb_n = pbb_nn;
    }
}
