// cpt_tiled_kernel.cuh -- interp_4th_cpt (traadv_fct.F90:517-616) as a tiled kernel: k_interp_4th_cpt_tiled.
//
// The Thomas solve of the compact 4th-order vertical interpolation is two sequential sweeps per column.  The column kernel
// of fct_column_kernels.cuh writes the forward sweep to global memory and reads it back (ncu: 6.1 GB of DRAM traffic for a
// 4.2 GB job, 168 instructions per point with two IEEE divisions, latency bound at 2.1 ms for ORCA025 T+S).  Here a block
// owns a 32 x 4 tile of columns: ptn arrives as TMA boxes of CKL levels in a CNB-deep ring (cp.async.bulk.tensor.3d, up to
// three boxes in flight per block), the forward sweep stays in shared memory, and the back substitution streams ztw out.
// On "simple" columns (no cavity, wet from level 1 to mbkt: verified per column by k_cpt_classify) the pivots follow from
// mbkt and a jpk-entry table, the ratio zwi/zwt(k-1) of the forward sweep is a table entry and the division by the pivot uses
// a once-refined reciprocal (div_by): per point the kernel reads 8 bytes, writes 8 and issues ~25 instructions.
// Same operations in the same order as the reference: bit-identical to the column kernel and to the oracle.
//
// Included by fct_kernels.cu INSIDE namespace nemo's anonymous namespace after fct_fused_kernel.cuh (div_rn / div_by /
// rcp_refined, mbarrier + TMA helpers) and fct_column_kernels.cuh (cpt_row); also compiled for the host by tests/emu.
#pragma once

constexpr int CTX = 32, CTY = 4, CKL = 8, CNB = 4;
constexpr int kCptBoxBytes = CTX * CTY * CKL * 8;
inline size_t cpt_tiled_smem_bytes(int jpk) { return (size_t)(jpk + 1) * CTX * CTY * 8 + (size_t)CNB * kCptBoxBytes + 3 * (size_t)(jpk + 2) * 8 + 64; }

struct CptMap { CUtensorMap m; };

__global__ void __launch_bounds__(CTX * CTY) k_interp_4th_cpt_tiled(int jpi, int jpj, int jpk, const double *__restrict__ wmask,
                                                                    const int *__restrict__ mikt, const int *__restrict__ mbkt,
                                                                    const double *__restrict__ zwt, const unsigned char *__restrict__ simple,
                                                                    const double *__restrict__ utab, double *__restrict__ pt_out_all,
                                                                    const __grid_constant__ CptMap map, int jlo, int jhi)
{
    NEMO_DYN_SMEM_ALIGNED(unsigned char, cpt_smem, 128);
    constexpr int NT = CTX * CTY;
    double *ring = reinterpret_cast<double *>(cpt_smem);                                  // [CNB][CKL][CTY][CTX]
    double *zbuf = ring + (size_t)CNB * CKL * NT;                                         // [jpk + 1][NT]: forward sweep
    double *ut = zbuf + (size_t)(jpk + 1) * NT, *rt = ut + (jpk + 2), *r2 = rt + (jpk + 2);  // pivots, 1/pivot(k-1), refined reciprocals
    unsigned long long *full = reinterpret_cast<unsigned long long *>(r2 + (jpk + 2));
    const int tid = (int)threadIdx.x, tx = tid % CTX, ty = tid / CTX;
    const int X0 = (int)blockIdx.x * CTX, Y0 = jlo - 1 + (int)blockIdx.y * CTY;           // 0-based origin of the tile (even in ji); rows jlo..jhi
    const int jn = (int)blockIdx.z;
    const int gi = X0 + tx + 1, gj = Y0 + ty + 1;                                         // 1-based column
    const bool valid = gi >= 2 && gi <= jpi - 1 && gj <= jhi;
    const int ci = min(max(gi, 2), jpi - 1), cj = min(gj, jpj - 1);
    const size_t jpij = (size_t)jpi * jpj, c2 = (size_t)(cj - 1) * jpi + (ci - 1);
    double *__restrict__ pt_out = pt_out_all + (size_t)jn * jpij * jpk;
    const int jpkm1 = jpk - 1;
    const int nbox = (jpkm1 + CKL - 1) / CKL;                                             // levels 1..jpkm1 of ptn
    const CUtensorMap *mp = &map.m;
    auto issue = [=](int b) {
        unsigned long long *bar = &full[b % CNB];
        mbar_expect_tx(bar, kCptBoxBytes);
        tma_load_3d(ring + (size_t)(b % CNB) * CKL * NT, mp, bar, X0, Y0, jn * jpk + b * CKL);
    };
    if (tid == 0) {
        for (int s = 0; s < CNB; ++s) mbar_init(&full[s], 1);
        mbar_init_fence();
        for (int b = 0; b < CNB - 1 && b < nbox; ++b) issue(b);
    }
    for (int k = tid; k <= jpk; k += NT) {                                               // tables of the simple columns
        const double u = (k >= 2 && k <= jpkm1) ? utab[k] : 1.0;
        ut[k] = u; r2[k] = rcp_refined(u);
        rt[k] = (k >= 3 && k <= jpk) ? 1.0 / utab[k - 1] : 1.0;                           // zwi(k)/zwt(k-1) with zwi = 1  (:598)
    }
    const int ikt = mikt[c2] + 1, ikb = mbkt[c2];
    const bool smp = simple && simple[c2] != 0;
    __syncthreads();

    // forward sweep: pt_out(k) = zwrm(k) - zwi(k)/zwt(k-1)*pt_out(k-1)   (:590-601)
    // Simple columns (ikt = 2): every row is rhs = c*(pt_in(k) + pt_in(k-1)) with c = 0.5 on the rows 2 and ikb, 3*wmask = 3 above
    // ikb and 3*wmask = 0 below it (:538-542, :563-571), and the elimination factor zwi/zwt(k-1) is the table entry rt(k) on the
    // regular rows 3 <= k < ikb and 0/zwt = +0 elsewhere; the product with +0 and the subtraction are still carried out, so the
    // signs of zeros come out as in the reference.
    double t_km1 = 0.0, z_m = 0.0, zwt_m = 1.0;
    double *zp = zbuf + tid;
    for (int b = 0; b < nbox; ++b) {
        if (tid == 0 && b + CNB - 1 < nbox) issue(b + CNB - 1);                           // its slot was released by the barrier ending box b-1
        mbar_wait(&full[b % CNB], (b / CNB) & 1);
        const double *box = ring + (size_t)(b % CNB) * CKL * NT + tid;
        const int kend = min(CKL, jpkm1 - b * CKL);
        if (smp) {
#pragma unroll
            for (int q = 0; q < CKL; ++q) {
                if (q >= kend) break;
                const int k = b * CKL + q + 1;
                const double t_k = box[q * NT];
                if (k >= 2) {
                    const bool reg = k >= 3 && k < ikb;
                    const double c = (k == 2 || k == ikb) ? 0.5 : (k < ikb ? 3.0 : 0.0);
                    const double rhs = c * (t_k + t_km1);
                    const double z = (k >= 3) ? rhs - (reg ? rt[k] : 0.0) * z_m : rhs;
                    zp[k * NT] = z;
                    z_m = z;
                }
                t_km1 = t_k;
            }
        } else {
            for (int q = 0; q < kend; ++q) {
                const int k = b * CKL + q + 1;
                const double t_k = box[q * NT];
                if (k >= 2) {
                    const double wm = wmask[c2 + (size_t)(k - 1) * jpij];
                    double rhs, wi;
                    if (k == ikt || k == ikb) { rhs = 0.5 * (t_km1 + t_k); wi = 0.0; }    // (:563-571)
                    else if (k == 2)          { rhs = 0.0; wi = 0.0; }                     // ln_isfcav preset (:554-556)
                    else                      { rhs = 3.0 * wm * (t_k + t_km1); wi = wm; } // (:538-542)
                    double z = rhs;
                    if (k >= 3) z = rhs - div_rn(wi, zwt_m) * z_m;
                    zp[k * NT] = z;
                    z_m = z;
                    zwt_m = zwt[c2 + (size_t)(k - 1) * jpij];
                }
                t_km1 = t_k;
            }
        }
        __syncthreads();                                                                  // box b is consumed by every thread
    }
    // back substitution (:603-614): level jpkm1 is still in registers.  Simple columns: zws = 1 on the regular rows, else 0; the
    // pivot is utab(k) above the bottom row (division through the once-refined reciprocal) and 1 from there down (x / 1 = x)
    double *po = pt_out + c2 + (size_t)(jpkm1 - 1) * jpij;
    if (smp) {
        double x = (jpkm1 < ikb) ? div_by(z_m, ut[jpkm1], r2[jpkm1]) : z_m;
        if (valid) *po = x;
        for (int k = jpk - 2; k >= 2; --k) {
            po -= jpij;
            const bool reg = k >= 3 && k < ikb;
            const double num = zp[k * NT] - (reg ? 1.0 : 0.0) * x;
            x = (k < ikb) ? div_by(num, ut[k], r2[k]) : num;
            if (valid) *po = x;
        }
    } else {
        double x = div_rn(z_m, zwt_m);
        if (valid) *po = x;
        for (int k = jpk - 2; k >= 2; --k) {
            po -= jpij;
            double d, s;
            cpt_row(k, ikt, ikb, wmask[c2 + (size_t)(k - 1) * jpij], d, s);
            x = div_rn(zp[k * NT] - s * x, zwt[c2 + (size_t)(k - 1) * jpij]);
            if (valid) *po = x;
        }
    }
}
