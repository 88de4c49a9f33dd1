// host emulation of the one-kernel FCT step k_fct_fused (fct_fused_kernel.cuh), see emu_block.h and emu_tma_helpers.h.
// TEST INFRASTRUCTURE ONLY.  One host thread per CUDA thread (512 per block), std::barrier for __syncthreads.
#include "emu_block.h"

#include <cuda.h>
#include <cstdio>
#include <cstdlib>
#include <map>
#include <mutex>
#include <thread>

#include "../../nemo-fmi-devel_b200/csrc/kernels.cuh"

#undef __grid_constant__
#define __grid_constant__

namespace nemo { namespace {

#include "emu_tma_helpers.h"
#include "../../nemo-fmi-devel_b200/csrc/nonosc_final.cuh"        // dmax / dmin / bup_bdo / limit_coef_sel
#include "../../nemo-fmi-devel_b200/csrc/fct_fused_kernel.cuh"
constexpr int kThreads = 128;
#include "../../nemo-fmi-devel_b200/csrc/fct_column_kernels.cuh"   // cpt_row, k_cpt_pivots, k_cpt_classify
#include "../../nemo-fmi-devel_b200/csrc/cpt_tiled_kernel.cuh"

} }  // namespace

extern "C" {

// arrays as in emu_fct (same table); out = i0, i1, j0, j1 (1-based output rectangle).  Returns -1 when the product would
// refuse the kernel (prepare_fct_fused), else the number of hardware-rule violations seen (0 = fine).
int emu_fct_fused(int jpi, int jpj, int jpk, int kjpt, int h, int v, int ln_linssh, int ln_isfcav, const int *out, int nkchunk,
                  double p2dt, double *const *arr, const int *mikt, const int *mbkt, int masks_from_t)
{
    using namespace nemo;
    FctArgs a;
    std::memset(&a, 0, sizeof a);
    a.out = Rect{out[0], out[1], out[2], out[3]};
    const int ni = a.out.i1 - a.out.i0 + 1, nj = a.out.j1 - a.out.j0 + 1;
    if (ni <= 0 || nj <= 0) return -1;
    if ((jpi & 1) || !(a.out.i0 & 1) || a.out.i0 - 1 - FHALO - 2 < 0 || a.out.j0 - 1 - FHALO - 1 < 0 || jpk < 3 || !masks_from_t) return -1;
    a.jpi = jpi; a.jpj = jpj; a.jpk = jpk; a.jpij = (size_t)jpi * jpj; a.n3 = a.jpij * jpk;
    a.tmask = arr[0]; a.umask = arr[1]; a.vmask = arr[2]; a.wmask = arr[3]; a.e3t_b = arr[4]; a.e3t_n = arr[5]; a.e3t_a = arr[6];
    a.e1e2t = arr[7]; a.r1_e1e2t = arr[8]; a.mikt = mikt; a.mbkt = mbkt;
    a.pun = arr[9]; a.pvn = arr[10]; a.pwn = arr[11]; a.ptb = arr[12]; a.ptn = arr[13]; a.pta = arr[14]; a.ztw = arr[21];
    a.p2dt = p2dt; a.kjpt = kjpt; a.kn_fct_h = h; a.kn_fct_v = v; a.ln_linssh = ln_linssh; a.ln_isfcav = ln_isfcav; a.nkchunk = nkchunk;
    a.masks_from_t = masks_from_t ? 3 : 0;
    FusedMaps tm;
    const long long n4 = (long long)jpk * kjpt, n3 = jpk;                           // as prepare_fct_fused builds them
    const double *hb[FH_COUNT] = {a.ptb, a.ptn, a.tmask, a.pun, a.pvn};
    const long long hn[FH_COUNT] = {n4, n4, n3, n3, n3};
    const double *pb[FP_COUNT] = {a.pta, v == 4 ? a.ztw : a.pta, a.pwn, a.e3t_b, a.e3t_n, a.e3t_a};
    const long long pn[FP_COUNT] = {n4, n4, n3, n3, n3, n3};
    for (int q = 0; q < FH_COUNT; ++q) set_map(&tm.h[q], hb[q], jpi, jpj, hn[q], FBW, FBH);
    for (int q = 0; q < FP_COUNT; ++q) set_map(&tm.p[q], pb[q], jpi, jpj, pn[q], FX, FY);
    const int gx = ((ni + FOX - 1) / FOX) * kjpt, gy = (nj + FOY - 1) / FOY;
    emu_tma_violations = 0;
    // persistent grid: 3 blocks share the (tile, tracer, chunk) work items round-robin, as on the device with one block per SM
    const int nwork = gx * gy * nkchunk, nblk = nwork < 3 ? nwork : 3;
#define LFU(H, V) emu_run_blocks3(nblk, 1, 1, FX * FY, kFusedSmemBytes, k_fct_fused<H, V, 0, true>, a, tm, gx, gy, nwork, 0)
    if (h == 2 && v == 2) LFU(2, 2); else if (h == 2) LFU(2, 4); else if (v == 2) LFU(4, 2); else LFU(4, 4);
#undef LFU
    return emu_tma_violations;
}


// interp_4th_cpt through the tiled kernel exactly as launch_interp_4th_cpt runs it: pivots, classification, TMA-fed solve.
// Returns -1 where the product falls back to the column kernel (odd jpi), else the number of TMA rule violations (0 = fine).
int emu_interp_4th_cpt_tiled(int jpi, int jpj, int jpk, int nfld, const double *wmask, const int *mikt, const int *mbkt,
                             const double *pt_in, double *pt_out, int use_simple)
{
    using namespace nemo;
    if ((jpi & 1) || jpk < 3) return -1;
    const size_t jpij = (size_t)jpi * jpj;
    std::vector<double> zwt(jpij * jpk, 0.0), utab(jpk + 1, 1.0);
    std::vector<unsigned char> simple(jpij, 0);
    const int ncol = (jpi - 2) * (jpj - 2);
    blockDim = {128, 1, 1};
    auto grid1 = [&](auto kernel) {
        for (int bx = 0; bx < (ncol + 127) / 128; ++bx)
            for (int t = 0; t < 128; ++t) { blockIdx = {(unsigned)bx, 0, 0}; threadIdx = {(unsigned)t, 0, 0}; kernel(); }
    };
    grid1([&]() { k_cpt_pivots(jpi, jpj, jpk, wmask, mikt, mbkt, zwt.data()); });
    { double t_m = 1.0, s_m = 0.0; for (int k = 3; k <= jpk - 1; ++k) { const double t = 4.0 - 1.0 * s_m / t_m; utab[k] = t; t_m = t; s_m = 1.0; } }
    if (use_simple) grid1([&]() { k_cpt_classify(jpi, jpj, jpk, wmask, mikt, mbkt, zwt.data(), utab.data(), simple.data()); });
    CptMap m;
    std::memset(&m, 0, sizeof m);
    set_map(&m.m, pt_in, jpi, jpj, (long long)jpk * nfld, CTX, CTY * 1);
    emu_tma_violations = 0;
    emu_box_depth = CKL;
    emu_run_blocks3((jpi - 1 + CTX - 1) / CTX, (jpj - 2 + CTY - 1) / CTY, nfld, CTX * CTY, cpt_tiled_smem_bytes(jpk), k_interp_4th_cpt_tiled,
                    jpi, jpj, jpk, wmask, mikt, mbkt, (const double *)zwt.data(), (const unsigned char *)simple.data(), (const double *)utab.data(), pt_out, m, 2, jpj - 1);
    emu_box_depth = 1;
    return emu_tma_violations;
}

}  // extern "C"
