// host emulation of the FCT column kernels (fct_column_kernels.cuh), see emu_common.h (test infrastructure only)
#include "emu_block.h"
#include "../../nemo-fmi-devel_b200/csrc/kernels.cuh"

#include "../../nemo-fmi-devel_b200/csrc/schedule.hpp"

namespace nemo { namespace {
constexpr int kThreads = 128;
#include "../../nemo-fmi-devel_b200/csrc/nonosc_final.cuh"
#include "../../nemo-fmi-devel_b200/csrc/fct_column_kernels.cuh"
// cp.async on the host: the copy happens at issue time (a thread only ever reads the slots it filled itself)
inline void cp_async8(double *smem_dst, const double *gsrc) { *smem_dst = *gsrc; }
inline void cp_async_commit() {}
template <int N> inline void cp_async_wait() {}
#include "../../nemo-fmi-devel_b200/csrc/fct_inner_kernel.cuh"
} }

extern "C" {

// schedule 4: rectangle inside which the band launch of k_fct_low_antidiff_inner leaves pta alone (FctArgs::out); empty = none
static int g_skip_rect[4] = {0, -1, 0, -1};
void emu_fct_set_skip_rect(int i0, int i1, int j0, int j1) { g_skip_rect[0] = i0; g_skip_rect[1] = i1; g_skip_rect[2] = j0; g_skip_rect[3] = j1; }

// which: 0 laplacian, 1 low_antidiff, 2 betas, 3 limit, 4 final, 5 trend-diagnostic hook, 6 low_antidiff_inner (fused P1-P5).
// arrays: tmask umask vmask wmask e3t_b e3t_n e3t_a e1e2t r1_e1e2t | pun pvn pwn ptb ptn pta | zwi zwx zwy zwz zltu zltv ztw zbetup zbetdo
//         | trdx trdy trdz (hook only) | zlx zly zlz (frame of the fused schedules: limited fluxes kept apart, else NULL)
int emu_fct(int which, int jpi, int jpj, int jpk, int kjpt, int h, int v, int ln_linssh, int ln_isfcav, const int *rect, int nkchunk,
            double p2dt, double *const *arr, const int *mikt, const int *mbkt, int masks_from_t)
{
    using namespace nemo;
    FctArgs a;
    std::memset(&a, 0, sizeof a);
    a.reg = Region(); a.reg.add(rect[0], rect[1], rect[2], rect[3]);
    a.out = Rect{g_skip_rect[0], g_skip_rect[1], g_skip_rect[2], g_skip_rect[3]};
    a.jpi = jpi; a.jpj = jpj; a.jpk = jpk; a.jpij = (size_t)jpi * jpj; a.n3 = a.jpij * jpk;
    a.tmask = arr[0]; a.umask = arr[1]; a.vmask = arr[2]; a.wmask = arr[3]; a.e3t_b = arr[4]; a.e3t_n = arr[5]; a.e3t_a = arr[6];
    a.e1e2t = arr[7]; a.r1_e1e2t = arr[8]; a.mikt = mikt; a.mbkt = mbkt;
    a.pun = arr[9]; a.pvn = arr[10]; a.pwn = arr[11]; a.ptb = arr[12]; a.ptn = arr[13]; a.pta = arr[14];
    a.zwi = arr[15]; a.zwx = arr[16]; a.zwy = arr[17]; a.zwz = arr[18]; a.zltu = arr[19]; a.zltv = arr[20]; a.ztw = arr[21];
    a.zbetup = arr[22]; a.zbetdo = arr[23];
    a.zlx = arr[27]; a.zly = arr[28]; a.zlz = arr[29]; a.masks_from_t = masks_from_t;
    a.p2dt = p2dt; a.kjpt = kjpt; a.kn_fct_h = h; a.kn_fct_v = v; a.ln_linssh = ln_linssh; a.ln_isfcav = ln_isfcav; a.nkchunk = nkchunk;
    const int ncol = a.reg.ncol();
    switch (which) {
    case 0: emu_run_grid(a, k_fct_laplacian, ncol, nkchunk, kjpt); break;
    case 1:
        if (h == 2 && v == 2) emu_run_grid(a, k_fct_low_antidiff<2, 2>, ncol, nkchunk, kjpt);
        else if (h == 2)      emu_run_grid(a, k_fct_low_antidiff<2, 4>, ncol, nkchunk, kjpt);
        else if (v == 2)      emu_run_grid(a, k_fct_low_antidiff<4, 2>, ncol, nkchunk, kjpt);
        else                  emu_run_grid(a, k_fct_low_antidiff<4, 4>, ncol, nkchunk, kjpt);
        break;
    case 2: emu_run_grid(a, k_fct_betas, ncol, nkchunk, kjpt); break;
    case 3: emu_run_grid(a, k_fct_limit, ncol, nkchunk, kjpt); break;
    case 4: emu_run_grid(a, k_fct_final, ncol, nkchunk, kjpt); break;
    case 5: {
        double *tx = arr[24], *ty = arr[25], *tz = arr[26];
        auto k = [&](const FctArgs &x) { k_fct_diag(x, tx, ty, tz); };
        emu_run_grid(a, k, (int)a.jpij, 1, kjpt);
        break;
    }
    case 6: {
        static std::vector<double> ring((size_t)kPfStages * F_LOW_COUNT * kThreads);     // the block's dynamic shared memory
        emu_block_smem = reinterpret_cast<unsigned char *>(ring.data());
#define LAI(H, V) do { if (masks_from_t) emu_run_grid(a, k_fct_low_antidiff_inner<H, V, true>, ncol, nkchunk, kjpt); \
                       else emu_run_grid(a, k_fct_low_antidiff_inner<H, V, false>, ncol, nkchunk, kjpt); } while (0)
        if (h == 2 && v == 2) LAI(2, 2); else if (h == 2) LAI(2, 4); else if (v == 2) LAI(4, 2); else LAI(4, 4);
#undef LAI
        break;
    }
    default: return 1;
    }
    return 0;
}

// the column sets of the fused schedules (schedule.hpp).  Each region = count + 4 rectangles (i0, i1, j0, j1) = 17 ints.
static int *put_region(int *o, const nemo::Region &r)
{
    *o++ = r.n;
    for (int q = 0; q < 4; ++q) { const nemo::Rect x = q < r.n ? r.r[q] : nemo::Rect{0, -1, 0, -1}; *o++ = x.i0; *o++ = x.i1; *o++ = x.j0; *o++ = x.j1; }
    return o;
}
// out: k1, k1_band, k1_centre, lowf, lap, bet, lim, fin (8 x 17), k2_out (4), split (1), min fused size (1)
void emu_fct_fused_plan(int jpi, int jpj, int fold, int want_split, int *out)
{
    const nemo::FctFusedPlan p = nemo::fct_fused_plan(jpi, jpj, fold != 0, want_split != 0);
    for (const nemo::Region *r : {&p.k1, &p.k1_band, &p.k1_centre, &p.lowf, &p.lap, &p.bet, &p.lim, &p.fin}) out = put_region(out, *r);
    *out++ = p.k2_out.i0; *out++ = p.k2_out.i1; *out++ = p.k2_out.j0; *out++ = p.k2_out.j1;
    *out++ = p.split ? 1 : 0; *out++ = nemo::kMinFusedSize;
}
// which: 0 = default (semi-fused), 1 = fully fused.  out: inner, grad, hflux, trend (4 x 17)
void emu_mus_plan(int which, int jpi, int jpj, int fold, int *out)
{
    const nemo::MusPlan p = which == 0 ? nemo::mus_semi_plan(jpi, jpj, fold != 0) : nemo::mus_fused_plan(jpi, jpj, fold != 0);
    for (const nemo::Region *r : {&p.inner, &p.grad, &p.hflux, &p.trend}) out = put_region(out, *r);
}

// interp_4th_cpt on nfld fields exactly as the product does it: pivots once, classification of the simple columns, solve
int emu_interp_4th_cpt(int jpi, int jpj, int jpk, int nfld, int ln_isfcav, const double *wmask, const int *mikt, const int *mbkt,
                       const double *pt_in, double *pt_out, int use_simple)
{
    using namespace nemo;
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
    {   // pivots of a full-depth, cavity-free column (nemo_fct_set_domain_arrays)
        double t_m = 1.0, s_m = 0.0;
        for (int k = 3; k <= jpk - 1; ++k) { const double t = 4.0 - 1.0 * s_m / t_m; utab[k] = t; t_m = t; s_m = 1.0; }
    }
    if (use_simple) grid1([&]() { k_cpt_classify(jpi, jpj, jpk, wmask, mikt, mbkt, zwt.data(), utab.data(), simple.data()); });
    (void)ln_isfcav;
    for (int bz = 0; bz < nfld; ++bz)
        for (int bx = 0; bx < (ncol + 127) / 128; ++bx)
            for (int t = 0; t < 128; ++t) {
                blockIdx = {(unsigned)bx, 0, (unsigned)bz}; threadIdx = {(unsigned)t, 0, 0};
                k_interp_4th_cpt(jpi, jpj, jpk, wmask, mikt, mbkt, zwt.data(), simple.data(), utab.data(), pt_in, pt_out, Region(), nullptr);
            }
    int nsimple = 0;
    for (unsigned char c : simple) nsimple += c;
    return nsimple;
}

}  // extern "C"
