// host emulation of the TMA-tiled P1-P5 kernel k_fct_low_antidiff_tma (fct_tma_kernel.cuh), see emu_block.h.
// TEST INFRASTRUCTURE ONLY.  One host thread per CUDA thread; an mbarrier is an atomic count of completed phases; a bulk
// tensor copy is a box copy with zero fill on the high side that completes before the issuing thread goes on.  The
// hardware rules the kernel was designed around are CHECKED here: box origins >= 0 and even in the innermost coordinate.
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

#include "../../nemo-fmi-devel_b200/csrc/fct_tma_kernel.cuh"

} }  // namespace

extern "C" {

// arrays as in emu_fct (same table).  rect = the single rectangle of the region (launch_fct_low_antidiff_tma needs one
// rectangle, an even jpi and an even first column).  Returns -1 when the product would refuse the TMA path, else the number
// of hardware-rule violations seen (0 = fine).
int emu_fct_low_antidiff_tma(int jpi, int jpj, int jpk, int kjpt, int h, int v, int ln_linssh, int ln_isfcav, const int *rect, int nkchunk,
                             double p2dt, double *const *arr, const int *mikt, const int *mbkt, int masks_from_t)
{
    using namespace nemo;
    if ((jpi & 1) || (rect[0] & 1)) return -1;
    FctArgs a;
    std::memset(&a, 0, sizeof a);
    a.jpi = jpi; a.jpj = jpj; a.jpk = jpk; a.jpij = (size_t)jpi * jpj; a.n3 = a.jpij * jpk;
    a.tmask = arr[0]; a.umask = arr[1]; a.vmask = arr[2]; a.wmask = arr[3]; a.e3t_b = arr[4]; a.e3t_n = arr[5]; a.e3t_a = arr[6];
    a.e1e2t = arr[7]; a.r1_e1e2t = arr[8]; a.mikt = mikt; a.mbkt = mbkt;
    a.pun = arr[9]; a.pvn = arr[10]; a.pwn = arr[11]; a.ptb = arr[12]; a.ptn = arr[13]; a.pta = arr[14];
    a.zwi = arr[15]; a.zwx = arr[16]; a.zwy = arr[17]; a.zwz = arr[18]; a.ztw = arr[21];
    a.p2dt = p2dt; a.kjpt = kjpt; a.kn_fct_h = h; a.kn_fct_v = v; a.ln_linssh = ln_linssh; a.ln_isfcav = ln_isfcav; a.nkchunk = nkchunk;
    a.masks_from_t = masks_from_t;
    const Rect rc{rect[0], rect[1], rect[2], rect[3]};
    TileMaps tm;
    const long long n4 = (long long)jpk * kjpt, n3 = jpk;                           // as launch_fct_low_antidiff_tma builds them
    set_map(&tm.h[TH_PTB], a.ptb, jpi, jpj, n4, TBW, TBH); set_map(&tm.h[TH_PTN], a.ptn, jpi, jpj, n4, TBW, TBH);
    set_map(&tm.h[TH_TM], a.tmask, jpi, jpj, n3, TBW, TBH); set_map(&tm.h[TH_PUN], a.pun, jpi, jpj, n3, TBW, TBH);
    set_map(&tm.h[TH_PVN], a.pvn, jpi, jpj, n3, TBW, TBH); set_map(&tm.p[TP_PTA], a.pta, jpi, jpj, n4, TPW, TTY);
    set_map(&tm.p[TP_ZTW], v == 4 ? a.ztw : a.pta, jpi, jpj, n4, TPW, TTY);
    set_map(&tm.p[TP_PWN], a.pwn, jpi, jpj, n3, TPW, TTY); set_map(&tm.p[TP_E3B], a.e3t_b, jpi, jpj, n3, TPW, TTY);
    set_map(&tm.p[TP_E3N], a.e3t_n, jpi, jpj, n3, TPW, TTY); set_map(&tm.p[TP_E3A], a.e3t_a, jpi, jpj, n3, TPW, TTY);
    const size_t smem = (size_t)TSTAGES * kTileStageBytes + 64;
    const int gx = ((rc.i1 - rc.i0 + 1 + TTX - 1) / TTX) * kjpt, gy = (rc.j1 - rc.j0 + 1 + TTY - 1) / TTY;
    emu_tma_violations = 0;
#define LAT(H, V) do { if (masks_from_t) emu_run_blocks3(gx, gy, nkchunk, TTX * TTY, smem, k_fct_low_antidiff_tma<H, V, true>, a, tm, rc); \
                       else emu_run_blocks3(gx, gy, nkchunk, TTX * TTY, smem, k_fct_low_antidiff_tma<H, V, false>, a, tm, rc); } while (0)
    if (h == 2 && v == 2) LAT(2, 2); else if (h == 2) LAT(2, 4); else if (v == 2) LAT(4, 2); else LAT(4, 4);
#undef LAT
    return emu_tma_violations;
}

}  // extern "C"
