// host emulation of the fused limiter kernel k_fct_nonosc_final (nonosc_final.cuh), see emu_block.h (test infrastructure only)
#include "emu_block.h"
#include "../../nemo-fmi-devel_b200/csrc/kernels.cuh"

namespace nemo { namespace {
#include "../../nemo-fmi-devel_b200/csrc/nonosc_final.cuh"
#include "../../experiments/nonosc_final_v3.cuh"      // experimental variant, not in the product library
} }

extern "C" {

// out = i0, i1, j0, j1 (1-based output rectangle); variant 0 = the product kernel, 3 = experiments/nonosc_final_v3.cuh
int emu_nonosc_final_variant(int variant, int jpi, int jpj, int jpk, int kjpt, const int *out, double p2dt, const double *tmask,
                             const double *e3t_n, const double *e1e2t, const double *r1_e1e2t, const double *ptb, const double *zwi,
                             const double *zwx, const double *zwy, const double *zwz, double *pta);
int emu_nonosc_final(int jpi, int jpj, int jpk, int kjpt, const int *out, double p2dt, const double *tmask,
                     const double *e3t_n, const double *e1e2t, const double *r1_e1e2t, const double *ptb, const double *zwi,
                     const double *zwx, const double *zwy, const double *zwz, double *pta)
{
    return emu_nonosc_final_variant(0, jpi, jpj, jpk, kjpt, out, p2dt, tmask, e3t_n, e1e2t, r1_e1e2t, ptb, zwi, zwx, zwy, zwz, pta);
}
int emu_nonosc_final_variant(int variant, int jpi, int jpj, int jpk, int kjpt, const int *out, double p2dt, const double *tmask,
                             const double *e3t_n, const double *e1e2t, const double *r1_e1e2t, const double *ptb, const double *zwi,
                             const double *zwx, const double *zwy, const double *zwz, double *pta)
{
    using namespace nemo;
    FctArgs a;
    std::memset(&a, 0, sizeof a);
    a.out = Rect{out[0], out[1], out[2], out[3]};
    a.jpi = jpi; a.jpj = jpj; a.jpk = jpk; a.jpij = (size_t)jpi * jpj; a.n3 = a.jpij * jpk;
    a.tmask = tmask; a.e3t_n = e3t_n; a.e1e2t = e1e2t; a.r1_e1e2t = r1_e1e2t;
    a.ptb = ptb; a.pta = pta; a.zwi = const_cast<double *>(zwi); a.zwx = const_cast<double *>(zwx);
    a.zwy = const_cast<double *>(zwy); a.zwz = const_cast<double *>(zwz);
    a.p2dt = p2dt; a.kjpt = kjpt;
    const int ox = NX - 2 * NHALO, oy = NY - 2 * NHALO;
    const int ni = out[1] - out[0] + 1, nj = out[3] - out[2] + 1;
    if (ni <= 0 || nj <= 0) return 0;
    const int gx = ((ni + ox - 1) / ox) * kjpt, gy = (nj + oy - 1) / oy;           // the grid of launch_fct_nonosc_final
    const size_t smem = (size_t)(3 * 4 + 2 * 2) * NY * NX * sizeof(double);
    if (variant == 3) emu_run_blocks(gx, gy, NX * NY, smem, k_fct_nonosc_final_v3, a);
    else              emu_run_blocks(gx, gy, NX * NY, smem, k_fct_nonosc_final, a);
    return 0;
}

}  // extern "C"
