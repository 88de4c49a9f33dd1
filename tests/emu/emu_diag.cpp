// host emulation of the diagnostic reductions (glob_sum.cu, stp_ctl.cu): warp shuffles + static shared memory, see emu_block.h
// (test infrastructure only)
#include "emu_block.h"
#undef __shared__
#define __shared__ static                  // blocks are emulated one at a time: one copy per kernel is the block's copy
#include "../../nemo-fmi-devel_b200/csrc/glob_sum.cu"
#include "../../nemo-fmi-devel_b200/csrc/stp_ctl.cu"

extern "C" {

// out_pairs[2 * f + {0, 1}] = double-double sum of field f, as launch_glob_sum with a grid of nblk blocks
int emu_glob_sum(const double *const *ptab, int nfld, const double *pw3d, const double *tmask_i, long long jpij, int ipk, int nblk, double *out_pairs)
{
    std::vector<double> partial((size_t)nfld * nblk * 2);
    emu_run_blocks(nblk, nfld, 256, 0, nemo::k_glob_sum_partial, ptab, pw3d, tmask_i, (size_t)jpij, ipk, partial.data());
    emu_run_blocks(nfld, 1, 256, 0, nemo::k_glob_sum_final, (const double *)partial.data(), nblk, out_pairs);
    return 0;
}

// rec: the StpCtlRec of launch_stp_ctl with a grid of nblk blocks, as 6 doubles + 6 indices + flags
int emu_stp_ctl(const double *sshn, const double *un, const double *tem, const double *sal, const double *tmask, long long jpij, long long n3,
                int nblk, double *vals, long long *idx, int *flags)
{
    std::vector<nemo::StpCtlRec> partial(nblk);
    nemo::StpCtlRec out;
    emu_run_blocks(nblk, 1, 256, 0, nemo::k_stp_ctl_partial, sshn, un, tem, sal, tmask, (size_t)jpij, (size_t)n3, partial.data());
    emu_run_blocks(1, 1, 256, 0, nemo::k_stp_ctl_final, (const nemo::StpCtlRec *)partial.data(), nblk, &out);
    const double v[6] = {out.z1, out.z2, out.smin, out.smax, out.tmin, out.tmax};
    const long long l[6] = {out.l1, out.l2, out.ls1, out.ls2, out.lt1, out.lt2};
    for (int q = 0; q < 6; ++q) { vals[q] = v[q]; idx[q] = l[q]; }
    *flags = out.flags;
    return 0;
}

}  // extern "C"
