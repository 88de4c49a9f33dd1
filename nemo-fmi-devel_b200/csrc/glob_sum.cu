// glob_sum.cu -- the conservation metric of the reference on the device: glob_sum (src/OCE/lib_fortran_generic.h90:32-65), a
// masked sum accumulated in double-double with Knuth's two-sum (DDPDD, src/OCE/lib_fortran.F90:300-332) so that the result does
// not depend on the decomposition (the reference combines the per-rank pairs with MPI_SUMDD, lib_mpp.F90:1158-1186).
//
//   local part:   ctmp = DDPDD( CMPLX( ptab(ji,jj,jk) [* pw3d(ji,jj,jk)] * tmask_i(ji,jj), 0 ), ctmp )      for all ji, jj, jk
//
// The optional pw3d stands for the array expression the callers multiply in before the call (e.g. tr * cvol in trcrad.F90).
// Summation order: the reference adds point after point; here every thread adds a strided subset, then lanes, warps and blocks
// are folded pairwise with the same DDPDD (fixed grid => deterministic).  The pair carries ~106 bits, so REAL(ctmp) -- what
// glob_sum returns -- is the correctly rounded sum whatever the order, save for pathological cancellations.
// Compiled with -fmad=false: the error terms must not be contracted.
#include "kernels.cuh"

namespace nemo {
void note_launch();

namespace {

__device__ __forceinline__ void ddpdd(double a_hi, double a_lo, double &b_hi, double &b_lo)      // b = a + b  (lib_fortran.F90:314-331)
{
    const double zt1 = a_hi + b_hi;
    const double zerr = zt1 - a_hi;
    const double zt2 = ((b_hi - zerr) + (a_hi - (zt1 - zerr))) + a_lo + b_lo;
    const double s = zt1 + zt2;
    b_lo = zt2 - (s - zt1);
    b_hi = s;
}

__device__ __forceinline__ void block_fold(double &hi, double &lo, double *sh)
{
    for (int off = 16; off > 0; off >>= 1) {
        const double ohi = __shfl_down_sync(0xffffffffu, hi, off), olo = __shfl_down_sync(0xffffffffu, lo, off);
        ddpdd(ohi, olo, hi, lo);
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    if (lane == 0) { sh[2 * warp] = hi; sh[2 * warp + 1] = lo; }
    __syncthreads();
    if (warp == 0) {
        hi = lane < nwarp ? sh[2 * lane] : 0.0; lo = lane < nwarp ? sh[2 * lane + 1] : 0.0;
        for (int off = 16; off > 0; off >>= 1) {
            const double ohi = __shfl_down_sync(0xffffffffu, hi, off), olo = __shfl_down_sync(0xffffffffu, lo, off);
            ddpdd(ohi, olo, hi, lo);
        }
    }
}

// partial[(f * gridDim.x + b) * 2 + {0,1}] = double-double sum of block b over field f
__global__ void __launch_bounds__(256) k_glob_sum_partial(const double *const *ptab, const double *pw3d, const double *tmask_i, size_t jpij, int ipk,
                                                          double *partial)
{
    __shared__ double sh[64];
    const double *p = ptab[blockIdx.y];
    const size_t n = jpij * (size_t)ipk;
    double hi = 0.0, lo = 0.0;
    for (size_t m = (size_t)blockIdx.x * blockDim.x + threadIdx.x; m < n; m += (size_t)gridDim.x * blockDim.x) {
        double v = p[m];
        if (pw3d) v = v * pw3d[m];
        v = v * tmask_i[m % jpij];
        ddpdd(v, 0.0, hi, lo);
    }
    block_fold(hi, lo, sh);
    if (threadIdx.x == 0) { partial[((size_t)blockIdx.y * gridDim.x + blockIdx.x) * 2] = hi; partial[((size_t)blockIdx.y * gridDim.x + blockIdx.x) * 2 + 1] = lo; }
}

__global__ void __launch_bounds__(256) k_glob_sum_final(const double *partial, int nblk, double *out)
{
    __shared__ double sh[64];
    double hi = 0.0, lo = 0.0;
    for (int b = threadIdx.x; b < nblk; b += blockDim.x)
        ddpdd(partial[((size_t)blockIdx.x * nblk + b) * 2], partial[((size_t)blockIdx.x * nblk + b) * 2 + 1], hi, lo);
    block_fold(hi, lo, sh);
    if (threadIdx.x == 0) { out[2 * blockIdx.x] = hi; out[2 * blockIdx.x + 1] = lo; }
}

}  // namespace

#ifndef NEMO_EMU_KERNELS_ONLY          // the launcher (CUDA launch syntax) is left out of the host emulation build of tests/emu
void launch_glob_sum(const double *const *ptab_dev, int nfld, const double *pw3d, const double *tmask_i, size_t jpij, int ipk, double *partial,
                     double *out_pairs, cudaStream_t s)
{
    const int nblk = kGlobSumBlocks;
    k_glob_sum_partial<<<dim3(nblk, nfld), 256, 0, s>>>(ptab_dev, pw3d, tmask_i, jpij, ipk, partial);
    note_launch();
    k_glob_sum_final<<<nfld, 256, 0, s>>>(partial, nblk, out_pairs);
    note_launch();
}

#endif  // NEMO_EMU_KERNELS_ONLY

}  // namespace nemo
