// stp_ctl.cu -- the extrema / NaN guard of the reference's time-stepping control on device-resident state
// (src/OCE/stpctl.F90:115-124: zmax(1:6) = max |sshn|, max |un|, max(-S), max(S), max(-T), max(T), the tracer ones where
// tmask == 1;  :162-165: MAXLOC / MINLOC of |sshn|, |un|, S).  One pass over un / tsn / tmask (32 B per point) and sshn.
//
// MAXVAL / MAXLOC over NaN operands are processor dependent in Fortran (gfortran skips them unless all are NaN); here a NaN never
// wins a comparison and is reported through a flag instead, which is what the reference's ISNAN( zmax(1)+zmax(2)+zmax(3) ) is for.
// Ties go to the first element in array order, as MAXLOC / MINLOC do.
#include "kernels.cuh"

#include <cfloat>

namespace nemo {
void note_launch();

namespace {

struct Best { double v; long long l; };                                 // value and linear index (-1: nothing yet)

template <bool MAX> __device__ __forceinline__ void take(Best &a, double v, long long l)
{
    if (l < 0 || v != v) return;
    const bool better = MAX ? v > a.v : v < a.v;
    if (a.l < 0 || better || (v == a.v && l < a.l)) { a.v = v; a.l = l; }
}

template <bool MAX> __device__ __forceinline__ void fold(Best &a, double *shv, long long *shl)
{
    for (int off = 16; off > 0; off >>= 1) {
        const double ov = __shfl_down_sync(0xffffffffu, a.v, off);
        const long long ol = __shfl_down_sync(0xffffffffu, a.l, off);
        take<MAX>(a, ov, ol);
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    __syncthreads();                                                    // the arrays are reused from one quantity to the next
    if (lane == 0) { shv[warp] = a.v; shl[warp] = a.l; }
    __syncthreads();
    if (warp == 0) {
        a.v = lane < nwarp ? shv[lane] : 0.0; a.l = lane < nwarp ? shl[lane] : -1;
        for (int off = 16; off > 0; off >>= 1) {
            const double ov = __shfl_down_sync(0xffffffffu, a.v, off);
            const long long ol = __shfl_down_sync(0xffffffffu, a.l, off);
            take<MAX>(a, ov, ol);
        }
    }
}

__device__ __forceinline__ void fold_all(StpCtlRec &r, double *shv, long long *shl, int *shf)
{
    Best b;
    b = {r.z1, r.l1};     fold<true>(b, shv, shl);  r.z1 = b.v;   r.l1 = b.l;
    b = {r.z2, r.l2};     fold<true>(b, shv, shl);  r.z2 = b.v;   r.l2 = b.l;
    b = {r.smin, r.ls1};  fold<false>(b, shv, shl); r.smin = b.v; r.ls1 = b.l;
    b = {r.smax, r.ls2};  fold<true>(b, shv, shl);  r.smax = b.v; r.ls2 = b.l;
    b = {r.tmin, r.lt1};  fold<false>(b, shv, shl); r.tmin = b.v; r.lt1 = b.l;
    b = {r.tmax, r.lt2};  fold<true>(b, shv, shl);  r.tmax = b.v; r.lt2 = b.l;
    int f = r.flags;
    for (int off = 16; off > 0; off >>= 1) f |= __shfl_down_sync(0xffffffffu, f, off);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) shf[threadIdx.x >> 5] = f;
    __syncthreads();
    if (threadIdx.x == 0) { f = 0; for (int w = 0; w < (int)(blockDim.x >> 5); ++w) f |= shf[w]; r.flags = f; }
}

__device__ __forceinline__ StpCtlRec empty_rec()
{
    StpCtlRec r;
    r.z1 = r.z2 = r.smin = r.smax = r.tmin = r.tmax = 0.0;
    r.l1 = r.l2 = r.ls1 = r.ls2 = r.lt1 = r.lt2 = -1;
    r.flags = 0; r.pad = 0;
    return r;
}

__global__ void __launch_bounds__(256) k_stp_ctl_partial(const double *sshn, const double *un, const double *tem, const double *sal,
                                                         const double *tmask, size_t jpij, size_t n3, StpCtlRec *partial)
{
    __shared__ double shv[8];
    __shared__ long long shl[8];
    __shared__ int shf[8];
    StpCtlRec r = empty_rec();
    Best z1{0.0, -1}, z2{0.0, -1}, s1{0.0, -1}, s2{0.0, -1}, t1{0.0, -1}, t2{0.0, -1};
    int flags = 0;
    const size_t stride = (size_t)gridDim.x * blockDim.x, first = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (size_t m = first; m < jpij; m += stride) {
        const double a = fabs(sshn[m]);
        if (a != a) flags |= 1;
        take<true>(z1, a, (long long)m);
    }
    for (size_t m = first; m < n3; m += stride) {
        const double a = fabs(un[m]);
        if (a != a) flags |= 1;
        take<true>(z2, a, (long long)m);
        if (tmask[m] == 1.0) {
            flags |= 2;                                                 // lsomeoce
            const double s = sal[m], t = tem[m];
            if (s != s) flags |= 1;
            take<false>(s1, s, (long long)m); take<true>(s2, s, (long long)m);
            take<false>(t1, t, (long long)m); take<true>(t2, t, (long long)m);
        }
    }
    r.z1 = z1.v; r.l1 = z1.l; r.z2 = z2.v; r.l2 = z2.l; r.smin = s1.v; r.ls1 = s1.l; r.smax = s2.v; r.ls2 = s2.l;
    r.tmin = t1.v; r.lt1 = t1.l; r.tmax = t2.v; r.lt2 = t2.l; r.flags = flags;
    fold_all(r, shv, shl, shf);
    if (threadIdx.x == 0) partial[blockIdx.x] = r;
}

__global__ void __launch_bounds__(256) k_stp_ctl_final(const StpCtlRec *partial, int nblk, StpCtlRec *out)
{
    __shared__ double shv[8];
    __shared__ long long shl[8];
    __shared__ int shf[8];
    Best z1{0.0, -1}, z2{0.0, -1}, s1{0.0, -1}, s2{0.0, -1}, t1{0.0, -1}, t2{0.0, -1};
    int flags = 0;
    for (int b = threadIdx.x; b < nblk; b += blockDim.x) {
        const StpCtlRec p = partial[b];
        take<true>(z1, p.z1, p.l1); take<true>(z2, p.z2, p.l2);
        take<false>(s1, p.smin, p.ls1); take<true>(s2, p.smax, p.ls2);
        take<false>(t1, p.tmin, p.lt1); take<true>(t2, p.tmax, p.lt2);
        flags |= p.flags;
    }
    StpCtlRec r = empty_rec();
    r.z1 = z1.v; r.l1 = z1.l; r.z2 = z2.v; r.l2 = z2.l; r.smin = s1.v; r.ls1 = s1.l; r.smax = s2.v; r.ls2 = s2.l;
    r.tmin = t1.v; r.lt1 = t1.l; r.tmax = t2.v; r.lt2 = t2.l; r.flags = flags;
    fold_all(r, shv, shl, shf);
    if (threadIdx.x == 0) *out = r;
}

}  // namespace

#ifndef NEMO_EMU_KERNELS_ONLY          // the launcher (CUDA launch syntax) is left out of the host emulation build of tests/emu
void launch_stp_ctl(const double *sshn, const double *un, const double *tem, const double *sal, const double *tmask, size_t jpij, size_t n3,
                    StpCtlRec *partial, StpCtlRec *out, cudaStream_t s)
{
    k_stp_ctl_partial<<<kGlobSumBlocks, 256, 0, s>>>(sshn, un, tem, sal, tmask, jpij, n3, partial);
    note_launch();
    k_stp_ctl_final<<<1, 256, 0, s>>>(partial, kGlobSumBlocks, out);
    note_launch();
}

#endif  // NEMO_EMU_KERNELS_ONLY

}  // namespace nemo
