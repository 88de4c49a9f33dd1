// lbc_kernels.cu -- device execution of the compiled lbc_lnk gather plan (layout.hpp).
//
// The reference's mpp_lnk (src/OCE/LBC/mpp_lnk_generic.h90:109-336) packs strided strips into freshly ALLOCATEd
// buffers four times per call (E-W, fold, N-S) and unpacks them after blocking receives.  Here every exchange is
//   pack   : buf[peer][field][k][cell] = field[src_cell + k*jpij]        (one launch, all peers and fields)
//   move   : ncclSend/ncclRecv per remote peer (or nothing for cells whose source is on this rank)
//   unpack : field[dst_cell + k*jpij] = psgn^spow * buf[...]  and  land-value fills (one launch each)
// Cell tables are static per grid-point type; buffers are persistent.  The cell index is the fastest-varying
// thread index so buffer accesses are fully coalesced and field accesses are coalesced along ji for row strips.
#include "kernels.cuh"

namespace nemo {

void note_launch();

namespace {

__global__ void __launch_bounds__(128) k_lbc_pack(const PackJob *jobs, int nlev, size_t jpij)
{
    const PackJob jb = jobs[blockIdx.z];
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= jb.ncell) return;
    const int s = jb.src[c];
    for (int k = blockIdx.y; k < nlev; k += gridDim.y)
        jb.buf[(size_t)k * jb.ncell + c] = jb.field[(size_t)s + (size_t)k * jpij];
}

__device__ __forceinline__ double apply_sgn(double v, double sgn, int spow)
{
    for (int p = 0; p < spow; ++p) v = sgn * v;      // SGN_IN(jf) * ARRAY_IN(...), lbc_nfd_generic.h90:83
    return v;
}

__global__ void __launch_bounds__(128) k_lbc_unpack(const UnpackJob *jobs, int nlev, size_t jpij)
{
    const UnpackJob jb = jobs[blockIdx.z];
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= jb.ncell) return;
    const int d = jb.dst[c], sp = jb.spow[c];
    for (int k = blockIdx.y; k < nlev; k += gridDim.y)
        jb.field[(size_t)d + (size_t)k * jpij] = apply_sgn(jb.buf[(size_t)k * jb.ncell + c], jb.sgn, sp);
}

__global__ void __launch_bounds__(128) k_lbc_fill(const FillJob *jobs, int nlev, size_t jpij)
{
    const FillJob jb = jobs[blockIdx.z];
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= jb.ncell) return;
    const int d = jb.dst[c];
    const double v = apply_sgn(jb.zland, jb.sgn, jb.spow[c]);
    for (int k = blockIdx.y; k < nlev; k += gridDim.y)
        jb.field[(size_t)d + (size_t)k * jpij] = v;
}

inline dim3 job_grid(int njobs, int maxcell, int nlev)
{
    return dim3((unsigned)((maxcell + 127) / 128), (unsigned)(nlev < 256 ? nlev : 256), (unsigned)njobs);
}

}  // namespace

void launch_lbc_pack(const PackJob *jobs_dev, int njobs, int maxcell, int nlev, size_t jpij, cudaStream_t s)
{
    if (njobs == 0 || maxcell == 0) return;
    k_lbc_pack<<<job_grid(njobs, maxcell, nlev), 128, 0, s>>>(jobs_dev, nlev, jpij); note_launch();
}
void launch_lbc_unpack(const UnpackJob *jobs_dev, int njobs, int maxcell, int nlev, size_t jpij, cudaStream_t s)
{
    if (njobs == 0 || maxcell == 0) return;
    k_lbc_unpack<<<job_grid(njobs, maxcell, nlev), 128, 0, s>>>(jobs_dev, nlev, jpij); note_launch();
}
void launch_lbc_fill(const FillJob *jobs_dev, int njobs, int maxcell, int nlev, size_t jpij, cudaStream_t s)
{
    if (njobs == 0 || maxcell == 0) return;
    k_lbc_fill<<<job_grid(njobs, maxcell, nlev), 128, 0, s>>>(jobs_dev, nlev, jpij); note_launch();
}

}  // namespace nemo
