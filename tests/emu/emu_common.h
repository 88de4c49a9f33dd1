// CPU emulation harness for the column kernels of mus_kernels.cu / cen_kernels.cu (TEST INFRASTRUCTURE ONLY).
//
// The kernels are plain fp64 C++ apart from the thread indices, so the very same source is compiled here for the host
// (g++ -ffp-contract=off == nvcc -fmad=false: IEEE +,-,*,/ only) with blockIdx / threadIdx provided as globals and the
// grid walked serially.  Purpose: exercise kernel code paths the GPU suite does not reach at its small jpk -- the jk loop
// split in chunks across blockIdx.y (what production sizes with jpk = 75 use) -- against the oracle, without a GPU.
// Nothing here is linked into libnemo_fct.so.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstring>
#include <cuda_runtime.h>

// CUDA function-space and launch qualifiers mean nothing on the host
#undef __global__
#undef __device__
#undef __host__
#undef __forceinline__
#undef __launch_bounds__
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define NEMO_NOINLINE

struct EmuIdx { unsigned x, y, z; };
static thread_local EmuIdx blockIdx, threadIdx, blockDim, gridDim;      // thread_local: emu_block.h runs one host thread per CUDA thread
using std::min;
using std::max;
#include <type_traits>
// high word of a double (device intrinsic of the same name)
static inline int __double2hiint(double x) { long long b; std::memcpy(&b, &x, 8); return (int)(b >> 32); }
#define NEMO_EMU_KERNELS_ONLY 1

// walk the launch grid of a column kernel serially: blockIdx.x = column block, .y = jk chunk, .z = tracer
template <typename Args, typename K>
static void emu_run_grid(const Args &a, K kernel, int ncol, int nkchunk, int kjpt)
{
    const int nthreads = 128;
    blockDim = {(unsigned)nthreads, 1, 1};
    const int nbx = (ncol + nthreads - 1) / nthreads;
    for (int bz = 0; bz < kjpt; ++bz)
        for (int by = 0; by < nkchunk; ++by)
            for (int bx = 0; bx < nbx; ++bx)
                for (int t = 0; t < nthreads; ++t) {
                    blockIdx = {(unsigned)bx, (unsigned)by, (unsigned)bz};
                    threadIdx = {(unsigned)t, 0, 0};
                    kernel(a);
                }
}
