// CUDA programming guide TMA example (2D), to check the toolchain/box
#include <cuda.h>
#include <cudaTypedefs.h>
#include <cuda/barrier>
#include <cstdio>
#include <cstdlib>
#include <vector>
using barrier = cuda::barrier<cuda::thread_scope_block>;
namespace cde = cuda::device::experimental;
constexpr int SMEM_W = 32, SMEM_H = 8;
__global__ void kernel(const __grid_constant__ CUtensorMap tensor_map, double *out, int x, int y)
{
    __shared__ alignas(128) double smem_buffer[SMEM_H][SMEM_W];
#pragma nv_diag_suppress static_var_with_dynamic_init
    __shared__ barrier bar;
    if (threadIdx.x == 0) { init(&bar, blockDim.x); cde::fence_proxy_async_shared_cta(); }
    __syncthreads();
    barrier::arrival_token token;
    if (threadIdx.x == 0) {
        cde::cp_async_bulk_tensor_2d_global_to_shared(&smem_buffer, &tensor_map, x, y, bar);
        token = cuda::device::barrier_arrive_tx(bar, 1, sizeof(smem_buffer));
    } else token = bar.arrive();
    bar.wait(std::move(token));
    for (int i = threadIdx.x; i < SMEM_W * SMEM_H; i += blockDim.x) out[i] = (&smem_buffer[0][0])[i];
}
int main(int argc, char **argv)
{
    const int W = 76, Hh = 45;
    std::vector<double> h((size_t)W * Hh);
    for (size_t i = 0; i < h.size(); ++i) h[i] = (double)i;
    double *d, *o; cudaMalloc(&d, h.size() * 8); cudaMalloc(&o, SMEM_W * SMEM_H * 8);
    cudaMemcpy(d, h.data(), h.size() * 8, cudaMemcpyHostToDevice);
    void *p = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    auto enc = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
    CUtensorMap tm;
    cuuint64_t dims[2] = {W, Hh}; cuuint64_t strides[1] = {W * 8};
    cuuint32_t box[2] = {SMEM_W, SMEM_H}, es[2] = {1, 1};
    CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                     CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode %d\n", (int)r);
    int x = argc > 1 ? atoi(argv[1]) : 4, y = 3;
    kernel<<<1, 128>>>(tm, o, x, y);
    cudaError_t e = cudaDeviceSynchronize();
    printf("x=%d kernel: %s\n", x, cudaGetErrorString(e));
    if (e == cudaSuccess) {
        std::vector<double> got(SMEM_W * SMEM_H); cudaMemcpy(got.data(), o, got.size() * 8, cudaMemcpyDeviceToHost);
        int bad = 0;
        for (int j = 0; j < SMEM_H; ++j) for (int i = 0; i < SMEM_W; ++i) {
            int gi = x + i, gj = y + j; double want = (gi < 0 || gi >= W || gj < 0 || gj >= Hh) ? 0.0 : h[(size_t)gj * W + gi];
            if (got[j * SMEM_W + i] != want) ++bad;
        }
        printf("mismatches %d\n", bad);
    }
    return 0;
}
