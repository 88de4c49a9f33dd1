// dev probe: one TMA 3-D box load of fp64 tiles, descriptor variants.  nvcc -gencode arch=compute_100a,code=sm_100a tma_probe.cu -o tma_probe
#include <cuda.h>
#include <cudaTypedefs.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
struct Maps { CUtensorMap m; };
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
template <int BW, int BH>
__global__ void probe(const __grid_constant__ Maps maps, double *out, int x, int y, int z)
{
    extern __shared__ __align__(128) unsigned char sm[];
    unsigned long long *bar = reinterpret_cast<unsigned long long *>(sm + BW * BH * 8);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(smem_u32(bar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(BW * BH * 8) : "memory");
        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];\n"
                     ::"r"(smem_u32(sm)), "l"(&maps.m), "r"(smem_u32(bar)), "r"(x), "r"(y), "r"(z) : "memory");
    }
    __syncthreads();
    asm volatile("{\n .reg .pred p;\n W: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n @p bra D;\n bra W;\n D:\n}\n" ::"r"(smem_u32(bar)) : "memory");
    const double *s = reinterpret_cast<const double *>(sm);
    for (int i = threadIdx.x; i < BW * BH; i += blockDim.x) out[i] = s[i];
}
template <int BW, int BH>
int run(const char *name, CUtensorMapDataType dt, int scale, CUtensorMapL2promotion l2)
{
    const int jpi = 76, jpj = 45, nlev = 22;
    std::vector<double> h((size_t)jpi * jpj * nlev);
    for (size_t i = 0; i < h.size(); ++i) h[i] = (double)i;
    double *d, *o; cudaMalloc(&d, h.size() * 8); cudaMalloc(&o, BW * BH * 8);
    cudaMemcpy(d, h.data(), h.size() * 8, cudaMemcpyHostToDevice);
    void *p = nullptr; cudaDriverEntryPointQueryResult q;
    if (getenv("BYVER")) cudaGetDriverEntryPointByVersion("cuTensorMapEncodeTiled", &p, 12000, cudaEnableDefault, &q);
    else cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    auto enc = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
    Maps mp;
    cuuint64_t dims[3] = {(cuuint64_t)jpi * scale, (cuuint64_t)jpj, (cuuint64_t)nlev};
    cuuint64_t strides[2] = {(cuuint64_t)jpi * 8, (cuuint64_t)jpi * jpj * 8};
    cuuint32_t box[3] = {(cuuint32_t)BW * scale, BH, 1}, es[3] = {1, 1, 1};
    CUresult r = enc(&mp.m, dt, 3, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, l2, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("%s: encode failed %d\n", name, (int)r); return 1; }
    const int x = 0, y = 15, z = 3;
    probe<BW, BH><<<1, 128, BW * BH * 8 + 64>>>(mp, o, x * scale, y, z);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%s: kernel failed: %s\n", name, cudaGetErrorString(e)); return 2; }
    std::vector<double> got(BW * BH); cudaMemcpy(got.data(), o, BW * BH * 8, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int j = 0; j < BH; ++j) for (int i = 0; i < BW; ++i) {
        int gi = x + i, gj = y + j;
        double want = (gi < 0 || gi >= jpi || gj < 0 || gj >= jpj) ? 0.0 : h[((size_t)z * jpj + gj) * jpi + gi];
        if (got[j * BW + i] != want) ++bad;
    }
    printf("%s: ok, %d mismatches\n", name, bad);
    return 0;
}
int main(int argc, char **argv)
{
    int v = argc > 1 ? atoi(argv[1]) : 0;
    if (v == 0) return run<32, 8>("f64 box32x8 l2-128", CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 1, CU_TENSOR_MAP_L2_PROMOTION_L2_128B);
    if (v == 1) return run<36, 12>("f64 box36x12 l2-128", CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 1, CU_TENSOR_MAP_L2_PROMOTION_L2_128B);
    if (v == 2) return run<36, 12>("f64 box36x12 l2-none", CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 1, CU_TENSOR_MAP_L2_PROMOTION_NONE);
    if (v == 3) return run<36, 12>("u32x2 box36x12", CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, CU_TENSOR_MAP_L2_PROMOTION_L2_128B);
    if (v == 4) return run<32, 8>("u32x2 box32x8", CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, CU_TENSOR_MAP_L2_PROMOTION_L2_128B);
    if (v == 5) return run<32, 8>("u64 box32x8", CU_TENSOR_MAP_DATA_TYPE_UINT64, 1, CU_TENSOR_MAP_L2_PROMOTION_L2_128B);
    return 0;
}
