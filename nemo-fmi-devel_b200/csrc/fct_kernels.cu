// fct_kernels.cu -- hand-written fp64 CUDA kernels (sm_100a) of the FCT tracer-advection step.
//
// What is computed follows src/OCE/TRA/traadv_fct.F90 (tra_adv_fct :54-327, nonosc :330-428, interp_4th_cpt
// :517-616); how it is computed does not: the reference's ~60 whole-array sweeps per tracer become a handful of
// column-marching kernels.  One thread owns one interior (ji,jj) column (chunk) and marches down jk, so every
// jk+-1 operand lives in registers; ji+-1 / jj+-1 operands are neighbouring threads' cache lines (reads are
// coalesced along ji).  All tracers are batched in one launch (blockIdx.z).  Memory-bound stencil: no tensor
// cores (fp64, ~2.5 flop/B).
//
// Bit-exactness: each expression keeps the reference's operation order and the file is compiled with
// -fmad=false, so results equal the IEEE-strict CPU restatement bit for bit.  MAX/MIN are written as selects
// (a > b ? a : b), SIGN(0.5,x) follows the key_nosignedzero override (x >= 0 -> +0.5, lib_fortran.F90:339-351).
#include "kernels.cuh"

#include <cuda.h>
#include <cudaTypedefs.h>

#include <algorithm>
#include <atomic>
#include <cstdlib>
#include <cstring>
#include <type_traits>

namespace nemo {

static std::atomic<long long> g_launches{0};
long long kernel_launch_count() { return g_launches.load(); }
void note_launch() { g_launches.fetch_add(1); }   // also called from glob_sum.cu

namespace {

constexpr int kThreads = 128;

#include "nonosc_final.cuh"   // dmax / dmin, bup_bdo, limit_coef_sel, k_fct_nonosc_final

// cp.async (LDGSTS) helpers: 8-byte asynchronous global -> shared copies, grouped
__device__ __forceinline__ void cp_async8(double *smem_dst, const double *gsrc)
{
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(s), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }


#include "fct_column_kernels.cuh"   // reference-structured column kernels, interp_4th_cpt, transports, trend hook

#include "fct_inner_kernel.cuh"   // Masks<FROM_T>, k_fct_low_antidiff_inner (cp.async prefetch ring)

// ---- PTX helpers of the TMA-tiled kernels: mbarrier phases and bulk tensor copies ---------------------------------
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity)
{
    asm volatile(
        "{\n .reg .pred p;\n WAIT_%=:\n"
        " mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        " @p bra DONE_%=;\n bra WAIT_%=;\n DONE_%=:\n}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_3d(void *dst, const CUtensorMap *map, unsigned long long *bar, int x, int y, int z)
{
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];\n"
                 ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(x), "r"(y), "r"(z) : "memory");
}
__device__ __forceinline__ void mbar_inval(unsigned long long *bar)
{
    asm volatile("mbarrier.inval.shared::cta.b64 [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
// makes the initialised mbarriers visible to the async proxy (TMA) before the first copy is issued
__device__ __forceinline__ void mbar_init_fence()
{
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
}

#include "fct_tma_kernel.cuh"     // tile geometry, TileMaps, k_fct_low_antidiff_tma

#include "fct_fused_kernel.cuh"   // k_fct_fused: P1-P8 in one kernel (schedule 4); div_rn
#include "cpt_tiled_kernel.cuh"   // k_interp_4th_cpt_tiled: TMA-fed Thomas solve with the forward sweep in shared memory

// ---- TMA-fed variant of the fused nonosc + final kernel ---------------------------------------------------------
// Same tile, same arithmetic; the six streamed arrays (ptb, zwi, tmask, zwz of level jk+1; zwx, zwy of level jk) arrive
// as 32x16 boxes through a 2-stage TMA ring one level ahead instead of per-thread loads consumed on arrival (ncu on the
// LDG variant: 2 barriers per level with every warp waiting for HBM).  zbup/zbdo/paa/pbb and the betas are exchanged
// through single-buffered shared planes (two barriers per level order the reuse).  Needs even jpi and an odd out.i0.
enum { NQ_PTB = 0, NQ_ZWI, NQ_TM, NQ_PCC, NQ_PAA, NQ_PBB, NQ_COUNT };
constexpr int kNqBoxBytes = NX * NY * 8, kNqStageBytes = NQ_COUNT * kNqBoxBytes;
struct NonoscMaps { CUtensorMap m[NQ_COUNT]; };

__global__ void __launch_bounds__(NX * NY, 2) k_fct_nonosc_final_tma(const FctArgs a, const __grid_constant__ NonoscMaps maps)
{
    extern __shared__ __align__(128) unsigned char nq_smem[];
    double(*sA)[NY][NX] = reinterpret_cast<double(*)[NY][NX]>(nq_smem + 2 * kNqStageBytes);                     // zbup, zbdo, paa, pbb
    double(*sB)[NY][NX] = reinterpret_cast<double(*)[NY][NX]>(nq_smem + 2 * kNqStageBytes + 4 * kNqBoxBytes);   // zbetup, zbetdo
    unsigned long long *full = reinterpret_cast<unsigned long long *>(nq_smem + 2 * kNqStageBytes + 6 * kNqBoxBytes);
    const CUtensorMap *mq = maps.m;
    const int tx = threadIdx.x % NX, ty = threadIdx.x / NX;
    const int ox = NX - 2 * NHALO, oy = NY - 2 * NHALO;
    const int bx0 = a.out.i0 - 1 + (int)blockIdx.x * ox - NHALO, by0 = a.out.j0 - 1 + (int)blockIdx.y * oy - NHALO;   // 0-based box origin
    const int gi = bx0 + tx + 1, gj = by0 + ty + 1;
    const int ji = min(gi, a.out.i1 + NHALO), jj = min(gj, a.out.j1 + NHALO);
    const bool is_out = tx >= NHALO && tx < NX - NHALO && ty >= NHALO && ty < NY - NHALO && gi <= a.out.i1 && gj <= a.out.j1;
    const bool is_beta = tx >= 1 && tx < NX - 1 && ty >= 1 && ty < NY - 1;
    const int jn = (int)blockIdx.z;
    const size_t toff = (size_t)jn * a.n3;
    double *__restrict__ pta = a.pta + toff;
    const double *__restrict__ e3t_n = a.e3t_n;
    const int jpi = a.jpi, jpk = a.jpk;
    const size_t jpij = a.jpij;
    const size_t c2 = (size_t)(jj - 1) * jpi + (ji - 1);
    const double zrtrn = 1.e-15;
    const double e12 = a.e1e2t[c2], r1 = a.r1_e1e2t[c2];
    const double p2dt = a.p2dt;
    const int cell = ty * NX + tx;

    auto issue = [=](int lev) {                          // one thread: the six boxes of level `lev` -> stage lev % 2
        unsigned char *st = nq_smem + (size_t)(lev & 1) * kNqStageBytes;
        unsigned long long *bar = &full[lev & 1];
        mbar_expect_tx(bar, kNqStageBytes);
        const int z3 = lev - 1, z4 = jn * jpk + lev - 1;
        tma_load_3d(st + NQ_PTB * kNqBoxBytes, &mq[NQ_PTB], bar, bx0, by0, z4);
        tma_load_3d(st + NQ_ZWI * kNqBoxBytes, &mq[NQ_ZWI], bar, bx0, by0, z4);
        tma_load_3d(st + NQ_TM * kNqBoxBytes, &mq[NQ_TM], bar, bx0, by0, z3);
        tma_load_3d(st + NQ_PCC * kNqBoxBytes, &mq[NQ_PCC], bar, bx0, by0, z4);
        tma_load_3d(st + NQ_PAA * kNqBoxBytes, &mq[NQ_PAA], bar, bx0, by0, z4);
        tma_load_3d(st + NQ_PBB * kNqBoxBytes, &mq[NQ_PBB], bar, bx0, by0, z4);
    };
    auto stg = [&](int lev, int q) -> const double * { return reinterpret_cast<const double *>(nq_smem + (size_t)(lev & 1) * kNqStageBytes + (size_t)q * kNqBoxBytes); };

    if (threadIdx.x == 0) {
        mbar_init(&full[0], 1); mbar_init(&full[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
        issue(1);
        issue(2);                                        // jpk >= 3
    }
    __syncthreads();
    mbar_wait(&full[1], 0);                              // level 1
    double up_m, do_m, up_c, do_c, up_p = 0.0, do_p = 0.0;
    double aft_c = stg(1, NQ_ZWI)[cell];
    bup_bdo(stg(1, NQ_PTB)[cell], aft_c, stg(1, NQ_TM)[cell], up_c, do_c);
    up_m = up_c; do_m = do_c;
    double pcc_k = stg(1, NQ_PCC)[cell];
    double bup_mm = 0.0, bdo_mm = 0.0, bup_m = 0.0, bdo_m = 0.0;
    double paa_m = 0.0, pbb_m = 0.0, pcc_m = 0.0, paa_w_m = 0.0, pbb_s_m = 0.0;
    double e3n_c = e3t_n[c2], e3n_m = 1.0, pta_m = 0.0;

    for (int k = 1; k <= jpk; ++k) {
        const size_t o = c2 + (size_t)(k - 1) * jpij;
        const bool lev = k <= jpk - 1;
        double paa_c = 0.0, pbb_c = 0.0, pcc_p = 0.0, aft_p = 0.0, e3n_p = 1.0, pta_c = 0.0;
        if (lev) {
            e3n_p = e3t_n[o + jpij];                     // consumed one level later
            if (is_out) pta_c = pta[o];
            mbar_wait(&full[(k + 1) & 1], (k / 2) & 1);  // level k+1: use index (k+1-1)/2 of its stage
            aft_p = stg(k + 1, NQ_ZWI)[cell];
            bup_bdo(stg(k + 1, NQ_PTB)[cell], aft_p, stg(k + 1, NQ_TM)[cell], up_p, do_p);
            pcc_p = stg(k + 1, NQ_PCC)[cell];
            paa_c = stg(k, NQ_PAA)[cell]; pbb_c = stg(k, NQ_PBB)[cell];
            sA[0][ty][tx] = up_c; sA[1][ty][tx] = do_c; sA[2][ty][tx] = paa_c; sA[3][ty][tx] = pbb_c;
        }
        __syncthreads();                                 // sA(k) visible; stage k%2 no longer read
        if (threadIdx.x == 0 && k + 2 <= jpk) issue(k + 2);
        double bup_c = 0.0, bdo_c = 0.0, paa_w = 0.0, pbb_s = 0.0;
        if (lev && is_beta) {
            const double zup = dmax(dmax(dmax(dmax(dmax(dmax(up_c, sA[0][ty][tx - 1]), sA[0][ty][tx + 1]), sA[0][ty - 1][tx]), sA[0][ty + 1][tx]), up_m), up_p);
            const double zdo = dmin(dmin(dmin(dmin(dmin(dmin(do_c, sA[1][ty][tx - 1]), sA[1][ty][tx + 1]), sA[1][ty - 1][tx]), sA[1][ty + 1][tx]), do_m), do_p);
            paa_w = sA[2][ty][tx - 1]; pbb_s = sA[3][ty - 1][tx];
            const double zpos = dmax(0., paa_w) - dmin(0., paa_c) + dmax(0., pbb_s) - dmin(0., pbb_c)
                              + dmax(0., pcc_p) - dmin(0., pcc_k);
            const double zneg = dmax(0., paa_c) - dmin(0., paa_w) + dmax(0., pbb_c) - dmin(0., pbb_s)
                              + dmax(0., pcc_k) - dmin(0., pcc_p);
            const double zbt = e12 * e3n_c / p2dt;
            bup_c = (zup - aft_c) / (zpos + zrtrn) * zbt;
            bdo_c = (aft_c - zdo) / (zneg + zrtrn) * zbt;
        }
        if (k >= 2 && is_out) {
            // final trend of level kk = k-1: the neighbours' betas(kk) are in sB (written at the end of iteration k-1)
            const int kk = k - 1;
            const double bup_e = sB[0][ty][tx + 1], bdo_e = sB[1][ty][tx + 1], bup_w = sB[0][ty][tx - 1], bdo_w = sB[1][ty][tx - 1];
            const double bup_n = sB[0][ty + 1][tx], bdo_n = sB[1][ty + 1][tx], bup_s = sB[0][ty - 1][tx], bdo_s = sB[1][ty - 1][tx];
            const double lx_e = paa_m * limit_coef_sel(paa_m, bdo_m, bup_e, bup_m, bdo_e);
            const double lx_w = paa_w_m * limit_coef_sel(paa_w_m, bdo_w, bup_m, bup_w, bdo_m);
            const double ly_n = pbb_m * limit_coef_sel(pbb_m, bdo_m, bup_n, bup_m, bdo_n);
            const double ly_s = pbb_s_m * limit_coef_sel(pbb_s_m, bdo_s, bup_m, bup_s, bdo_m);
            const double lz_t = (kk == 1) ? pcc_m : pcc_m * limit_coef_sel(pcc_m, bdo_m, bup_mm, bup_m, bdo_mm);
            const double lz_b = pcc_k * limit_coef_sel(pcc_k, bdo_c, bup_m, bup_c, bdo_m);
            pta[o - jpij] = pta_m - (lx_e - lx_w + ly_n - ly_s + lz_t - lz_b) * r1 / e3n_m;
        }
        __syncthreads();                                 // every read of sB(k-1) and sA(k) is done
        sB[0][ty][tx] = bup_c; sB[1][ty][tx] = bdo_c;
        up_m = up_c; do_m = do_c; up_c = up_p; do_c = do_p; aft_c = aft_p;
        bup_mm = bup_m; bdo_mm = bdo_m; bup_m = bup_c; bdo_m = bdo_c;
        paa_m = paa_c; pbb_m = pbb_c; paa_w_m = paa_w; pbb_s_m = pbb_s; pcc_m = pcc_k; pcc_k = pcc_p;
        e3n_m = e3n_c; e3n_c = e3n_p; pta_m = pta_c;
    }
}

// ---- self-test of div_rn ----------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long splitmix64(unsigned long long &x)
{
    unsigned long long z = (x += 0x9e3779b97f4a7c15ull);
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull; z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
    return z ^ (z >> 31);
}
// a double with random sign and mantissa and an exponent field drawn from [elo, ehi]
__device__ __forceinline__ double random_double(unsigned long long &st, int elo, int ehi)
{
    const unsigned long long r = splitmix64(st), e = (unsigned long long)(elo + (int)(splitmix64(st) % (unsigned)(ehi - elo + 1)));
    return __longlong_as_double((long long)((r & 0x800fffffffffffffull) | (e << 52)));
}
__global__ void k_div_selftest(long long n, unsigned long long seed, unsigned long long *nbad)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    unsigned long long st = seed + 0x1234567ull * (unsigned long long)i;
    // exponent-field ranges: ordinary; the FCT ranges (1e-15 .. 1e40); near the fast-path limits; subnormal / zero; huge / Inf / NaN
    const int cls[8][4] = {{900, 1150, 900, 1150}, {970, 1160, 970, 1160}, {1, 120, 900, 1100}, {0, 0, 1, 2046},
                           {1, 2046, 0, 0}, {1960, 2047, 1, 2047}, {1, 2047, 1960, 2047}, {0, 2047, 0, 2047}};
    unsigned long long bad = 0;
    for (int c = 0; c < 8; ++c) {
        double x = random_double(st, cls[c][0], cls[c][1]), y = random_double(st, cls[c][2], cls[c][3]);
        if (c == 3 && (i & 3) == 0) x = (i & 4) ? 0.0 : -0.0;
        if (c == 4 && (i & 3) == 0) y = (i & 4) ? 0.0 : -0.0;
        const double q1 = div_rn(x, y), q2 = x / y;
        const bool same = (q1 != q1 && q2 != q2) || __double_as_longlong(q1) == __double_as_longlong(q2);
        bad += same ? 0 : 1;
    }
    if (bad) atomicAdd(nbad, bad);
}

inline dim3 column_grid(const FctArgs &a)
{
    const long long ncol = a.reg.ncol();
    return dim3((unsigned)((ncol + kThreads - 1) / kThreads), (unsigned)a.nkchunk, (unsigned)a.kjpt);
}

}  // namespace

// Opt a kernel in to more than 48 KB of dynamic shared memory.  The attribute belongs to the (function, device) pair, so it
// is remembered per device: contexts of several GPUs may live in one process.
constexpr int kMaxDevices = 64;
template <typename Kernel>
static void allow_dynamic_smem(Kernel kernel, size_t smem, std::atomic<bool> (&done)[kMaxDevices])
{
    int dev = 0;
    cudaGetDevice(&dev);
    const bool known = dev >= 0 && dev < kMaxDevices;
    if (known && done[dev]) return;
    cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (known) done[dev] = true;
}

long long division_selftest(long long n, unsigned long long seed, cudaStream_t s)
{
    unsigned long long *d = nullptr, h = 0;
    if (cudaMalloc(&d, sizeof h) != cudaSuccess) return -1;
    cudaMemsetAsync(d, 0, sizeof h, s);
    k_div_selftest<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(n, seed, d);
    note_launch();
    const cudaError_t e = cudaMemcpyAsync(&h, d, sizeof h, cudaMemcpyDeviceToHost, s);
    const cudaError_t e2 = cudaStreamSynchronize(s);
    cudaFree(d);
    return (e == cudaSuccess && e2 == cudaSuccess) ? (long long)h : -1;
}

void launch_fct_laplacian(const FctArgs &a, cudaStream_t s)
{
    k_fct_laplacian<<<column_grid(a), kThreads, 0, s>>>(a); note_launch();
}

void launch_fct_low_antidiff(const FctArgs &a, cudaStream_t s)
{
    const dim3 g = column_grid(a);
    if (a.kn_fct_h == 2 && a.kn_fct_v == 2) k_fct_low_antidiff<2, 2><<<g, kThreads, 0, s>>>(a);
    else if (a.kn_fct_h == 2)               k_fct_low_antidiff<2, 4><<<g, kThreads, 0, s>>>(a);
    else if (a.kn_fct_v == 2)               k_fct_low_antidiff<4, 2><<<g, kThreads, 0, s>>>(a);
    else                                    k_fct_low_antidiff<4, 4><<<g, kThreads, 0, s>>>(a);
    note_launch();
}

void launch_fct_low_antidiff_inner(const FctArgs &a, cudaStream_t s)
{
    const dim3 g = column_grid(a);
    const bool ft = (a.masks_from_t & 1) != 0;
    const size_t smem = (size_t)kPfStages * F_LOW_COUNT * kThreads * sizeof(double);   // 33 KB: below the 48 KB opt-in limit
#define LAI(H, V) (ft ? k_fct_low_antidiff_inner<H, V, true><<<g, kThreads, smem, s>>>(a) : k_fct_low_antidiff_inner<H, V, false><<<g, kThreads, smem, s>>>(a))
    if (a.kn_fct_h == 2 && a.kn_fct_v == 2) LAI(2, 2);
    else if (a.kn_fct_h == 2)               LAI(2, 4);
    else if (a.kn_fct_v == 2)               LAI(4, 2);
    else                                    LAI(4, 4);
#undef LAI
    note_launch();
}

// cuTensorMapEncodeTiled through the runtime's driver entry point (no link-time dependency on libcuda)
static PFN_cuTensorMapEncodeTiled_v12000 tensor_map_encoder()
{
    static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
    }
    return fn;
}
static bool make_tile_map(CUtensorMap *m, const double *base, int jpi, int jpj, long long nlev, int bw, int bh)
{
    auto enc = tensor_map_encoder();
    if (!enc || (reinterpret_cast<uintptr_t>(base) & 15u)) return false;
    cuuint64_t dims[3] = {(cuuint64_t)jpi, (cuuint64_t)jpj, (cuuint64_t)nlev};
    cuuint64_t strides[2] = {(cuuint64_t)jpi * 8, (cuuint64_t)jpi * jpj * 8};
    cuuint32_t box[3] = {(cuuint32_t)bw, (cuuint32_t)bh, 1};
    cuuint32_t es[3] = {1, 1, 1};
    return enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, const_cast<double *>(base), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

bool launch_fct_low_antidiff_tma(const FctArgs &a, cudaStream_t s)
{
    if ((a.jpi & 1) || a.reg.n != 1 || (a.reg.r[0].i0 & 1)) return false;   // 16-byte strides and even box origins
    TileMaps tm;
    const long long n4 = (long long)a.jpk * a.kjpt, n3 = a.jpk;
    bool ok = make_tile_map(&tm.h[TH_PTB], a.ptb, a.jpi, a.jpj, n4, TBW, TBH) && make_tile_map(&tm.h[TH_PTN], a.ptn, a.jpi, a.jpj, n4, TBW, TBH) &&
              make_tile_map(&tm.h[TH_TM], a.tmask, a.jpi, a.jpj, n3, TBW, TBH) && make_tile_map(&tm.h[TH_PUN], a.pun, a.jpi, a.jpj, n3, TBW, TBH) &&
              make_tile_map(&tm.h[TH_PVN], a.pvn, a.jpi, a.jpj, n3, TBW, TBH) && make_tile_map(&tm.p[TP_PTA], a.pta, a.jpi, a.jpj, n4, TPW, TTY) &&
              make_tile_map(&tm.p[TP_ZTW], a.kn_fct_v == 4 ? a.ztw : a.pta, a.jpi, a.jpj, n4, TPW, TTY) &&
              make_tile_map(&tm.p[TP_PWN], a.pwn, a.jpi, a.jpj, n3, TPW, TTY) && make_tile_map(&tm.p[TP_E3B], a.e3t_b, a.jpi, a.jpj, n3, TPW, TTY) &&
              make_tile_map(&tm.p[TP_E3N], a.e3t_n, a.jpi, a.jpj, n3, TPW, TTY) && make_tile_map(&tm.p[TP_E3A], a.e3t_a, a.jpi, a.jpj, n3, TPW, TTY);
    if (!ok) return false;
    const Rect rc = a.reg.r[0];
    const size_t smem = (size_t)TSTAGES * kTileStageBytes + 64;
    const dim3 g((unsigned)(((rc.i1 - rc.i0 + 1 + TTX - 1) / TTX) * a.kjpt), (unsigned)((rc.j1 - rc.j0 + 1 + TTY - 1) / TTY), (unsigned)a.nkchunk);
    const bool ft = (a.masks_from_t & 1) != 0;
#define LAT(H, V, F) do { static std::atomic<bool> done[kMaxDevices]; allow_dynamic_smem(k_fct_low_antidiff_tma<H, V, F>, smem, done); \
                          k_fct_low_antidiff_tma<H, V, F><<<g, TTX * TTY, smem, s>>>(a, tm, rc); } while (0)
#define LAT2(H, V) do { if (ft) LAT(H, V, true); else LAT(H, V, false); } while (0)
    if (a.kn_fct_h == 2 && a.kn_fct_v == 2) LAT2(2, 2);
    else if (a.kn_fct_h == 2)               LAT2(2, 4);
    else if (a.kn_fct_v == 2)               LAT2(4, 2);
    else                                    LAT2(4, 4);
#undef LAT2
#undef LAT
    note_launch();
    return true;
}

bool launch_fct_nonosc_final_tma(const FctArgs &a, cudaStream_t s)
{
    const int ni = a.out.i1 - a.out.i0 + 1, nj = a.out.j1 - a.out.j0 + 1;
    if (ni <= 0 || nj <= 0) return true;
    if ((a.jpi & 1) || !(a.out.i0 & 1) || a.out.i0 - 1 - NHALO < 0 || a.out.j0 - 1 - NHALO < 0 || a.jpk < 3) return false;
    NonoscMaps tm;
    const long long n4 = (long long)a.jpk * a.kjpt, n3 = a.jpk;
    const bool ok = make_tile_map(&tm.m[NQ_PTB], a.ptb, a.jpi, a.jpj, n4, NX, NY) && make_tile_map(&tm.m[NQ_ZWI], a.zwi, a.jpi, a.jpj, n4, NX, NY) &&
                    make_tile_map(&tm.m[NQ_TM], a.tmask, a.jpi, a.jpj, n3, NX, NY) && make_tile_map(&tm.m[NQ_PCC], a.zwz, a.jpi, a.jpj, n4, NX, NY) &&
                    make_tile_map(&tm.m[NQ_PAA], a.zwx, a.jpi, a.jpj, n4, NX, NY) && make_tile_map(&tm.m[NQ_PBB], a.zwy, a.jpi, a.jpj, n4, NX, NY);
    if (!ok) return false;
    static std::atomic<bool> done[kMaxDevices];
    const size_t smem = (size_t)2 * kNqStageBytes + 6 * kNqBoxBytes + 64;
    allow_dynamic_smem(k_fct_nonosc_final_tma, smem, done);
    const int ox = NX - 2 * NHALO, oy = NY - 2 * NHALO;
    const dim3 g((unsigned)((ni + ox - 1) / ox), (unsigned)((nj + oy - 1) / oy), (unsigned)a.kjpt);
    k_fct_nonosc_final_tma<<<g, NX * NY, smem, s>>>(a, tm);
    note_launch();
    return true;
}

bool prepare_fct_fused(const FctArgs &a, TmaMapCache *cache)
{
    const int ni = a.out.i1 - a.out.i0 + 1, nj = a.out.j1 - a.out.j0 + 1;
    if (ni <= 0 || nj <= 0) return false;
    // 16-byte global strides, even box origins >= 0, masks as tmask products (the kernel derives umask/vmask/wmask)
    if ((a.jpi & 1) || !(a.out.i0 & 1) || a.out.i0 - 1 - FHALO - 2 < 0 || a.out.j0 - 1 - FHALO - 1 < 0 || a.jpk < 3 || (a.masks_from_t & 3) != 3) return false;
    static_assert(sizeof(FusedMaps) == 11 * 128, "tensor-map cache size");
    const void *key[12] = {a.ptb, a.ptn, a.tmask, a.pun, a.pvn, a.pta, a.kn_fct_v == 4 ? a.ztw : a.pta, a.pwn, a.e3t_b, a.e3t_n, a.e3t_a, nullptr};
    const int dims[4] = {a.jpi, a.jpj, a.jpk, a.kjpt};
    if (!cache->valid || memcmp(cache->key, key, sizeof key) || memcmp(cache->dims, dims, sizeof dims)) {
        FusedMaps tm;
        const long long n4 = (long long)a.jpk * a.kjpt, n3 = a.jpk;
        const double *hb[FH_COUNT] = {a.ptb, a.ptn, a.tmask, a.pun, a.pvn};
        const long long hn[FH_COUNT] = {n4, n4, n3, n3, n3};
        const double *pb[FP_COUNT] = {a.pta, a.kn_fct_v == 4 ? a.ztw : a.pta, a.pwn, a.e3t_b, a.e3t_n, a.e3t_a};
        const long long pn[FP_COUNT] = {n4, n4, n3, n3, n3, n3};
        bool ok = true;
        for (int q = 0; q < FH_COUNT && ok; ++q) ok = make_tile_map(&tm.h[q], hb[q], a.jpi, a.jpj, hn[q], FBW, FBH);
        for (int q = 0; q < FP_COUNT && ok; ++q) ok = make_tile_map(&tm.p[q], pb[q], a.jpi, a.jpj, pn[q], FX, FY);
        if (!ok) return false;
        memcpy(cache->maps, &tm, sizeof tm);
        memcpy(cache->key, key, sizeof key); memcpy(cache->dims, dims, sizeof dims);
        cache->valid = true;
    }
    return true;
}

void launch_fct_fused(const FctArgs &a, cudaStream_t s, const TmaMapCache *cache, int max_blocks)
{
    const int ni = a.out.i1 - a.out.i0 + 1, nj = a.out.j1 - a.out.j0 + 1;
    FusedMaps tm;
    memcpy(&tm, cache->maps, sizeof tm);
    const int gx = ((ni + FOX - 1) / FOX) * a.kjpt, gy = (nj + FOY - 1) / FOY, nwork = gx * gy * a.nkchunk;
    const int nblk = std::max(1, std::min(nwork, max_blocks));
    static const int skew_env = getenv("NEMO_FCT_SKEW_NS") ? atoi(getenv("NEMO_FCT_SKEW_NS")) : 0;   // experiment knob: measured to hurt
    const int skew_ns = nblk < nwork ? skew_env : 0;                  // persistent blocks only
#define LFU(H, V, A) do { static std::atomic<bool> done[kMaxDevices], donep[kMaxDevices]; \
                          if (nblk < nwork) { allow_dynamic_smem(k_fct_fused<H, V, A, true>, kFusedSmemBytes, donep); \
                                              k_fct_fused<H, V, A, true><<<nblk, FX * FY, kFusedSmemBytes, s>>>(a, tm, gx, gy, nwork, skew_ns); } \
                          else              { allow_dynamic_smem(k_fct_fused<H, V, A, false>, kFusedSmemBytes, done); \
                                              k_fct_fused<H, V, A, false><<<dim3(gx, gy, a.nkchunk), FX * FY, kFusedSmemBytes, s>>>(a, tm, gx, gy, nwork, 0); } } while (0)
#define LFU2(H, V) do { if (a.arith == 0) LFU(H, V, 0); else LFU(H, V, 1); } while (0)
    if (a.kn_fct_h == 2 && a.kn_fct_v == 2) LFU2(2, 2);
    else if (a.kn_fct_h == 2)               LFU2(2, 4);
    else if (a.kn_fct_v == 2)               LFU2(4, 2);
    else                                    LFU2(4, 4);
#undef LFU2
#undef LFU
    note_launch();
}

void launch_fct_nonosc_final(const FctArgs &a, cudaStream_t s)
{
    static std::atomic<bool> done[kMaxDevices];
    const size_t smem = (size_t)(3 * 4 + 2 * 2) * NY * NX * sizeof(double);
    allow_dynamic_smem(k_fct_nonosc_final, smem, done);
    const int ox = NX - 2 * NHALO, oy = NY - 2 * NHALO;
    const int ni = a.out.i1 - a.out.i0 + 1, nj = a.out.j1 - a.out.j0 + 1;
    if (ni <= 0 || nj <= 0) return;
    const dim3 g((unsigned)(((ni + ox - 1) / ox) * a.kjpt), (unsigned)((nj + oy - 1) / oy), 1);
    k_fct_nonosc_final<<<g, NX * NY, smem, s>>>(a);
    note_launch();
}

void launch_fct_diag(const FctArgs &a, double *trdx, double *trdy, double *trdz, cudaStream_t s)
{
    const dim3 g((unsigned)((a.jpij + kThreads - 1) / kThreads), 1, (unsigned)a.kjpt);
    k_fct_diag<<<g, kThreads, 0, s>>>(a, trdx, trdy, trdz);
    note_launch();
}

void launch_fct_betas(const FctArgs &a, cudaStream_t s) { k_fct_betas<<<column_grid(a), kThreads, 0, s>>>(a); note_launch(); }
void launch_fct_limit(const FctArgs &a, cudaStream_t s) { k_fct_limit<<<column_grid(a), kThreads, 0, s>>>(a); note_launch(); }
void launch_fct_final(const FctArgs &a, cudaStream_t s) { k_fct_final<<<column_grid(a), kThreads, 0, s>>>(a); note_launch(); }

void launch_cpt_pivots(int jpi, int jpj, int jpk, const double *wmask, const int *mikt, const int *mbkt,
                       int ln_isfcav, double *zwt, cudaStream_t s)
{
    (void)ln_isfcav;
    const long long ncol = (long long)(jpi - 2) * (jpj - 2);
    k_cpt_pivots<<<(unsigned)((ncol + kThreads - 1) / kThreads), kThreads, 0, s>>>(jpi, jpj, jpk, wmask, mikt, mbkt, zwt);
    note_launch();
}

void launch_cpt_classify(int jpi, int jpj, int jpk, const double *wmask, const int *mikt, const int *mbkt, const double *zwt,
                         const double *utab, unsigned char *simple, cudaStream_t s)
{
    const long long ncol = (long long)(jpi - 2) * (jpj - 2);
    k_cpt_classify<<<(unsigned)((ncol + kThreads - 1) / kThreads), kThreads, 0, s>>>(jpi, jpj, jpk, wmask, mikt, mbkt, zwt, utab, simple);
    note_launch();
}

void launch_interp_4th_cpt(int jpi, int jpj, int jpk, int nfld, const double *wmask, const int *mikt, const int *mbkt,
                           int ln_isfcav, const double *zwt, const unsigned char *simple, const double *utab,
                           const double *pt_in, double *pt_out, cudaStream_t s, TmaMapCache *cache, int jlo, int jhi)
{
    if (jlo < 2 || jhi > jpj - 1 || jhi < jlo) { jlo = 2; jhi = jpj - 1; }                // rows to solve (1-based); default: the whole interior
    (void)ln_isfcav;
    // tiled kernel: even jpi (16-byte global strides), 16-byte aligned input, the forward sweep of a tile fits in shared memory.
    // (Tried: parking the upper levels of the forward sweep in pt_out for 12-20 resident warps per SM instead of 8 -- 1.24-1.37 ms
    // against 1.21 ms at ORCA025; 64x2 and 128x1 tiles for longer DRAM bursts -- no gain.  Dropped.)
    const size_t smem = cpt_tiled_smem_bytes(jpk);
    if (cache && !(jpi & 1) && jpk >= 3 && smem <= 200 * 1024 && utab && simple) {
        const void *key[12] = {pt_in};
        const int dims[4] = {jpi, jpj, jpk, nfld};
        bool ok = cache->valid && !memcmp(cache->key, key, sizeof key) && !memcmp(cache->dims, dims, sizeof dims);
        if (!ok) {
            CptMap m;
            auto enc = tensor_map_encoder();
            if (enc && !(reinterpret_cast<uintptr_t>(pt_in) & 15u)) {
                cuuint64_t gd[3] = {(cuuint64_t)jpi, (cuuint64_t)jpj, (cuuint64_t)jpk * nfld};
                cuuint64_t gs[2] = {(cuuint64_t)jpi * 8, (cuuint64_t)jpi * jpj * 8};
                cuuint32_t box[3] = {CTX, CTY, CKL};
                cuuint32_t es[3] = {1, 1, 1};
                ok = enc(&m.m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, const_cast<double *>(pt_in), gd, gs, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
            }
            if (ok) {
                memcpy(cache->maps, &m, sizeof m);
                memcpy(cache->key, key, sizeof key); memcpy(cache->dims, dims, sizeof dims);
                cache->valid = true;
            }
        }
        if (ok) {
            CptMap m;
            memcpy(&m, cache->maps, sizeof m);
            static std::atomic<bool> done[kMaxDevices];
            allow_dynamic_smem(k_interp_4th_cpt_tiled, 200 * 1024, done);
            const dim3 g((unsigned)((jpi - 1 + CTX - 1) / CTX), (unsigned)((jhi - jlo + 1 + CTY - 1) / CTY), (unsigned)nfld);
            k_interp_4th_cpt_tiled<<<g, CTX * CTY, smem, s>>>(jpi, jpj, jpk, wmask, mikt, mbkt, zwt, simple, utab, pt_out, m, jlo, jhi);
            note_launch();
            return;
        }
    }
    const long long ncol = (long long)(jpi - 2) * (jpj - 2);
    const dim3 g((unsigned)((ncol + kThreads - 1) / kThreads), 1, (unsigned)nfld);
    k_interp_4th_cpt<<<g, kThreads, 0, s>>>(jpi, jpj, jpk, wmask, mikt, mbkt, zwt, simple, utab, pt_in, pt_out, Region(), nullptr);
    note_launch();
}

void launch_interp_4th_cpt_region(const Region &reg, int jpi, int jpj, int jpk, int nfld, const double *wmask, const int *mikt, const int *mbkt,
                                  const double *zwt, const unsigned char *simple, const double *utab, const double *pt_in, double *pt_out,
                                  double *scratch, cudaStream_t s)
{
    const long long ncol = reg.ncol();
    if (ncol <= 0) return;
    const dim3 g((unsigned)((ncol + kThreads - 1) / kThreads), 1, (unsigned)nfld);
    k_interp_4th_cpt<<<g, kThreads, 0, s>>>(jpi, jpj, jpk, wmask, mikt, mbkt, zwt, simple, utab, pt_in, pt_out, reg, scratch);
    note_launch();
}

void launch_transports(int jpi, int jpj, int jpk, const double *e2u, const double *e1v, const double *e1e2t,
                       const double *e3u_n, const double *e3v_n, const double *un, const double *vn,
                       const double *wn, double *zun, double *zvn, double *zwn, cudaStream_t s)
{
    const size_t jpij = (size_t)jpi * jpj;
    k_transports<<<148 * 8, 256, 0, s>>>(jpk, jpij, e2u, e1v, e1e2t, e3u_n, e3v_n, un, vn, wn, zun, zvn, zwn);
    note_launch();
}

}  // namespace nemo
