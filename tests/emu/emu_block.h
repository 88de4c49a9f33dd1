// Emulation of ONE thread block of a cooperative kernel (shared memory + __syncthreads): one host thread per CUDA thread,
// std::barrier for __syncthreads(), a per-block host buffer for the dynamic shared memory.  TEST INFRASTRUCTURE ONLY.
#pragma once
#include "emu_common.h"

#include <barrier>
#include <cstdint>
#include <memory>
#include <thread>
#include <vector>

static thread_local std::barrier<> *emu_block_barrier = nullptr;
static thread_local unsigned char *emu_block_smem = nullptr;
// one warp: its lanes meet at `bar` to exchange registers (__shfl_down_sync); n = lanes of this warp that exist
struct EmuWarp { std::barrier<> bar; int n; unsigned long long slot[32]; explicit EmuWarp(int lanes) : bar(lanes), n(lanes) {} };
static thread_local EmuWarp *emu_warp = nullptr;
// __shfl_down_sync of a full-warp mask: lane l receives the value of lane l + off, its own when that lane does not exist
template <typename T> static inline T __shfl_down_sync(unsigned, T v, int off)
{
    static_assert(sizeof(T) <= 8, "register-sized operands only");
    const int lane = (int)(threadIdx.x & 31u);
    unsigned long long raw = 0, got;
    std::memcpy(&raw, &v, sizeof(T));
    emu_warp->slot[lane] = raw;
    emu_warp->bar.arrive_and_wait();
    got = lane + off < emu_warp->n ? emu_warp->slot[lane + off] : raw;
    emu_warp->bar.arrive_and_wait();                                   // the slots may be rewritten
    T out;
    std::memcpy(&out, &got, sizeof(T));
    return out;
}
#define __syncthreads() emu_block_barrier->arrive_and_wait()
#define NEMO_DYN_SMEM(type, name) type *name = reinterpret_cast<type *>(emu_block_smem)
#define NEMO_DYN_SMEM_ALIGNED(type, name, alignment) type *name = reinterpret_cast<type *>(emu_block_smem)

// run kernel(args...) for every block of a (gx, gy, gz) grid with nthreads threads per block and smem_bytes of shared memory
template <typename K, typename... A>
static void emu_run_blocks3(int gx, int gy, int gz, int nthreads, size_t smem_bytes, K kernel, const A &...args)
{
    std::vector<unsigned char> smem(smem_bytes + 128);
    unsigned char *base = smem.data() + (128 - reinterpret_cast<uintptr_t>(smem.data()) % 128) % 128;    // 128-byte aligned, like the device
    for (int bz = 0; bz < gz; ++bz)
    for (int by = 0; by < gy; ++by)
        for (int bx = 0; bx < gx; ++bx) {
            std::barrier<> bar(nthreads);
            std::vector<std::unique_ptr<EmuWarp>> warps;
            for (int w0 = 0; w0 < nthreads; w0 += 32) warps.emplace_back(new EmuWarp(std::min(32, nthreads - w0)));
            std::vector<std::thread> th;
            th.reserve(nthreads);
            for (int t = 0; t < nthreads; ++t)
                th.emplace_back([&, t, bx, by, bz]() {
                    blockIdx = {(unsigned)bx, (unsigned)by, (unsigned)bz};
                    threadIdx = {(unsigned)t, 0, 0};
                    blockDim = {(unsigned)nthreads, 1, 1};
                    gridDim = {(unsigned)gx, (unsigned)gy, (unsigned)gz};
                    emu_block_barrier = &bar;
                    emu_warp = warps[t / 32].get();
                    emu_block_smem = base;
                    kernel(args...);
                });
            for (auto &x : th) x.join();
        }
}
template <typename K, typename... A>
static void emu_run_blocks(int gx, int gy, int nthreads, size_t smem_bytes, K kernel, const A &...args)
{
    emu_run_blocks3(gx, gy, 1, nthreads, smem_bytes, kernel, args...);
}
