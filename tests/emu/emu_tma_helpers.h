// Host stand-ins for the PTX helpers of the TMA-tiled kernels (TEST INFRASTRUCTURE ONLY): an mbarrier is an atomic count of
// completed phases; a bulk tensor copy is a box copy with zero fill outside the array that completes before the issuing
// thread goes on.  The hardware rules the kernels were designed around are CHECKED here: box origins >= 0 and even in the
// innermost coordinate.  Include inside namespace nemo { namespace { ... } } after emu_block.h and kernels.cuh.
#pragma once
struct EmuMap { const double *base; int jpi, jpj; long long nlev; int bw, bh; };   // what make_tile_map encodes
static_assert(sizeof(EmuMap) <= sizeof(CUtensorMap), "descriptor does not fit");

static thread_local std::map<unsigned long long *, long long> emu_pending;          // bytes still expected per barrier (issuing thread)
static int emu_tma_violations = 0;

inline void mbar_init(unsigned long long *bar, unsigned) { __atomic_store_n(bar, 0ull, __ATOMIC_SEQ_CST); }
inline void mbar_init_fence() {}
inline void mbar_expect_tx(unsigned long long *bar, unsigned bytes) { emu_pending[bar] = bytes; }
inline void mbar_wait(unsigned long long *bar, unsigned parity)
{
    while ((__atomic_load_n(bar, __ATOMIC_SEQ_CST) & 1ull) == parity) std::this_thread::yield();
}
inline void tma_load_3d(void *dst, const CUtensorMap *map, unsigned long long *bar, int x, int y, int z)
{
    const EmuMap m = *reinterpret_cast<const EmuMap *>(map);
    if (x < 0 || y < 0 || z < 0 || (x & 1) || z >= m.nlev) __atomic_fetch_add(&emu_tma_violations, 1, __ATOMIC_SEQ_CST);
    double *d = static_cast<double *>(dst);
    for (int yy = 0; yy < m.bh; ++yy)
        for (int xx = 0; xx < m.bw; ++xx) {
            const int gx = x + xx, gy = y + yy;
            const bool in = gx >= 0 && gx < m.jpi && gy >= 0 && gy < m.jpj && z >= 0 && z < m.nlev;
            d[yy * m.bw + xx] = in ? m.base[(size_t)z * m.jpi * m.jpj + (size_t)gy * m.jpi + gx] : 0.0;
        }
    long long &left = emu_pending[bar];
    left -= (long long)m.bw * m.bh * 8;
    if (left == 0) __atomic_fetch_add(bar, 1ull, __ATOMIC_SEQ_CST);                 // phase complete
}


static void set_map(CUtensorMap *m, const double *base, int jpi, int jpj, long long nlev, int bw, int bh)
{
    std::memset(m, 0, sizeof *m);
    const EmuMap e{base, jpi, jpj, nlev, bw, bh};
    std::memcpy(m, &e, sizeof e);
}

