// Host stand-ins for the PTX helpers of the TMA-tiled kernels (TEST INFRASTRUCTURE ONLY): an mbarrier is an atomic count of
// completed phases; a bulk tensor copy is a box copy with zero fill outside the array that completes before the issuing
// thread goes on.  The hardware rules the kernels were designed around are CHECKED here: box origins >= 0 and even in the
// innermost coordinate.  Include inside namespace nemo { namespace { ... } } after emu_block.h and kernels.cuh.
#pragma once
struct EmuMap { const double *base; int jpi, jpj; long long nlev; int bw, bh; };   // what make_tile_map encodes
static_assert(sizeof(EmuMap) <= sizeof(CUtensorMap), "descriptor does not fit");

// per-barrier transaction state, shared by the issuing threads (several warps may issue copies on one barrier): the phase
// completes when the expect_tx arrival has happened and the SIGNED byte count is back to zero
struct EmuBar { long long pending = 0; bool arrived = false; };
static std::map<unsigned long long *, EmuBar> emu_bars;
static std::mutex emu_bars_mutex;
static int emu_tma_violations = 0;
static int emu_box_depth = 1;                            // levels per box (boxDim[2]) of the maps in use: 1 for the FCT tiles

inline void emu_bar_check(unsigned long long *bar, EmuBar &b)
{
    if (b.arrived && b.pending == 0) { b.arrived = false; __atomic_fetch_add(bar, 1ull, __ATOMIC_SEQ_CST); }   // phase complete
}
inline void mbar_init(unsigned long long *bar, unsigned)
{
    std::lock_guard<std::mutex> g(emu_bars_mutex);
    emu_bars[bar] = EmuBar();
    __atomic_store_n(bar, 0ull, __ATOMIC_SEQ_CST);
}
inline void mbar_init_fence() {}
inline void mbar_inval(unsigned long long *) {}
inline void mbar_expect_tx(unsigned long long *bar, unsigned bytes)
{
    std::lock_guard<std::mutex> g(emu_bars_mutex);
    EmuBar &b = emu_bars[bar];
    b.pending += bytes; b.arrived = true;
    emu_bar_check(bar, b);
}
inline void mbar_wait(unsigned long long *bar, unsigned parity)
{
    while ((__atomic_load_n(bar, __ATOMIC_SEQ_CST) & 1ull) == parity) std::this_thread::yield();
}
inline void tma_load_3d(void *dst, const CUtensorMap *map, unsigned long long *bar, int x, int y, int z)
{
    const EmuMap m = *reinterpret_cast<const EmuMap *>(map);
    if (x < 0 || y < 0 || z < 0 || (x & 1) || z >= m.nlev) __atomic_fetch_add(&emu_tma_violations, 1, __ATOMIC_SEQ_CST);
    double *d = static_cast<double *>(dst);
    const int depth = emu_box_depth;
    for (int zz = 0; zz < depth; ++zz)
    for (int yy = 0; yy < m.bh; ++yy)
        for (int xx = 0; xx < m.bw; ++xx) {
            const int gx = x + xx, gy = y + yy, gz = z + zz;
            const bool in = gx >= 0 && gx < m.jpi && gy >= 0 && gy < m.jpj && gz >= 0 && gz < m.nlev;
            d[((size_t)zz * m.bh + yy) * m.bw + xx] = in ? m.base[(size_t)gz * m.jpi * m.jpj + (size_t)gy * m.jpi + gx] : 0.0;
        }
    std::lock_guard<std::mutex> g(emu_bars_mutex);
    EmuBar &b = emu_bars[bar];
    b.pending -= (long long)m.bw * m.bh * depth * 8;
    emu_bar_check(bar, b);
}

static void set_map(CUtensorMap *m, const double *base, int jpi, int jpj, long long nlev, int bw, int bh)
{
    std::memset(m, 0, sizeof *m);
    const EmuMap e{base, jpi, jpj, nlev, bw, bh};
    std::memcpy(m, &e, sizeof e);
}

