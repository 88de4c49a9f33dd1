// schedule.hpp -- which (ji,jj) columns each kernel of the fused schedules covers (host only, pure functions).
//
// The fused schedules split the interior (2:jpim1, 2:jpjm1) of a subdomain into an exchange-free inner region, computed by
// fused kernels from the input fields alone, and a boundary frame, computed by the reference-structured kernels restricted
// to bands together with the real exchanges.  Every band holds exactly what the next kernel of the frame chain and the
// following exchange read (dependency radius of one FCT step: 3 cells).  `fold` = the subdomain touches
// the north fold, which also rewrites part of the interior row jpj-1.  Kept apart from the launch code so that the CPU
// test-suite can run the same plans with the host emulation of the kernels (tests/emu) at any size, minimum sizes included.
#pragma once
#include "kernels.cuh"

namespace nemo {

constexpr int kMinFusedSize = 20;      // subdomains smaller than this in ji or jj use the reference structure only

// bands of the interior along the four edges: W = 2..1+wW, E = jpi-wE..jpi-1 (all rows), S = 2..1+wS, N = jpj-wN..jpj-1
inline Region frame_band(int jpi, int jpj, int wW, int wE, int wS, int wN)
{
    Region r;
    r.add(2, 1 + wW, 2, jpj - 1);
    r.add(jpi - wE, jpi - 1, 2, jpj - 1);
    r.add(2 + wW, jpi - wE - 1, 2, 1 + wS);
    r.add(2 + wW, jpi - wE - 1, jpj - wN, jpj - 1);
    return r;
}

struct FctFusedPlan {
    Region k1;                 // P1-P5 inner kernel: (2:jpi-2, 2:jpj-2-f)
    Region k1_band, k1_centre; // its split: an 8-wide band (side stream, feeds X2 early) and the centre (main stream)
    bool split;                // false: the rectangle is too small to split, or no split was asked for; k1_centre = k1
    Rect k2_out;               // nonosc + final inner kernel: output rectangle (5:jpi-4, 4:jpj-4-f), i0 odd (even TMA box origin)
    Region lowf;               // frame P1-P5: E column and N rows left out by k1
    Region lap, bet, lim, fin; // frame Laplacian, betas, limiter, final trend
    Region cptb;               // schedule 4: columns whose ztw the frame chain reads (k1_band + lowf), solved on the side stream
};

inline FctFusedPlan fct_fused_plan(int jpi, int jpj, bool fold, bool want_split)
{
    const int f = fold ? 1 : 0;
    FctFusedPlan p;
    p.k1.add(2, jpi - 2, 2, jpj - 2 - f);
    p.k2_out = Rect{5, jpi - 4, 4, jpj - 4 - f};
    p.lowf.add(jpi - 1, jpi - 1, 2, jpj - 1);
    p.lowf.add(2, jpi - 2, jpj - 1 - f, jpj - 1);
    p.lap = frame_band(jpi, jpj, 1, 1, 1, 2 + f);
    p.fin = frame_band(jpi, jpj, 3, 3, 2, 3 + f);
    p.lim = frame_band(jpi, jpj, 4, 4, 3, 4 + f);
    p.bet = frame_band(jpi, jpj, 5, 5, 4, 5 + f);
    p.cptb = frame_band(jpi, jpj, 8, 9, 7, 9 + f);      // = k1_band + lowf exactly (their P1-P5 launches rewrite the zwz scratch)
    const Rect r1 = p.k1.r[0];
    const int w = 8;
    p.split = want_split && !(r1.i1 - r1.i0 + 1 < 2 * w + 4 || r1.j1 - r1.j0 + 1 < 2 * w + 3);
    if (p.split) {
        p.k1_band.add(r1.i0, r1.i0 + w - 1, r1.j0, r1.j1);                      // i0 = 2: the centre starts at an even column (TMA)
        p.k1_band.add(r1.i1 - w + 1, r1.i1, r1.j0, r1.j1);
        p.k1_band.add(r1.i0 + w, r1.i1 - w, r1.j0, r1.j0 + w - 2);
        p.k1_band.add(r1.i0 + w, r1.i1 - w, r1.j1 - w + 1, r1.j1);
        p.k1_centre.add(r1.i0 + w, r1.i1 - w, r1.j0 + w - 1, r1.j1 - w);
    } else {
        p.k1_centre = p.k1;
    }
    return p;
}

// bands for the MUSCL schedules: W = i_lo..wW, E = jpi-wE..jpi-1 (rows j_lo..jpj-1), S = j_lo..wS, N = jpj-wN..jpj-1
inline Region mus_band(int jpi, int jpj, int i_lo, int wW, int wE, int j_lo, int wS, int wN)
{
    Region r;
    r.add(i_lo, wW, j_lo, jpj - 1);
    r.add(jpi - wE, jpi - 1, j_lo, jpj - 1);
    r.add(wW + 1, jpi - wE - 1, j_lo, wS);
    r.add(wW + 1, jpi - wE - 1, jpj - wN, jpj - 1);
    return r;
}

struct MusPlan { Region inner, grad, hflux, trend; };

// default MUSCL schedule: fluxes straight from ptb on `inner`; the rest of the interior (`hflux`) from exchanged differences,
// which k_mus_grad forms on `grad` (what those columns and the first exchange read); trend on the whole interior
inline MusPlan mus_semi_plan(int jpi, int jpj, bool fold)
{
    const int f = fold ? 1 : 0;
    MusPlan p;
    p.inner.add(3, jpi - 2, 3, jpj - 2 - f);
    p.hflux = mus_band(jpi, jpj, 2, 2, 1, 2, 2, 1 + f);
    p.grad = mus_band(jpi, jpj, 1, 3, 2, 1, 3, 3 + f);
    p.trend.add(2, jpi - 1, 2, jpj - 1);
    return p;
}

// schedule 1: everything fused on `inner` (4:jpi-2, 4:jpj-2-f); grad / hflux / trend bands of the two-cell frame
inline MusPlan mus_fused_plan(int jpi, int jpj, bool fold)
{
    const int f = fold ? 1 : 0;
    MusPlan p;
    p.inner.add(4, jpi - 2, 4, jpj - 2 - f);
    p.trend = mus_band(jpi, jpj, 2, 3, 1, 2, 3, 1 + f);
    p.hflux = mus_band(jpi, jpj, 2, 4, 2, 2, 4, 3 + f);
    p.grad = mus_band(jpi, jpj, 1, 5, 3, 1, 5, 4 + f);
    return p;
}

}  // namespace nemo
