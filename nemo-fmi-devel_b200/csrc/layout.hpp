// layout.hpp -- host-side domain decomposition and lbc_lnk exchange-plan compiler (no CUDA in this header).
//
// Replaces, for the FCT path, the reference's
//   mpp_init / mpp_basic_decomposition / mpp_init_nfdcom   src/OCE/LBC/mppini.F90:110-692, 695-798, 1180-1240
//   mpp_lnk + lbc_nfd / mpp_nfd index logic                src/OCE/LBC/mpp_lnk_generic.h90:85-336,
//                                                          lbc_nfd_generic.h90:66-163, mpp_nfd_generic.h90:222-298
// Design (not a port): instead of executing pack / send / recv / unpack phases (E-W, fold, N-S) on values every
// call, the three phases are executed ONCE, symbolically, on cell references.  The result is a flat gather plan:
// every halo cell (and every interior cell the north fold rewrites) is  sgn^p * (one pre-exchange cell of one
// rank)  or the land value.  The device then needs one pack kernel, one grouped send/recv and one unpack kernel
// per exchange, with no ordering constraints between directions, and corner cells still come out exactly as the
// reference's two-phase ordering produces them.
#pragma once
#include <cstdint>
#include <string>
#include <vector>
#include "../../include/nemo_fct.h"

namespace nemo {

struct Layout {
    int jpiglo = 0, jpjglo = 0, jpk = 0, jperio = 0, jpni = 1, jpnj = 1, jpnij = 1;
    int jpimax = 0, jpjmax = 0;
    bool mpi = true;                       // key_mpp_mpi semantics
    std::vector<int> nimppt, njmppt, nlcit, nlcjt;   // per rank r = (ij-1)*jpni + (ii-1), 0-based
    bool ew_cyclic() const { return jperio == 1 || jperio == 4 || jperio == 6 || jperio == 7; }
    bool ns_cyclic() const { return jperio == 2 || jperio == 7; }
    bool fold() const { return jperio >= 3 && jperio <= 6; }
    bool t_pivot() const { return jperio == 3 || jperio == 4; }
};

// throws std::runtime_error on an impossible layout (same conditions as the reference's ctl_stop calls)
Layout make_layout(int jpiglo, int jpjglo, int jpk, int jperio, int jpni, int jpnj, bool mpi);
nemo_fct_domain rank_domain(const Layout &L, int narea);
// "" if `d` agrees with rank_domain(L, d.narea), else a description of the first mismatch
std::string check_domain(const Layout &L, const nemo_fct_domain &d);

// One received cell: dst (local index on this rank) = psgn^spow * src (local index on rank `peer`)
struct PlanCell { int dst; int src; int spow; };
struct PeerList { int peer; std::vector<PlanCell> cells; };   // message order == vector order on both sides

struct RankPlan {
    std::vector<PlanCell> fill;            // dst = psgn^spow * zland (src = -1)
    std::vector<PeerList> recv;            // what this rank receives, grouped by source rank (self included)
    std::vector<PeerList> send;            // what this rank sends: cells[i].src = local source index; same order
                                           // as the peer's recv list from this rank
};

// Compile lbc_lnk for grid-point type nat in {T,U,V,W,F} for ALL ranks (W is treated as T, as in the reference).
std::vector<RankPlan> compile_lbc_plan(const Layout &L, char nat);

}  // namespace nemo
