// layout.cpp -- domain decomposition + symbolic lbc_lnk plan compiler (host only).  See layout.hpp.
#include "layout.hpp"

#include <algorithm>
#include <map>
#include <sstream>
#include <stdexcept>
#include <unordered_map>

namespace nemo {

static const int kHls = 1;   // nn_hls, src/OCE/par_oce.F90:76

// ------------------------------------------------------------------------------------------------------------
// Decomposition.  Closed-form restatement of mpp_basic_decomposition (mppini.F90:723-796): subdomains overlap
// by 2*nn_hls, the first `rest` ones along an axis get the larger size; with a north fold the last row is
// shrunk first (min 5 rows for a T-point pivot, 4 for an F-point pivot, mppini.F90:755-764).
// ------------------------------------------------------------------------------------------------------------
static void split_axis(int nglo, int nparts, int &nmax, std::vector<int> &sizes)
{
    nmax = (nglo - 2 * kHls + (nparts - 1)) / nparts + 2 * kHls;
    const int rest = 1 + (nglo - 2 * kHls - 1) % nparts;
    sizes.assign(nparts, nmax - 1);
    for (int p = 0; p < rest; ++p) sizes[p] = nmax;
}

Layout make_layout(int jpiglo, int jpjglo, int jpk, int jperio, int jpni, int jpnj, bool mpi)
{
    if (jpiglo < 3 || jpjglo < 3 || jpk < 1) throw std::runtime_error("mpp_init: global domain too small");
    if (jperio < 0 || jperio > 7) throw std::runtime_error("mpp_init: jperio must be in 0..7");
    if (jpni < 1 || jpnj < 1) throw std::runtime_error("mpp_init: jpni, jpnj must be >= 1");
    if (!mpi && (jpni != 1 || jpnj != 1))
        throw std::runtime_error("mpp_init: equality jpni = jpnj = jpnij = 1 is not satisfied (no key_mpp_mpi)");
    Layout L;
    L.jpiglo = jpiglo; L.jpjglo = jpjglo; L.jpk = jpk; L.jperio = jperio;
    L.jpni = jpni; L.jpnj = jpnj; L.jpnij = jpni * jpnj; L.mpi = mpi;
    std::vector<int> isz, jsz;
    if (!mpi) { L.jpimax = jpiglo; L.jpjmax = jpjglo; isz = {jpiglo}; jsz = {jpjglo}; }
    else {
        split_axis(jpiglo, jpni, L.jpimax, isz);
        split_axis(jpjglo, jpnj, L.jpjmax, jsz);
        int jmin = 3;
        if (L.fold()) {
            jmin = L.t_pivot() ? 5 : 4;
            const int rest = 1 + (jpjglo - 2 * kHls - 1) % jpnj;
            int remove = jpnj - rest;                         // lines to take away in total
            const int last = std::max(jmin, L.jpjmax - remove);
            remove -= L.jpjmax - last;                        // what the other rows still have to give
            const int nfull = jpnj - 1 - remove;
            for (int q = 0; q < jpnj - 1; ++q) jsz[q] = (q < nfull) ? L.jpjmax : L.jpjmax - 1;
            jsz[jpnj - 1] = last;
        }
        if (*std::min_element(isz.begin(), isz.end()) < 3)
            throw std::runtime_error("mpp_basic_decomposition: minimum value of jpi must be >= 3");
        if (*std::min_element(jsz.begin(), jsz.end()) < jmin)
            throw std::runtime_error("mpp_basic_decomposition: minimum value of jpj too small for this jperio");
    }
    std::vector<int> i0(jpni, 1), j0(jpnj, 1);
    for (int p = 1; p < jpni; ++p) i0[p] = i0[p - 1] + isz[p - 1] - 2 * kHls;
    for (int q = 1; q < jpnj; ++q) j0[q] = j0[q - 1] + jsz[q - 1] - 2 * kHls;
    L.nimppt.resize(L.jpnij); L.njmppt.resize(L.jpnij); L.nlcit.resize(L.jpnij); L.nlcjt.resize(L.jpnij);
    for (int q = 0; q < jpnj; ++q)
        for (int p = 0; p < jpni; ++p) {
            const int r = q * jpni + p;
            L.nimppt[r] = i0[p]; L.njmppt[r] = j0[q]; L.nlcit[r] = isz[p]; L.nlcjt[r] = jsz[q];
        }
    return L;
}

nemo_fct_domain rank_domain(const Layout &L, int narea)
{
    if (narea < 1 || narea > L.jpnij) throw std::runtime_error("mpp_init: narea out of range");
    nemo_fct_domain d{};
    const int r = narea - 1, p = r % L.jpni, q = r / L.jpni;
    d.jpiglo = L.jpiglo; d.jpjglo = L.jpjglo; d.jpk = L.jpk; d.jperio = L.jperio;
    d.jpni = L.jpni; d.jpnj = L.jpnj; d.narea = narea; d.key_mpp_mpi = L.mpi ? 1 : 0;
    d.jpimax = L.jpimax; d.jpjmax = L.jpjmax;
    d.nimpp = L.nimppt[r]; d.njmpp = L.njmppt[r]; d.nlci = L.nlcit[r]; d.nlcj = L.nlcjt[r];
    d.jpi = d.nlci; d.jpj = d.nlcj;
    d.l_Iperio = (L.jpni == 1 && L.ew_cyclic()) ? 1 : 0;            // mppini.F90:345
    d.l_Jperio = (L.jpnj == 1 && L.ns_cyclic()) ? 1 : 0;            // mppini.F90:346
    d.noea = d.nowe = d.noso = d.nono = -1;
    if (!L.mpi) {                                                   // mppini.F90:53-102
        d.nldi = 1; d.nlei = d.jpi; d.nldj = 1; d.nlej = d.jpj; d.nbondi = d.nbondj = 2; d.npolj = L.jperio;
        return d;
    }
    // neighbour flags: -1 first (only e/n nbr), 1 last (only w/s nbr), 0 both, 2 none  (mppini.F90:355-362)
    auto bond = [](int pos, int n) { return n == 1 ? 2 : (pos == 0 ? -1 : (pos == n - 1 ? 1 : 0)); };
    d.nbondi = bond(p, L.jpni); d.nbondj = bond(q, L.jpnj);
    int we = r - 1, ea = r + 1, so = r - L.jpni, no = r + L.jpni;
    if (L.ew_cyclic()) {                                            // mppini.F90:374-379
        if (L.jpni != 1) d.nbondi = 0;
        if (p == 0) we = r + (L.jpni - 1);
        if (p == L.jpni - 1) ea = r - (L.jpni - 1);
    }
    if (L.ns_cyclic()) {                                            // mppini.F90:381-386
        if (L.jpnj != 1) d.nbondj = 0;
        if (q == 0) so = r + L.jpni * (L.jpnj - 1);
        if (q == L.jpnj - 1) no = r - L.jpni * (L.jpnj - 1);
    }
    d.npolj = 0;
    if (L.fold() && q == L.jpnj - 1) {                              // mppini.F90:388-403, 621-628
        d.npolj = L.t_pivot() ? 3 : 5;
        // mirror rank across the fold; the middle subdomain of an odd jpni is its own mirror (ipolj = 4 / 6 in
        // mppini.F90:393,400) and keeps the out-of-range default
        const bool middle = (L.jpni % 2 == 1) && (p == (L.jpni + 1) / 2 - 1);
        if (!middle) no = L.jpni * L.jpnj - narea + L.jpni * (L.jpnj - 1);
    }
    auto inside = [&](int z) { return (z >= 0 && z < L.jpnij) ? z : -1; };
    d.nowe = inside(we); d.noea = inside(ea); d.noso = inside(so); d.nono = inside(no);
    d.nldi = (d.nbondi == -1 || d.nbondi == 2) ? 1 : 1 + kHls;      // mppini.F90:369-372, 480-487
    d.nlei = (d.nbondi == 1 || d.nbondi == 2) ? d.nlci : d.nlci - kHls;
    d.nldj = (d.nbondj == -1 || d.nbondj == 2) ? 1 : 1 + kHls;
    d.nlej = (d.nbondj == 1 || d.nbondj == 2) ? d.nlcj : d.nlcj - kHls;
    // no-gather fold partners: top-row subdomains whose columns overlap my mirrored range (mppini.F90:1201-1227)
    d.nsndto = 0;
    if (L.fold() && L.jpni > 1 && q == L.jpnj - 1) {
        const int lo = L.jpiglo - d.nimpp - d.nlci + 1, hi = L.jpiglo - d.nimpp + 2;
        for (int c = 0; c < L.jpni; ++c) {
            const int rr = q * L.jpni + c, a = L.nimppt[rr], b = a + L.nlcit[rr] - 1;
            const bool hit = (a < lo && lo < b) || (lo <= a && hi >= b) || (hi < b && a < hi);
            if (hit) {      // informational only: the gather plan serves any number of partners, while the
                            // reference's no-gather path is limited to jpmaxngh = 3 (nsndto > 3 flags such a layout)
                if (d.nsndto < NEMO_FCT_JPMAXNGH) d.isendto[d.nsndto] = c + 1;
                d.nsndto++;
            }
        }
    }
    return d;
}

std::string check_domain(const Layout &L, const nemo_fct_domain &d)
{
    nemo_fct_domain e;
    try { e = rank_domain(L, d.narea); } catch (const std::exception &ex) { return ex.what(); }
    std::ostringstream os;
#define CHK(f) if (e.f != d.f) { os << "nemo_fct_domain." #f " = " << d.f << " but this layout gives " << e.f; return os.str(); }
    CHK(jpi) CHK(jpj) CHK(nimpp) CHK(njmpp) CHK(nlci) CHK(nlcj) CHK(nldi) CHK(nlei) CHK(nldj) CHK(nlej)
    CHK(nbondi) CHK(nbondj) CHK(noea) CHK(nowe) CHK(noso) CHK(nono) CHK(npolj) CHK(l_Iperio) CHK(l_Jperio)
#undef CHK
    return "";
}

// ------------------------------------------------------------------------------------------------------------
// Symbolic execution of lbc_lnk.
// ------------------------------------------------------------------------------------------------------------
namespace {

struct Ref { int rank; int idx; int spow; bool cst; };

struct SymRank {
    int rank = 0, jpi = 0, jpj = 0;
    std::unordered_map<int, Ref> m;                        // only cells that were assigned
    Ref get(int i, int j) const {                          // 1-based
        const int k = (i - 1) + (j - 1) * jpi;
        auto it = m.find(k);
        return it == m.end() ? Ref{rank, k, 0, false} : it->second;
    }
    void set(int i, int j, const Ref &r) { m[(i - 1) + (j - 1) * jpi] = r; }
};

inline Ref land() { return Ref{-1, -1, 0, true}; }
inline Ref times_sgn(Ref r) { r.spow += 1; return r; }

// The north-fold index rules of lbc_nfd (lbc_nfd_generic.h90:75-151) as data: rows are offsets from the folded
// row ipj; for ji in [i0,i1]:  A(ji, ipj+drow) = sgn * A(mirror - ji, ipj+srow); mirror < 0 => fixed source -mirror.
struct FoldRule { int drow, i0, i1, srow, mirror; };

std::vector<FoldRule> fold_rules(bool t_pivot, char nat, int G)
{
    const int h = G / 2;
    if (t_pivot) switch (nat) {
        case 'T': case 'W': return {{0, 2, G, -2, G + 2}, {0, 1, 1, -2, -3}, {-1, h + 1, G, -1, G + 2}};
        case 'U': return {{0, 1, G - 1, -2, G + 1}, {0, 1, 1, -2, -2}, {0, G, G, -2, -(G - 1)}, {-1, h, G - 1, -1, G + 1}};
        // V and F copy rows ipj-1 and ipj inside ONE loop over ji; the two targets never feed each other, so two
        // rules in sequence are equivalent
        case 'V': return {{-1, 2, G, -2, G + 2}, {0, 2, G, -3, G + 2}, {0, 1, 1, -3, -3}};
        case 'F': return {{-1, 1, G - 1, -2, G + 1}, {0, 1, G - 1, -3, G + 1}, {0, 1, 1, -3, -2}, {0, G, G, -3, -(G - 1)}};
    } else switch (nat) {
        case 'T': case 'W': return {{0, 1, G, -1, G + 1}};
        case 'U': return {{0, 1, G - 1, -1, G}, {0, G, G, -1, -(G - 2)}};
        case 'V': return {{0, 1, G, -2, G + 1}, {-1, h + 1, G, -1, G + 1}};
        case 'F': return {{0, 1, G - 1, -2, G}, {0, G, G, -2, -(G - 2)}, {-1, h + 1, G - 1, -1, G}};
    }
    throw std::runtime_error("lbc_nfd: unknown grid-point type");
}

template <class Get, class Set>
void apply_fold(const std::vector<FoldRule> &rules, int ipj, Get get, Set set)
{
    for (const FoldRule &r : rules)
        for (int ji = r.i0; ji <= r.i1; ++ji) {                 // sequential, in place: later reads see earlier writes
            const int si = r.mirror < 0 ? -r.mirror : r.mirror - ji;
            set(ji, ipj + r.drow, times_sgn(get(si, ipj + r.srow)));
        }
}

}  // namespace

std::vector<RankPlan> compile_lbc_plan(const Layout &L, char nat)
{
    if (nat != 'T' && nat != 'U' && nat != 'V' && nat != 'W' && nat != 'F')
        throw std::runtime_error("lbc_lnk: cd_nat must be one of T,U,V,W,F");
    const int n = L.jpnij;
    std::vector<nemo_fct_domain> dom(n);
    std::vector<SymRank> S(n);
    for (int r = 0; r < n; ++r) {
        dom[r] = rank_domain(L, r + 1);
        S[r].rank = r; S[r].jpi = dom[r].jpi; S[r].jpj = dom[r].jpj;
    }
    // --- 1. closed / locally cyclic edges on every rank (mpp_lnk_generic.h90:85-107)
    for (int r = 0; r < n; ++r) {
        const nemo_fct_domain &d = dom[r]; SymRank &s = S[r];
        const int jpi = d.jpi, jpj = d.jpj;
        if (d.l_Iperio) {
            for (int j = 1; j <= jpj; ++j) s.set(1, j, s.get(jpi - 1, j));
            for (int j = 1; j <= jpj; ++j) s.set(jpi, j, s.get(2, j));
        } else {
            if (nat != 'F') for (int j = 1; j <= jpj; ++j) for (int i = 1; i <= kHls; ++i) s.set(i, j, land());
            for (int j = 1; j <= jpj; ++j) for (int i = d.nlci - kHls + 1; i <= jpi; ++i) s.set(i, j, land());
        }
        if (d.l_Jperio) {
            for (int i = 1; i <= jpi; ++i) s.set(i, 1, s.get(i, jpj - 1));
            for (int i = 1; i <= jpi; ++i) s.set(i, jpj, s.get(i, 2));
        } else {
            if (nat != 'F') for (int i = 1; i <= jpi; ++i) for (int j = 1; j <= kHls; ++j) s.set(i, j, land());
            for (int i = 1; i <= jpi; ++i) for (int j = d.nlcj - kHls + 1; j <= jpj; ++j) s.set(i, j, land());
        }
    }
    // --- 2. east-west neighbours (mpp_lnk_generic.h90:109-215): col 2 goes west, col nlci-1 goes east
    {
        std::vector<std::vector<std::pair<int, Ref>>> upd(n);
        for (int r = 0; r < n; ++r) {
            const nemo_fct_domain &d = dom[r];
            if (d.nbondi == 2) continue;
            for (int j = 1; j <= d.jpj; ++j) {
                if (d.nbondi == 0 || d.nbondi == 1)       // has a west neighbour: my col 1 <- its col nlci-1
                    upd[r].push_back({(1 - 1) + (j - 1) * d.jpi, S[d.nowe].get(dom[d.nowe].nlci - 1, j)});
                if (d.nbondi == 0 || d.nbondi == -1)      // has an east neighbour: my col nlci <- its col 2
                    upd[r].push_back({(d.nlci - 1) + (j - 1) * d.jpi, S[d.noea].get(2, j)});
            }
        }
        for (int r = 0; r < n; ++r) for (auto &u : upd[r]) S[r].m[u.first] = u.second;
    }
    // --- 3. north fold (mpp_lnk_generic.h90:217-228)
    if (L.fold()) {
        const std::vector<FoldRule> rules = fold_rules(L.t_pivot(), nat, L.jpiglo);
        if (L.jpni == 1) {                                 // lbc_nfd on the single top subdomain, ipj = nlcj
            const int r = n - 1; SymRank &s = S[r];
            apply_fold(rules, dom[r].nlcj, [&](int i, int j) { return s.get(i, j); },
                       [&](int i, int j, const Ref &v) { s.set(i, j, v); });
        } else {                                           // mpp_nfd, gather semantics (mpp_nfd_generic.h90:222-298)
            const int G = L.jpiglo, ipj = 4;
            std::vector<Ref> ztab((size_t)G * ipj, Ref{-2, -1, 0, true});      // rank -2: never defined
            for (int p = 0; p < L.jpni; ++p) {
                const int r = (L.jpnj - 1) * L.jpni + p; const nemo_fct_domain &d = dom[r];
                int ildi = d.nldi, ilei = d.nlei;
                if (d.nimpp == 1) ildi = 1;                                     // e-w boundary already done
                if (d.nimpp + d.nlci - 1 == G) ilei = d.nlci;
                for (int jj = 1; jj <= ipj; ++jj)
                    for (int ji = ildi; ji <= ilei; ++ji)
                        ztab[(size_t)(jj - 1) * G + (ji + d.nimpp - 1 - 1)] = S[r].get(ji, d.nlcj - ipj + jj);
            }
            apply_fold(rules, ipj, [&](int i, int j) { return ztab[(size_t)(j - 1) * G + (i - 1)]; },
                       [&](int i, int j, const Ref &v) { ztab[(size_t)(j - 1) * G + (i - 1)] = v; });
            for (int p = 0; p < L.jpni; ++p) {
                const int r = (L.jpnj - 1) * L.jpni + p; const nemo_fct_domain &d = dom[r];
                for (int jj = 1; jj <= ipj; ++jj)
                    for (int ji = 1; ji <= d.nlci; ++ji) {
                        const Ref &v = ztab[(size_t)(jj - 1) * G + (ji + d.nimpp - 1 - 1)];
                        if (v.rank == -2) throw std::runtime_error("mpp_nfd: undefined cell in the gathered north rows");
                        S[r].set(ji, d.nlcj - ipj + jj, v);
                    }
            }
        }
    }
    // --- 4. north-south neighbours (mpp_lnk_generic.h90:230-336): row 2 goes south, row nlcj-1 goes north
    {
        std::vector<std::vector<std::pair<int, Ref>>> upd(n);
        for (int r = 0; r < n; ++r) {
            const nemo_fct_domain &d = dom[r];
            if (d.nbondj == 2) continue;
            for (int i = 1; i <= d.jpi; ++i) {
                if (d.nbondj == 0 || d.nbondj == 1)       // has a south neighbour: my row 1 <- its row nlcj-1
                    upd[r].push_back({(i - 1) + (1 - 1) * d.jpi, S[d.noso].get(i, dom[d.noso].nlcj - 1)});
                if (d.nbondj == 0 || d.nbondj == -1)      // has a north neighbour: my row nlcj <- its row 2
                    upd[r].push_back({(i - 1) + (d.nlcj - 1) * d.jpi, S[d.nono].get(i, 2)});
            }
        }
        for (int r = 0; r < n; ++r) for (auto &u : upd[r]) S[r].m[u.first] = u.second;
    }
    // --- flatten
    std::vector<RankPlan> plan(n);
    for (int r = 0; r < n; ++r) {
        std::vector<std::pair<int, Ref>> cells(S[r].m.begin(), S[r].m.end());
        std::sort(cells.begin(), cells.end(), [](const auto &a, const auto &b) { return a.first < b.first; });
        std::map<int, std::vector<PlanCell>> by_peer;
        for (auto &c : cells) {
            const Ref &v = c.second;
            if (v.cst) { plan[r].fill.push_back(PlanCell{c.first, -1, v.spow}); continue; }
            if (v.rank == r && v.idx == c.first && v.spow == 0) continue;  // unchanged
            by_peer[v.rank].push_back(PlanCell{c.first, v.idx, v.spow});
        }
        for (auto &kv : by_peer) plan[r].recv.push_back(PeerList{kv.first, kv.second});
    }
    for (int r = 0; r < n; ++r)
        for (const PeerList &pl : plan[r].recv)
            plan[pl.peer].send.push_back(PeerList{r, pl.cells});
    for (int r = 0; r < n; ++r)
        std::sort(plan[r].send.begin(), plan[r].send.end(), [](const PeerList &a, const PeerList &b) { return a.peer < b.peer; });
    return plan;
}

}  // namespace nemo
