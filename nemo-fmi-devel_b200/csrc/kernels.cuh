// kernels.cuh -- launch interfaces of the sm_100a kernels of the FCT path (definitions in fct_kernels.cu and
// lbc_kernels.cu).  All arrays are Fortran column-major fp64 (ji fastest); indices in comments are 1-based as in
// the reference (src/OCE/TRA/traadv_fct.F90).  Everything is compiled with -fmad=false: the reference builds are
// IEEE-strict without FMA contraction (arch/arch-linux_gfortran.fcm), and with the reference's operation order
// kept inside each expression the results are bit-identical to the op-by-op CPU restatement (oracle/).
#pragma once
#include <cuda_runtime.h>
#include <cstddef>

namespace nemo {

// A set of up to 4 disjoint rectangles of (ji,jj) columns (1-based, inclusive) that one launch covers: the whole
// interior (2:jpim1, 2:jpjm1), the fused inner region, or the boundary frame around it.
struct Rect { int i0, i1, j0, j1; };
struct Region {
    int n = 0;
    Rect r[4];
    int start[5] = {0, 0, 0, 0, 0};      // prefix sums of the column counts
    int ncol() const { return start[n]; }
    void add(int i0, int i1, int j0, int j1) {
        if (i1 < i0 || j1 < j0) return;
        r[n] = Rect{i0, i1, j0, j1};
        start[n + 1] = start[n] + (i1 - i0 + 1) * (j1 - j0 + 1);
        ++n;
    }
};

struct FctArgs {
    Region reg;                            // columns this launch works on
    Rect out;                              // fused nonosc+final kernel: output rectangle
    int masks_from_t;                      // bit 0: umask/vmask/wmask are products of tmask (dommsk.F90:176-177,193): derive them; bit 1: tmask in {+0, 1}
    double *zlx, *zly, *zlz;               // schedule 1: limited fluxes of the frame path (separate from zwx/zwy/zwz)
    int jpi, jpj, jpk;
    size_t jpij, n3;                       // jpi*jpj, jpi*jpj*jpk
    // dom_oce module arrays
    const double *tmask, *umask, *vmask, *wmask, *e3t_b, *e3t_n, *e3t_a, *e1e2t, *r1_e1e2t;
    const int *mikt, *mbkt;
    const double *cpt_zwt;                 // time-invariant Thomas pivots of interp_4th_cpt (traadv_fct.F90:577-588)
    // tra_adv_fct arguments
    const double *pun, *pvn, *pwn, *ptb, *ptn;
    double *pta;
    // work arrays (jpi,jpj,jpk,kjpt) -- the automatic arrays of traadv_fct.F90:86 and :351, batched over tracers
    double *zwi, *zwx, *zwy, *zwz, *zltu, *zltv, *ztw, *zbetup, *zbetdo;
    double p2dt;
    int kjpt, kn_fct_h, kn_fct_v, ln_linssh, ln_isfcav;
    int nkchunk;                           // the jk loop is split in nkchunk chunks across blockIdx.y
    int arith;                             // k_fct_fused: 0 = IEEE divisions (default), 1 = relaxed (nemo_fct_set_arithmetic)
};

// P4 (order 4): zltu, zltv on (2:jpim1,2:jpjm1,1:jpkm1)                     traadv_fct.F90:195-208
void launch_fct_laplacian(const FctArgs &a, cudaStream_t s);
// P1-P5: upstream fluxes, low-order update (pta, zwi), anti-diffusive fluxes  traadv_fct.F90:123-278
void launch_fct_low_antidiff(const FctArgs &a, cudaStream_t s);
// nonosc P6: zbetup, zbetdo                                                 traadv_fct.F90:356-399
void launch_fct_betas(const FctArgs &a, cudaStream_t s);
// nonosc P7: limit zwx, zwy, zwz in place                                   traadv_fct.F90:404-425
void launch_fct_limit(const FctArgs &a, cudaStream_t s);
// P8: final trend                                                           traadv_fct.F90:288-297
void launch_fct_final(const FctArgs &a, cudaStream_t s);
// schedule 1, inner region: P1-P5 with the 4th-order Laplacian computed in place (no zltu/zltv arrays, no X1)
void launch_fct_low_antidiff_inner(const FctArgs &a, cudaStream_t s);
// same, with TMA-staged 32x8 tiles (+halo) in a 3-stage shared-memory ring; false if TMA cannot be used (odd jpi)
bool launch_fct_low_antidiff_tma(const FctArgs &a, cudaStream_t s);
// schedule 1, inner region: nonosc (P6, P7) + final trend (P8) in one kernel, betas shared through shared memory
void launch_fct_nonosc_final(const FctArgs &a, cudaStream_t s);
// same, streamed arrays through a 2-stage TMA ring; false if TMA cannot be used
bool launch_fct_nonosc_final_tma(const FctArgs &a, cudaStream_t s);

// schedule 4, inner region: the whole step (P1-P8) in ONE kernel on the rectangle a.out -- no intermediate array in HBM.
// The 11 tensor maps depend only on (pointers, shape): they are encoded once and kept in `cache` (one per context).
struct TmaMapCache {
    bool valid = false;
    const void *key[12] = {};
    int dims[4] = {};
    alignas(64) unsigned char maps[11 * 128];
};
// prepare: false if the TMA path cannot be used (odd jpi, even out.i0, masks not derived from tmask, unaligned arrays, no
// driver entry point) -- nothing has been launched then and the caller falls back to the three-kernel schedule
bool prepare_fct_fused(const FctArgs &a, TmaMapCache *cache);
// max_blocks: persistent grid size (<= number of SMs to occupy)
void launch_fct_fused(const FctArgs &a, cudaStream_t s, const TmaMapCache *cache, int max_blocks);

// interp_4th_cpt                                                            traadv_fct.F90:517-616
void launch_cpt_pivots(int jpi, int jpj, int jpk, const double *wmask, const int *mikt, const int *mbkt,
                       int ln_isfcav, double *zwt, cudaStream_t s);
// per-column check that pivots and wmask follow from mbkt alone (utab = pivots of a full-depth column)
void launch_cpt_classify(int jpi, int jpj, int jpk, const double *wmask, const int *mikt, const int *mbkt, const double *zwt,
                         const double *utab, unsigned char *simple, cudaStream_t s);
void launch_interp_4th_cpt(int jpi, int jpj, int jpk, int nfld, const double *wmask, const int *mikt, const int *mbkt,
                           int ln_isfcav, const double *zwt, const unsigned char *simple, const double *utab,
                           const double *pt_in, double *pt_out, cudaStream_t s, struct TmaMapCache *cache = nullptr, int jlo = 0, int jhi = 0);   // jlo..jhi: rows to solve (tiled kernel; 0 = all)
// the same on the columns of `reg` only (the frame bands of schedule 4), column kernel.  The forward sweep is parked in
// `scratch` (same shape), never in pt_out: another stream may be reading pt_out on these columns (same final values)
void launch_interp_4th_cpt_region(const Region &reg, int jpi, int jpj, int jpk, int nfld, const double *wmask, const int *mikt, const int *mbkt,
                                  const double *zwt, const unsigned char *simple, const double *utab, const double *pt_in, double *pt_out,
                                  double *scratch, cudaStream_t s);
// tra_adv transports                                                        traadv.F90:100-124
void launch_transports(int jpi, int jpj, int jpk, const double *e2u, const double *e1v, const double *e1e2t,
                       const double *e3u_n, const double *e3v_n, const double *un, const double *vn,
                       const double *wn, double *zun, double *zvn, double *zwn, cudaStream_t s);

// ---- tra_adv_mus (MUSCL)                                                    traadv_mus.F90:55-273 ----
struct MusArgs {
    Region reg;                            // columns this launch works on
    int jpi, jpj, jpk;
    size_t jpij, n3;
    const double *tmask, *umask, *vmask, *wmask, *e3t_n, *r1_e1e2t;
    const double *r1_e1e2u, *r1_e1e2v, *e3u_n, *e3v_n, *e3w_n;
    const double *xind;                    // upstream indicator, NULL = 1 everywhere (ld_msc_ups = .FALSE.)
    const int *mikt;
    const double *pun, *pvn, *pwn, *ptb;
    double *pta;
    double *zwx, *zwy, *fx, *fy;           // first-guess differences and horizontal fluxes (jpi,jpj,jpk,kjpt)
    double p2dt;
    int kjpt, ln_linssh, ln_isfcav, nkchunk;
};
void launch_mus_grad(const MusArgs &a, cudaStream_t s);     // first guess of the slopes         :134-141
void launch_mus_hflux(const MusArgs &a, cudaStream_t s, bool from_ptb = false);   // slopes, limitation, fluxes :145-191
void launch_mus_trend(const MusArgs &a, cudaStream_t s);    // horizontal trend + vertical part  :194-272
void launch_mus_inner(const MusArgs &a, cudaStream_t s);    // everything, exchange-free columns
void launch_mus_xind(int jpi, int jpj, int jpk, const double *rnfmsk, const double *rnfmsk_z, const double *tmask, double *xind,
                     cudaStream_t s);                        // :99-113

// ---- tra_adv_cen                                                            traadv_cen.F90:46-204 ----
struct CenArgs {
    Region reg;
    int jpi, jpj, jpk;
    size_t jpij, n3;
    const double *wmask, *e3t_n, *r1_e1e2t;
    const int *mikt;
    const double *pun, *pvn, *pwn, *ptn;
    double *pta;
    const double *ztu, *ztv, *ztw;         // exchanged masked gradients (order 4), compact interpolation (vertical order 4)
    int kjpt, kn_cen_h, kn_cen_v, ln_linssh, ln_isfcav, nkchunk;
};
void launch_cen(const CenArgs &a, cudaStream_t s);

// l_trd / l_hst / l_ptr hooks of tra_adv_fct (traadv_fct.F90:172-176, 299-303): total advective fluxes = upstream fluxes
// (recomputed from ptb and the transports) + the limited anti-diffusive fluxes left in zwx, zwy, zwz by the
// reference-structured schedule.  trdx / trdy on (1:jpim1, 1:jpjm1, 1:jpk), trdz on (1:jpi, 1:jpj, 1:jpk).
void launch_fct_diag(const FctArgs &a, double *trdx, double *trdy, double *trdz, cudaStream_t s);

// ---- tra_nxt_fix / tra_nxt_vvl / Euler swap                                 tranxt.F90:148-153, 190-380 ----
struct NxtArgs {
    int jpi, jpj, jpk, kjpt;
    size_t jpij, n3;
    const double *e3t_b, *e3t_n, *e3t_a;
    const int *mikt;
    double *ptb, *ptn, *pta;
    double atfp, zfact1, zfact2;           // atfp, atfp*p2dt, atfp*p2dt*r1_rau0   (:283-284)
    int ll_traqsr, ll_rnf, ll_isf, ln_rnf_depth, nksr;
    const double *sbc_tc, *sbc_tc_b;       // (jpi,jpj,kjpt) or NULL
    const double *emp_b, *emp, *fwfisf_b, *fwfisf, *rnf_b, *rnf;     // (jpi,jpj) or NULL (= zeros)
    const double *qsr_hc, *qsr_hc_b;       // (jpi,jpj,jpk)
    const int *nk_rnf; const double *h_rnf, *rnf_tsc, *rnf_tsc_b;
    const int *misfkt, *misfkb; const double *risf_tsc, *risf_tsc_b, *r1_hisf_tbl, *ralpha;
};
void launch_nxt_fix(const NxtArgs &a, cudaStream_t s);
void launch_nxt_vvl(const NxtArgs &a, cudaStream_t s);
void launch_nxt_euler(const NxtArgs &a, int also_before, cudaStream_t s);

// ---- lbc_lnk on the device: gather plan execution -------------------------------------------------------------
// One job = one (peer, field) pair: `ncell` cells x `nlev` levels.
struct PackJob   { const double *field; const int *src; double *buf; int ncell; };                       // buf[k*ncell+c] = field[src[c] + k*jpij]
struct UnpackJob { double *field; const int *dst; const int *spow; const double *buf; int ncell; double sgn; }; // field[dst[c]+k*jpij] = sgn^spow[c] * buf[k*ncell+c]
struct FillJob   { double *field; const int *dst; const int *spow; int ncell; double sgn; double zland; };
void launch_lbc_pack(const PackJob *jobs_dev, int njobs, int maxcell, int nlev, size_t jpij, cudaStream_t s);
void launch_lbc_unpack(const UnpackJob *jobs_dev, int njobs, int maxcell, int nlev, size_t jpij, cudaStream_t s);
void launch_lbc_fill(const FillJob *jobs_dev, int njobs, int maxcell, int nlev, size_t jpij, cudaStream_t s);

// ---- glob_sum (lib_fortran_generic.h90:32-65): double-double masked sums of nfld fields, out_pairs[2*f] = REAL, [2*f+1] = AIMAG
constexpr int kGlobSumBlocks = 148 * 4;
void launch_glob_sum(const double *const *ptab_dev, int nfld, const double *pw3d, const double *tmask_i, size_t jpij, int ipk, double *partial,
                     double *out_pairs, cudaStream_t s);

// ---- stp_ctl (stpctl.F90:115-124, 162-165): extrema of |sshn|, |un|, S, T (tracers where tmask == 1) with the linear index of
// the first occurrence (-1: no operand), flags bit 0 = NaN met in sshn / un / masked S, bit 1 = some tmask == 1
struct StpCtlRec { double z1, z2, smin, smax, tmin, tmax; long long l1, l2, ls1, ls2, lt1, lt2; int flags, pad; };
void launch_stp_ctl(const double *sshn, const double *un, const double *tem, const double *sal, const double *tmask, size_t jpij, size_t n3,
                    StpCtlRec *partial, StpCtlRec *out, cudaStream_t s);

long long kernel_launch_count();
// div_rn (fct_fused_kernel.cuh) against x / y on n operand pairs per class; returns the number of differing results, < 0 on error
long long division_selftest(long long n, unsigned long long seed, cudaStream_t s);

}  // namespace nemo
