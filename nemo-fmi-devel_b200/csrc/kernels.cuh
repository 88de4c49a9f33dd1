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
    int masks_from_t;                      // umask/vmask/wmask are products of tmask (dommsk.F90:176-177,193): derive them
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

// interp_4th_cpt                                                            traadv_fct.F90:517-616
void launch_cpt_pivots(int jpi, int jpj, int jpk, const double *wmask, const int *mikt, const int *mbkt,
                       int ln_isfcav, double *zwt, cudaStream_t s);
// per-column check that pivots and wmask follow from mbkt alone (utab = pivots of a full-depth column)
void launch_cpt_classify(int jpi, int jpj, int jpk, const double *wmask, const int *mikt, const int *mbkt, const double *zwt,
                         const double *utab, unsigned char *simple, cudaStream_t s);
void launch_interp_4th_cpt(int jpi, int jpj, int jpk, int nfld, const double *wmask, const int *mikt, const int *mbkt,
                           int ln_isfcav, const double *zwt, const unsigned char *simple, const double *utab,
                           const double *pt_in, double *pt_out, cudaStream_t s);
// tra_adv transports                                                        traadv.F90:100-124
void launch_transports(int jpi, int jpj, int jpk, const double *e2u, const double *e1v, const double *e1e2t,
                       const double *e3u_n, const double *e3v_n, const double *un, const double *vn,
                       const double *wn, double *zun, double *zvn, double *zwn, cudaStream_t s);

// ---- lbc_lnk on the device: gather plan execution -------------------------------------------------------------
// One job = one (peer, field) pair: `ncell` cells x `nlev` levels.
struct PackJob   { const double *field; const int *src; double *buf; int ncell; };                       // buf[k*ncell+c] = field[src[c] + k*jpij]
struct UnpackJob { double *field; const int *dst; const int *spow; const double *buf; int ncell; double sgn; }; // field[dst[c]+k*jpij] = sgn^spow[c] * buf[k*ncell+c]
struct FillJob   { double *field; const int *dst; const int *spow; int ncell; double sgn; double zland; };
void launch_lbc_pack(const PackJob *jobs_dev, int njobs, int maxcell, int nlev, size_t jpij, cudaStream_t s);
void launch_lbc_unpack(const UnpackJob *jobs_dev, int njobs, int maxcell, int nlev, size_t jpij, cudaStream_t s);
void launch_lbc_fill(const FillJob *jobs_dev, int njobs, int maxcell, int nlev, size_t jpij, cudaStream_t s);

long long kernel_launch_count();

}  // namespace nemo
