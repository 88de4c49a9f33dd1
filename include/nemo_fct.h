/*
 * nemo_fct.h -- C ABI of the B200-native FCT tracer-advection path (libnemo_fct.so).
 *
 * This is the drop-in boundary: exactly the entry points a NEMO host (Fortran, through ISO_C_BINDING) binds to
 * replace its own routines on the tra_adv_fct path.  Plain pointers and sizes only, no torch / C++ types.
 * Every entry point cites the reference interface it replaces (paths relative to the NEMO tree).
 *
 *   reference routine                                             ->  entry point here
 *   ---------------------------------------------------------------------------------------------------------
 *   mpp_init              src/OCE/LBC/mppini.F90:110-692          ->  nemo_mpp_init            (host only)
 *   nemo_alloc/dom_init   src/OCE/nemogcm.F90:640-673 (+dom_oce)  ->  nemo_fct_create / nemo_fct_set_domain_arrays
 *   dom_vvl_sf_swp        src/OCE/DOM/domvvl.F90:620 (e3t swap)    ->  nemo_fct_set_e3t
 *   tra_adv_fct           src/OCE/TRA/traadv_fct.F90:54-80        ->  nemo_tra_adv_fct[_dev]
 *     (call sites         src/OCE/TRA/traadv.F90:150, src/TOP/TRP/trcadv.F90:127)
 *   interp_4th_cpt        src/OCE/TRA/traadv_fct.F90:517-527      ->  nemo_interp_4th_cpt[_dev]
 *   tra_adv transports    src/OCE/TRA/traadv.F90:100-124          ->  nemo_tra_adv_transports_dev
 *   tra_adv / trc_adv     src/OCE/TRA/traadv.F90:77, src/TOP/TRP/trcadv.F90:70 -> nemo_tra_adv_dev / nemo_trc_adv_dev
 *   tra_adv_mus           src/OCE/TRA/traadv_mus.F90:55-79        ->  nemo_tra_adv_mus[_dev] (+ nemo_fct_set_mus_*, nemo_fct_set_e3uvw)
 *   tra_adv_cen           src/OCE/TRA/traadv_cen.F90:46-77        ->  nemo_tra_adv_cen_dev
 *   tra_nxt / trc_nxt     src/OCE/TRA/tranxt.F90:65-380, src/TOP/TRP/trcnxt.F90:56-183 -> nemo_tra_nxt_dev
 *   l_trd/l_hst/l_ptr     src/OCE/TRA/traadv_fct.F90:96-112,172-176,299-316 -> nemo_fct_set_trend_diag
 *   lbc_lnk_multi         src/OCE/LBC/lbc_lnk_multi_generic.h90:16-29 -> nemo_lbc_lnk_multi[_dev]
 *   mynode / MPI_Init     src/OCE/LBC/lib_mpp.F90:197-331         ->  nemo_fct_comm_unique_id / nemo_fct_comm_init
 *   glob_sum              src/OCE/lib_fortran_generic.h90:32-65    ->  nemo_glob_sum_dev
 *   stp_ctl (extrema)     src/OCE/stpctl.F90:115-186               ->  nemo_stp_ctl_dev
 *   ctl_stop              src/OCE/LBC/lib_mpp.F90:1868-1907       ->  non-zero return + nemo_fct_last_error
 *
 * Conventions
 *   - All arrays are Fortran column-major REAL(wp)=fp64: a(jpi,jpj,jpk[,kjpt]) == double[kjpt][jpk][jpj][jpi].
 *     INTEGER arrays are 32-bit.  Indices documented 1-based as in the reference.
 *   - "host" pointers are ordinary CPU memory (pageable or pinned); "_dev" variants take CUDA device pointers
 *     valid on the context's device and are asynchronous on the context's stream.
 *   - Every function returns 0 on success, non-zero on error; nemo_fct_last_error() gives the message.  The
 *     Fortran shim maps non-zero to CALL ctl_stop('STOP', msg).  There is NO CPU fallback: if no CUDA device is
 *     usable, nemo_fct_create fails.
 *   - SPMD contract as in the reference (lib_mpp.F90:1497-1506): with jpnij > 1 every rank calls the same
 *     entry points in the same order.
 */
#ifndef NEMO_FCT_H
#define NEMO_FCT_H

#include <stddef.h>   /* size_t */

#ifdef __cplusplus
extern "C" {
#endif

#define NEMO_FCT_ABI_VERSION 1
#define NEMO_FCT_JPMAXNGH 3            /* lbcnfd.F90:53 */
#define NEMO_FCT_UNIQUE_ID_BYTES 128   /* sizeof(ncclUniqueId) */

typedef struct nemo_fct_ctx *nemo_fct_handle;

/* Decomposition scalars of one subdomain: the PUBLIC variables of par_oce.F90 / dom_oce.F90 that mpp_init sets
 * (mppini.F90:548-580, 621-628) plus the global description.  Filled by nemo_mpp_init, or by the host from its
 * own mpp_init -- nemo_fct_create cross-checks them against its own decomposition and fails on mismatch.       */
typedef struct nemo_fct_domain {
    int jpiglo, jpjglo, jpk;          /* global horizontal size, number of levels                               */
    int jperio;                       /* lateral boundary type 0..7 (dom_oce.F90, chap_LBC.tex:149-260)         */
    int jpni, jpnj;                   /* processor grid                                                         */
    int narea;                        /* this rank, 1-based (nproc = narea-1)                                   */
    int jpi, jpj;                     /* local array extents (= nlci, nlcj)                                     */
    int jpimax, jpjmax;
    int nimpp, njmpp;                 /* global index of local (1,1)                                            */
    int nlci, nlcj, nldi, nlei, nldj, nlej;
    int nbondi, nbondj;               /* -1 / 0 / 1 / 2 neighbour flags                                         */
    int noea, nowe, noso, nono;       /* neighbour ranks (0-based), -1 if none                                  */
    int npolj;                        /* north fold type of this rank: 0 / 3,4 (T pivot) / 5,6 (F pivot)        */
    int l_Iperio, l_Jperio;           /* periodicity handled locally (mppini.F90:345-346)                       */
    int nsndto, isendto[NEMO_FCT_JPMAXNGH];   /* no-gather fold partners (mppini.F90:1180-1240)                  */
    int key_mpp_mpi;                  /* 0: single-domain build semantics (mppini.F90:53-102, npolj = jperio)   */
} nemo_fct_domain;

/* ---- decomposition (host only, no GPU needed) -------------------------------------------------------------- */
/* mpp_init for rank `narea` of a jpni x jpnj layout (all-ocean: no land-subdomain elimination).                 */
int nemo_mpp_init(int jpiglo, int jpjglo, int jpk, int jperio, int jpni, int jpnj, int narea, int key_mpp_mpi,
                  nemo_fct_domain *out);
/* mpp_basic_decomposition (mppini.F90:695-798): tables are (jpni,jpnj) column-major; any may be NULL.           */
int nemo_mpp_basic_decomposition(int jpiglo, int jpjglo, int jperio, int jpni, int jpnj, int *jpimax, int *jpjmax,
                                 int *nimppt, int *njmppt, int *nlcit, int *nlcjt);
/* The compiled lbc_lnk exchange plan of one rank for grid-point type cd_nat in "TUVWF": number of halo / fold
 * cells this rank receives from `peer` (0-based; peer == own rank counts local copies), or fills with the land
 * value when peer == -1.  With non-NULL outputs (each sized to the returned count) also returns, per cell and
 * in message order: the destination local index (i-1)+(j-1)*jpi, the source local index on the peer, and the
 * power of psgn applied.  Host only; used by the CPU tests to execute the plan over gloo and compare with the
 * oracle's mpp_lnk.                                                                                             */
int nemo_lbc_plan_query(const nemo_fct_domain *dom, char cd_nat, int peer, int *dst_index, int *src_index,
                        int *sgn_power);

/* ---- life cycle ------------------------------------------------------------------------------------------------ */
/* Create the device context of one subdomain on CUDA device `device` (-1: use LOCAL_RANK, else 0).  Allocates
 * streams, the lbc_lnk plans and (lazily, sized by the first call) the work arrays, ONCE, as nemo_alloc does
 * (nemogcm.F90:640-673).  Fails if no CUDA device is available.                                                  */
int nemo_fct_create(const nemo_fct_domain *dom, int device, nemo_fct_handle *out);
int nemo_fct_destroy(nemo_fct_handle h);
/* Time-invariant module arrays of dom_oce.F90 (host pointers, copied to the device):
 * tmask,umask,vmask,wmask (jpi,jpj,jpk); e1e2t,r1_e1e2t (jpi,jpj); mikt,mbkt INTEGER (jpi,jpj); flags of
 * dom_oce.F90 (ln_linssh) and of the ice-shelf-cavity option (ln_isfcav).                                         */
int nemo_fct_set_domain_arrays(nemo_fct_handle h, const double *tmask, const double *umask, const double *vmask,
                               const double *wmask, const double *e1e2t, const double *r1_e1e2t, const int *mikt,
                               const int *mbkt, int ln_linssh, int ln_isfcav);
/* Vertical scale factors e3t_b, e3t_n, e3t_a (jpi,jpj,jpk), time-varying when .NOT.ln_linssh.
 * is_device != 0: the pointers are device pointers that the context borrows (no copy).                            */
int nemo_fct_set_e3t(nemo_fct_handle h, const double *e3t_b, const double *e3t_n, const double *e3t_a, int is_device);
/* Run all subsequent work of this context on the caller's CUDA stream (cudaStream_t passed as void*; NULL
 * restores the context's own stream).                                                                             */
int nemo_fct_set_stream(nemo_fct_handle h, void *cuda_stream);
int nemo_fct_synchronize(nemo_fct_handle h);

/* ---- communicator (jpnij > 1, one process per GPU) -------------------------------------------------------------- */
/* Rank 0 obtains an id, the host broadcasts the 128 bytes (MPI_Bcast in NEMO, torch.distributed in bench.py),
 * then every rank calls nemo_fct_comm_init.  The halo strips and the north fold then move with ncclSend/ncclRecv. */
int nemo_fct_comm_unique_id(void *id128);
int nemo_fct_comm_init(nemo_fct_handle h, const void *id128, int nranks, int rank);
/* In-process communicator: all `n` subdomains live in this process on the SAME device (decomposition tests on
 * one GPU); collective entry points are then the nemo_group_* calls below.                                        */
int nemo_fct_comm_init_local(nemo_fct_handle *hs, int n);

/* ---- the hot path --------------------------------------------------------------------------------------------------- */
/* tra_adv_fct (traadv_fct.F90:54-80), same argument list.  pun,pvn,pwn (jpi,jpj,jpk) are TRANSPORTS; ptb,ptn,pta
 * (jpi,jpj,jpk,kjpt).  Only pta(2:jpim1,2:jpjm1,1:jpkm1,:) is modified.  kn_fct_h, kn_fct_v in {2,4}.
 * Host variant: synchronous, copies inputs to the device and pta back.                                            */
int nemo_tra_adv_fct(nemo_fct_handle h, int kt, int kit000, const char *cdtype, double p2dt, const double *pun,
                     const double *pvn, const double *pwn, const double *ptb, const double *ptn, double *pta,
                     int kjpt, int kn_fct_h, int kn_fct_v);
/* Device-resident variant: all pointers are device pointers; asynchronous on the context's stream.               */
int nemo_tra_adv_fct_dev(nemo_fct_handle h, int kt, int kit000, const char *cdtype, double p2dt, const double *pun,
                         const double *pvn, const double *pwn, const double *ptb, const double *ptn, double *pta,
                         int kjpt, int kn_fct_h, int kn_fct_v);
/* In-process group variant (nemo_fct_comm_init_local): argument tables indexed like hs[].                         */
int nemo_group_tra_adv_fct_dev(nemo_fct_handle *hs, int n, int kt, int kit000, const char *cdtype, double p2dt,
                               const double *const *pun, const double *const *pvn, const double *const *pwn,
                               const double *const *ptb, const double *const *ptn, double *const *pta, int kjpt,
                               int kn_fct_h, int kn_fct_v);

/* interp_4th_cpt (traadv_fct.F90:517-527; also called by traadv_cen.F90:158): pt_in, pt_out (jpi,jpj,jpk);
 * pt_out is defined on (2:jpim1, 2:jpjm1, 2:jpkm1) only.                                                          */
int nemo_interp_4th_cpt(nemo_fct_handle h, const double *pt_in, double *pt_out);
int nemo_interp_4th_cpt_dev(nemo_fct_handle h, const double *pt_in, double *pt_out);

/* Effective transports of tra_adv / trc_adv (traadv.F90:100-124, trcadv.F90:93-108), Eulerian branch:
 * zun = e2u*e3u_n*un, zvn = e1v*e3v_n*vn, zwn = e1e2t*wn, level jpk zeroed.  Device pointers.                      */
int nemo_tra_adv_transports_dev(nemo_fct_handle h, const double *e2u, const double *e1v, const double *e3u_n,
                                const double *e3v_n, const double *un, const double *vn, const double *wn,
                                double *zun, double *zvn, double *zwn);

/* tra_adv (traadv.F90:77-175) for device-resident state, FCT branch (nadv = np_FCT): sets r2dt from (kt, nit000, neuler,
 * rdt) as :95-97, builds the effective transports from un, vn, wn (Eulerian branch :110-124; Stokes drift, z-tilde, eiv and
 * mle additions are not applied) into work arrays owned by the context, and calls tra_adv_fct on tsb, tsn, tsa (jpts tracers).
 * nemo_trc_adv_dev is trc_adv (trcadv.F90:70-145) for the passive tracers: it REUSES the transports of the last
 * nemo_tra_adv_dev call instead of rebuilding them (the reference recomputes the same three arrays, trcadv.F90:93-108).   */
/* The additions of tra_adv to the Eulerian transports that nemo_tra_adv_dev does NOT apply: Stokes drift (ln_wave .AND. ln_sdw,
 * traadv.F90:103-107), the z-tilde / layer thickness transports (ln_vvl_ztilde .OR. ln_vvl_layer, :116-119), the eddy-induced
 * transport (ln_ldfeiv .AND. .NOT. ln_ldfeiv_dia, :126-127) and the mixed-layer eddy transport (ln_mle, :129).  The host
 * declares its namelist switches once; with any of them set nemo_tra_adv_dev / nemo_trc_adv_dev fail (no silent omission):
 * such a host builds zun, zvn, zwn itself and calls nemo_tra_adv_fct_dev.                                                */
int nemo_fct_declare_transport_options(nemo_fct_handle h, int ln_wave_sdw, int ln_vvl_ztilde_or_layer, int ln_ldfeiv, int ln_mle);
int nemo_tra_adv_dev(nemo_fct_handle h, int kt, int nit000, int neuler, double rdt, const double *e2u, const double *e1v,
                     const double *e3u_n, const double *e3v_n, const double *un, const double *vn, const double *wn,
                     const double *tsb, const double *tsn, double *tsa, int jpts, int nn_fct_h, int nn_fct_v);
int nemo_trc_adv_dev(nemo_fct_handle h, int kt, int nittrc000, double r2dttrc, const double *trb, const double *trn,
                     double *tra, int jptra, int nn_fct_h, int nn_fct_v);

/* ---- tra_adv_mus: MUSCL scheme (traadv_mus.F90:55-273), the scheme BENCH and ORCA2_ICE_PISCES select for TOP ---------- */
/* Extra module arrays it reads: r1_e1e2u, r1_e1e2v (jpi,jpj) (dom_oce.F90:118; HOST pointers, time-invariant, copied) and
 * e3u_n, e3v_n, e3w_n (jpi,jpj,jpk) (dom_oce.F90:132-136; time-varying with ln_linssh = F: is_device != 0 borrows device
 * pointers that must stay valid until replaced, else host arrays are copied).                                           */
int nemo_fct_set_mus_metrics(nemo_fct_handle h, const double *r1_e1e2u, const double *r1_e1e2v);
int nemo_fct_set_e3uvw(nemo_fct_handle h, const double *e3u_n, const double *e3v_n, const double *e3w_n, int is_device);
/* Upstream indicator xind (traadv_mus.F90:99-113, built by the reference at kt == kit000): ld_msc_ups = 0 -> xind = 1
 * (default, nothing stored); else xind = 1 - MAX(rnfmsk*rnfmsk_z(jk), upsmsk = 0)*tmask from HOST rnfmsk (jpi,jpj), rnfmsk_z (jpk) */
int nemo_fct_set_mus_upstream(nemo_fct_handle h, int ld_msc_ups, const double *rnfmsk, const double *rnfmsk_z);
/* CALL tra_adv_mus( kt, kit000, cdtype, p2dt, pun, pvn, pwn, ptb, pta, kjpt, ld_msc_ups ) -- ld_msc_ups is the state set by
 * nemo_fct_set_mus_upstream.  Only pta(2:jpim1, 2:jpjm1, 1:jpkm1, :) is modified.  Collective like tra_adv_fct (2 exchanges).
 * Schedules (nemo_fct_set_schedule), identical results bit for bit: 0 = three reference-structured kernels on the whole
 * interior, both exchanges on the main stream; 1 = one fused kernel on the exchange-free inner columns + the
 * reference-structured kernels on the two-cell frame, overlapped on a side stream (measured slower: kept for study);
 * >= 2 (default) = the flux kernel forms the first-guess differences in place from ptb on the columns that need no
 * exchanged value, so the first exchange only serves the one-cell frame (side stream); second exchange and trend kernel as 0.
 * Subdomains smaller than 20 x 20 always use 0.                                                                           */
int nemo_tra_adv_mus(nemo_fct_handle h, int kt, int kit000, const char *cdtype, double p2dt, const double *pun,
                     const double *pvn, const double *pwn, const double *ptb, double *pta, int kjpt);
int nemo_tra_adv_mus_dev(nemo_fct_handle h, int kt, int kit000, const char *cdtype, double p2dt, const double *pun,
                         const double *pvn, const double *pwn, const double *ptb, double *pta, int kjpt);
int nemo_group_tra_adv_mus_dev(nemo_fct_handle *hs, int n, int kt, int kit000, const char *cdtype, double p2dt,
                               const double *const *pun, const double *const *pvn, const double *const *pwn,
                               const double *const *ptb, double *const *pta, int kjpt);

/* ---- tra_adv_cen: centred scheme (traadv_cen.F90:46-204) -------------------------------------------------------------- */
/* CALL tra_adv_cen( kt, kit000, cdtype, pun, pvn, pwn, ptn, pta, kjpt, kn_cen_h, kn_cen_v ), kn_cen_h, kn_cen_v in {2, 4}
 * (4 in the vertical = the compact scheme of interp_4th_cpt).  Device pointers.  kn_cen_h = 2 needs no exchange at all;
 * kn_cen_h = 4 does the reference's lbc_lnk on the masked gradients (:126) and reproduces what its loop bounds do at the
 * first interior row / column (:128-137 read ztu(0,jj,jk) and the never-assigned zwy(:,1,:), taken as 0 -- DESIGN.md).    */
int nemo_tra_adv_cen_dev(nemo_fct_handle h, int kt, int kit000, const char *cdtype, const double *pun, const double *pvn,
                         const double *pwn, const double *ptn, double *pta, int kjpt, int kn_cen_h, int kn_cen_v);
int nemo_group_tra_adv_cen_dev(nemo_fct_handle *hs, int n, int kt, int kit000, const char *cdtype,
                               const double *const *pun, const double *const *pvn, const double *const *pwn,
                               const double *const *ptn, double *const *pta, int kjpt, int kn_cen_h, int kn_cen_v);

/* ---- l_trd / l_hst / l_ptr hooks of tra_adv_fct (traadv_fct.F90:96-112, 172-176, 299-316) ------------------------------- */
/* With three non-NULL DEVICE arrays (jpi,jpj,jpk,kjpt) every following nemo_tra_adv_fct[_dev] call also returns the total
 * advective fluxes the reference hands to trd_tra / dia_ar5_hst / dia_ptr_hst: ztrdx, ztrdy (= zptry) = upstream + limited
 * anti-diffusive flux on (1:jpim1, 1:jpjm1, 1:jpk) (column jpi / row jpj are undefined in the reference and left
 * untouched), ztrdz on the whole (jpi,jpj,jpk).  The step then runs the reference pass structure (schedule 0), where the
 * limited fluxes exist in memory.  Three NULLs switch the hooks off.  The consumers themselves (trd_tra, dia_*) stay on
 * the host side of the boundary.                                                                                          */
int nemo_fct_set_trend_diag(nemo_fct_handle h, double *ztrdx, double *ztrdy, double *ztrdz);

/* ---- tra_nxt / trc_nxt: lateral boundary conditions on the after field, Asselin filter, swap ---------------------------- */
/* Module variables read by tra_nxt_vvl (tranxt.F90:262-343; sbc_oce, sbcrnf, sbcisf, traqsr, phycst).  All pointers are
 * DEVICE pointers; a NULL 2-D flux array stands for zeros.  The arrays behind an enabled switch must be non-NULL.          */
typedef struct nemo_nxt_forcing {
    double atfp, r1_rau0;                              /* dom_oce.F90:58, phycst.F90:42 */
    int ln_traqsr, ln_rnf, ln_isf, ln_rnf_depth, nksr;
    const double *emp_b, *emp, *fwfisf_b, *fwfisf, *rnf_b, *rnf;        /* (jpi,jpj) */
    const double *qsr_hc, *qsr_hc_b;                                    /* (jpi,jpj,jpk) */
    const int *nk_rnf; const double *h_rnf, *rnf_tsc, *rnf_tsc_b;       /* (jpi,jpj), (jpi,jpj,jpts) */
    const int *misfkt, *misfkb;                                         /* (jpi,jpj) */
    const double *risf_tsc, *risf_tsc_b, *r1_hisf_tbl, *ralpha;         /* (jpi,jpj,jpts), (jpi,jpj) */
} nemo_nxt_forcing;
/* tra_nxt (cdtype "TRA", tranxt.F90:65-187) / trc_nxt (cdtype "TRC", trcnxt.F90:56-183) on device-resident ptb, ptn, pta
 * (jpi,jpj,jpk,kjpt): lbc_lnk on pta; then, if l_euler (neuler == 0 .AND. kt == nit000 [.OR. ln_top_euler]) the Euler swap
 * ptn = pta (and ptb = ptn for TRC), else tra_nxt_fix (ln_linssh) or tra_nxt_vvl(p2dt = rdt; psbc_tc, psbc_tc_b (jpi,jpj,kjpt)
 * device pointers or NULL) followed by lbc_lnk on ptb, ptn, pta.  AGRIF, ln_bdy and the l_trdtra trends are not applied.   */
int nemo_tra_nxt_dev(nemo_fct_handle h, int kt, int nit000, int l_euler, double rdt, const char *cdtype,
                     const nemo_nxt_forcing *f, double *ptb, double *ptn, double *pta, const double *psbc_tc,
                     const double *psbc_tc_b, int kjpt);
int nemo_group_tra_nxt_dev(nemo_fct_handle *hs, int n, int kt, int nit000, int l_euler, double rdt, const char *cdtype,
                           const nemo_nxt_forcing *const *f, double *const *ptb, double *const *ptn, double *const *pta,
                           const double *const *psbc_tc, const double *const *psbc_tc_b, int kjpt);

/* lbc_lnk_multi (lbc_lnk_multi_generic.h90:16-29): nfld fields ptab[f] of (jpi,jpj,ipk) each (a 4-D field is a
 * 3-D field with ipk*ipl levels), grid-point type cd_nat[f] in "TUVWF", fold sign psgn[f]; has_pval/pval = the
 * optional land value.  cd_mpp is not supported (not used on this path).                                          */
int nemo_lbc_lnk_multi(nemo_fct_handle h, const char *cdname, int nfld, double *const *ptab, const char *cd_nat,
                       const double *psgn, int ipk, int has_pval, double pval);
int nemo_lbc_lnk_multi_dev(nemo_fct_handle h, const char *cdname, int nfld, double *const *ptab,
                           const char *cd_nat, const double *psgn, int ipk, int has_pval, double pval);
int nemo_group_lbc_lnk_multi_dev(nemo_fct_handle *hs, int n, const char *cdname, int nfld,
                                 double *const *const *ptab, const char *cd_nat, const double *psgn, int ipk,
                                 int has_pval, double pval);

/* ---- diagnostics ------------------------------------------------------------------------------------------------------ */
const char *nemo_fct_last_error(void);
int nemo_fct_abi_version(void);
/* number of CUDA kernels this library has launched in the process so far (bench.py's gpu_launches)                */
long long nemo_fct_launch_count(void);
/* Page-lock host arrays the host-pointer entry points will be given (NEMO's module arrays are ALLOCATABLE, i.e. pageable: without
 * this the copies are staged through the driver's bounce buffer at a fraction of the PCIe rate and cannot overlap the step).
 * Call once after nemo_alloc for tsb, tsn, tsa, un, vn, wn, e3t_b/n/a ... (nemogcm.F90:659-664); unregister before deallocation. */
int nemo_fct_host_register(void *ptr, size_t bytes);
int nemo_fct_host_unregister(void *ptr);
/* glob_sum (src/OCE/lib_fortran_generic.h90:32-65; DDPDD lib_fortran.F90:300-332; MPI_SUMDD lib_mpp.F90:1158-1186): for each of
 * the nfld device fields ptab[f](jpi,jpj,ipk), out[f] = REAL( SUM in double-double of ptab[f] [* pw3d] * tmask_i ) over this
 * subdomain and, through the communicator, over all ranks -- the same bits on every rank and for every decomposition.  pw3d
 * (device, same shape, or NULL) is multiplied in point by point: the array expression the reference's callers form before the
 * call (e.g. tr * cvol, trcrad.F90).  tmask_i: dom_oce's interior mask (jpi,jpj), device.  out: host, nfld values.  Collective. */
int nemo_glob_sum_dev(nemo_fct_handle h, const char *cdname, int nfld, const double *const *ptab, const double *pw3d,
                      const double *tmask_i, int ipk, double *out);
int nemo_group_glob_sum_dev(nemo_fct_handle *hs, int n, const char *cdname, int nfld, const double *const *const *ptab,
                            const double *const *pw3d, const double *const *tmask_i, int ipk, double *out);
/* stp_ctl's extrema test on device-resident state (src/OCE/stpctl.F90:115-124 zmax(1:6), :149-156 the condition, :162-165 the
 * locations, :184 kindic).  sshn (jpi,jpj), un (jpi,jpj,jpk), tsn (jpi,jpj,jpk,2) with jp_tem = 1, jp_sal = 2: device pointers;
 * tmask is the one given to nemo_fct_set_domain_arrays.  zmax = max|sshn|, max|un|, max(-S), max(S), max(-T), max(T), the tracer
 * ones where tmask == 1 (MAXVAL of an empty set: -HUGE); ih, iu, is1, is2 = MAXLOC|sshn|, MAXLOC|un|, MINLOC S, MAXLOC S as 1-based
 * GLOBAL indices (first occurrence in array order).  kindic = -3 when the reference's condition holds (the ctl_stop text is
 * then what nemo_fct_last_error returns), else 0; the return value is 0 unless the call itself failed.  A NaN never wins a
 * comparison (MAXVAL over NaN is processor dependent in Fortran); nan_found reports one in sshn, un or the masked S, which is what
 * the reference's ISNAN( zmax(1)+zmax(2)+zmax(3) ) tests for.  collective != 0 (ln_ctl / sn_cfctl%l_runstat, :126-129, 157-160):
 * maxima and locations over all ranks of the communicator, identical on every rank, ties to the lowest rank as MPI_MAXLOC (the
 * locations are the unmasked ones of the local branch; mpp_maxloc's ssmask / umask change nothing while land values are zero).
 * The files stp_ctl writes (time.step, run.stat, output.abort) and zmax(8:9) (ln_zad_Aimp) stay with the host.               */
typedef struct nemo_stp_ctl_result {
    double zmax[6];
    int ih[2], iu[3], is1[3], is2[3];
    int nan_found;
    int kindic;
} nemo_stp_ctl_result;
int nemo_stp_ctl_dev(nemo_fct_handle h, int kt, const double *sshn, const double *un, const double *tsn, int collective,
                     nemo_stp_ctl_result *res);
/* Self-test of the inlined IEEE division of the fused kernel (csrc/fct_fused_kernel.cuh: div_rn): n pseudo-random operand
 * pairs per class (ordinary magnitudes, the FCT ranges 1e-15 .. 1e40, zeros, subnormals, huge, Inf/NaN) are divided on device
 * `device` by div_rn and by the compiler's x / y; *nbad = number of pairs whose bit patterns differ (NaN payloads aside).
 * Every REAL(wp) division of traadv_fct.F90 (:164, :166, :394-396, :294) goes through it.                              */
int nemo_fct_selftest_division(int device, long long n, unsigned long long seed, long long *nbad);
/* communication report in the spirit of mpp_report (lib_mpp.F90:1471-1587): exchanges and bytes sent so far       */
int nemo_fct_comm_report(nemo_fct_handle h, long long *n_exchanges, long long *bytes_sent);
/* Per-kernel device timing with CUDA events recorded on the launching stream around every launch of this context
 * (bench.py's roofline figures).  nemo_fct_set_profiling(h, 1) clears the counters and starts recording;
 * nemo_fct_profile_read synchronises and returns the number of kernels seen, their names (name_stride bytes
 * each, NUL-terminated), accumulated milliseconds and launch counts.  Returns -1 on error.                         */
int nemo_fct_set_profiling(nemo_fct_handle h, int on);
int nemo_fct_profile_read(nemo_fct_handle h, int max_entries, char *names, int name_stride, double *total_ms,
                          long long *calls);
/* Select the kernel schedule.  Schedules 0 - 3 return the same bits as the CPU restatement of the reference (checked by sha256 for 0
 * and 2; 1 and 3 share their arithmetic); schedule 4 returns the
 * same VALUE at every cell and the same bits at every ocean cell -- on land cells, where the trend is +-0, it may return the zero
 * with the other sign (its sign-bit forms of MAX(0,f) / MIN(0,f) / SIGN take -0.0 for negative; the reference built with
 * key_nosignedzero takes it for positive): tests/test_cpu_reference_exec.py.
 *   0 = reference pass structure: one kernel per pass group on the whole interior, exchanges X1..X4 as in
 *       traadv_fct.F90:209,280,400,426;
 *   1 = fused inner region (P1-P5 with in-place Laplacian; nonosc + final trend in one shared-memory kernel) on the main
 *       stream, boundary frame + X1..X4 on a side stream; per-thread cp.async prefetch;
 *   2 = as 1 with TMA-staged tiles for the inner kernels when jpi is even, else falls back to 1 (round-1 default).
 *   3 = as 2, and the fused nonosc kernel also fed by a 2-stage TMA ring (measured slower than 2 on B200: kept for study).
 *   4 = (default) the whole step of the inner region in ONE kernel (k_fct_fused, csrc/fct_fused_kernel.cuh) + the tiled
 *       interp_4th_cpt; falls back to 2 where its tiles do not fit (odd jpi, unaligned arrays, no room to split).
 * Subdomains smaller than 20 x 20 always use 0.                                                                    */
int nemo_fct_set_schedule(nemo_fct_handle h, int schedule);
/* Arithmetic of the one-kernel schedule (4).  STRICT (default): every REAL(wp) operation of traadv_fct.F90 as an IEEE operation
 * in the reference's order, no FMA contraction: the values of the CPU restatement (see nemo_fct_set_schedule for the sign of zero
 * on land).  FAST: the six divisions per point (:164,
 * :166, :393-396, :294) become a multiplication by a refined reciprocal (<= 1 ulp instead of <= 0.5 ulp); measured -12 % .. -32 %
 * kernel time, max relative difference 5e-11 on near-zero trends (above the 1e-12 bar: opt-in only).                      */
#define NEMO_FCT_ARITH_STRICT 0
#define NEMO_FCT_ARITH_FAST   1
int nemo_fct_set_arithmetic(nemo_fct_handle h, int mode);

#ifdef __cplusplus
}
#endif
#endif
