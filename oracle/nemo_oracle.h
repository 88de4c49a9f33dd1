/*
 * nemo_oracle.h -- CPU ORACLE (TEST INFRASTRUCTURE ONLY, never shipped, never on the product path).
 *
 * Plain-C, loop-for-loop restatement of the reference NEMO 4.0 routines on the FCT tracer-advection path:
 *   tra_adv_fct / nonosc / interp_4th_cpt   src/OCE/TRA/traadv_fct.F90:54-327, 330-428, 517-616
 *   tra_adv (transports)                    src/OCE/TRA/traadv.F90:95-124
 *   lbc_lnk (no-MPI)                        src/OCE/LBC/lbc_lnk_generic.h90:48-108
 *   mpp_lnk (MPI halo exchange)             src/OCE/LBC/mpp_lnk_generic.h90:48-338
 *   lbc_nfd (north fold)                    src/OCE/LBC/lbc_nfd_generic.h90:46-167
 *   mpp_nfd (multi-rank north fold)         src/OCE/LBC/mpp_nfd_generic.h90:48-302 (+ lbc_nfd_nogather_generic.h90)
 *   mpp_init / mpp_basic_decomposition      src/OCE/LBC/mppini.F90:110-692, 695-798, 1180-1240
 *   dom_msk (masks)                         src/OCE/DOM/dommsk.F90:135-238
 *   tra_adv_cen (centred 2nd / 4th order)   src/OCE/TRA/traadv_cen.F90:46-204
 *   tra_adv_mus (MUSCL)                     src/OCE/TRA/traadv_mus.F90:55-273
 *   tra_nxt / tra_nxt_fix / tra_nxt_vvl     src/OCE/TRA/tranxt.F90:65-380 (+ trc_nxt swap, src/TOP/TRP/trcnxt.F90:56-183)
 *   SIGN / DDPDD / glob_sum                 src/OCE/lib_fortran.F90:300-351, lib_fortran_generic.h90:32-65
 *
 * PARITY PIN: the reference tree holds no golden vectors / known-answer tests for this path and the reference
 * (Fortran) cannot be COMPILED in this image or on the GPU box (no Fortran compiler; oracle/_ref_recipe/ is the
 * unexecuted recipe).  Instead the reference's own source text is EXECUTED here by translation (oracle/f90exec.py,
 * oracle/ref_exec.py: IEEE binary64, key_nosignedzero as arch/arch-linux_gfortran.fcm:46 builds): tra_adv (driver,
 * transports), trc_adv, tra_adv_fct + nonosc + interp_4th_cpt, tra_adv_mus, tra_adv_cen, tra_nxt (+ _fix / _vvl), the mono-
 * processor lbc_lnk + lbc_nfd, the multi-rank mpp_lnk + mpp_nfd (gather and no-gather fold, on emulated MPI ranks),
 * mpp_init (+ mpp_basic_decomposition, mpp_init_nfdcom; all-ocean layouts), dom_msk, stp_ctl, glob_sum + DDPDD and SIGN agree with this restatement BIT FOR BIT
 * (tests/test_cpu_reference_exec.py), and
 * the committed golden vectors carry the hashes of those executions (tests/golden/ref_exec_pins.json).  Not pinned
 * that way: compiled-binary effects, land-subdomain elimination in mpp_init (out of scope).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may use this.
 *
 * Conventions: arrays are Fortran column-major (ji fastest), indices in the code are 1-based through the
 * IDX macros, fp64 everywhere (wp = dp, src/OCE/par_kind.F90:24-26). Build with -ffp-contract=off (no FMA),
 * matching the IEEE-strict reference builds (arch/arch-linux_gfortran.fcm).
 */
#ifndef NEMO_ORACLE_H
#define NEMO_ORACLE_H
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define JPMAXNGH 3 /* lbcnfd.F90:53 */

struct oce_world;

/* One (sub)domain = the module variables of par_oce / dom_oce that the path reads. */
typedef struct oce_dom {
    /* par_oce.F90 */
    int jpi, jpj, jpk, jpim1, jpjm1, jpkm1;
    int jpiglo, jpjglo, jpni, jpnj, jpnij, jpimax, jpjmax;
    int jperio;
    /* dom_oce.F90 decomposition scalars (mppini.F90:548-580) */
    int narea, nproc;                 /* narea = nproc+1 */
    int nimpp, njmpp, nlci, nlcj, nldi, nlei, nldj, nlej;
    int nbondi, nbondj, noea, nowe, noso, nono, npolj;
    int l_Iperio, l_Jperio;
    int nreci, nrecj;
    /* north-fold comm tables (mppini.F90:1180-1240, lbcnfd.F90:53-55) */
    int nsndto, isendto[JPMAXNGH], nfsloop, nfeloop;
    int ln_nnogather;                 /* nammpp, cfgs/SHARED/namelist_ref:1299 */
    /* flags */
    int ln_linssh, ln_isfcav;
    /* masks / metrics (owned by caller, may be NULL until set) */
    const double *tmask, *umask, *vmask, *wmask;       /* (jpi,jpj,jpk) */
    const double *e3t_b, *e3t_n, *e3t_a;               /* (jpi,jpj,jpk) */
    const double *e1e2t, *r1_e1e2t;                    /* (jpi,jpj)     */
    const int *mikt, *mbkt;                            /* (jpi,jpj)     */
    /* "MPI world" this subdomain lives in; NULL => key_mpp_mpi undefined (pure lbc_lnk_generic) */
    struct oce_world *world;
    /* optional capture of intermediates (debugging the device path); any may be NULL */
    int dbg_jn;                        /* tracer (1-based) to capture, 0 = none */
    double *dbg_zwi, *dbg_zwx, *dbg_zwy, *dbg_zwz;     /* after X2 (traadv_fct.F90:280) */
    double *dbg_zbetup, *dbg_zbetdo;                   /* after X3 (:400) */
    double *dbg_paa, *dbg_pbb, *dbg_pcc;               /* after X4 (:426) */
    double *dbg_ztw;                                   /* interp_4th_cpt output */
    double *dbg_zltu, *dbg_zltv;                       /* after X1 (:209) */
    /* extra module arrays read by tra_adv_mus (dom_oce.F90:118, 132-136); NULL until set */
    const double *r1_e1e2u, *r1_e1e2v;                 /* (jpi,jpj)     */
    const double *e3u_n, *e3v_n, *e3w_n;               /* (jpi,jpj,jpk) */
    /* l_trd / l_hst / l_ptr hooks of tra_adv_fct (traadv_fct.F90:104-112, 172-176, 299-316): total advective fluxes
     * ztrdx, ztrdy (= zptry), ztrdz per tracer, (jpi,jpj,jpk,kjpt); NULL = hooks off */
    double *diag_trdx, *diag_trdy, *diag_trdz;
} oce_dom;

/* Module variables read by tra_nxt_vvl (tranxt.F90:262-343): sbc_oce, sbcrnf, sbcisf, traqsr, phycst.
 * A NULL 2-D flux array stands for zeros (that forcing is switched off).                                        */
typedef struct oce_nxt_forcing {
    double atfp, r1_rau0;                              /* dom_oce.F90:58, phycst.F90:42 */
    int ln_traqsr, ln_rnf, ln_isf, ln_rnf_depth, nksr; /* traqsr.F90:54, sbcrnf.F90, sbcisf.F90 */
    const double *emp_b, *emp, *fwfisf_b, *fwfisf, *rnf_b, *rnf;        /* (jpi,jpj) */
    const double *qsr_hc, *qsr_hc_b;                                    /* (jpi,jpj,jpk) */
    const int *nk_rnf; const double *h_rnf, *rnf_tsc, *rnf_tsc_b;       /* (jpi,jpj), (jpi,jpj,jpts) */
    const int *misfkt, *misfkb;                                         /* (jpi,jpj) */
    const double *risf_tsc, *risf_tsc_b, *r1_hisf_tbl, *ralpha;         /* (jpi,jpj,jpts), (jpi,jpj) */
} oce_nxt_forcing;

/* The set of subdomains of one run ("mpi_comm_oce"), with global tables (mppini.F90). */
typedef struct oce_world {
    int jpnij, jpni, jpnj, jpiglo, jpjglo, jperio, jpimax, jpjmax;
    oce_dom *dom;                      /* [jpnij] */
    int *nimppt, *njmppt, *nlcit, *nlcjt, *nldit, *nleit, *nldjt, *nlejt, *ibonit, *ibonjt; /* [jpnij] */
    int *nfiimpp, *nfilcit, *nfipproc; /* (jpni,jpnj) column-major */
    int ndim_rank_north, *nrank_north, njmppmax;
    void *mail;                        /* mailbox state of the MPI emulation (lbclnk.c) */
    int invalid;                       /* layout the reference cannot run (more than jpmaxngh fold partners) */
} oce_world;

/* ---- mppini.c ---- */
void mpp_basic_decomposition(int jpiglo, int jpjglo, int jperio, int knbi, int knbj, int *kimax, int *kjmax,
                             int *kimppt, int *kjmppt, int *klci, int *klcj);   /* mppini.F90:695-798 */
oce_world *mpp_init(int jpiglo, int jpjglo, int jpk, int jperio, int jpni, int jpnj, int ln_nnogather,
                    int key_mpp_mpi);                                           /* mppini.F90:110-692 */
void mpp_finalize(oce_world *w);
int  oce_world_size(const oce_world *w);
oce_dom *oce_world_dom(oce_world *w, int rank0);

/* ---- lbclnk.c ---- */
/* lbc_lnk_multi: nfld fields, each (jpi,jpj,ipk) fp64; cd_nat[f] in "TUVWF"; psgn[f]; has_pval/pval optional land
 * value.  Dispatches like lbclnk.F90:32-37 / 84-86: mpp_lnk when d->world has MPI semantics, else lbc_lnk.       */
void lbc_lnk_multi(oce_dom *d, const char *cdname, int nfld, double **ptab, const char *cd_nat,
                   const double *psgn, int ipk, int has_pval, double pval);
void lbc_lnk_generic(oce_dom *d, int nfld, double **ptab, const char *cd_nat, const double *psgn, int ipk,
                     int has_pval, double pval);                                /* lbc_lnk_generic.h90 */
void mpp_lnk_generic(oce_dom *d, int nfld, double **ptab, const char *cd_nat, const double *psgn, int ipk,
                     int has_pval, double pval);                                /* mpp_lnk_generic.h90 */
void lbc_nfd_generic(const oce_dom *d, int ipi, int ipj_arr, int ipj, int nfld, double **ptab,
                     const char *cd_nat, const double *psgn, int ipk);          /* lbc_nfd_generic.h90 */
/* run fn(dom, arg) once per subdomain, one thread per subdomain ("mpirun -np jpnij") */
void oce_world_run(oce_world *w, void (*fn)(oce_dom *, void *), void *arg);
/* convenience: collective lbc_lnk_multi over all subdomains; ptabs[rank][f] */
void oce_world_lbc_lnk(oce_world *w, int nfld, double ***ptabs, const char *cd_nat, const double *psgn, int ipk);

/* ---- traadv_fct.c ---- */
void tra_adv_fct(oce_dom *d, int kt, int kit000, const char *cdtype, double p2dt,
                 const double *pun, const double *pvn, const double *pwn,
                 const double *ptb, const double *ptn, double *pta, int kjpt, int kn_fct_h, int kn_fct_v);
void nonosc(oce_dom *d, const double *pbef, double *paa, double *pbb, double *pcc, const double *paft,
            double p2dt, int jn);
void interp_4th_cpt(const oce_dom *d, const double *pt_in, double *pt_out);
void oracle_poison_workspace(int on);  /* fill automatic arrays with NaN before use (catches undefined reads) */
int  oracle_poison_enabled(void);

/* ---- traadv_cen.c ---- */
void tra_adv_cen(oce_dom *d, int kt, int kit000, const char *cdtype, const double *pun, const double *pvn,
                 const double *pwn, const double *ptn, double *pta, int kjpt, int kn_cen_h, int kn_cen_v);  /* traadv_cen.F90:46-204 */

/* ---- traadv_mus.c ---- */
void tra_adv_mus_xind(const oce_dom *d, int ld_msc_ups, const double *rnfmsk, const double *rnfmsk_z, double *xind);
void tra_adv_mus(oce_dom *d, int kt, int kit000, const char *cdtype, double p2dt,
                 const double *pun, const double *pvn, const double *pwn,
                 const double *ptb, double *pta, int kjpt, const double *xind);  /* traadv_mus.F90:55-273 */

/* ---- tranxt.c ---- */
void tra_nxt_fix(oce_dom *d, int kt, int kit000, const char *cdtype, double atfp,
                 double *ptb, double *ptn, double *pta, int kjpt);              /* tranxt.F90:190-234 */
void tra_nxt_vvl(oce_dom *d, int kt, int kit000, double p2dt, const char *cdtype, const oce_nxt_forcing *f,
                 double *ptb, double *ptn, double *pta, const double *psbc_tc, const double *psbc_tc_b,
                 int kjpt);                                                     /* tranxt.F90:237-380 */
void tra_nxt(oce_dom *d, int kt, int kit000, int l_euler, double rdt, const char *cdtype, const oce_nxt_forcing *f,
             double *ptb, double *ptn, double *pta, const double *psbc_tc, const double *psbc_tc_b, int kjpt);

/* ---- traadv.c ---- */
void tra_adv_transports(const oce_dom *d, const double *e2u, const double *e1v, const double *e3u_n,
                        const double *e3v_n, const double *un, const double *vn, const double *wn,
                        double *zun, double *zvn, double *zwn);                 /* traadv.F90:100-124 */

/* ---- dommsk.c ---- */
void dom_msk(oce_dom *d, const int *k_top, const int *k_bot, double *tmask, double *umask, double *vmask,
             double *wmask, double *tmask_i, int *mikt, int *mbkt);             /* dommsk.F90:135-238 */

/* ---- lib_fortran.c ---- */
double sign_nosignedzero(double pa, double pb);                                 /* lib_fortran.F90:339-351 */
void   ddpdd(const double ydda[2], double yddb[2]);                             /* lib_fortran.F90:300-332 */
/* local part of glob_sum (lib_fortran_generic.h90:52-63): sum over ptab*tmask_i in double-double, accumulated
 * into ctmp (caller zero-inits); the cross-rank MPI_SUMDD (lib_mpp.F90:1158-1186) is ddpdd over ranks.       */
void   glob_sum_local(const double *ptab, const double *tmask_i, int jpi, int jpj, int ipk, double ctmp[2]);

/* ---- stpctl.c ---- */
/* extrema test + error condition of stp_ctl (stpctl.F90:115-124, 149-166, 184), local domain, branch without ln_ctl */
void   stp_ctl_local(const oce_dom *d, const double *sshn, const double *un, const double *tsn, const double *tmask,
                     double zmax[6], int ih[2], int iu[3], int is1[3], int is2[3], int *nan_found, int *kindic);

#ifdef __cplusplus
}
#endif
#endif
