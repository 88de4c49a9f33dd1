/*
 * driver.c -- ORACLE (test infrastructure only; see nemo_oracle.h).
 * Thin helpers so that tests/ and bench.py's cpu_baseline leg can drive the restated routines through ctypes:
 * accessors for the decomposition scalars, setters for the module arrays, and "mpirun"-style collective calls
 * that run one thread per subdomain (oce_world_run).
 */
#include "nemo_oracle.h"
#include <string.h>

/* fixed-order dump of the decomposition scalars of one subdomain (order is mirrored in oracle/oracle.py) */
void oce_dom_info(const oce_dom *d, int *o)
{
    int n = 0;
    o[n++] = d->jpi;    o[n++] = d->jpj;    o[n++] = d->jpk;    o[n++] = d->jpiglo; o[n++] = d->jpjglo;
    o[n++] = d->jpni;   o[n++] = d->jpnj;   o[n++] = d->jpnij;  o[n++] = d->jpimax; o[n++] = d->jpjmax;
    o[n++] = d->jperio; o[n++] = d->narea;  o[n++] = d->nproc;  o[n++] = d->nimpp;  o[n++] = d->njmpp;
    o[n++] = d->nlci;   o[n++] = d->nlcj;   o[n++] = d->nldi;   o[n++] = d->nlei;   o[n++] = d->nldj;
    o[n++] = d->nlej;   o[n++] = d->nbondi; o[n++] = d->nbondj; o[n++] = d->noea;   o[n++] = d->nowe;
    o[n++] = d->noso;   o[n++] = d->nono;   o[n++] = d->npolj;  o[n++] = d->l_Iperio; o[n++] = d->l_Jperio;
    o[n++] = d->nsndto; o[n++] = d->isendto[0]; o[n++] = d->isendto[1]; o[n++] = d->isendto[2];
    o[n++] = d->nfsloop; o[n++] = d->nfeloop;
}

void oce_dom_set_fields(oce_dom *d, const double *tmask, const double *umask, const double *vmask,
                        const double *wmask, const double *e3t_b, const double *e3t_n, const double *e3t_a,
                        const double *e1e2t, const double *r1_e1e2t, const int *mikt, const int *mbkt,
                        int ln_linssh, int ln_isfcav)
{
    d->tmask = tmask; d->umask = umask; d->vmask = vmask; d->wmask = wmask;
    d->e3t_b = e3t_b; d->e3t_n = e3t_n; d->e3t_a = e3t_a; d->e1e2t = e1e2t; d->r1_e1e2t = r1_e1e2t;
    d->mikt = mikt; d->mbkt = mbkt; d->ln_linssh = ln_linssh; d->ln_isfcav = ln_isfcav;
}

/* capture pointers, order: zwi zwx zwy zwz zbetup zbetdo paa pbb pcc ztw zltu zltv (NULL = skip) */
void oce_dom_set_dbg(oce_dom *d, int jn, double **p)
{
    d->dbg_jn = jn;
    d->dbg_zwi = p[0]; d->dbg_zwx = p[1]; d->dbg_zwy = p[2]; d->dbg_zwz = p[3];
    d->dbg_zbetup = p[4]; d->dbg_zbetdo = p[5]; d->dbg_paa = p[6]; d->dbg_pbb = p[7]; d->dbg_pcc = p[8];
    d->dbg_ztw = p[9]; d->dbg_zltu = p[10]; d->dbg_zltv = p[11];
}

typedef struct {
    double p2dt; const double **pun, **pvn, **pwn, **ptb, **ptn; double **pta; int kjpt, kn_fct_h, kn_fct_v;
} fct_arg;
static void fct_thr(oce_dom *d, void *p)
{
    fct_arg *a = (fct_arg *)p; int r = d->nproc;
    tra_adv_fct(d, 1, 1, "TRA", a->p2dt, a->pun[r], a->pvn[r], a->pwn[r], a->ptb[r], a->ptn[r], a->pta[r],
                a->kjpt, a->kn_fct_h, a->kn_fct_v);
}
/* tra_adv_fct on every subdomain of the world, one thread per subdomain (the per-rank pointer tables are
 * indexed by nproc) */
void oce_world_tra_adv_fct(oce_world *w, double p2dt, const double **pun, const double **pvn, const double **pwn,
                           const double **ptb, const double **ptn, double **pta, int kjpt, int kn_fct_h, int kn_fct_v)
{
    fct_arg a = { p2dt, pun, pvn, pwn, ptb, ptn, pta, kjpt, kn_fct_h, kn_fct_v };
    oce_world_run(w, fct_thr, &a);
}

typedef struct { const int **k_top, **k_bot; double **tmask, **umask, **vmask, **wmask, **tmask_i; int **mikt, **mbkt; } msk_arg;
static void msk_thr(oce_dom *d, void *p)
{
    msk_arg *a = (msk_arg *)p; int r = d->nproc;
    dom_msk(d, a->k_top[r], a->k_bot[r], a->tmask[r], a->umask[r], a->vmask[r], a->wmask[r],
            a->tmask_i ? a->tmask_i[r] : 0, a->mikt[r], a->mbkt[r]);
}
void oce_world_dom_msk(oce_world *w, const int **k_top, const int **k_bot, double **tmask, double **umask,
                       double **vmask, double **wmask, double **tmask_i, int **mikt, int **mbkt)
{
    msk_arg a = { k_top, k_bot, tmask, umask, vmask, wmask, tmask_i, mikt, mbkt };
    oce_world_run(w, msk_thr, &a);
}

/* l_trd / l_hst / l_ptr hooks of tra_adv_fct: (jpi,jpj,jpk,kjpt) output arrays, or NULLs to switch them off */
void oce_dom_set_diag(oce_dom *d, double *trdx, double *trdy, double *trdz)
{
    d->diag_trdx = trdx; d->diag_trdy = trdy; d->diag_trdz = trdz;
}

typedef struct { const double **pun, **pvn, **pwn, **ptn; double **pta; int kjpt, h, v; } cen_arg;
static void cen_thr(oce_dom *d, void *p)
{
    cen_arg *a = (cen_arg *)p; int r = d->nproc;
    tra_adv_cen(d, 1, 1, "TRA", a->pun[r], a->pvn[r], a->pwn[r], a->ptn[r], a->pta[r], a->kjpt, a->h, a->v);
}
void oce_world_tra_adv_cen(oce_world *w, const double **pun, const double **pvn, const double **pwn,
                           const double **ptn, double **pta, int kjpt, int kn_cen_h, int kn_cen_v)
{
    cen_arg a = { pun, pvn, pwn, ptn, pta, kjpt, kn_cen_h, kn_cen_v };
    oce_world_run(w, cen_thr, &a);
}

void oce_dom_set_mus_fields(oce_dom *d, const double *r1_e1e2u, const double *r1_e1e2v, const double *e3u_n,
                            const double *e3v_n, const double *e3w_n)
{
    d->r1_e1e2u = r1_e1e2u; d->r1_e1e2v = r1_e1e2v; d->e3u_n = e3u_n; d->e3v_n = e3v_n; d->e3w_n = e3w_n;
}

typedef struct { double p2dt; const double **pun, **pvn, **pwn, **ptb; double **pta; int kjpt; const double **xind; } mus_arg;
static void mus_thr(oce_dom *d, void *p)
{
    mus_arg *a = (mus_arg *)p; int r = d->nproc;
    tra_adv_mus(d, 1, 1, "TRA", a->p2dt, a->pun[r], a->pvn[r], a->pwn[r], a->ptb[r], a->pta[r], a->kjpt, a->xind[r]);
}
void oce_world_tra_adv_mus(oce_world *w, double p2dt, const double **pun, const double **pvn, const double **pwn,
                           const double **ptb, double **pta, int kjpt, const double **xind)
{
    mus_arg a = { p2dt, pun, pvn, pwn, ptb, pta, kjpt, xind };
    oce_world_run(w, mus_thr, &a);
}

typedef struct {
    int kt, kit000, l_euler; double rdt; const char *cdtype; const oce_nxt_forcing **f;
    double **ptb, **ptn, **pta; const double **sbc, **sbc_b; int kjpt;
} nxt_arg;
static void nxt_thr(oce_dom *d, void *p)
{
    nxt_arg *a = (nxt_arg *)p; int r = d->nproc;
    tra_nxt(d, a->kt, a->kit000, a->l_euler, a->rdt, a->cdtype, a->f[r], a->ptb[r], a->ptn[r], a->pta[r],
            a->sbc ? a->sbc[r] : 0, a->sbc_b ? a->sbc_b[r] : 0, a->kjpt);
}
/* tra_nxt / trc_nxt on every subdomain (per-rank forcing structs and pointer tables indexed by nproc) */
void oce_world_tra_nxt(oce_world *w, int kt, int kit000, int l_euler, double rdt, const char *cdtype,
                       const oce_nxt_forcing **f, double **ptb, double **ptn, double **pta,
                       const double **sbc, const double **sbc_b, int kjpt)
{
    nxt_arg a = { kt, kit000, l_euler, rdt, cdtype, f, ptb, ptn, pta, sbc, sbc_b, kjpt };
    oce_world_run(w, nxt_thr, &a);
}

/* world-level tables for tests of mpp_init */
void oce_world_tables(const oce_world *w, int *nimppt, int *njmppt, int *nlcit, int *nlcjt)
{
    for (int r = 0; r < w->jpnij; ++r) {
        if (w->jpnij == 1) { nimppt[0] = w->dom[0].nimpp; njmppt[0] = w->dom[0].njmpp; nlcit[0] = w->dom[0].nlci; nlcjt[0] = w->dom[0].nlcj; }
        else { nimppt[r] = w->nimppt[r]; njmppt[r] = w->njmppt[r]; nlcit[r] = w->nlcit[r]; nlcjt[r] = w->nlcjt[r]; }
    }
}
