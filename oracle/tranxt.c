/*
 * tranxt.c -- ORACLE (test infrastructure only; see nemo_oracle.h).
 * Loop-for-loop C restatement of src/OCE/TRA/tranxt.F90 (tra_nxt :65-187, tra_nxt_fix :190-234,
 * tra_nxt_vvl :237-380) and of the swap part of src/TOP/TRP/trcnxt.F90:56-183 (trc_nxt).
 * PARITY PIN: bit-identical to the reference's own source executed by translation (oracle/f90exec.py); see nemo_oracle.h.
 *
 * Not restated (optional hooks, off by default in the reference): AGRIF, ln_bdy, l_trdtra / l_trdtrc trends
 * (:121-146, :169-178, :278-281, :359-378), prt_ctl, trc_nxt_off (l_offline).
 */
#include "nemo_oracle.h"
#include <stdlib.h>
#include <string.h>

#define I3(ji, jj, jk) ((size_t)((jk) - 1) * jpij + (size_t)((jj) - 1) * jpi + (size_t)((ji) - 1))
#define I2(ji, jj)     ((size_t)((jj) - 1) * jpi + (size_t)((ji) - 1))

/* value of an optional 2-D module array: NULL stands for an array of zeros (forcing switched off) */
static inline double v2(const double *p, size_t i) { return p ? p[i] : 0.0; }

void tra_nxt_fix(oce_dom *d, int kt, int kit000, const char *cdtype, double atfp,
                 double *ptb_all, double *ptn_all, double *pta_all, int kjpt)
{
    (void)kt; (void)kit000; (void)cdtype;                                       /* banner only (:211-215) */
    const int jpi = d->jpi, jpim1 = d->jpim1, jpjm1 = d->jpjm1, jpkm1 = d->jpkm1;
    const size_t jpij = (size_t)jpi * d->jpj, n3 = jpij * d->jpk;
    int ji, jj, jk, jn;
    double ztn, ztd;
    for (jn = 1; jn <= kjpt; ++jn) {                                            /* :217 */
        double *ptb = ptb_all + (size_t)(jn - 1) * n3, *ptn = ptn_all + (size_t)(jn - 1) * n3;
        double *pta = pta_all + (size_t)(jn - 1) * n3;
        for (jk = 1; jk <= jpkm1; ++jk)
            for (jj = 2; jj <= jpjm1; ++jj)
                for (ji = 2; ji <= jpim1; ++ji) {
                    ztn = ptn[I3(ji, jj, jk)];
                    ztd = pta[I3(ji, jj, jk)] - 2.0 * ztn + ptb[I3(ji, jj, jk)];    /* time laplacian on tracers */
                    ptb[I3(ji, jj, jk)] = ztn + atfp * ztd;                         /* ptb <-- filtered ptn */
                    ptn[I3(ji, jj, jk)] = pta[I3(ji, jj, jk)];                      /* ptn <-- pta */
                }
    }
}

void tra_nxt_vvl(oce_dom *d, int kt, int kit000, double p2dt, const char *cdtype, const oce_nxt_forcing *f,
                 double *ptb_all, double *ptn_all, double *pta_all,
                 const double *psbc_tc, const double *psbc_tc_b, int kjpt)
{
    (void)kt; (void)kit000;
    const int jpi = d->jpi, jpim1 = d->jpim1, jpjm1 = d->jpjm1, jpkm1 = d->jpkm1;
    const size_t jpij = (size_t)jpi * d->jpj, n3 = jpij * d->jpk;
    const double *e3t_b = d->e3t_b, *e3t_n = d->e3t_n, *e3t_a = d->e3t_a;
    const int *mikt = d->mikt;
    const int jp_tem = 1;
    int ll_traqsr, ll_rnf, ll_isf;
    int ji, jj, jk, jn;
    double zfact1, ztc_a, ztc_n, ztc_b, ztc_f, ztc_d;
    double zfact2, ze3t_b, ze3t_n, ze3t_a, ze3t_f, ze3t_d;

    if (strncmp(cdtype, "TRA", 3) == 0) {                                       /* :269-277 */
        ll_traqsr = f->ln_traqsr; ll_rnf = f->ln_rnf; ll_isf = f->ln_isf;
    } else {
        ll_traqsr = 0; ll_rnf = 0; ll_isf = 0;
    }
    zfact1 = f->atfp * p2dt;                                                    /* :283-284 */
    zfact2 = zfact1 * f->r1_rau0;
    for (jn = 1; jn <= kjpt; ++jn) {                                            /* :285 */
        double *ptb = ptb_all + (size_t)(jn - 1) * n3, *ptn = ptn_all + (size_t)(jn - 1) * n3;
        double *pta = pta_all + (size_t)(jn - 1) * n3;
        const double *sbc = psbc_tc ? psbc_tc + (size_t)(jn - 1) * jpij : NULL;
        const double *sbc_b = psbc_tc_b ? psbc_tc_b + (size_t)(jn - 1) * jpij : NULL;
        const double *rnf_tsc = f->rnf_tsc ? f->rnf_tsc + (size_t)(jn - 1) * jpij : NULL;
        const double *rnf_tsc_b = f->rnf_tsc_b ? f->rnf_tsc_b + (size_t)(jn - 1) * jpij : NULL;
        const double *risf_tsc = f->risf_tsc ? f->risf_tsc + (size_t)(jn - 1) * jpij : NULL;
        const double *risf_tsc_b = f->risf_tsc_b ? f->risf_tsc_b + (size_t)(jn - 1) * jpij : NULL;
        for (jk = 1; jk <= jpkm1; ++jk)
            for (jj = 2; jj <= jpjm1; ++jj)
                for (ji = 2; ji <= jpim1; ++ji) {
                    const size_t o = I3(ji, jj, jk), o2 = I2(ji, jj);
                    ze3t_b = e3t_b[o];
                    ze3t_n = e3t_n[o];
                    ze3t_a = e3t_a[o];
                    /* tracer content at Before, now and after */
                    ztc_b = ptb[o] * ze3t_b;
                    ztc_n = ptn[o] * ze3t_n;
                    ztc_a = pta[o] * ze3t_a;

                    ze3t_d = ze3t_a - 2. * ze3t_n + ze3t_b;
                    ztc_d  = ztc_a  - 2. * ztc_n  + ztc_b;

                    ze3t_f = ze3t_n + f->atfp * ze3t_d;
                    ztc_f  = ztc_n  + f->atfp * ztc_d;

                    if (jk == mikt[o2]) {                                       /* first level (:304-308) */
                        ze3t_f = ze3t_f - zfact2 * ((v2(f->emp_b, o2) - v2(f->emp, o2))
                                                    + (v2(f->fwfisf_b, o2) - v2(f->fwfisf, o2)));
                        ztc_f  = ztc_f  - zfact1 * (v2(sbc, o2) - v2(sbc_b, o2));
                    }
                    if (f->ln_rnf_depth) {                                      /* :309-319 */
                        if (mikt[o2] <= jk && jk <= f->nk_rnf[o2])
                            ze3t_f = ze3t_f - zfact2 * (-(v2(f->rnf_b, o2) - v2(f->rnf, o2)))
                                                     * (e3t_n[o] / f->h_rnf[o2]);
                    } else {
                        if (jk == mikt[o2])
                            ze3t_f = ze3t_f - zfact2 * (-(v2(f->rnf_b, o2) - v2(f->rnf, o2)));
                    }
                    /* solar penetration (temperature only)  (:323-325) */
                    if (ll_traqsr && jn == jp_tem && jk <= f->nksr)
                        ztc_f = ztc_f - zfact1 * (f->qsr_hc[o] - f->qsr_hc_b[o]);
                    /* river runoff  (:327-330) */
                    if (ll_rnf && jk <= f->nk_rnf[o2])
                        ztc_f = ztc_f - zfact1 * (rnf_tsc[o2] - rnf_tsc_b[o2]) * e3t_n[o] / f->h_rnf[o2];
                    /* ice shelf  (:332-343) */
                    if (ll_isf) {
                        if (jk >= f->misfkt[o2] && jk < f->misfkb[o2])
                            ztc_f = ztc_f - zfact1 * (risf_tsc[o2] - risf_tsc_b[o2]) * e3t_n[o] * f->r1_hisf_tbl[o2];
                        if (jk == f->misfkb[o2])
                            ztc_f = ztc_f - zfact1 * (risf_tsc[o2] - risf_tsc_b[o2]) * e3t_n[o] * f->r1_hisf_tbl[o2]
                                                   * f->ralpha[o2];
                    }
                    ze3t_f = 1.e0 / ze3t_f;                                     /* :345-347 */
                    ptb[o] = ztc_f * ze3t_f;                                    /* ptb <-- ptn filtered */
                    ptn[o] = pta[o];                                            /* ptn <-- pta */
                }
    }
}

/* tra_nxt (cdtype "TRA", tranxt.F90:65-187) / trc_nxt (cdtype "TRC", trcnxt.F90:56-183).
 * l_euler = (neuler == 0 .AND. kt == nit000) [.OR. ln_top_euler for TRC]; rdt = rdt / rdttrc.                    */
void tra_nxt(oce_dom *d, int kt, int kit000, int l_euler, double rdt, const char *cdtype, const oce_nxt_forcing *f,
             double *ptb, double *ptn, double *pta, const double *psbc_tc, const double *psbc_tc_b, int kjpt)
{
    const int jpi = d->jpi, jpj = d->jpj, jpk = d->jpk, jpkm1 = d->jpkm1;
    const size_t jpij = (size_t)jpi * jpj, n3 = jpij * jpk;
    const int is_trc = strncmp(cdtype, "TRC", 3) == 0;
    double **tab = (double **)malloc(sizeof(double *) * 3 * (size_t)kjpt);
    char *nat = (char *)malloc(3 * (size_t)kjpt + 1);
    double *sgn = (double *)malloc(sizeof(double) * 3 * (size_t)kjpt);
    int jn, jk;
    for (jn = 0; jn < 3 * kjpt; ++jn) { nat[jn] = 'T'; sgn[jn] = 1.0; }
    nat[3 * kjpt] = 0;

    /* Update after tracer on domain lateral boundaries  (tranxt.F90:108, trcnxt.F90:100) */
    for (jn = 0; jn < kjpt; ++jn) tab[jn] = pta + (size_t)jn * n3;
    lbc_lnk_multi(d, "tranxt", kjpt, tab, nat, sgn, jpk, 0, 0.0);

    if (l_euler) {                                  /* Euler time-stepping at first time-step (only swap) */
        for (jn = 0; jn < kjpt; ++jn)
            for (jk = 1; jk <= jpkm1; ++jk) {       /* tranxt.F90:148-153; trcnxt.F90:147-153 also resets trb */
                memcpy(ptn + (size_t)jn * n3 + (size_t)(jk - 1) * jpij, pta + (size_t)jn * n3 + (size_t)(jk - 1) * jpij,
                       jpij * sizeof(double));
                if (is_trc)
                    memcpy(ptb + (size_t)jn * n3 + (size_t)(jk - 1) * jpij, ptn + (size_t)jn * n3 + (size_t)(jk - 1) * jpij,
                           jpij * sizeof(double));
            }
    } else {                                        /* Leap-Frog + Asselin filter time stepping (:162-171) */
        if (d->ln_linssh) tra_nxt_fix(d, kt, kit000, cdtype, f->atfp, ptb, ptn, pta, kjpt);
        else              tra_nxt_vvl(d, kt, kit000, rdt, cdtype, f, ptb, ptn, pta, psbc_tc, psbc_tc_b, kjpt);
        /* tranxt.F90:168-170 orders the fields (tsb T,S, tsn T,S, tsa T,S); trcnxt.F90:171 (trb, trn, tra) */
        for (jn = 0; jn < kjpt; ++jn) {
            tab[jn] = ptb + (size_t)jn * n3; tab[kjpt + jn] = ptn + (size_t)jn * n3; tab[2 * kjpt + jn] = pta + (size_t)jn * n3;
        }
        lbc_lnk_multi(d, "tranxt", 3 * kjpt, tab, nat, sgn, jpk, 0, 0.0);
    }
    free(tab); free(nat); free(sgn);
}
