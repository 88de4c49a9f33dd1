/*
 * traadv_fct.c -- ORACLE (test infrastructure only; see nemo_oracle.h).
 * Loop-for-loop C restatement of src/OCE/TRA/traadv_fct.F90 (tra_adv_fct :54-327, nonosc :330-428,
 * interp_4th_cpt :517-616).  Same loop bounds, same operation order, same automatic work arrays living
 * across the tracer loop.  PARITY PIN: bit-identical to the reference's own source executed by translation (oracle/f90exec.py); see nemo_oracle.h.
 *
 * The l_trd / l_hst / l_ptr hooks (:96-112, :172-176, :299-316) are restated up to the point where the reference hands
 * ztrdx / ztrdy / ztrdz (zptry = ztrdy) to trd_tra, dia_ar5_hst and dia_ptr_hst: d->diag_trd* receive those arrays.
 * Not restated: kn_fct_h = 41 (:223-250, "coding attempt, need to be tested").
 */
#include "nemo_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>

static int g_poison = 0;
void oracle_poison_workspace(int on) { g_poison = on; }
int  oracle_poison_enabled(void) { return g_poison; }

/* automatic array (jpi,jpj,jpk): undefined on entry in Fortran; optionally NaN-poisoned here */
static double *auto3d(const oce_dom *d)
{
    size_t n = (size_t)d->jpi * d->jpj * d->jpk;
    double *p = (double *)malloc(n * sizeof(double));
    if (!p) { fprintf(stderr, "oracle: out of memory\n"); abort(); }
    if (g_poison) { for (size_t i = 0; i < n; ++i) p[i] = NAN; }
    else          { memset(p, 0, n * sizeof(double)); }
    return p;
}

/* 1-based Fortran indexing of a (jpi,jpj,jpk) array */
#define I3(ji, jj, jk) ((size_t)((jk) - 1) * jpij + (size_t)((jj) - 1) * jpi + (size_t)((ji) - 1))
#define I2(ji, jj)     ((size_t)((jj) - 1) * jpi + (size_t)((ji) - 1))

static inline double dmax(double a, double b) { return a > b ? a : b; }   /* Fortran MAX (no NaN on path) */
static inline double dmin(double a, double b) { return a < b ? a : b; }   /* Fortran MIN */

static void dbg_copy(double *dst, const double *src, size_t n) { if (dst) memcpy(dst, src, n * sizeof(double)); }

void tra_adv_fct(oce_dom *d, int kt, int kit000, const char *cdtype, double p2dt,
                 const double *pun, const double *pvn, const double *pwn,
                 const double *ptb_all, const double *ptn_all, double *pta_all, int kjpt, int kn_fct_h, int kn_fct_v)
{
    (void)kt; (void)kit000; (void)cdtype;   /* banner + diagnostics switches only (:90-112) */
    const int jpi = d->jpi, jpj = d->jpj, jpk = d->jpk;
    const int jpim1 = d->jpim1, jpjm1 = d->jpjm1, jpkm1 = d->jpkm1;
    const size_t jpij = (size_t)jpi * jpj, n3 = jpij * jpk;
    const double *tmask = d->tmask, *umask = d->umask, *vmask = d->vmask, *wmask = d->wmask;
    const double *e3t_b = d->e3t_b, *e3t_n = d->e3t_n, *e3t_a = d->e3t_a, *r1_e1e2t = d->r1_e1e2t;
    const double r1_6 = 1.0 / 6.0;                                              /* :39 */
    int ji, jj, jk, jn;
    double ztra, zfp_ui, zfp_vj, zfp_wk, zC2t_u, zfm_ui, zfm_vj, zfm_wk, zC2t_v;

    /* REAL(wp), DIMENSION(jpi,jpj,jpk) :: zwi, zwx, zwy, zwz, ztu, ztv, zltu, zltv, ztw   (:86) */
    double *zwi = auto3d(d), *zwx = auto3d(d), *zwy = auto3d(d), *zwz = auto3d(d);
    double *ztu = NULL, *ztv = NULL, *zltu = NULL, *zltv = NULL, *ztw = NULL;
    if (kn_fct_h == 4) { ztu = auto3d(d); ztv = auto3d(d); zltu = auto3d(d); zltv = auto3d(d); }
    if (kn_fct_v == 4) { ztw = auto3d(d); }

    /* surface & bottom value : flux set to zero one for all   (:113-117) */
    for (jj = 1; jj <= jpj; ++jj) for (ji = 1; ji <= jpi; ++ji) zwz[I3(ji, jj, 1)] = 0.0;
    for (jj = 1; jj <= jpj; ++jj) for (ji = 1; ji <= jpi; ++ji) {
        zwx[I3(ji, jj, jpk)] = 0.0; zwy[I3(ji, jj, jpk)] = 0.0; zwz[I3(ji, jj, jpk)] = 0.0;
    }
    for (size_t n = 0; n < n3; ++n) zwi[n] = 0.0;

    for (jn = 1; jn <= kjpt; ++jn) {                                            /* :119 */
        const double *ptb = ptb_all + (size_t)(jn - 1) * n3;
        const double *ptn = ptn_all + (size_t)(jn - 1) * n3;
        double *pta = pta_all + (size_t)(jn - 1) * n3;

        /* upstream tracer flux in the i and j direction  (:123-135) */
        for (jk = 1; jk <= jpkm1; ++jk)
            for (jj = 1; jj <= jpjm1; ++jj)
                for (ji = 1; ji <= jpim1; ++ji) {
                    zfp_ui = pun[I3(ji, jj, jk)] + fabs(pun[I3(ji, jj, jk)]);
                    zfm_ui = pun[I3(ji, jj, jk)] - fabs(pun[I3(ji, jj, jk)]);
                    zfp_vj = pvn[I3(ji, jj, jk)] + fabs(pvn[I3(ji, jj, jk)]);
                    zfm_vj = pvn[I3(ji, jj, jk)] - fabs(pvn[I3(ji, jj, jk)]);
                    zwx[I3(ji, jj, jk)] = 0.5 * (zfp_ui * ptb[I3(ji, jj, jk)] + zfm_ui * ptb[I3(ji + 1, jj, jk)]);
                    zwy[I3(ji, jj, jk)] = 0.5 * (zfp_vj * ptb[I3(ji, jj, jk)] + zfm_vj * ptb[I3(ji, jj + 1, jk)]);
                }
        /* upstream tracer flux in the k direction  (:137-145) */
        for (jk = 2; jk <= jpkm1; ++jk)
            for (jj = 1; jj <= jpj; ++jj)
                for (ji = 1; ji <= jpi; ++ji) {
                    zfp_wk = pwn[I3(ji, jj, jk)] + fabs(pwn[I3(ji, jj, jk)]);
                    zfm_wk = pwn[I3(ji, jj, jk)] - fabs(pwn[I3(ji, jj, jk)]);
                    zwz[I3(ji, jj, jk)] = 0.5 * (zfp_wk * ptb[I3(ji, jj, jk)] + zfm_wk * ptb[I3(ji, jj, jk - 1)])
                                          * wmask[I3(ji, jj, jk)];
                }
        if (d->ln_linssh) {                                                     /* :146-156 */
            if (d->ln_isfcav) {
                for (jj = 1; jj <= jpj; ++jj)
                    for (ji = 1; ji <= jpi; ++ji) {
                        int ik = d->mikt[I2(ji, jj)];
                        zwz[I3(ji, jj, ik)] = pwn[I3(ji, jj, ik)] * ptb[I3(ji, jj, ik)];
                    }
            } else {
                for (jj = 1; jj <= jpj; ++jj)
                    for (ji = 1; ji <= jpi; ++ji)
                        zwz[I3(ji, jj, 1)] = pwn[I3(ji, jj, 1)] * ptb[I3(ji, jj, 1)];
            }
        }
        /* trend and after field with monotonic scheme  (:158-170) */
        for (jk = 1; jk <= jpkm1; ++jk)
            for (jj = 2; jj <= jpjm1; ++jj)
                for (ji = 2; ji <= jpim1; ++ji) {
                    ztra = -(zwx[I3(ji, jj, jk)] - zwx[I3(ji - 1, jj, jk)]
                             + zwy[I3(ji, jj, jk)] - zwy[I3(ji, jj - 1, jk)]
                             + zwz[I3(ji, jj, jk)] - zwz[I3(ji, jj, jk + 1)]) * r1_e1e2t[I2(ji, jj)];
                    pta[I3(ji, jj, jk)] = pta[I3(ji, jj, jk)] + ztra / e3t_n[I3(ji, jj, jk)] * tmask[I3(ji, jj, jk)];
                    zwi[I3(ji, jj, jk)] = (e3t_b[I3(ji, jj, jk)] * ptb[I3(ji, jj, jk)] + p2dt * ztra)
                                          / e3t_a[I3(ji, jj, jk)] * tmask[I3(ji, jj, jk)];
                }

        /* trend diagnostics / poleward transports: contribution of the upstream fluxes  (:172-176, l_trd .OR. l_hst, l_ptr) */
        if (d->diag_trdx) {
            double *tx = d->diag_trdx + (size_t)(jn - 1) * n3, *ty = d->diag_trdy + (size_t)(jn - 1) * n3;
            double *tz = d->diag_trdz + (size_t)(jn - 1) * n3;
            for (size_t n = 0; n < n3; ++n) { tx[n] = zwx[n]; ty[n] = zwy[n]; tz[n] = zwz[n]; }
        }

        /* anti-diffusive flux : high order minus low order  (:180-251) */
        switch (kn_fct_h) {
        case 2:                                                                 /* :182-190 */
            for (jk = 1; jk <= jpkm1; ++jk)
                for (jj = 1; jj <= jpjm1; ++jj)
                    for (ji = 1; ji <= jpim1; ++ji) {
                        zwx[I3(ji, jj, jk)] = 0.5 * pun[I3(ji, jj, jk)] * (ptn[I3(ji, jj, jk)] + ptn[I3(ji + 1, jj, jk)])
                                              - zwx[I3(ji, jj, jk)];
                        zwy[I3(ji, jj, jk)] = 0.5 * pvn[I3(ji, jj, jk)] * (ptn[I3(ji, jj, jk)] + ptn[I3(ji, jj + 1, jk)])
                                              - zwy[I3(ji, jj, jk)];
                    }
            break;
        case 4:                                                                 /* :192-221 */
            for (jj = 1; jj <= jpj; ++jj) for (ji = 1; ji <= jpi; ++ji) {
                zltu[I3(ji, jj, jpk)] = 0.0; zltv[I3(ji, jj, jpk)] = 0.0;
            }
            for (jk = 1; jk <= jpkm1; ++jk) {
                for (jj = 1; jj <= jpjm1; ++jj)
                    for (ji = 1; ji <= jpim1; ++ji) {
                        ztu[I3(ji, jj, jk)] = (ptn[I3(ji + 1, jj, jk)] - ptn[I3(ji, jj, jk)]) * umask[I3(ji, jj, jk)];
                        ztv[I3(ji, jj, jk)] = (ptn[I3(ji, jj + 1, jk)] - ptn[I3(ji, jj, jk)]) * vmask[I3(ji, jj, jk)];
                    }
                for (jj = 2; jj <= jpjm1; ++jj)
                    for (ji = 2; ji <= jpim1; ++ji) {
                        /* NB: '+' is the reference (traadv_fct.F90:204-205) */
                        zltu[I3(ji, jj, jk)] = (ztu[I3(ji, jj, jk)] + ztu[I3(ji - 1, jj, jk)]) * r1_6;
                        zltv[I3(ji, jj, jk)] = (ztv[I3(ji, jj, jk)] + ztv[I3(ji, jj - 1, jk)]) * r1_6;
                    }
            }
            {   /* X1  (:209) */
                double *pt[2] = { zltu, zltv }; const double sg[2] = { 1.0, 1.0 };
                lbc_lnk_multi(d, "traadv_fct", 2, pt, "TT", sg, jpk, 0, 0.0);
            }
            if (jn == d->dbg_jn) { dbg_copy(d->dbg_zltu, zltu, n3); dbg_copy(d->dbg_zltv, zltv, n3); }
            for (jk = 1; jk <= jpkm1; ++jk)
                for (jj = 1; jj <= jpjm1; ++jj)
                    for (ji = 1; ji <= jpim1; ++ji) {
                        zC2t_u = ptn[I3(ji, jj, jk)] + ptn[I3(ji + 1, jj, jk)];
                        zC2t_v = ptn[I3(ji, jj, jk)] + ptn[I3(ji, jj + 1, jk)];
                        zwx[I3(ji, jj, jk)] = 0.5 * pun[I3(ji, jj, jk)]
                                              * (zC2t_u + zltu[I3(ji, jj, jk)] - zltu[I3(ji + 1, jj, jk)])
                                              - zwx[I3(ji, jj, jk)];
                        zwy[I3(ji, jj, jk)] = 0.5 * pvn[I3(ji, jj, jk)]
                                              * (zC2t_v + zltv[I3(ji, jj, jk)] - zltv[I3(ji, jj + 1, jk)])
                                              - zwy[I3(ji, jj, jk)];
                    }
            break;
        default:
            fprintf(stderr, "oracle tra_adv_fct: kn_fct_h=%d not restated (41 is untested in the reference)\n", kn_fct_h);
            abort();
        }

        switch (kn_fct_v) {                                                     /* :253-275 */
        case 2:
            for (jk = 2; jk <= jpkm1; ++jk)
                for (jj = 2; jj <= jpjm1; ++jj)
                    for (ji = 2; ji <= jpim1; ++ji)
                        zwz[I3(ji, jj, jk)] = (pwn[I3(ji, jj, jk)] * 0.5 * (ptn[I3(ji, jj, jk)] + ptn[I3(ji, jj, jk - 1)])
                                               - zwz[I3(ji, jj, jk)]) * wmask[I3(ji, jj, jk)];
            break;
        case 4:
            interp_4th_cpt(d, ptn, ztw);
            if (jn == d->dbg_jn) dbg_copy(d->dbg_ztw, ztw, n3);
            for (jk = 2; jk <= jpkm1; ++jk)
                for (jj = 2; jj <= jpjm1; ++jj)
                    for (ji = 2; ji <= jpim1; ++ji)
                        zwz[I3(ji, jj, jk)] = (pwn[I3(ji, jj, jk)] * ztw[I3(ji, jj, jk)] - zwz[I3(ji, jj, jk)])
                                              * wmask[I3(ji, jj, jk)];
            break;
        default:
            fprintf(stderr, "oracle tra_adv_fct: kn_fct_v=%d invalid\n", kn_fct_v);
            abort();
        }
        if (d->ln_linssh)                                                       /* :276-278 */
            for (jj = 1; jj <= jpj; ++jj) for (ji = 1; ji <= jpi; ++ji) zwz[I3(ji, jj, 1)] = 0.0;

        {   /* X2  (:280) */
            double *pt[4] = { zwi, zwx, zwy, zwz }; const double sg[4] = { 1.0, -1.0, -1.0, 1.0 };
            lbc_lnk_multi(d, "traadv_fct", 4, pt, "TUVW", sg, jpk, 0, 0.0);
        }
        if (jn == d->dbg_jn) {
            dbg_copy(d->dbg_zwi, zwi, n3); dbg_copy(d->dbg_zwx, zwx, n3);
            dbg_copy(d->dbg_zwy, zwy, n3); dbg_copy(d->dbg_zwz, zwz, n3);
        }

        nonosc(d, ptb, zwx, zwy, zwz, zwi, p2dt, jn);                           /* :284 */

        /* final trend with corrected fluxes  (:288-297) */
        for (jk = 1; jk <= jpkm1; ++jk)
            for (jj = 2; jj <= jpjm1; ++jj)
                for (ji = 2; ji <= jpim1; ++ji)
                    pta[I3(ji, jj, jk)] = pta[I3(ji, jj, jk)]
                        - (zwx[I3(ji, jj, jk)] - zwx[I3(ji - 1, jj, jk)]
                           + zwy[I3(ji, jj, jk)] - zwy[I3(ji, jj - 1, jk)]
                           + zwz[I3(ji, jj, jk)] - zwz[I3(ji, jj, jk + 1)])
                          * r1_e1e2t[I2(ji, jj)] / e3t_n[I3(ji, jj, jk)];
        /* add the (limited) anti-diffusive fluxes to the upstream ones  (:299-303, :313) */
        if (d->diag_trdx) {
            double *tx = d->diag_trdx + (size_t)(jn - 1) * n3, *ty = d->diag_trdy + (size_t)(jn - 1) * n3;
            double *tz = d->diag_trdz + (size_t)(jn - 1) * n3;
            for (size_t n = 0; n < n3; ++n) { tx[n] = tx[n] + zwx[n]; ty[n] = ty[n] + zwy[n]; tz[n] = tz[n] + zwz[n]; }
        }
    }

    free(zwi); free(zwx); free(zwy); free(zwz);
    free(ztu); free(ztv); free(zltu); free(zltv); free(ztw);
}

void nonosc(oce_dom *d, const double *pbef, double *paa, double *pbb, double *pcc, const double *paft,
            double p2dt, int jn)
{
    const int jpi = d->jpi, jpj = d->jpj, jpk = d->jpk;
    const int jpim1 = d->jpim1, jpjm1 = d->jpjm1, jpkm1 = d->jpkm1;
    const size_t jpij = (size_t)jpi * jpj, n3 = jpij * jpk;
    const double *tmask = d->tmask, *e3t_n = d->e3t_n, *e1e2t = d->e1e2t;
    int ji, jj, jk, ikm1;
    double zpos, zneg, zbt, za, zb, zc, zbig, zrtrn;
    double zau, zbu, zcu, zav, zbv, zcv, zup, zdo;
    /* REAL(wp), DIMENSION(jpi,jpj,jpk) :: zbetup, zbetdo, zbup, zbdo   (:351) */
    double *zbetup = auto3d(d), *zbetdo = auto3d(d), *zbup = auto3d(d), *zbdo = auto3d(d);

    zbig  = 1.e+40;                                                             /* :354 */
    zrtrn = 1.e-15;                                                             /* :355 */
    for (size_t n = 0; n < n3; ++n) { zbetup[n] = 0.0; zbetdo[n] = 0.0; }       /* :356 */

    /* Search local extrema: whole-array expressions  (:361-364) */
    for (size_t n = 0; n < n3; ++n) {
        zbup[n] = dmax(pbef[n] * tmask[n] - zbig * (1.0 - tmask[n]), paft[n] * tmask[n] - zbig * (1.0 - tmask[n]));
        zbdo[n] = dmin(pbef[n] * tmask[n] + zbig * (1.0 - tmask[n]), paft[n] * tmask[n] + zbig * (1.0 - tmask[n]));
    }

    for (jk = 1; jk <= jpkm1; ++jk) {                                           /* :366-399 */
        ikm1 = (jk - 1 > 1) ? jk - 1 : 1;
        for (jj = 2; jj <= jpjm1; ++jj)
            for (ji = 2; ji <= jpim1; ++ji) {
                /* search maximum in neighbourhood (Fortran MAX of 7 args: left to right) */
                zup = dmax(dmax(dmax(dmax(dmax(dmax(zbup[I3(ji, jj, jk)],
                      zbup[I3(ji - 1, jj, jk)]), zbup[I3(ji + 1, jj, jk)]),
                      zbup[I3(ji, jj - 1, jk)]), zbup[I3(ji, jj + 1, jk)]),
                      zbup[I3(ji, jj, ikm1)]), zbup[I3(ji, jj, jk + 1)]);
                zdo = dmin(dmin(dmin(dmin(dmin(dmin(zbdo[I3(ji, jj, jk)],
                      zbdo[I3(ji - 1, jj, jk)]), zbdo[I3(ji + 1, jj, jk)]),
                      zbdo[I3(ji, jj - 1, jk)]), zbdo[I3(ji, jj + 1, jk)]),
                      zbdo[I3(ji, jj, ikm1)]), zbdo[I3(ji, jj, jk + 1)]);
                /* positive part of the flux */
                zpos = dmax(0., paa[I3(ji - 1, jj, jk)]) - dmin(0., paa[I3(ji, jj, jk)])
                     + dmax(0., pbb[I3(ji, jj - 1, jk)]) - dmin(0., pbb[I3(ji, jj, jk)])
                     + dmax(0., pcc[I3(ji, jj, jk + 1)]) - dmin(0., pcc[I3(ji, jj, jk)]);
                /* negative part of the flux */
                zneg = dmax(0., paa[I3(ji, jj, jk)]) - dmin(0., paa[I3(ji - 1, jj, jk)])
                     + dmax(0., pbb[I3(ji, jj, jk)]) - dmin(0., pbb[I3(ji, jj - 1, jk)])
                     + dmax(0., pcc[I3(ji, jj, jk)]) - dmin(0., pcc[I3(ji, jj, jk + 1)]);
                /* up & down beta terms */
                zbt = e1e2t[I2(ji, jj)] * e3t_n[I3(ji, jj, jk)] / p2dt;
                zbetup[I3(ji, jj, jk)] = (zup - paft[I3(ji, jj, jk)]) / (zpos + zrtrn) * zbt;
                zbetdo[I3(ji, jj, jk)] = (paft[I3(ji, jj, jk)] - zdo) / (zneg + zrtrn) * zbt;
            }
    }
    {   /* X3  (:400) */
        double *pt[2] = { zbetup, zbetdo }; const double sg[2] = { 1.0, 1.0 };
        lbc_lnk_multi(d, "traadv_fct", 2, pt, "TT", sg, jpk, 0, 0.0);
    }
    if (jn == d->dbg_jn) { dbg_copy(d->dbg_zbetup, zbetup, n3); dbg_copy(d->dbg_zbetdo, zbetdo, n3); }

    /* 3. monotonic flux in the i, j & k directions  (:404-425) */
    for (jk = 1; jk <= jpkm1; ++jk)
        for (jj = 2; jj <= jpjm1; ++jj)
            for (ji = 2; ji <= jpim1; ++ji) {
                zau = dmin(dmin(1.0, zbetdo[I3(ji, jj, jk)]), zbetup[I3(ji + 1, jj, jk)]);
                zbu = dmin(dmin(1.0, zbetup[I3(ji, jj, jk)]), zbetdo[I3(ji + 1, jj, jk)]);
                zcu = (0.5 + sign_nosignedzero(0.5, paa[I3(ji, jj, jk)]));
                paa[I3(ji, jj, jk)] = paa[I3(ji, jj, jk)] * (zcu * zau + (1.0 - zcu) * zbu);

                zav = dmin(dmin(1.0, zbetdo[I3(ji, jj, jk)]), zbetup[I3(ji, jj + 1, jk)]);
                zbv = dmin(dmin(1.0, zbetup[I3(ji, jj, jk)]), zbetdo[I3(ji, jj + 1, jk)]);
                zcv = (0.5 + sign_nosignedzero(0.5, pbb[I3(ji, jj, jk)]));
                pbb[I3(ji, jj, jk)] = pbb[I3(ji, jj, jk)] * (zcv * zav + (1.0 - zcv) * zbv);

                /* monotonic flux in the k direction, i.e. pcc */
                za = dmin(dmin(1., zbetdo[I3(ji, jj, jk + 1)]), zbetup[I3(ji, jj, jk)]);
                zb = dmin(dmin(1., zbetup[I3(ji, jj, jk + 1)]), zbetdo[I3(ji, jj, jk)]);
                zc = (0.5 + sign_nosignedzero(0.5, pcc[I3(ji, jj, jk + 1)]));
                pcc[I3(ji, jj, jk + 1)] = pcc[I3(ji, jj, jk + 1)] * (zc * za + (1.0 - zc) * zb);
            }
    {   /* X4  (:426) */
        double *pt[2] = { paa, pbb }; const double sg[2] = { -1.0, -1.0 };
        lbc_lnk_multi(d, "traadv_fct", 2, pt, "UV", sg, jpk, 0, 0.0);
    }
    if (jn == d->dbg_jn) { dbg_copy(d->dbg_paa, paa, n3); dbg_copy(d->dbg_pbb, pbb, n3); dbg_copy(d->dbg_pcc, pcc, n3); }

    free(zbetup); free(zbetdo); free(zbup); free(zbdo);
}

void interp_4th_cpt(const oce_dom *d, const double *pt_in, double *pt_out)
{
    const int jpi = d->jpi, jpj = d->jpj, jpk = d->jpk;
    const int jpim1 = d->jpim1, jpjm1 = d->jpjm1, jpkm1 = d->jpkm1;
    const size_t jpij = (size_t)jpi * jpj;
    const double *wmask = d->wmask;
    int ji, jj, jk, ikt, ikb;
    /* REAL(wp),DIMENSION(jpi,jpj,jpk) :: zwd, zwi, zws, zwrm, zwt   (:530) */
    double *zwd = auto3d(d), *zwi = auto3d(d), *zws = auto3d(d), *zwrm = auto3d(d), *zwt = auto3d(d);

    for (jk = 3; jk <= jpkm1; ++jk)                                             /* :535-545 */
        for (jj = 2; jj <= jpjm1; ++jj)
            for (ji = 2; ji <= jpim1; ++ji) {
                zwd[I3(ji, jj, jk)]  = 3.0 * wmask[I3(ji, jj, jk)] + 1.0;
                zwi[I3(ji, jj, jk)]  = wmask[I3(ji, jj, jk)];
                zws[I3(ji, jj, jk)]  = wmask[I3(ji, jj, jk)];
                zwrm[I3(ji, jj, jk)] = 3.0 * wmask[I3(ji, jj, jk)] * (pt_in[I3(ji, jj, jk)] + pt_in[I3(ji, jj, jk - 1)]);
            }
    if (d->ln_isfcav) {                                                         /* :554-556 */
        for (jj = 1; jj <= jpj; ++jj) for (ji = 1; ji <= jpi; ++ji) {
            zwd[I3(ji, jj, 2)] = 1.0; zwi[I3(ji, jj, 2)] = 0.0; zws[I3(ji, jj, 2)] = 0.0; zwrm[I3(ji, jj, 2)] = 0.0;
        }
    }
    for (jj = 2; jj <= jpjm1; ++jj)                                             /* :558-573 */
        for (ji = 2; ji <= jpim1; ++ji) {
            ikt = d->mikt[I2(ji, jj)] + 1;
            ikb = d->mbkt[I2(ji, jj)];
            zwd[I3(ji, jj, ikt)] = 1.0;
            zwi[I3(ji, jj, ikt)] = 0.0;
            zws[I3(ji, jj, ikt)] = 0.0;
            zwrm[I3(ji, jj, ikt)] = 0.5 * (pt_in[I3(ji, jj, ikt - 1)] + pt_in[I3(ji, jj, ikt)]);
            zwd[I3(ji, jj, ikb)] = 1.0;
            zwi[I3(ji, jj, ikb)] = 0.0;
            zws[I3(ji, jj, ikb)] = 0.0;
            /* On a land column mbkt = MAX(k_bot,1) = 1 (domzgr.F90:292-294) so the reference reads pt_in(ji,jj,0):
             * out of bounds, and stores to level 1 which the solver below never reads.  Skip the undefined read. */
            if (ikb >= 2) zwrm[I3(ji, jj, ikb)] = 0.5 * (pt_in[I3(ji, jj, ikb - 1)] + pt_in[I3(ji, jj, ikb)]);
            else          zwrm[I3(ji, jj, ikb)] = 0.0;
        }
    /* tridiagonal solver */
    for (jj = 2; jj <= jpjm1; ++jj)                                             /* :577-581 */
        for (ji = 2; ji <= jpim1; ++ji)
            zwt[I3(ji, jj, 2)] = zwd[I3(ji, jj, 2)];
    for (jk = 3; jk <= jpkm1; ++jk)                                             /* :582-588 */
        for (jj = 2; jj <= jpjm1; ++jj)
            for (ji = 2; ji <= jpim1; ++ji)
                zwt[I3(ji, jj, jk)] = zwd[I3(ji, jj, jk)]
                                      - zwi[I3(ji, jj, jk)] * zws[I3(ji, jj, jk - 1)] / zwt[I3(ji, jj, jk - 1)];
    for (jj = 2; jj <= jpjm1; ++jj)                                             /* :590-594 */
        for (ji = 2; ji <= jpim1; ++ji)
            pt_out[I3(ji, jj, 2)] = zwrm[I3(ji, jj, 2)];
    for (jk = 3; jk <= jpkm1; ++jk)                                             /* :595-601 */
        for (jj = 2; jj <= jpjm1; ++jj)
            for (ji = 2; ji <= jpim1; ++ji)
                pt_out[I3(ji, jj, jk)] = zwrm[I3(ji, jj, jk)]
                                         - zwi[I3(ji, jj, jk)] / zwt[I3(ji, jj, jk - 1)] * pt_out[I3(ji, jj, jk - 1)];
    for (jj = 2; jj <= jpjm1; ++jj)                                             /* :603-607 */
        for (ji = 2; ji <= jpim1; ++ji)
            pt_out[I3(ji, jj, jpkm1)] = pt_out[I3(ji, jj, jpkm1)] / zwt[I3(ji, jj, jpkm1)];
    for (jk = jpk - 2; jk >= 2; --jk)                                           /* :608-614 */
        for (jj = 2; jj <= jpjm1; ++jj)
            for (ji = 2; ji <= jpim1; ++ji)
                pt_out[I3(ji, jj, jk)] = (pt_out[I3(ji, jj, jk)] - zws[I3(ji, jj, jk)] * pt_out[I3(ji, jj, jk + 1)])
                                         / zwt[I3(ji, jj, jk)];

    free(zwd); free(zwi); free(zws); free(zwrm); free(zwt);
}
