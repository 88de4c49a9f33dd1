/*
 * traadv_mus.c -- ORACLE (test infrastructure only; see nemo_oracle.h).
 * Loop-for-loop C restatement of src/OCE/TRA/traadv_mus.F90:55-273 (tra_adv_mus, MUSCL scheme): same loop
 * bounds, same operation order, the two automatic pairs zwx/zslpx, zwy/zslpy reused as in the reference.
 * PARITY PIN: bit-identical to the reference's own source executed by translation (oracle/f90exec.py); see nemo_oracle.h.
 *
 * Not restated: l_trd / l_hst / l_ptr diagnostics (:117-124, :207-214, :268), off by default.
 * Module arrays read besides those of tra_adv_fct: r1_e1e2u, r1_e1e2v (dom_oce.F90:118), e3u_n, e3v_n, e3w_n
 * (dom_oce.F90:132-136), xind (traadv_mus.F90:41, built at kt == kit000 by tra_adv_mus_xind below).
 */
#include "nemo_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>

#define I3(ji, jj, jk) ((size_t)((jk) - 1) * jpij + (size_t)((jj) - 1) * jpi + (size_t)((ji) - 1))
#define I2(ji, jj)     ((size_t)((jj) - 1) * jpi + (size_t)((ji) - 1))
#define SIGN(a, b) sign_nosignedzero((a), (b))       /* key_nosignedzero override, lib_fortran.F90:339-351 */

static inline double dmax(double a, double b) { return a > b ? a : b; }
static inline double dmin(double a, double b) { return a < b ? a : b; }

static double *work3d(const oce_dom *d)
{
    size_t n = (size_t)d->jpi * d->jpj * d->jpk;
    double *p = (double *)calloc(n, sizeof(double));
    if (!p) { fprintf(stderr, "oracle: out of memory\n"); abort(); }
    return p;
}

/* Upstream / MUSCL scheme indicator (traadv_mus.F90:99-113).  rnfmsk (jpi,jpj), rnfmsk_z (jpk); upsmsk = 0. */
void tra_adv_mus_xind(const oce_dom *d, int ld_msc_ups, const double *rnfmsk, const double *rnfmsk_z, double *xind)
{
    const int jpi = d->jpi, jpj = d->jpj, jpk = d->jpk, jpkm1 = d->jpkm1;
    const size_t jpij = (size_t)jpi * jpj;
    int ji, jj, jk;
    for (size_t n = 0; n < jpij * jpk; ++n) xind[n] = 1.0;   /* set equal to 1 where up-stream is not needed */
    if (ld_msc_ups) {
        const double upsmsk = 0.0;                           /* not upstream by default */
        for (jk = 1; jk <= jpkm1; ++jk)
            for (jj = 1; jj <= jpj; ++jj)
                for (ji = 1; ji <= jpi; ++ji)
                    xind[I3(ji, jj, jk)] = 1.0 - dmax(rnfmsk[I2(ji, jj)] * rnfmsk_z[jk - 1], upsmsk) * d->tmask[I3(ji, jj, jk)];
    }
}

void tra_adv_mus(oce_dom *d, int kt, int kit000, const char *cdtype, double p2dt,
                 const double *pun, const double *pvn, const double *pwn,
                 const double *ptb_all, double *pta_all, int kjpt, const double *xind)
{
    (void)kt; (void)kit000; (void)cdtype;
    const int jpi = d->jpi, jpj = d->jpj, jpk = d->jpk;
    const int jpim1 = d->jpim1, jpjm1 = d->jpjm1, jpkm1 = d->jpkm1;
    const size_t jpij = (size_t)jpi * jpj, n3 = jpij * jpk;
    const double *tmask = d->tmask, *umask = d->umask, *vmask = d->vmask, *wmask = d->wmask;
    const double *e3t_n = d->e3t_n, *r1_e1e2t = d->r1_e1e2t;
    const double *r1_e1e2u = d->r1_e1e2u, *r1_e1e2v = d->r1_e1e2v, *e3u_n = d->e3u_n, *e3v_n = d->e3v_n, *e3w_n = d->e3w_n;
    int ji, jj, jk, jn;
    double zu, z0u, zzwx, zw, zalpha, zv, z0v, zzwy, z0w;
    double *zwx = work3d(d), *zslpx = work3d(d), *zwy = work3d(d), *zslpy = work3d(d);   /* :89-90 */
    double *tab[2]; const double sgn[2] = { -1.0, -1.0 };

    for (jn = 1; jn <= kjpt; ++jn) {                                            /* :126 */
        const double *ptb = ptb_all + (size_t)(jn - 1) * n3;
        double *pta = pta_all + (size_t)(jn - 1) * n3;
        /* Horizontal advective fluxes: first guess of the slopes  (:131-141) */
        for (jj = 1; jj <= jpj; ++jj) for (ji = 1; ji <= jpi; ++ji) { zwx[I3(ji, jj, jpk)] = 0.0; zwy[I3(ji, jj, jpk)] = 0.0; }
        for (jk = 1; jk <= jpkm1; ++jk)
            for (jj = 1; jj <= jpjm1; ++jj)
                for (ji = 1; ji <= jpim1; ++ji) {
                    zwx[I3(ji, jj, jk)] = umask[I3(ji, jj, jk)] * (ptb[I3(ji + 1, jj, jk)] - ptb[I3(ji, jj, jk)]);
                    zwy[I3(ji, jj, jk)] = vmask[I3(ji, jj, jk)] * (ptb[I3(ji, jj + 1, jk)] - ptb[I3(ji, jj, jk)]);
                }
        tab[0] = zwx; tab[1] = zwy;
        lbc_lnk_multi(d, "traadv_mus", 2, tab, "UV", sgn, jpk, 0, 0.0);         /* :143 (changed sign) */
        /* Slopes of tracer  (:145-156) */
        for (jj = 1; jj <= jpj; ++jj) for (ji = 1; ji <= jpi; ++ji) { zslpx[I3(ji, jj, jpk)] = 0.0; zslpy[I3(ji, jj, jpk)] = 0.0; }
        for (jk = 1; jk <= jpkm1; ++jk)
            for (jj = 2; jj <= jpj; ++jj)
                for (ji = 2; ji <= jpi; ++ji) {
                    zslpx[I3(ji, jj, jk)] = (zwx[I3(ji, jj, jk)] + zwx[I3(ji - 1, jj, jk)])
                                          * (0.25 + SIGN(0.25, zwx[I3(ji, jj, jk)] * zwx[I3(ji - 1, jj, jk)]));
                    zslpy[I3(ji, jj, jk)] = (zwy[I3(ji, jj, jk)] + zwy[I3(ji, jj - 1, jk)])
                                          * (0.25 + SIGN(0.25, zwy[I3(ji, jj, jk)] * zwy[I3(ji, jj - 1, jk)]));
                }
        /* Slopes limitation  (:158-169) */
        for (jk = 1; jk <= jpkm1; ++jk)
            for (jj = 2; jj <= jpj; ++jj)
                for (ji = 2; ji <= jpi; ++ji) {
                    zslpx[I3(ji, jj, jk)] = SIGN(1., zslpx[I3(ji, jj, jk)])
                        * dmin(dmin(fabs(zslpx[I3(ji, jj, jk)]), 2. * fabs(zwx[I3(ji - 1, jj, jk)])), 2. * fabs(zwx[I3(ji, jj, jk)]));
                    zslpy[I3(ji, jj, jk)] = SIGN(1., zslpy[I3(ji, jj, jk)])
                        * dmin(dmin(fabs(zslpy[I3(ji, jj, jk)]), 2. * fabs(zwy[I3(ji, jj - 1, jk)])), 2. * fabs(zwy[I3(ji, jj, jk)]));
                }
        /* MUSCL horizontal advective fluxes  (:171-191) */
        for (jk = 1; jk <= jpkm1; ++jk)
            for (jj = 2; jj <= jpjm1; ++jj)
                for (ji = 2; ji <= jpim1; ++ji) {
                    z0u = SIGN(0.5, pun[I3(ji, jj, jk)]);
                    zalpha = 0.5 - z0u;
                    zu = z0u - 0.5 * pun[I3(ji, jj, jk)] * p2dt * r1_e1e2u[I2(ji, jj)] / e3u_n[I3(ji, jj, jk)];
                    zzwx = ptb[I3(ji + 1, jj, jk)] + xind[I3(ji, jj, jk)] * zu * zslpx[I3(ji + 1, jj, jk)];
                    zzwy = ptb[I3(ji, jj, jk)] + xind[I3(ji, jj, jk)] * zu * zslpx[I3(ji, jj, jk)];
                    zwx[I3(ji, jj, jk)] = pun[I3(ji, jj, jk)] * (zalpha * zzwx + (1. - zalpha) * zzwy);

                    z0v = SIGN(0.5, pvn[I3(ji, jj, jk)]);
                    zalpha = 0.5 - z0v;
                    zv = z0v - 0.5 * pvn[I3(ji, jj, jk)] * p2dt * r1_e1e2v[I2(ji, jj)] / e3v_n[I3(ji, jj, jk)];
                    zzwx = ptb[I3(ji, jj + 1, jk)] + xind[I3(ji, jj, jk)] * zv * zslpy[I3(ji, jj + 1, jk)];
                    zzwy = ptb[I3(ji, jj, jk)] + xind[I3(ji, jj, jk)] * zv * zslpy[I3(ji, jj, jk)];
                    zwy[I3(ji, jj, jk)] = pvn[I3(ji, jj, jk)] * (zalpha * zzwx + (1. - zalpha) * zzwy);
                }
        lbc_lnk_multi(d, "traadv_mus", 2, tab, "UV", sgn, jpk, 0, 0.0);         /* :192 */
        /* Tracer advective trend  (:194-202) */
        for (jk = 1; jk <= jpkm1; ++jk)
            for (jj = 2; jj <= jpjm1; ++jj)
                for (ji = 2; ji <= jpim1; ++ji)
                    pta[I3(ji, jj, jk)] = pta[I3(ji, jj, jk)] - (zwx[I3(ji, jj, jk)] - zwx[I3(ji - 1, jj, jk)]
                                                               + zwy[I3(ji, jj, jk)] - zwy[I3(ji, jj - 1, jk)])
                                                              * r1_e1e2t[I2(ji, jj)] / e3t_n[I3(ji, jj, jk)];

        /* Vertical advective fluxes: first guess of the slopes  (:219-223) */
        for (jj = 1; jj <= jpj; ++jj) for (ji = 1; ji <= jpi; ++ji) { zwx[I3(ji, jj, 1)] = 0.0; zwx[I3(ji, jj, jpk)] = 0.0; }
        for (jk = 2; jk <= jpkm1; ++jk)
            for (jj = 1; jj <= jpj; ++jj)
                for (ji = 1; ji <= jpi; ++ji)
                    zwx[I3(ji, jj, jk)] = tmask[I3(ji, jj, jk)] * (ptb[I3(ji, jj, jk - 1)] - ptb[I3(ji, jj, jk)]);
        /* Slopes of tracer  (:225-233) */
        for (jj = 1; jj <= jpj; ++jj) for (ji = 1; ji <= jpi; ++ji) zslpx[I3(ji, jj, 1)] = 0.0;
        for (jk = 2; jk <= jpkm1; ++jk)
            for (jj = 1; jj <= jpj; ++jj)
                for (ji = 1; ji <= jpi; ++ji)
                    zslpx[I3(ji, jj, jk)] = (zwx[I3(ji, jj, jk)] + zwx[I3(ji, jj, jk + 1)])
                                          * (0.25 + SIGN(0.25, zwx[I3(ji, jj, jk)] * zwx[I3(ji, jj, jk + 1)]));
        /* Slopes limitation  (:234-242) */
        for (jk = 2; jk <= jpkm1; ++jk)
            for (jj = 1; jj <= jpj; ++jj)
                for (ji = 1; ji <= jpi; ++ji)
                    zslpx[I3(ji, jj, jk)] = SIGN(1., zslpx[I3(ji, jj, jk)])
                        * dmin(dmin(fabs(zslpx[I3(ji, jj, jk)]), 2. * fabs(zwx[I3(ji, jj, jk + 1)])), 2. * fabs(zwx[I3(ji, jj, jk)]));
        /* vertical advective flux  (:243-253) */
        for (jk = 1; jk <= jpk - 2; ++jk)
            for (jj = 2; jj <= jpjm1; ++jj)
                for (ji = 2; ji <= jpim1; ++ji) {
                    z0w = SIGN(0.5, pwn[I3(ji, jj, jk + 1)]);
                    zalpha = 0.5 + z0w;
                    zw = z0w - 0.5 * pwn[I3(ji, jj, jk + 1)] * p2dt * r1_e1e2t[I2(ji, jj)] / e3w_n[I3(ji, jj, jk + 1)];
                    zzwx = ptb[I3(ji, jj, jk + 1)] + xind[I3(ji, jj, jk)] * zw * zslpx[I3(ji, jj, jk + 1)];
                    zzwy = ptb[I3(ji, jj, jk)] + xind[I3(ji, jj, jk)] * zw * zslpx[I3(ji, jj, jk)];
                    zwx[I3(ji, jj, jk + 1)] = pwn[I3(ji, jj, jk + 1)] * (zalpha * zzwx + (1. - zalpha) * zzwy) * wmask[I3(ji, jj, jk)];
                }
        if (d->ln_linssh) {                                                     /* top values, linear free surface only (:254-264) */
            if (d->ln_isfcav) {
                for (jj = 1; jj <= jpj; ++jj)
                    for (ji = 1; ji <= jpi; ++ji) {
                        int ik = d->mikt[I2(ji, jj)];
                        zwx[I3(ji, jj, ik)] = pwn[I3(ji, jj, ik)] * ptb[I3(ji, jj, ik)];
                    }
            } else {
                for (jj = 1; jj <= jpj; ++jj)
                    for (ji = 1; ji <= jpi; ++ji)
                        zwx[I3(ji, jj, 1)] = pwn[I3(ji, jj, 1)] * ptb[I3(ji, jj, 1)];
            }
        }
        /* vertical advective trend  (:266-272) */
        for (jk = 1; jk <= jpkm1; ++jk)
            for (jj = 2; jj <= jpjm1; ++jj)
                for (ji = 2; ji <= jpim1; ++ji)
                    pta[I3(ji, jj, jk)] = pta[I3(ji, jj, jk)] - (zwx[I3(ji, jj, jk)] - zwx[I3(ji, jj, jk + 1)])
                                                              * r1_e1e2t[I2(ji, jj)] / e3t_n[I3(ji, jj, jk)];
    }
    free(zwx); free(zslpx); free(zwy); free(zslpy);
}
