/*
 * traadv_cen.c -- ORACLE (test infrastructure only; see nemo_oracle.h).
 * Loop-for-loop C restatement of src/OCE/TRA/traadv_cen.F90:46-204 (tra_adv_cen, 2nd / 4th order centred scheme,
 * 2nd order or 4th-order COMPACT in the vertical).  PARITY PIN: bit-identical to the reference's own source executed by translation (oracle/f90exec.py); see nemo_oracle.h.
 *
 * REFERENCE DEFECT kept visible (kn_cen_h = 4 only, traadv_cen.F90:124-137): the flux loop runs ji = 1..jpim1,
 * jj = 2..jpjm1, so (i) at ji = 1 it reads ztu(0,jj,jk), one element before the row -- in the contiguous automatic
 * array that is ztu(jpi,jj-1,jk), which is what this restatement reads -- and (ii) zwy(:,1,:) is never assigned although
 * the divergence at jj = 2 reads it; the automatic array is undefined there in Fortran, here it is the work array's
 * initial value (0, or NaN under oracle_poison_workspace).  Rows jj = 2 (all of it) and column ji = 2 of the 4th-order
 * result are therefore not decomposition-invariant in the reference either.  kn_cen_h = 2 has no such problem.
 *
 * Not restated: l_trd / l_hst / l_ptr diagnostics (:89-96, :191-199), off by default.
 */
#include "nemo_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>

#define I3(ji, jj, jk) ((size_t)((jk) - 1) * jpij + (size_t)((jj) - 1) * jpi + (size_t)((ji) - 1))
#define I2(ji, jj)     ((size_t)((jj) - 1) * jpi + (size_t)((ji) - 1))

static double *auto3d(const oce_dom *d)
{
    size_t n = (size_t)d->jpi * d->jpj * d->jpk;
    double *p = (double *)malloc(n * sizeof(double));
    if (!p) { fprintf(stderr, "oracle: out of memory\n"); abort(); }
    if (oracle_poison_enabled()) { for (size_t i = 0; i < n; ++i) p[i] = NAN; }
    else                         { memset(p, 0, n * sizeof(double)); }
    return p;
}

void tra_adv_cen(oce_dom *d, int kt, int kit000, const char *cdtype, const double *pun, const double *pvn,
                 const double *pwn, const double *ptn_all, double *pta_all, int kjpt, int kn_cen_h, int kn_cen_v)
{
    (void)kt; (void)kit000; (void)cdtype;
    const int jpi = d->jpi, jpj = d->jpj, jpk = d->jpk;
    const int jpim1 = d->jpim1, jpjm1 = d->jpjm1, jpkm1 = d->jpkm1;
    const size_t jpij = (size_t)jpi * jpj, n3 = jpij * jpk;
    const double *umask = d->umask, *vmask = d->vmask, *wmask = d->wmask, *e3t_n = d->e3t_n, *r1_e1e2t = d->r1_e1e2t;
    const double r1_6 = 1.0 / 6.0;
    int ji, jj, jk, jn;
    double zC2t_u, zC4t_u, zC2t_v, zC4t_v;
    double *zwx = auto3d(d), *zwy = auto3d(d), *zwz = auto3d(d), *ztu = NULL, *ztv = NULL, *ztw = NULL;   /* :78 */
    if (kn_cen_h == 4) { ztu = auto3d(d); ztv = auto3d(d); }
    if (kn_cen_v == 4) { ztw = auto3d(d); }

    for (jj = 1; jj <= jpj; ++jj) for (ji = 1; ji <= jpi; ++ji) { zwz[I3(ji, jj, 1)] = 0.0; zwz[I3(ji, jj, jpk)] = 0.0; }   /* :99-100 */

    for (jn = 1; jn <= kjpt; ++jn) {                                            /* :102 */
        const double *ptn = ptn_all + (size_t)(jn - 1) * n3;
        double *pta = pta_all + (size_t)(jn - 1) * n3;
        switch (kn_cen_h) {
        case 2:                                                                 /* :106-114 */
            for (jk = 1; jk <= jpkm1; ++jk)
                for (jj = 1; jj <= jpjm1; ++jj)
                    for (ji = 1; ji <= jpim1; ++ji) {
                        zwx[I3(ji, jj, jk)] = 0.5 * pun[I3(ji, jj, jk)] * (ptn[I3(ji, jj, jk)] + ptn[I3(ji + 1, jj, jk)]);
                        zwy[I3(ji, jj, jk)] = 0.5 * pvn[I3(ji, jj, jk)] * (ptn[I3(ji, jj, jk)] + ptn[I3(ji, jj + 1, jk)]);
                    }
            break;
        case 4:                                                                 /* :116-141 */
            for (jj = 1; jj <= jpj; ++jj) for (ji = 1; ji <= jpi; ++ji) { ztu[I3(ji, jj, jpk)] = 0.0; ztv[I3(ji, jj, jpk)] = 0.0; }
            for (jk = 1; jk <= jpkm1; ++jk)
                for (jj = 2; jj <= jpjm1; ++jj)
                    for (ji = 2; ji <= jpim1; ++ji) {
                        ztu[I3(ji, jj, jk)] = (ptn[I3(ji + 1, jj, jk)] - ptn[I3(ji, jj, jk)]) * umask[I3(ji, jj, jk)];
                        ztv[I3(ji, jj, jk)] = (ptn[I3(ji, jj + 1, jk)] - ptn[I3(ji, jj, jk)]) * vmask[I3(ji, jj, jk)];
                    }
            {
                double *pt[2] = { ztu, ztv }; const double sg[2] = { -1.0, -1.0 };
                lbc_lnk_multi(d, "traadv_cen", 2, pt, "UV", sg, jpk, 0, 0.0);   /* :126 */
            }
            for (jk = 1; jk <= jpkm1; ++jk)
                for (jj = 2; jj <= jpjm1; ++jj)
                    for (ji = 1; ji <= jpim1; ++ji) {                           /* ji = 1 reads ztu(0,jj,jk): see the header */
                        zC2t_u = ptn[I3(ji, jj, jk)] + ptn[I3(ji + 1, jj, jk)];
                        zC2t_v = ptn[I3(ji, jj, jk)] + ptn[I3(ji, jj + 1, jk)];
                        zC4t_u = zC2t_u + r1_6 * (ztu[I3(ji - 1, jj, jk)] - ztu[I3(ji + 1, jj, jk)]);
                        zC4t_v = zC2t_v + r1_6 * (ztv[I3(ji, jj - 1, jk)] - ztv[I3(ji, jj + 1, jk)]);
                        zwx[I3(ji, jj, jk)] = 0.5 * pun[I3(ji, jj, jk)] * zC4t_u;
                        zwy[I3(ji, jj, jk)] = 0.5 * pvn[I3(ji, jj, jk)] * zC4t_v;
                    }
            break;
        default:
            fprintf(stderr, "oracle tra_adv_cen: wrong value for nn_cen_h = %d\n", kn_cen_h); abort();
        }
        switch (kn_cen_v) {                                                     /* :147-168 */
        case 2:
            for (jk = 2; jk <= jpk; ++jk)
                for (jj = 2; jj <= jpjm1; ++jj)
                    for (ji = 2; ji <= jpim1; ++ji)
                        zwz[I3(ji, jj, jk)] = 0.5 * pwn[I3(ji, jj, jk)] * (ptn[I3(ji, jj, jk)] + ptn[I3(ji, jj, jk - 1)])
                                              * wmask[I3(ji, jj, jk)];
            break;
        case 4:
            interp_4th_cpt(d, ptn, ztw);
            for (jk = 2; jk <= jpkm1; ++jk)
                for (jj = 2; jj <= jpjm1; ++jj)
                    for (ji = 2; ji <= jpim1; ++ji)
                        zwz[I3(ji, jj, jk)] = pwn[I3(ji, jj, jk)] * ztw[I3(ji, jj, jk)] * wmask[I3(ji, jj, jk)];
            break;
        default:
            fprintf(stderr, "oracle tra_adv_cen: wrong value for nn_cen_v = %d\n", kn_cen_v); abort();
        }
        if (d->ln_linssh) {                                                     /* :170-180 */
            if (d->ln_isfcav) {
                for (jj = 1; jj <= jpj; ++jj)
                    for (ji = 1; ji <= jpi; ++ji) {
                        int ik = d->mikt[I2(ji, jj)];
                        zwz[I3(ji, jj, ik)] = pwn[I3(ji, jj, ik)] * ptn[I3(ji, jj, ik)];
                    }
            } else {
                for (jj = 1; jj <= jpj; ++jj)
                    for (ji = 1; ji <= jpi; ++ji)
                        zwz[I3(ji, jj, 1)] = pwn[I3(ji, jj, 1)] * ptn[I3(ji, jj, 1)];
            }
        }
        for (jk = 1; jk <= jpkm1; ++jk)                                         /* :182-190 */
            for (jj = 2; jj <= jpjm1; ++jj)
                for (ji = 2; ji <= jpim1; ++ji)
                    pta[I3(ji, jj, jk)] = pta[I3(ji, jj, jk)]
                        - (zwx[I3(ji, jj, jk)] - zwx[I3(ji - 1, jj, jk)]
                           + zwy[I3(ji, jj, jk)] - zwy[I3(ji, jj - 1, jk)]
                           + zwz[I3(ji, jj, jk)] - zwz[I3(ji, jj, jk + 1)]) * r1_e1e2t[I2(ji, jj)] / e3t_n[I3(ji, jj, jk)];
    }
    free(zwx); free(zwy); free(zwz); free(ztu); free(ztv); free(ztw);
}
