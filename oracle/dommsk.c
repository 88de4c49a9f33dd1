/*
 * dommsk.c -- ORACLE (test infrastructure only; see nemo_oracle.h).
 * Restates the mask derivation of dom_msk, src/OCE/DOM/dommsk.F90:135-147 (tmask), :173-185 (umask, vmask),
 * :189-196 (wmask), :206-238 (tmask_h / tmask_i) and mikt/mbkt of src/OCE/DOM/domzgr.F90:292-294.
 * fmask and the wu/wv masks are not on the FCT path and are not built.
 */
#include "nemo_oracle.h"
#include <stdlib.h>

#define I3(ji, jj, jk) ((size_t)((jk) - 1) * jpij + (size_t)((jj) - 1) * jpi + (size_t)((ji) - 1))
#define I2(ji, jj)     ((size_t)((jj) - 1) * jpi + (size_t)((ji) - 1))

void dom_msk(oce_dom *d, const int *k_top, const int *k_bot, double *tmask, double *umask, double *vmask,
             double *wmask, double *tmask_i, int *mikt, int *mbkt)
{
    const int jpi = d->jpi, jpj = d->jpj, jpk = d->jpk, jpim1 = d->jpim1, jpjm1 = d->jpjm1;
    const size_t jpij = (size_t)jpi * jpj, n3 = jpij * jpk;
    int ji, jj, jk;

    for (size_t n = 0; n < jpij; ++n) {                                         /* domzgr.F90:292-294 */
        mikt[n] = k_top[n] > 1 ? k_top[n] : 1;
        mbkt[n] = k_bot[n] > 1 ? k_bot[n] : 1;
    }
    for (size_t n = 0; n < n3; ++n) tmask[n] = 0.0;                             /* dommsk.F90:135-144 */
    for (jj = 1; jj <= jpj; ++jj)
        for (ji = 1; ji <= jpi; ++ji) {
            int iktop = k_top[I2(ji, jj)], ikbot = k_bot[I2(ji, jj)];
            if (iktop != 0) for (jk = iktop; jk <= ikbot; ++jk) tmask[I3(ji, jj, jk)] = 1.0;
        }
    { double *pt[1] = { tmask }; const double sg[1] = { 1.0 };                  /* :147 */
      lbc_lnk_multi(d, "dommsk", 1, pt, "T", sg, jpk, 0, 0.0); }

    for (size_t n = 0; n < n3; ++n) { umask[n] = 0.0; vmask[n] = 0.0; }         /* ALLOCATE'd arrays: define them */
    for (jk = 1; jk <= jpk; ++jk)                                               /* :173-183 */
        for (jj = 1; jj <= jpjm1; ++jj)
            for (ji = 1; ji <= jpim1; ++ji) {
                umask[I3(ji, jj, jk)] = tmask[I3(ji, jj, jk)] * tmask[I3(ji + 1, jj, jk)];
                vmask[I3(ji, jj, jk)] = tmask[I3(ji, jj, jk)] * tmask[I3(ji, jj + 1, jk)];
            }
    { double *pt[2] = { umask, vmask }; const double sg[2] = { 1.0, 1.0 };      /* :185 */
      lbc_lnk_multi(d, "dommsk", 2, pt, "UV", sg, jpk, 0, 0.0); }

    for (size_t n = 0; n < jpij; ++n) wmask[n] = tmask[n];                      /* :189 */
    for (jk = 2; jk <= jpk; ++jk)                                               /* :192-196 */
        for (size_t n = 0; n < jpij; ++n)
            wmask[(size_t)(jk - 1) * jpij + n] = tmask[(size_t)(jk - 1) * jpij + n] * tmask[(size_t)(jk - 2) * jpij + n];

    if (tmask_i) {                                                              /* :200-238 */
        const int nn_hls = 1;
        int iif = nn_hls, iil = d->nlci - nn_hls + 1, ijf = nn_hls, ijl = d->nlcj - nn_hls + 1;
        double *tmask_h = (double *)malloc(jpij * sizeof(double));
        for (size_t n = 0; n < jpij; ++n) tmask_h[n] = 1.0;
        for (jj = 1; jj <= jpj; ++jj) {
            for (ji = 1; ji <= iif; ++ji) tmask_h[I2(ji, jj)] = 0.0;
            for (ji = iil; ji <= jpi; ++ji) tmask_h[I2(ji, jj)] = 0.0;
        }
        for (ji = 1; ji <= jpi; ++ji) {
            for (jj = 1; jj <= ijf; ++jj) tmask_h[I2(ji, jj)] = 0.0;
            for (jj = ijl; jj <= jpj; ++jj) tmask_h[I2(ji, jj)] = 0.0;
        }
        if (d->jperio == 3 || d->jperio == 4) {                                 /* T-point pivot: tpol */
            if (d->nlej + d->njmpp - 1 == d->jpjglo) {                          /* mjg(nlej) == jpjglo */
                for (ji = iif + 1; ji <= iil - 1; ++ji) {
                    int mig = ji + d->nimpp - 1;
                    double tpol = (mig >= d->jpiglo / 2 + 1) ? 0.0 : 1.0;
                    tmask_h[I2(ji, d->nlej - 1)] = tmask_h[I2(ji, d->nlej - 1)] * tpol;
                }
            }
        }
        for (jj = 1; jj <= jpj; ++jj)
            for (ji = 1; ji <= jpi; ++ji) {
                double ssmask = 0.0;                                            /* MAXVAL( tmask, DIM=3 ) */
                for (jk = 1; jk <= jpk; ++jk) if (tmask[I3(ji, jj, jk)] > ssmask) ssmask = tmask[I3(ji, jj, jk)];
                tmask_i[I2(ji, jj)] = ssmask * tmask_h[I2(ji, jj)];
            }
        free(tmask_h);
    }
}
