/*
 * traadv.c -- ORACLE (test infrastructure only; see nemo_oracle.h).
 * Restates the effective-transport build of tra_adv, src/OCE/TRA/traadv.F90:100-124 (Eulerian branch, no Stokes
 * drift / z-tilde / eiv / mle additions, which are optional and off in tests/BENCH).
 */
#include "nemo_oracle.h"

void tra_adv_transports(const oce_dom *d, const double *e2u, const double *e1v, const double *e3u_n,
                        const double *e3v_n, const double *un, const double *vn, const double *wn,
                        double *zun, double *zvn, double *zwn)
{
    const int jpi = d->jpi, jpj = d->jpj, jpk = d->jpk, jpkm1 = d->jpkm1;
    const size_t jpij = (size_t)jpi * jpj;
    for (size_t n = 0; n < jpij; ++n) {                                         /* :101-103 */
        zun[(size_t)(jpk - 1) * jpij + n] = 0.0; zvn[(size_t)(jpk - 1) * jpij + n] = 0.0;
        zwn[(size_t)(jpk - 1) * jpij + n] = 0.0;
    }
    for (int jk = 1; jk <= jpkm1; ++jk)                                         /* :110-114 */
        for (size_t n = 0; n < jpij; ++n) {
            size_t m = (size_t)(jk - 1) * jpij + n;
            zun[m] = e2u[n] * e3u_n[m] * un[m];
            zvn[m] = e1v[n] * e3v_n[m] * vn[m];
            zwn[m] = d->e1e2t[n] * wn[m];
        }
    for (size_t n = 0; n < jpij; ++n) {                                         /* :122-124 */
        zun[(size_t)(jpk - 1) * jpij + n] = 0.0; zvn[(size_t)(jpk - 1) * jpij + n] = 0.0;
        zwn[(size_t)(jpk - 1) * jpij + n] = 0.0;
    }
}
