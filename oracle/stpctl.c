/*
 * stpctl.c -- ORACLE (test infrastructure only; see nemo_oracle.h).
 * Restates the extrema test and the error condition of stp_ctl, src/OCE/stpctl.F90:115-124 (zmax(1:6), ll_wd = .false.),
 * :149-166 (the condition and the MAXLOC / MINLOC of the branch without ln_ctl), :184 (kindic = -3).  The files it opens
 * (time.step, run.stat, output.abort) and the ln_zad_Aimp pair zmax(8:9) are not restated.
 * PARITY PIN: bit-identical to the reference's own stp_ctl text executed by translation (tests/test_cpu_reference_exec.py).
 * MAXVAL / MAXLOC with NaN operands are processor dependent in Fortran; gfortran skips NaN unless every operand is one.  The
 * restatement does the same (comparisons with NaN are false) and reports separately whether a NaN was met.
 */
#include "nemo_oracle.h"
#include <float.h>
#include <math.h>

void stp_ctl_local(const oce_dom *d, const double *sshn, const double *un, const double *tsn, const double *tmask,
                   double zmax[6], int ih[2], int iu[3], int is1[3], int is2[3], int *nan_found, int *kindic)
{
    const int jpi = d->jpi, jpj = d->jpj, jpk = d->jpk;
    const size_t jpij = (size_t)jpi * jpj, n3 = jpij * jpk;
    const double *tem = tsn, *sal = tsn + n3;                                   /* jp_tem = 1, jp_sal = 2 */
    double z1 = -DBL_MAX, z2 = -DBL_MAX, smin = DBL_MAX, smax = -DBL_MAX, tmin = DBL_MAX, tmax = -DBL_MAX;
    long l1 = -1, l2 = -1, ls1 = -1, ls2 = -1;
    int someoce = 0, isnan_ = 0;

    for (size_t n = 0; n < jpij; ++n) {                                          /* :116-120  MAXVAL( ABS( sshn ) ) */
        double a = fabs(sshn[n]);
        if (a != a) isnan_ = 1;
        if (a > z1 || l1 < 0) { if (a == a) { z1 = a; l1 = (long)n; } }
    }
    for (size_t n = 0; n < n3; ++n) {                                            /* :121  MAXVAL( ABS( un ) ) */
        double a = fabs(un[n]);
        if (a != a) isnan_ = 1;
        if (a > z2 || l2 < 0) { if (a == a) { z2 = a; l2 = (long)n; } }
    }
    for (size_t n = 0; n < n3; ++n) {                                            /* :122-125, mask = tmask == 1 */
        if (tmask[n] != 1.0) continue;
        someoce = 1;                                                             /* lsomeoce: ssmask = MAXVAL( tmask, DIM=3 ) */
        double s = sal[n], t = tem[n];
        if (s != s) isnan_ = 1;
        if (s < smin || (ls1 < 0 && s == s)) { smin = s; ls1 = (long)n; }
        if (s > smax || (ls2 < 0 && s == s)) { smax = s; ls2 = (long)n; }
        if (t < tmin) tmin = t;
        if (t > tmax) tmax = t;
    }
    zmax[0] = l1 < 0 ? -DBL_MAX : z1;  zmax[1] = l2 < 0 ? -DBL_MAX : z2;
    zmax[2] = ls1 < 0 ? -DBL_MAX : -smin;  zmax[3] = ls2 < 0 ? -DBL_MAX : smax;
    zmax[4] = tmin == DBL_MAX ? -DBL_MAX : -tmin;  zmax[5] = tmax;
    /* MAXLOC / MINLOC + (/ nimpp - 1, njmpp - 1, 0 /)   (:162-165); an empty mask gives (0, 0, 0) before the offset */
#define LOC3(l, o) do { long m_ = (l); if (m_ < 0) { o[0] = d->nimpp - 1; o[1] = d->njmpp - 1; o[2] = 0; } else { \
        o[0] = (int)(m_ % jpi) + 1 + d->nimpp - 1; o[1] = (int)((m_ / jpi) % jpj) + 1 + d->njmpp - 1; o[2] = (int)(m_ / (long)jpij) + 1; } } while (0)
    { int t3[3]; LOC3(l1, t3); ih[0] = t3[0]; ih[1] = t3[1]; }
    LOC3(l2, iu); LOC3(ls1, is1); LOC3(ls2, is2);
#undef LOC3
    *nan_found = isnan_;
    *kindic = 0;
    if (someoce && (zmax[0] > 20.0 || zmax[1] > 10.0 || zmax[2] >= 0.0 || zmax[3] >= 100.0 || zmax[3] < 0.0 || isnan_))   /* :149-156 */
        *kindic = -3;                                                            /* :184 */
}
