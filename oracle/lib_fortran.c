/*
 * lib_fortran.c -- ORACLE (test infrastructure only; see nemo_oracle.h).
 * Restates src/OCE/lib_fortran.F90:300-332 (DDPDD), :339-351 (SIGN with key_nosignedzero) and the local part of
 * glob_sum, src/OCE/lib_fortran_generic.h90:52-63.
 */
#include "nemo_oracle.h"
#include <math.h>

/* SIGN_SCALAR (lib_fortran.F90:339-351): IF( pb >= 0.e0 ) |pa| ELSE -|pa|   => -0.0 counts as positive */
double sign_nosignedzero(double pa, double pb)
{
    if (pb >= 0.e0) return fabs(pa);
    else            return -fabs(pa);
}

/* DDPDD (lib_fortran.F90:300-332): yddb = ydda + yddb in double-double, Knuth's trick.  [0]=REAL, [1]=AIMAG */
void ddpdd(const double ydda[2], double yddb[2])
{
    volatile double zerr, zt1, zt2;   /* volatile: forbid any algebraic simplification of the error term */
    zt1  = ydda[0] + yddb[0];
    zerr = zt1 - ydda[0];
    zt2  = ((yddb[0] - zerr) + (ydda[0] - (zt1 - zerr))) + ydda[1] + yddb[1];
    double s = zt1 + zt2;
    yddb[1] = zt2 - (s - zt1);
    yddb[0] = s;
}

/* lib_fortran_generic.h90:52-63:  DO jk; DO jj; DO ji:  ztmp = ptab*tmask_i ; ctmp = DDPDD( CMPLX(ztmp,0), ctmp ) */
void glob_sum_local(const double *ptab, const double *tmask_i, int jpi, int jpj, int ipk, double ctmp[2])
{
    for (int jk = 0; jk < ipk; ++jk)
        for (int jj = 0; jj < jpj; ++jj)
            for (int ji = 0; ji < jpi; ++ji) {
                double z[2];
                z[0] = ptab[((size_t)jk * jpj + jj) * jpi + ji] * tmask_i[(size_t)jj * jpi + ji];
                z[1] = 0.0;
                ddpdd(z, ctmp);
            }
}
