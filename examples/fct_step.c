/*
 * examples/fct_step.c -- the C ABI of include/nemo_fct.h used from plain C (what a Fortran host does through
 * ISO_C_BINDING): decomposition, context, module arrays, one tra_adv_fct call with host pointers.
 *
 *   gcc -std=c99 -I include examples/fct_step.c -L nemo-fmi-devel_b200 -lnemo_fct -Wl,-rpath,$PWD/nemo-fmi-devel_b200 -o fct_step
 *   ./fct_step            (needs a B200; without a GPU nemo_fct_create fails with a message -- there is no CPU fallback)
 *
 * A closed 64 x 48 x 31 box, all ocean, uniform metrics, a smooth tracer advected by a uniform zonal transport.
 */
#include <stdio.h>
#include <stdlib.h>

#include "nemo_fct.h"

#define CHECK(call) do { if ((call) != 0) { fprintf(stderr, "%s failed: %s\n", #call, nemo_fct_last_error()); return 1; } } while (0)

int main(void)
{
    const int jpiglo = 64, jpjglo = 48, jpk = 31, jperio = 0, kjpt = 1;
    nemo_fct_domain dom;
    CHECK(nemo_mpp_init(jpiglo, jpjglo, jpk, jperio, 1, 1, 1, 1, &dom));          /* mppini.F90:110, one subdomain */
    const size_t jpij = (size_t)dom.jpi * dom.jpj, n3 = jpij * jpk;

    double *tmask = calloc(n3, sizeof(double)), *umask = calloc(n3, sizeof(double)), *vmask = calloc(n3, sizeof(double));
    double *wmask = calloc(n3, sizeof(double)), *e3t = malloc(n3 * sizeof(double));
    double *e1e2t = malloc(jpij * sizeof(double)), *r1_e1e2t = malloc(jpij * sizeof(double));
    int *mikt = malloc(jpij * sizeof(int)), *mbkt = malloc(jpij * sizeof(int));
    double *pun = calloc(n3, sizeof(double)), *pvn = calloc(n3, sizeof(double)), *pwn = calloc(n3, sizeof(double));
    double *ptb = calloc(n3, sizeof(double)), *ptn, *pta = calloc(n3, sizeof(double));
    if (!tmask || !umask || !vmask || !wmask || !e3t || !e1e2t || !r1_e1e2t || !mikt || !mbkt || !pun || !pvn || !pwn || !ptb || !pta) return 2;

#define IDX(i, j, k) ((size_t)(k) * jpij + (size_t)(j) * dom.jpi + (size_t)(i))     /* 0-based here */
    for (size_t p = 0; p < jpij; ++p) { e1e2t[p] = 1.0e10; r1_e1e2t[p] = 1.0e-10; mikt[p] = 1; mbkt[p] = jpk - 1; }
    for (size_t p = 0; p < n3; ++p) e3t[p] = 100.0;
    for (int k = 0; k < jpk - 1; ++k)                                             /* closed box: land on the rim (dommsk.F90:135-147) */
        for (int j = 1; j < dom.jpj - 1; ++j)
            for (int i = 1; i < dom.jpi - 1; ++i) tmask[IDX(i, j, k)] = 1.0;
    for (int k = 0; k < jpk; ++k)
        for (int j = 0; j < dom.jpj - 1; ++j)
            for (int i = 0; i < dom.jpi - 1; ++i) {
                umask[IDX(i, j, k)] = tmask[IDX(i, j, k)] * tmask[IDX(i + 1, j, k)];
                vmask[IDX(i, j, k)] = tmask[IDX(i, j, k)] * tmask[IDX(i, j + 1, k)];
                wmask[IDX(i, j, k)] = k == 0 ? tmask[IDX(i, j, k)] : tmask[IDX(i, j, k)] * tmask[IDX(i, j, k - 1)];
            }
    for (int k = 0; k < jpk - 1; ++k)
        for (int j = 0; j < dom.jpj; ++j)
            for (int i = 0; i < dom.jpi; ++i) {
                ptb[IDX(i, j, k)] = (10.0 + 0.1 * i + 0.05 * j) * tmask[IDX(i, j, k)];
                pun[IDX(i, j, k)] = 0.05 * 1.0e5 * 100.0 * umask[IDX(i, j, k)];      /* e2u * e3u_n * un */
            }
    ptn = ptb;

    nemo_fct_handle h;
    CHECK(nemo_fct_create(&dom, 0, &h));
    CHECK(nemo_fct_set_domain_arrays(h, tmask, umask, vmask, wmask, e1e2t, r1_e1e2t, mikt, mbkt, /*ln_linssh*/ 1, /*ln_isfcav*/ 0));
    CHECK(nemo_fct_set_e3t(h, e3t, e3t, e3t, /*is_device*/ 0));
    CHECK(nemo_tra_adv_fct(h, 1, 1, "TRA", 2.0 * 3600.0, pun, pvn, pwn, ptb, ptn, pta, kjpt, 2, 2));

    double s = 0.0, amax = 0.0;
    for (size_t p = 0; p < n3; ++p) { s += pta[p] * e1e2t[p % jpij] * e3t[p]; if (pta[p] > amax) amax = pta[p]; if (-pta[p] > amax) amax = -pta[p]; }
    printf("tra_adv_fct: max |trend| = %.6e, volume integral of the trend = %.3e, %lld kernel launches\n", amax, s, nemo_fct_launch_count());
    CHECK(nemo_fct_destroy(h));
    return 0;
}
