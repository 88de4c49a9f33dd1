/*
 * mppini.c -- ORACLE (test infrastructure only; see nemo_oracle.h).
 * Restates src/OCE/LBC/mppini.F90: mpp_init (:110-692; the non-MPI variant :53-102), mpp_basic_decomposition
 * (:695-798), mpp_ini_north (lib_mpp.F90:1038-1095, the rank list only) and mpp_init_nfdcom (:1180-1240).
 * Land-subdomain elimination (:407-488) is restated for the all-ocean case only (tests/BENCH never has land
 * subdomains: mpp_init_isoce returns all-ocean without a bathymetry file, :1057-1060), so ipproc(ii,ij) is the
 * zone number itself and jpnij = jpni*jpnj.
 */
#include "nemo_oracle.h"
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define NN_HLS 1   /* src/OCE/par_oce.F90:76 */
#define IX(ii, ij) ((size_t)((ij) - 1) * knbi + (size_t)((ii) - 1))   /* (knbi,knbj) column-major, 1-based */

static int imax(int a, int b) { return a > b ? a : b; }
static int imod(int a, int b) { return a % b; }   /* arguments are non-negative on this path */

void mpp_basic_decomposition(int jpiglo, int jpjglo, int jperio, int knbi, int knbj, int *kimax, int *kjmax,
                             int *kimppt, int *kjmppt, int *klci, int *klcj)
{
    int ji, jj, iresti, irestj, irm, ijpjmin, ireci, irecj;
    *kimax = (jpiglo - 2 * NN_HLS + (knbi - 1)) / knbi + 2 * NN_HLS;            /* :723 */
    *kjmax = (jpjglo - 2 * NN_HLS + (knbj - 1)) / knbj + 2 * NN_HLS;            /* :724 */
    if (!kimppt) return;                                                        /* :726 */
    ireci = 2 * NN_HLS; irecj = 2 * NN_HLS;
    iresti = 1 + imod(jpiglo - ireci - 1, knbi);                                /* :737 */
    irestj = 1 + imod(jpjglo - irecj - 1, knbj);                                /* :738 */
    for (jj = 1; jj <= knbj; ++jj) {                                            /* :748-749 */
        for (ji = 1; ji <= iresti; ++ji) klci[IX(ji, jj)] = *kimax;
        for (ji = iresti + 1; ji <= knbi; ++ji) klci[IX(ji, jj)] = *kimax - 1;
    }
    for (ji = 1; ji <= knbi; ++ji) for (jj = 1; jj <= knbj; ++jj)
        if (klci[IX(ji, jj)] < 3) { fprintf(stderr, "mpp_basic_decomposition: minimum value of jpi must be >= 3\n"); abort(); }
    if (jperio == 3 || jperio == 4 || jperio == 5 || jperio == 6) {             /* :755-764 */
        ijpjmin = (jperio == 3 || jperio == 4) ? 5 : 4;
        irm = knbj - irestj;
        for (ji = 1; ji <= knbi; ++ji) klcj[IX(ji, knbj)] = imax(ijpjmin, *kjmax - irm);
        irm = irm - (*kjmax - klcj[IX(1, knbj)]);
        irestj = knbj - 1 - irm;
        for (ji = 1; ji <= knbi; ++ji) {
            for (jj = 1; jj <= irestj; ++jj) klcj[IX(ji, jj)] = *kjmax;
            for (jj = irestj + 1; jj <= knbj - 1; ++jj) klcj[IX(ji, jj)] = *kjmax - 1;
        }
    } else {                                                                    /* :765-769 */
        ijpjmin = 3;
        for (ji = 1; ji <= knbi; ++ji) {
            for (jj = 1; jj <= irestj; ++jj) klcj[IX(ji, jj)] = *kjmax;
            for (jj = irestj + 1; jj <= knbj; ++jj) klcj[IX(ji, jj)] = *kjmax - 1;
        }
    }
    for (ji = 1; ji <= knbi; ++ji) for (jj = 1; jj <= knbj; ++jj)
        if (klcj[IX(ji, jj)] < ijpjmin) { fprintf(stderr, "mpp_basic_decomposition: minimum value of jpj must be >= %d\n", ijpjmin); abort(); }
    for (ji = 1; ji <= knbi; ++ji) for (jj = 1; jj <= knbj; ++jj) { kimppt[IX(ji, jj)] = 1; kjmppt[IX(ji, jj)] = 1; }
    if (knbi > 1)                                                               /* :782-788 */
        for (jj = 1; jj <= knbj; ++jj)
            for (ji = 2; ji <= knbi; ++ji)
                kimppt[IX(ji, jj)] = kimppt[IX(ji - 1, jj)] + klci[IX(ji - 1, jj)] - ireci;
    if (knbj > 1)                                                               /* :790-796 */
        for (jj = 2; jj <= knbj; ++jj)
            for (ji = 1; ji <= knbi; ++ji)
                kjmppt[IX(ji, jj)] = kjmppt[IX(ji, jj - 1)] + klcj[IX(ji, jj - 1)] - irecj;
}

static void set_sizes(oce_dom *d, int jpk)
{
    d->jpi = d->nlci; d->jpj = d->nlcj; d->jpk = jpk;                           /* :568-570 */
    d->jpim1 = d->jpi - 1; d->jpjm1 = d->jpj - 1; d->jpkm1 = imax(1, jpk - 1);  /* :577-579 */
}

oce_world *mpp_init(int jpiglo, int jpjglo, int jpk, int jperio, int jpni, int jpnj, int ln_nnogather,
                    int key_mpp_mpi)
{
    oce_world *w = (oce_world *)calloc(1, sizeof(oce_world));
    w->jpiglo = jpiglo; w->jpjglo = jpjglo; w->jperio = jperio;

    if (!key_mpp_mpi) {                                                         /* mppini.F90:53-102 */
        if (jpni != 1 || jpnj != 1) { fprintf(stderr, "mpp_init: jpni = jpnj = 1 required without key_mpp_mpi\n"); abort(); }
        w->jpni = w->jpnj = w->jpnij = 1; w->jpimax = jpiglo; w->jpjmax = jpjglo;
        w->dom = (oce_dom *)calloc(1, sizeof(oce_dom));
        oce_dom *d = &w->dom[0];
        d->jpiglo = jpiglo; d->jpjglo = jpjglo; d->jperio = jperio;
        d->jpimax = jpiglo; d->jpjmax = jpjglo;
        d->jpni = d->jpnj = d->jpnij = 1;
        d->nimpp = d->njmpp = 1; d->nlci = jpiglo; d->nlcj = jpjglo;
        d->nldi = d->nldj = 1; d->nlei = jpiglo; d->nlej = jpjglo;
        d->nbondi = d->nbondj = 2; d->npolj = jperio;
        d->narea = 1; d->nproc = 0; d->noea = d->nowe = d->noso = d->nono = -1;
        d->l_Iperio = (jperio == 1 || jperio == 4 || jperio == 6 || jperio == 7);
        d->l_Jperio = (jperio == 2 || jperio == 7);
        d->nreci = d->nrecj = 2 * NN_HLS;
        d->world = NULL;                                                         /* lbc_lnk = lbc_lnk_generic */
        set_sizes(d, jpk);
        return w;
    }

    const int knbi = jpni, jpnij = jpni * jpnj;
    w->jpni = jpni; w->jpnj = jpnj; w->jpnij = jpnij;
    int *iimppt = calloc(jpnij, sizeof(int)), *ijmppt = calloc(jpnij, sizeof(int));
    int *ilci = calloc(jpnij, sizeof(int)), *ilcj = calloc(jpnij, sizeof(int));
    int *ibondi = calloc(jpnij, sizeof(int)), *ibondj = calloc(jpnij, sizeof(int)), *ipolj = calloc(jpnij, sizeof(int));
    int *ildi = calloc(jpnij, sizeof(int)), *ilei = calloc(jpnij, sizeof(int));
    int *ildj = calloc(jpnij, sizeof(int)), *ilej = calloc(jpnij, sizeof(int));
    int *iono = calloc(jpnij, sizeof(int)), *ioea = calloc(jpnij, sizeof(int));
    int *ioso = calloc(jpnij, sizeof(int)), *iowe = calloc(jpnij, sizeof(int));
    int *ipproc = calloc(jpnij, sizeof(int));
    int jarea, iarea0, ii, ij, ili, ilj, ijm1, imil, jproc;

    mpp_basic_decomposition(jpiglo, jpjglo, jperio, jpni, jpnj, &w->jpimax, &w->jpjmax, iimppt, ijmppt, ilci, ilcj); /* :325 */
    w->nfiimpp = calloc(jpnij, sizeof(int)); w->nfilcit = calloc(jpnij, sizeof(int)); w->nfipproc = calloc(jpnij, sizeof(int));
    memcpy(w->nfiimpp, iimppt, jpnij * sizeof(int));                            /* :326 */
    memcpy(w->nfilcit, ilci, jpnij * sizeof(int));                              /* :327 */

    const int l_Iperio = (jpni == 1) && (jperio == 1 || jperio == 4 || jperio == 6 || jperio == 7);   /* :345 */
    const int l_Jperio = (jpnj == 1) && (jperio == 2 || jperio == 7);                                 /* :346 */

    for (jarea = 1; jarea <= jpni * jpnj; ++jarea) {                            /* :348-405 */
        iarea0 = jarea - 1;
        ii = 1 + imod(iarea0, jpni);
        ij = 1 + iarea0 / jpni;
        ili = ilci[IX(ii, ij)];
        ilj = ilcj[IX(ii, ij)];
        ibondi[IX(ii, ij)] = 0;
        if (ii == 1)    ibondi[IX(ii, ij)] = -1;
        if (ii == jpni) ibondi[IX(ii, ij)] = 1;
        if (jpni == 1)  ibondi[IX(ii, ij)] = 2;
        ibondj[IX(ii, ij)] = 0;
        if (ij == 1)    ibondj[IX(ii, ij)] = -1;
        if (ij == jpnj) ibondj[IX(ii, ij)] = 1;
        if (jpnj == 1)  ibondj[IX(ii, ij)] = 2;
        ioso[IX(ii, ij)] = iarea0 - jpni;
        iowe[IX(ii, ij)] = iarea0 - 1;
        ioea[IX(ii, ij)] = iarea0 + 1;
        iono[IX(ii, ij)] = iarea0 + jpni;
        ildi[IX(ii, ij)] = 1 + NN_HLS;
        ilei[IX(ii, ij)] = ili - NN_HLS;
        ildj[IX(ii, ij)] = 1 + NN_HLS;
        ilej[IX(ii, ij)] = ilj - NN_HLS;
        if (jperio == 1 || jperio == 4 || jperio == 6 || jperio == 7) {         /* :374-379 */
            if (jpni != 1)  ibondi[IX(ii, ij)] = 0;
            if (ii == 1)    iowe[IX(ii, ij)] = iarea0 + (jpni - 1);
            if (ii == jpni) ioea[IX(ii, ij)] = iarea0 - (jpni - 1);
        }
        if (jperio == 2 || jperio == 7) {                                       /* :381-386 */
            if (jpnj != 1)  ibondj[IX(ii, ij)] = 0;
            if (ij == 1)    ioso[IX(ii, ij)] = iarea0 + jpni * (jpnj - 1);
            if (ij == jpnj) iono[IX(ii, ij)] = iarea0 - jpni * (jpnj - 1);
        }
        ipolj[IX(ii, ij)] = 0;                                                  /* :388-403 */
        if (jperio == 3 || jperio == 4) {
            ijm1 = jpni * (jpnj - 1);
            imil = ijm1 + (jpni + 1) / 2;
            if (jarea > ijm1) ipolj[IX(ii, ij)] = 3;
            if (imod(jpni, 2) == 1 && jarea == imil) ipolj[IX(ii, ij)] = 4;
            if (ipolj[IX(ii, ij)] == 3) iono[IX(ii, ij)] = jpni * jpnj - jarea + ijm1;
        }
        if (jperio == 5 || jperio == 6) {
            ijm1 = jpni * (jpnj - 1);
            imil = ijm1 + (jpni + 1) / 2;
            if (jarea > ijm1) ipolj[IX(ii, ij)] = 5;
            if (imod(jpni, 2) == 1 && jarea == imil) ipolj[IX(ii, ij)] = 6;
            if (ipolj[IX(ii, ij)] == 5) iono[IX(ii, ij)] = jpni * jpnj - jarea + ijm1;
        }
    }
    /* 4. land subdomains: all ocean  =>  ipproc = zone number  (:410-434) */
    for (jarea = 1; jarea <= jpnij; ++jarea) ipproc[jarea - 1] = jarea - 1;
    memcpy(w->nfipproc, ipproc, jpnij * sizeof(int));                           /* :434 */
    /* Update il[de][ij] according to ibond[ij]  (:480-487) */
    for (jproc = 1; jproc <= jpnij; ++jproc) {
        ii = 1 + imod(jproc - 1, jpni); ij = 1 + (jproc - 1) / jpni;
        if (ibondi[IX(ii, ij)] == -1 || ibondi[IX(ii, ij)] == 2) ildi[IX(ii, ij)] = 1;
        if (ibondi[IX(ii, ij)] ==  1 || ibondi[IX(ii, ij)] == 2) ilei[IX(ii, ij)] = ilci[IX(ii, ij)];
        if (ibondj[IX(ii, ij)] == -1 || ibondj[IX(ii, ij)] == 2) ildj[IX(ii, ij)] = 1;
        if (ibondj[IX(ii, ij)] ==  1 || ibondj[IX(ii, ij)] == 2) ilej[IX(ii, ij)] = ilcj[IX(ii, ij)];
    }

    w->dom = (oce_dom *)calloc(jpnij, sizeof(oce_dom));
    w->nimppt = calloc(jpnij, sizeof(int)); w->njmppt = calloc(jpnij, sizeof(int));
    w->nlcit = calloc(jpnij, sizeof(int));  w->nlcjt = calloc(jpnij, sizeof(int));
    w->nldit = calloc(jpnij, sizeof(int));  w->nleit = calloc(jpnij, sizeof(int));
    w->nldjt = calloc(jpnij, sizeof(int));  w->nlejt = calloc(jpnij, sizeof(int));
    w->ibonit = calloc(jpnij, sizeof(int)); w->ibonjt = calloc(jpnij, sizeof(int));
    for (jproc = 1; jproc <= jpnij; ++jproc) {                                  /* :581-593 */
        ii = 1 + imod(jproc - 1, jpni); ij = 1 + (jproc - 1) / jpni;
        w->nlcit[jproc - 1] = ilci[IX(ii, ij)]; w->nldit[jproc - 1] = ildi[IX(ii, ij)]; w->nleit[jproc - 1] = ilei[IX(ii, ij)];
        w->nlcjt[jproc - 1] = ilcj[IX(ii, ij)]; w->nldjt[jproc - 1] = ildj[IX(ii, ij)]; w->nlejt[jproc - 1] = ilej[IX(ii, ij)];
        w->ibonit[jproc - 1] = ibondi[IX(ii, ij)]; w->ibonjt[jproc - 1] = ibondj[IX(ii, ij)];
        w->nimppt[jproc - 1] = iimppt[IX(ii, ij)]; w->njmppt[jproc - 1] = ijmppt[IX(ii, ij)];
    }
    /* mpp_ini_north (lib_mpp.F90:1063-1082) */
    w->njmppmax = 0;
    for (jproc = 0; jproc < jpnij; ++jproc) if (w->njmppt[jproc] > w->njmppmax) w->njmppmax = w->njmppt[jproc];
    w->ndim_rank_north = 0;
    for (jproc = 0; jproc < jpnij; ++jproc) if (w->njmppt[jproc] == w->njmppmax) w->ndim_rank_north++;
    w->nrank_north = calloc(w->ndim_rank_north, sizeof(int));
    { int n = 0; for (jproc = 0; jproc < jpnij; ++jproc) if (w->njmppt[jproc] == w->njmppmax) w->nrank_north[n++] = jproc; }

    for (int narea = 1; narea <= jpnij; ++narea) {                              /* 6. (:548-580) per rank */
        oce_dom *d = &w->dom[narea - 1];
        ii = 1 + imod(narea - 1, jpni); ij = 1 + (narea - 1) / jpni;
        d->jpiglo = jpiglo; d->jpjglo = jpjglo; d->jperio = jperio;
        d->jpni = jpni; d->jpnj = jpnj; d->jpnij = jpnij; d->jpimax = w->jpimax; d->jpjmax = w->jpjmax;
        d->narea = narea; d->nproc = narea - 1;
        /* ii_noso etc. (:527-546): neighbour zone -> proc number, -1 when outside */
        d->noso = (0 <= ioso[IX(ii, ij)] && ioso[IX(ii, ij)] <= jpnij - 1) ? ipproc[ioso[IX(ii, ij)]] : -1;
        d->nowe = (0 <= iowe[IX(ii, ij)] && iowe[IX(ii, ij)] <= jpnij - 1) ? ipproc[iowe[IX(ii, ij)]] : -1;
        d->noea = (0 <= ioea[IX(ii, ij)] && ioea[IX(ii, ij)] <= jpnij - 1) ? ipproc[ioea[IX(ii, ij)]] : -1;
        d->nono = (0 <= iono[IX(ii, ij)] && iono[IX(ii, ij)] <= jpnij - 1) ? ipproc[iono[IX(ii, ij)]] : -1;
        d->nlci = ilci[IX(ii, ij)]; d->nldi = ildi[IX(ii, ij)]; d->nlei = ilei[IX(ii, ij)];
        d->nlcj = ilcj[IX(ii, ij)]; d->nldj = ildj[IX(ii, ij)]; d->nlej = ilej[IX(ii, ij)];
        d->nbondi = ibondi[IX(ii, ij)]; d->nbondj = ibondj[IX(ii, ij)];
        d->nimpp = iimppt[IX(ii, ij)]; d->njmpp = ijmppt[IX(ii, ij)];
        set_sizes(d, jpk);
        d->npolj = 0;                                                           /* :621-628 */
        if ((jperio == 3 || jperio == 4) && ij == jpnj) d->npolj = 3;
        if ((jperio == 5 || jperio == 6) && ij == jpnj) d->npolj = 5;
        d->l_Iperio = l_Iperio; d->l_Jperio = l_Jperio;
        d->nreci = d->nrecj = 2 * NN_HLS;
        d->ln_nnogather = ln_nnogather;
        d->world = w;
        /* mpp_init_nfdcom (:1180-1240) */
        d->nsndto = 0; for (int k = 0; k < JPMAXNGH; ++k) d->isendto[k] = 0;
        d->nfsloop = 1; d->nfeloop = d->nlci;
        if (jperio >= 3 && jperio <= 6 && jpni > 1 && d->njmpp == w->njmppmax) {
            int sxM = jpiglo - w->nimppt[narea - 1] - w->nlcit[narea - 1] + 1;
            int dxM = jpiglo - w->nimppt[narea - 1] + 2;
            for (int jn = 1; jn <= jpni; ++jn) {
                int sxT = w->nfiimpp[IX(jn, jpnj)];
                int dxT = w->nfiimpp[IX(jn, jpnj)] + w->nfilcit[IX(jn, jpnj)] - 1;
                int hit = 0;
                if (sxT < sxM && sxM < dxT) hit = 1;
                else if (sxM <= sxT && dxM >= dxT) hit = 1;
                else if (dxM < dxT && sxT < dxM) hit = 1;
                if (hit) {
                    /* the reference would overrun isendto(jpmaxngh) here (lbcnfd.F90:53-55): refuse the layout */
                    if (d->nsndto >= JPMAXNGH) { w->invalid = 1; continue; }
                    d->isendto[d->nsndto++] = jn;
                }
            }
        }
    }
    free(iimppt); free(ijmppt); free(ilci); free(ilcj); free(ibondi); free(ibondj); free(ipolj);
    free(ildi); free(ilei); free(ildj); free(ilej); free(iono); free(ioea); free(ioso); free(iowe); free(ipproc);
    if (w->invalid && ln_nnogather) { mpp_finalize(w); return NULL; }
    return w;
}

extern void oce_mail_free(oce_world *w);

void mpp_finalize(oce_world *w)
{
    if (!w) return;
    oce_mail_free(w);
    free(w->dom); free(w->nimppt); free(w->njmppt); free(w->nlcit); free(w->nlcjt); free(w->nldit); free(w->nleit);
    free(w->nldjt); free(w->nlejt); free(w->ibonit); free(w->ibonjt); free(w->nfiimpp); free(w->nfilcit);
    free(w->nfipproc); free(w->nrank_north);
    free(w);
}

int oce_world_size(const oce_world *w) { return w->jpnij; }
oce_dom *oce_world_dom(oce_world *w, int rank0) { return &w->dom[rank0]; }
