/*
 * lbclnk.c -- ORACLE (test infrastructure only; see nemo_oracle.h).
 * Loop-for-loop C restatement of the lateral-boundary / halo-exchange layer under tra_adv_fct:
 *   lbc_lnk_multi -> lbc_lnk_ptr dispatch     src/OCE/LBC/lbc_lnk_multi_generic.h90:16-54, lbclnk.F90:32-37,84-86
 *   lbc_lnk (no key_mpp_mpi)                  src/OCE/LBC/lbc_lnk_generic.h90:48-108
 *   mpp_lnk (key_mpp_mpi)                     src/OCE/LBC/mpp_lnk_generic.h90:48-338
 *   lbc_nfd                                   src/OCE/LBC/lbc_nfd_generic.h90:46-167
 *   mpp_nfd  (allgather and no-gather paths)  src/OCE/LBC/mpp_nfd_generic.h90:48-302
 *   lbc_nfd_nogather                          src/OCE/LBC/lbc_nfd_nogather_generic.h90:52-350
 *   mppsend / mpprecv                         src/OCE/LBC/lib_mpp.F90:483-536
 * MPI is emulated inside one process: each subdomain ("rank") runs on its own thread (oce_world_run) and
 * mppsend/mpprecv move copies of the buffers through mailboxes keyed by (destination, source, tag), i.e. the
 * two-sided semantics of mpi_isend + blocking mpi_recv with explicit source.
 * Fields are (jpi,jpj,ipk) fp64; ipl (4th dim) is folded into ipk by the callers (contiguous layout).
 */
#include "nemo_oracle.h"
#include <math.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define NN_HLS 1
#define NTAG 8

/* --------------------------------------------------------------------------------------------------------- */
/*  MPI emulation                                                                                            */
/* --------------------------------------------------------------------------------------------------------- */
typedef struct { double *buf; size_t n; int full; } slot_t;
typedef struct {
    pthread_mutex_t mtx; pthread_cond_t cv;
    slot_t *slot;                 /* [dest][src][tag] */
    int nrank;
    /* allgather / barrier on the north communicator */
    int bar_count, bar_gen;
    const double **gather_ptr;    /* [nrank] */
} mail_t;

static mail_t *mail_get(oce_world *w)
{
    if (!w->mail) {
        mail_t *m = (mail_t *)calloc(1, sizeof(mail_t));
        pthread_mutex_init(&m->mtx, NULL); pthread_cond_init(&m->cv, NULL);
        m->nrank = w->jpnij;
        m->slot = (slot_t *)calloc((size_t)w->jpnij * w->jpnij * NTAG, sizeof(slot_t));
        m->gather_ptr = (const double **)calloc(w->jpnij, sizeof(double *));
        w->mail = m;
    }
    return (mail_t *)w->mail;
}
void oce_mail_free(oce_world *w)
{
    if (!w->mail) return;
    mail_t *m = (mail_t *)w->mail;
    free(m->slot); free((void *)m->gather_ptr);
    pthread_mutex_destroy(&m->mtx); pthread_cond_destroy(&m->cv);
    free(m); w->mail = NULL;
}
static slot_t *slot_of(mail_t *m, int dest, int src, int tag) { return &m->slot[((size_t)dest * m->nrank + src) * NTAG + tag]; }

/* mppsend (lib_mpp.F90:483-508): non-blocking for the caller as with mpi_isend (the payload is copied) */
static void mppsend(oce_dom *d, int ktyp, const double *pmess, size_t kbytes, int kdest)
{
    mail_t *m = (mail_t *)d->world->mail;
    slot_t *s = slot_of(m, kdest, d->nproc, ktyp);
    pthread_mutex_lock(&m->mtx);
    while (s->full) pthread_cond_wait(&m->cv, &m->mtx);
    s->buf = (double *)malloc(kbytes * sizeof(double));
    memcpy(s->buf, pmess, kbytes * sizeof(double));
    s->n = kbytes; s->full = 1;
    pthread_cond_broadcast(&m->cv);
    pthread_mutex_unlock(&m->mtx);
}
/* mpprecv (lib_mpp.F90:512-536): blocking receive from an explicit source */
static void mpprecv(oce_dom *d, int ktyp, double *pmess, size_t kbytes, int ksource)
{
    mail_t *m = (mail_t *)d->world->mail;
    slot_t *s = slot_of(m, d->nproc, ksource, ktyp);
    pthread_mutex_lock(&m->mtx);
    while (!s->full) pthread_cond_wait(&m->cv, &m->mtx);
    if (s->n != kbytes) { fprintf(stderr, "oracle mpprecv: size mismatch (%zu vs %zu)\n", s->n, kbytes); abort(); }
    memcpy(pmess, s->buf, kbytes * sizeof(double));
    free(s->buf); s->buf = NULL; s->full = 0;
    pthread_cond_broadcast(&m->cv);
    pthread_mutex_unlock(&m->mtx);
}
static void north_barrier(oce_world *w)
{
    mail_t *m = (mail_t *)w->mail;
    pthread_mutex_lock(&m->mtx);
    int gen = m->bar_gen;
    if (++m->bar_count == w->ndim_rank_north) { m->bar_count = 0; m->bar_gen++; pthread_cond_broadcast(&m->cv); }
    else while (gen == m->bar_gen) pthread_cond_wait(&m->cv, &m->mtx);
    pthread_mutex_unlock(&m->mtx);
}

typedef struct { oce_world *w; void (*fn)(oce_dom *, void *); void *arg; int rank; } thr_arg;
static void *thr_main(void *p) { thr_arg *a = (thr_arg *)p; a->fn(&a->w->dom[a->rank], a->arg); return NULL; }

void oce_world_run(oce_world *w, void (*fn)(oce_dom *, void *), void *arg)
{
    if (w->jpnij == 1) { fn(&w->dom[0], arg); return; }
    mail_get(w);
    pthread_t *th = (pthread_t *)malloc(w->jpnij * sizeof(pthread_t));
    thr_arg *ta = (thr_arg *)malloc(w->jpnij * sizeof(thr_arg));
    for (int r = 0; r < w->jpnij; ++r) {
        ta[r].w = w; ta[r].fn = fn; ta[r].arg = arg; ta[r].rank = r;
        if (pthread_create(&th[r], NULL, thr_main, &ta[r])) { fprintf(stderr, "oracle: pthread_create failed\n"); abort(); }
    }
    for (int r = 0; r < w->jpnij; ++r) pthread_join(th[r], NULL);
    free(th); free(ta);
}

/* --------------------------------------------------------------------------------------------------------- */
/*  indexing helpers                                                                                         */
/* --------------------------------------------------------------------------------------------------------- */
/* element (ji,jj,jk) of a (n1,n2,*) array, 1-based */
#define A3(p, n1, n2, ji, jj, jk) (p)[(size_t)((jk) - 1) * (n1) * (n2) + (size_t)((jj) - 1) * (n1) + (size_t)((ji) - 1)]

/* --------------------------------------------------------------------------------------------------------- */
/*  lbc_nfd  (lbc_nfd_generic.h90:46-167).  ipi/ipjdim = array extents, ipj = folded row index.               */
/* --------------------------------------------------------------------------------------------------------- */
void lbc_nfd_generic(const oce_dom *d, int ipi, int ipjdim, int ipj, int nfld, double **ptab,
                     const char *cd_nat, const double *psgn, int ipk)
{
    const int jpiglo = d->jpiglo, npolj = d->npolj;
    const int ipjm1 = ipj - 1;
    int ji, jk, jf, ijt, iju;
#define P(i, j) A3(a, ipi, ipjdim, i, j, jk)
    for (jf = 0; jf < nfld; ++jf) {
        double *a = ptab[jf]; const double sg = psgn[jf]; const char nat = cd_nat[jf];
        for (jk = 1; jk <= ipk; ++jk) {
            switch (npolj) {
            case 3: case 4:                                                     /* T-point pivot (:75-113) */
                switch (nat) {
                case 'T': case 'W':
                    for (ji = 2; ji <= jpiglo; ++ji) { ijt = jpiglo - ji + 2; P(ji, ipj) = sg * P(ijt, ipj - 2); }
                    P(1, ipj) = sg * P(3, ipj - 2);
                    for (ji = jpiglo / 2 + 1; ji <= jpiglo; ++ji) { ijt = jpiglo - ji + 2; P(ji, ipjm1) = sg * P(ijt, ipjm1); }
                    break;
                case 'U':
                    for (ji = 1; ji <= jpiglo - 1; ++ji) { iju = jpiglo - ji + 1; P(ji, ipj) = sg * P(iju, ipj - 2); }
                    P(1, ipj) = sg * P(2, ipj - 2);
                    P(jpiglo, ipj) = sg * P(jpiglo - 1, ipj - 2);
                    for (ji = jpiglo / 2; ji <= jpiglo - 1; ++ji) { iju = jpiglo - ji + 1; P(ji, ipjm1) = sg * P(iju, ipjm1); }
                    break;
                case 'V':
                    for (ji = 2; ji <= jpiglo; ++ji) {
                        ijt = jpiglo - ji + 2;
                        P(ji, ipj - 1) = sg * P(ijt, ipj - 2);
                        P(ji, ipj)     = sg * P(ijt, ipj - 3);
                    }
                    P(1, ipj) = sg * P(3, ipj - 3);
                    break;
                case 'F':
                    for (ji = 1; ji <= jpiglo - 1; ++ji) {
                        iju = jpiglo - ji + 1;
                        P(ji, ipj - 1) = sg * P(iju, ipj - 2);
                        P(ji, ipj)     = sg * P(iju, ipj - 3);
                    }
                    P(1, ipj) = sg * P(2, ipj - 3);
                    P(jpiglo, ipj) = sg * P(jpiglo - 1, ipj - 3);
                    break;
                }
                break;
            case 5: case 6:                                                     /* F-point pivot (:115-150) */
                switch (nat) {
                case 'T': case 'W':
                    for (ji = 1; ji <= jpiglo; ++ji) { ijt = jpiglo - ji + 1; P(ji, ipj) = sg * P(ijt, ipj - 1); }
                    break;
                case 'U':
                    for (ji = 1; ji <= jpiglo - 1; ++ji) { iju = jpiglo - ji; P(ji, ipj) = sg * P(iju, ipj - 1); }
                    P(jpiglo, ipj) = sg * P(jpiglo - 2, ipj - 1);
                    break;
                case 'V':
                    for (ji = 1; ji <= jpiglo; ++ji) { ijt = jpiglo - ji + 1; P(ji, ipj) = sg * P(ijt, ipj - 2); }
                    for (ji = jpiglo / 2 + 1; ji <= jpiglo; ++ji) { ijt = jpiglo - ji + 1; P(ji, ipjm1) = sg * P(ijt, ipjm1); }
                    break;
                case 'F':
                    for (ji = 1; ji <= jpiglo - 1; ++ji) { iju = jpiglo - ji; P(ji, ipj) = sg * P(iju, ipj - 2); }
                    P(jpiglo, ipj) = sg * P(jpiglo - 2, ipj - 2);
                    for (ji = jpiglo / 2 + 1; ji <= jpiglo - 1; ++ji) { iju = jpiglo - ji; P(ji, ipjm1) = sg * P(iju, ipjm1); }
                    break;
                }
                break;
            default:                                                            /* closed (:152-160) */
                switch (nat) {
                case 'T': case 'U': case 'V': case 'W':
                    for (ji = 1; ji <= ipi; ++ji) { P(ji, 1) = 0.0; P(ji, ipj) = 0.0; }
                    break;
                case 'F':
                    for (ji = 1; ji <= ipi; ++ji) P(ji, ipj) = 0.0;
                    break;
                }
            }
        }
    }
#undef P
}

/* --------------------------------------------------------------------------------------------------------- */
/*  lbc_lnk without MPI  (lbc_lnk_generic.h90:48-108)                                                         */
/* --------------------------------------------------------------------------------------------------------- */
void lbc_lnk_generic(oce_dom *d, int nfld, double **ptab, const char *cd_nat, const double *psgn, int ipk,
                     int has_pval, double pval)
{
    const int jpi = d->jpi, jpj = d->jpj, jpim1 = d->jpim1, jpjm1 = d->jpjm1, jperio = d->jperio;
    const int ll_nfd = (jperio == 3 || jperio == 4 || jperio == 5 || jperio == 6);
    const double zland = has_pval ? pval : 0.0;
    int ji, jj, jk, jf;
#define P(i, j) A3(a, jpi, jpj, i, j, jk)
    for (jf = 0; jf < nfld; ++jf) {
        double *a = ptab[jf]; const char nat = cd_nat[jf];
        /* East-West boundaries (:85-90) */
        if (d->l_Iperio) {
            for (jk = 1; jk <= ipk; ++jk) for (jj = 1; jj <= jpj; ++jj) P(1, jj) = P(jpim1, jj);
            for (jk = 1; jk <= ipk; ++jk) for (jj = 1; jj <= jpj; ++jj) P(jpi, jj) = P(2, jj);
        } else {
            if (nat != 'F') for (jk = 1; jk <= ipk; ++jk) for (jj = 1; jj <= jpj; ++jj) P(1, jj) = zland;
            for (jk = 1; jk <= ipk; ++jk) for (jj = 1; jj <= jpj; ++jj) P(jpi, jj) = zland;
        }
        /* North-South boundaries (:92-102) */
        if (d->l_Jperio) {
            for (jk = 1; jk <= ipk; ++jk) for (ji = 1; ji <= jpi; ++ji) P(ji, 1) = P(ji, jpjm1);
            for (jk = 1; jk <= ipk; ++jk) for (ji = 1; ji <= jpi; ++ji) P(ji, jpj) = P(ji, 2);
        } else if (ll_nfd) {
            if (nat != 'F') for (jk = 1; jk <= ipk; ++jk) for (ji = 1; ji <= jpi; ++ji) P(ji, 1) = zland;
            /* CALL lbc_nfd( ptab, NAT_IN(:), SGN_IN(:) ): with a pointer list it treats ALL fields on every pass of
             * this jf loop (:97); the operation is idempotent only field by field, so restate it literally.        */
            lbc_nfd_generic(d, jpi, jpj, d->nlcj, nfld, ptab, cd_nat, psgn, ipk);
        } else {
            if (nat != 'F') for (jk = 1; jk <= ipk; ++jk) for (ji = 1; ji <= jpi; ++ji) P(ji, 1) = zland;
            for (jk = 1; jk <= ipk; ++jk) for (ji = 1; ji <= jpi; ++ji) P(ji, jpj) = zland;
        }
    }
#undef P
}

/* --------------------------------------------------------------------------------------------------------- */
/*  lbc_nfd_nogather  (lbc_nfd_nogather_generic.h90:52-350), one field.  ptab2 = ztabr(:,1:ipj2,:) with first  */
/*  extent n1b = jpimax*jpmaxngh and second extent 2 (the allocated ipj of mpp_nfd).                          */
/* --------------------------------------------------------------------------------------------------------- */
static void lbc_nfd_nogather(const oce_dom *d, double *a, const double *b, int n1b, int ipj2, char nat, double sg, int ipk)
{
    const oce_world *w = d->world;
    const int jpi = d->jpi, jpj = d->jpj, jpiglo = d->jpiglo, nimpp = d->nimpp, nlci = d->nlci, nlcj = d->nlcj;
    const int nf1 = w->nfiimpp[(size_t)(d->jpnj - 1) * d->jpni + (d->isendto[0] - 1)];   /* nfiimpp(isendto(1),jpnj) */
    const int ijpj = 1, ijpjp1 = 2;
    const int l_fast_exchanges = (ipj2 == 1);
    int ji, jk, ijt, iju, ijta, ijua, jia, startloop, endloop;
#define P(i, j)  A3(a, jpi, jpj, i, j, jk)
#define Q(i, j)  A3(b, n1b, 2, i, j, jk)
    switch (d->npolj) {
    case 3: case 4:
        switch (nat) {
        case 'T': case 'W':                                                     /* :101-139 */
            startloop = (nimpp != 1) ? 1 : 2;
            for (jk = 1; jk <= ipk; ++jk)
                for (ji = startloop; ji <= nlci; ++ji) { ijt = jpiglo - ji - nimpp - nf1 + 4; P(ji, nlcj) = sg * Q(ijt, ijpj); }
            if (nimpp == 1) for (jk = 1; jk <= ipk; ++jk) P(1, nlcj) = sg * P(3, nlcj - 2);
            if (!l_fast_exchanges) {
                if (nimpp >= jpiglo / 2 + 1) startloop = 1;
                else if (nimpp + nlci - 1 >= jpiglo / 2 + 1 && nimpp < jpiglo / 2 + 1) startloop = jpiglo / 2 + 1 - nimpp + 1;
                else startloop = nlci + 1;
                if (startloop <= nlci)
                    for (jk = 1; jk <= ipk; ++jk)
                        for (ji = startloop; ji <= nlci; ++ji) {
                            ijt = jpiglo - ji - nimpp - nf1 + 4;
                            jia = ji + nimpp - 1;
                            ijta = jpiglo - jia + 2;
                            if (ijta >= startloop + nimpp - 1 && ijta < jia) P(ji, nlcj - 1) = sg * P(ijta - nimpp + 1, nlcj - 1);
                            else                                             P(ji, nlcj - 1) = sg * Q(ijt, ijpjp1);
                        }
            }
            break;
        case 'U':                                                               /* :141-187 */
            endloop = (nimpp + nlci - 1 != jpiglo) ? nlci : nlci - 1;
            for (jk = 1; jk <= ipk; ++jk)
                for (ji = 1; ji <= endloop; ++ji) { iju = jpiglo - ji - nimpp - nf1 + 3; P(ji, nlcj) = sg * Q(iju, ijpj); }
            if (nimpp == 1) for (jk = 1; jk <= ipk; ++jk) P(1, nlcj) = sg * P(2, nlcj - 2);
            if (nimpp + nlci - 1 == jpiglo) for (jk = 1; jk <= ipk; ++jk) P(nlci, nlcj) = sg * P(nlci - 1, nlcj - 2);
            if (!l_fast_exchanges) {
                endloop = (nimpp + nlci - 1 != jpiglo) ? nlci : nlci - 1;
                if (nimpp >= jpiglo / 2) startloop = 1;
                else if (nimpp + nlci - 1 >= jpiglo / 2 && nimpp < jpiglo / 2) startloop = jpiglo / 2 - nimpp + 1;
                else startloop = endloop + 1;
                if (startloop <= endloop)
                    for (jk = 1; jk <= ipk; ++jk)
                        for (ji = startloop; ji <= endloop; ++ji) {
                            iju = jpiglo - ji - nimpp - nf1 + 3;
                            jia = ji + nimpp - 1;
                            ijua = jpiglo - jia + 1;
                            if (ijua >= startloop + nimpp - 1 && ijua < jia) P(ji, nlcj - 1) = sg * P(ijua - nimpp + 1, nlcj - 1);
                            else                                             P(ji, nlcj - 1) = sg * Q(iju, ijpjp1);
                        }
            }
            break;
        case 'V':                                                               /* :189-211 */
            startloop = (nimpp != 1) ? 1 : 2;
            if (!l_fast_exchanges)
                for (jk = 1; jk <= ipk; ++jk)
                    for (ji = startloop; ji <= nlci; ++ji) { ijt = jpiglo - ji - nimpp - nf1 + 4; P(ji, nlcj - 1) = sg * Q(ijt, ijpjp1); }
            for (jk = 1; jk <= ipk; ++jk)
                for (ji = startloop; ji <= nlci; ++ji) { ijt = jpiglo - ji - nimpp - nf1 + 4; P(ji, nlcj) = sg * Q(ijt, ijpj); }
            if (nimpp == 1) for (jk = 1; jk <= ipk; ++jk) P(1, nlcj) = sg * P(3, nlcj - 3);
            break;
        case 'F':                                                               /* :212-243 */
            endloop = (nimpp + nlci - 1 != jpiglo) ? nlci : nlci - 1;
            if (!l_fast_exchanges)
                for (jk = 1; jk <= ipk; ++jk)
                    for (ji = 1; ji <= endloop; ++ji) { iju = jpiglo - ji - nimpp - nf1 + 3; P(ji, nlcj - 1) = sg * Q(iju, ijpjp1); }
            for (jk = 1; jk <= ipk; ++jk)
                for (ji = 1; ji <= endloop; ++ji) { iju = jpiglo - ji - nimpp - nf1 + 3; P(ji, nlcj) = sg * Q(iju, ijpj); }
            if (nimpp == 1) for (jk = 1; jk <= ipk; ++jk) {
                P(1, nlcj) = sg * P(2, nlcj - 3);
                if (!l_fast_exchanges) P(1, nlcj - 1) = sg * P(2, nlcj - 2);
            }
            if (nimpp + nlci - 1 == jpiglo) for (jk = 1; jk <= ipk; ++jk) {
                P(nlci, nlcj) = sg * P(nlci - 1, nlcj - 3);
                if (!l_fast_exchanges) P(nlci, nlcj - 1) = sg * P(nlci - 1, nlcj - 2);
            }
            break;
        }
        break;
    case 5: case 6:
        switch (nat) {
        case 'T': case 'W':                                                     /* :250-256 */
            for (jk = 1; jk <= ipk; ++jk)
                for (ji = 1; ji <= nlci; ++ji) { ijt = jpiglo - ji - nimpp - nf1 + 3; P(ji, nlcj) = sg * Q(ijt, ijpj); }
            break;
        case 'U':                                                               /* :258-275 */
            endloop = (nimpp + nlci - 1 != jpiglo) ? nlci : nlci - 1;
            for (jk = 1; jk <= ipk; ++jk)
                for (ji = 1; ji <= endloop; ++ji) { iju = jpiglo - ji - nimpp - nf1 + 2; P(ji, nlcj) = sg * Q(iju, ijpj); }
            if (nimpp + nlci - 1 == jpiglo) for (jk = 1; jk <= ipk; ++jk) P(nlci, nlcj) = sg * P(nlci - 2, nlcj - 1);
            break;
        case 'V':                                                               /* :277-302 */
            for (jk = 1; jk <= ipk; ++jk)
                for (ji = 1; ji <= nlci; ++ji) { ijt = jpiglo - ji - nimpp - nf1 + 3; P(ji, nlcj) = sg * Q(ijt, ijpj); }
            if (!l_fast_exchanges) {
                if (nimpp >= jpiglo / 2 + 1) startloop = 1;
                else if (nimpp + nlci - 1 >= jpiglo / 2 + 1 && nimpp < jpiglo / 2 + 1) startloop = jpiglo / 2 + 1 - nimpp + 1;
                else startloop = nlci + 1;
                if (startloop <= nlci)
                    for (jk = 1; jk <= ipk; ++jk)
                        for (ji = startloop; ji <= nlci; ++ji) { ijt = jpiglo - ji - nimpp - nf1 + 3; P(ji, nlcj - 1) = sg * Q(ijt, ijpjp1); }
            }
            break;
        case 'F':                                                               /* :304-340 */
            endloop = (nimpp + nlci - 1 != jpiglo) ? nlci : nlci - 1;
            for (jk = 1; jk <= ipk; ++jk)
                for (ji = 1; ji <= endloop; ++ji) { iju = jpiglo - ji - nimpp - nf1 + 2; P(ji, nlcj) = sg * Q(iju, ijpj); }
            if (nimpp + nlci - 1 == jpiglo) for (jk = 1; jk <= ipk; ++jk) P(nlci, nlcj) = sg * P(nlci - 2, nlcj - 2);
            if (!l_fast_exchanges) {
                endloop = (nimpp + nlci - 1 != jpiglo) ? nlci : nlci - 1;
                if (nimpp >= jpiglo / 2 + 1) startloop = 1;
                else if (nimpp + nlci - 1 >= jpiglo / 2 + 1 && nimpp < jpiglo / 2 + 1) startloop = jpiglo / 2 + 1 - nimpp + 1;
                else startloop = endloop + 1;
                if (startloop <= endloop)
                    for (jk = 1; jk <= ipk; ++jk)
                        for (ji = startloop; ji <= endloop; ++ji) { iju = jpiglo - ji - nimpp - nf1 + 2; P(ji, nlcj - 1) = sg * Q(iju, ijpjp1); }
            }
            break;
        }
        break;
    default:
        fprintf(stderr, "lbc_nfd_nogather: npolj=%d\n", d->npolj); abort();
    }
#undef P
#undef Q
}

/* --------------------------------------------------------------------------------------------------------- */
/*  mpp_nfd  (mpp_nfd_generic.h90:48-302)                                                                     */
/* --------------------------------------------------------------------------------------------------------- */
static void mpp_nfd(oce_dom *d, int nfld, double **ptab, const char *cd_nat, const double *psgn, int ipk)
{
    oce_world *w = d->world;
    const int jpi = d->jpi, jpj = d->jpj, jpni = d->jpni, jpnj = d->jpnj, jpimax = d->jpimax, jpiglo = d->jpiglo;
    const int nlci = d->nlci, nlcj = d->nlcj, nimpp = d->nimpp, me = d->nproc;
    int ji, jj, jk, jf, jr, js, ij, iproc, iilb = 0, ilci = 0, ildi = 0, ilei = 0;
    (void)ilci;
#define NFIPPROC(i) w->nfipproc[(size_t)(jpnj - 1) * jpni + ((i) - 1)]
#define NFIIMPP(i)  w->nfiimpp[(size_t)(jpnj - 1) * jpni + ((i) - 1)]

    if (d->ln_nnogather) {                                                      /* l_north_nogather (:78-221) */
        int ipj = 2, ipf_j = 0;
        int *ipj_s = (int *)malloc(nfld * sizeof(int));
        int (*jj_s)[2] = (int (*)[2])malloc(nfld * sizeof(int[2]));
        for (jf = 0; jf < nfld; ++jf) ipj_s[jf] = 2;                            /* l_full_nf_update = .TRUE. (:95-102) */
        for (jf = 0; jf < nfld; ++jf) {                                         /* :105-133 */
            jj_s[jf][0] = jj_s[jf][1] = 0;
            switch (d->npolj) {
            case 3: case 4:
                switch (cd_nat[jf]) {
                case 'T': case 'W': case 'U': jj_s[jf][0] = nlcj - 2; jj_s[jf][1] = nlcj - 1; break;
                case 'V': case 'F':           jj_s[jf][0] = nlcj - 3; jj_s[jf][1] = nlcj - 2; break;
                }
                break;
            case 5: case 6:
                switch (cd_nat[jf]) {
                case 'T': case 'W': case 'U': jj_s[jf][0] = nlcj - 1; ipj_s[jf] = 1; break;
                case 'V': case 'F':           jj_s[jf][0] = nlcj - 2; jj_s[jf][1] = nlcj - 1; break;
                }
                break;
            }
        }
        for (jf = 0; jf < nfld; ++jf) ipf_j += ipj_s[jf];                       /* :135 */
        const size_t ibuffsize = (size_t)jpimax * ipf_j * ipk;
        double *znorthloc = (double *)calloc(ibuffsize, sizeof(double));        /* (jpimax,ipf_j,ipk) */
        double *zfoldwk   = (double *)calloc(ibuffsize, sizeof(double));
        const int n1b = jpimax * JPMAXNGH;
        double **ztabr = (double **)malloc(nfld * sizeof(double *));            /* (n1b,ipj,ipk) per field */
        for (jf = 0; jf < nfld; ++jf) {
            size_t n = (size_t)n1b * ipj * ipk;
            ztabr[jf] = (double *)malloc(n * sizeof(double));
            for (size_t q = 0; q < n; ++q) ztabr[jf][q] = NAN;                  /* undefined in the reference */
        }
        js = 0;                                                                 /* :139-149 */
        for (jf = 0; jf < nfld; ++jf)
            for (jj = 1; jj <= ipj_s[jf]; ++jj) {
                js++;
                for (jk = 1; jk <= ipk; ++jk)
                    for (ji = 1; ji <= jpi; ++ji)
                        A3(znorthloc, jpimax, ipf_j, ji, js, jk) = A3(ptab[jf], jpi, jpj, ji, jj_s[jf][jj - 1], jk);
            }
        for (jr = 1; jr <= d->nsndto; ++jr) {                                   /* :164-168 */
            int p = NFIPPROC(d->isendto[jr - 1]);
            if (p != me && p != -1) mppsend(d, 5, znorthloc, ibuffsize, p);
        }
        for (jr = 1; jr <= d->nsndto; ++jr) {                                   /* :170-205 */
            iproc = NFIPPROC(d->isendto[jr - 1]);
            if (iproc != -1) {
                iilb = w->nimppt[iproc]; ilci = w->nlcit[iproc]; ildi = w->nldit[iproc]; ilei = w->nleit[iproc];
                if (iilb == 1) ildi = 1;
                if (iilb + ilci - 1 == jpiglo) ilei = ilci;
                iilb = NFIIMPP(d->isendto[jr - 1]) - NFIIMPP(d->isendto[0]);
            }
            if (iproc != me && iproc != -1) {
                mpprecv(d, 5, zfoldwk, ibuffsize, iproc);
                js = 0;
                for (jf = 0; jf < nfld; ++jf)
                    for (jj = 1; jj <= ipj_s[jf]; ++jj) {
                        js++;
                        for (jk = 1; jk <= ipk; ++jk)
                            for (ji = ildi; ji <= ilei; ++ji)
                                A3(ztabr[jf], n1b, ipj, iilb + ji, jj, jk) = A3(zfoldwk, jpimax, ipf_j, ji, js, jk);
                    }
            } else if (iproc == me) {
                for (jf = 0; jf < nfld; ++jf)
                    for (jj = 1; jj <= ipj_s[jf]; ++jj)
                        for (jk = 1; jk <= ipk; ++jk)
                            for (ji = ildi; ji <= ilei; ++ji)
                                A3(ztabr[jf], n1b, ipj, iilb + ji, jj, jk) = A3(ptab[jf], jpi, jpj, ji, jj_s[jf][jj - 1], jk);
            }
        }
        for (jf = 0; jf < nfld; ++jf)                                           /* :217-219 */
            lbc_nfd_nogather(d, ptab[jf], ztabr[jf], n1b, ipj_s[jf], cd_nat[jf], psgn[jf], ipk);
        for (jf = 0; jf < nfld; ++jf) free(ztabr[jf]);
        free(ztabr); free(zfoldwk); free(znorthloc); free(ipj_s); free(jj_s);
    } else {                                                                    /* allgather path (:222-298) */
        const int ipj = 4;
        const size_t nloc = (size_t)jpimax * ipj * ipk * nfld;
        double *znorthloc = (double *)calloc(nloc, sizeof(double));            /* (jpimax,ipj,ipk,ipf) */
        for (jf = 0; jf < nfld; ++jf)                                           /* :228-237 */
            for (jk = 1; jk <= ipk; ++jk)
                for (jj = nlcj - ipj + 1; jj <= nlcj; ++jj) {
                    ij = jj - nlcj + ipj;
                    for (ji = 1; ji <= jpi; ++ji)
                        znorthloc[(((size_t)jf * ipk + (jk - 1)) * ipj + (ij - 1)) * jpimax + (ji - 1)] =
                            A3(ptab[jf], jpi, jpj, ji, jj, jk);
                }
        double **ztab = (double **)malloc(nfld * sizeof(double *));             /* (jpiglo,ipj,ipk) per field */
        for (jf = 0; jf < nfld; ++jf) {
            size_t n = (size_t)jpiglo * ipj * ipk;
            ztab[jf] = (double *)malloc(n * sizeof(double));
            for (size_t q = 0; q < n; ++q) ztab[jf][q] = NAN;
        }
        /* MPI_ALLGATHER on ncomm_north (:252-253) */
        mail_t *m = (mail_t *)w->mail;
        m->gather_ptr[me] = znorthloc;
        north_barrier(w);
        for (jr = 1; jr <= w->ndim_rank_north; ++jr) {                          /* :257-277 */
            iproc = w->nrank_north[jr - 1] + 1;
            iilb = w->nimppt[iproc - 1]; ilci = w->nlcit[iproc - 1]; ildi = w->nldit[iproc - 1]; ilei = w->nleit[iproc - 1];
            if (iilb == 1) ildi = 1;
            if (iilb + ilci - 1 == jpiglo) ilei = ilci;
            const double *src = m->gather_ptr[iproc - 1];
            for (jf = 0; jf < nfld; ++jf)
                for (jk = 1; jk <= ipk; ++jk)
                    for (jj = 1; jj <= ipj; ++jj)
                        for (ji = ildi; ji <= ilei; ++ji)
                            A3(ztab[jf], jpiglo, ipj, ji + iilb - 1, jj, jk) =
                                src[(((size_t)jf * ipk + (jk - 1)) * ipj + (jj - 1)) * jpimax + (ji - 1)];
        }
        north_barrier(w);                                                       /* senders' buffers no longer read */
        lbc_nfd_generic(d, jpiglo, ipj, ipj, nfld, ztab, cd_nat, psgn, ipk);    /* :278-280, ipj = 4 since jpni > 1 */
        for (jf = 0; jf < nfld; ++jf)                                           /* :282-293 */
            for (jk = 1; jk <= ipk; ++jk)
                for (jj = nlcj - ipj + 1; jj <= nlcj; ++jj) {
                    ij = jj - nlcj + ipj;
                    for (ji = 1; ji <= nlci; ++ji)
                        A3(ptab[jf], jpi, jpj, ji, jj, jk) = A3(ztab[jf], jpiglo, ipj, ji + nimpp - 1, ij, jk);
                }
        for (jf = 0; jf < nfld; ++jf) free(ztab[jf]);
        free(ztab); free(znorthloc);
    }
#undef NFIPPROC
#undef NFIIMPP
}

/* --------------------------------------------------------------------------------------------------------- */
/*  mpp_lnk  (mpp_lnk_generic.h90:48-338)                                                                     */
/* --------------------------------------------------------------------------------------------------------- */
void mpp_lnk_generic(oce_dom *d, int nfld, double **ptab, const char *cd_nat, const double *psgn, int ipk,
                     int has_pval, double pval)
{
    const int jpi = d->jpi, jpj = d->jpj, jpim1 = d->jpim1, jpjm1 = d->jpjm1;
    const int nlci = d->nlci, nlcj = d->nlcj, nreci = d->nreci, nrecj = d->nrecj;
    const double zland = has_pval ? pval : 0.0;
    int ji, jj, jk, jf, jh, iihom, ijhom;
    size_t imigr;
#define P(i, j) A3(a, jpi, jpj, i, j, jk)

    /* standard boundary treatment (:85-107) */
    for (jf = 0; jf < nfld; ++jf) {
        double *a = ptab[jf]; const char nat = cd_nat[jf];
        if (d->l_Iperio) {
            for (jk = 1; jk <= ipk; ++jk) for (jj = 1; jj <= jpj; ++jj) P(1, jj) = P(jpim1, jj);
            for (jk = 1; jk <= ipk; ++jk) for (jj = 1; jj <= jpj; ++jj) P(jpi, jj) = P(2, jj);
        } else {
            if (nat != 'F')
                for (jk = 1; jk <= ipk; ++jk) for (jj = 1; jj <= jpj; ++jj) for (ji = 1; ji <= NN_HLS; ++ji) P(ji, jj) = zland;
            for (jk = 1; jk <= ipk; ++jk) for (jj = 1; jj <= jpj; ++jj) for (ji = nlci - NN_HLS + 1; ji <= jpi; ++ji) P(ji, jj) = zland;
        }
        if (d->l_Jperio) {
            for (jk = 1; jk <= ipk; ++jk) for (ji = 1; ji <= jpi; ++ji) P(ji, 1) = P(ji, jpjm1);
            for (jk = 1; jk <= ipk; ++jk) for (ji = 1; ji <= jpi; ++ji) P(ji, jpj) = P(ji, 2);
        } else {
            if (nat != 'F')
                for (jk = 1; jk <= ipk; ++jk) for (jj = 1; jj <= NN_HLS; ++jj) for (ji = 1; ji <= jpi; ++ji) P(ji, jj) = zland;
            for (jk = 1; jk <= ipk; ++jk) for (jj = nlcj - NN_HLS + 1; jj <= jpj; ++jj) for (ji = 1; ji <= jpi; ++ji) P(ji, jj) = zland;
        }
    }

    /* East and west exchange (:109-215).  zt3ew/zt3we(jpj,nn_hls,ipk,ipl,ipf,{1,2}) */
    imigr = (size_t)NN_HLS * jpj * ipk * nfld;
    if (d->nbondi != 2) {
        double *zt3ew1 = (double *)malloc(imigr * sizeof(double)), *zt3we1 = (double *)malloc(imigr * sizeof(double));
        double *zt3ew2 = (double *)malloc(imigr * sizeof(double)), *zt3we2 = (double *)malloc(imigr * sizeof(double));
#define B(buf, jj_, jh_, jk_, jf_) (buf)[(((size_t)(jf_) * ipk + ((jk_) - 1)) * NN_HLS + ((jh_) - 1)) * jpj + ((jj_) - 1)]
        iihom = nlci - nreci;
        for (jf = 0; jf < nfld; ++jf) {
            double *a = ptab[jf];
            for (jk = 1; jk <= ipk; ++jk)
                for (jh = 1; jh <= NN_HLS; ++jh)
                    for (jj = 1; jj <= jpj; ++jj) {
                        if (d->nbondi == 0 || d->nbondi == 1)  B(zt3ew1, jj, jh, jk, jf) = P(NN_HLS + jh, jj);
                        if (d->nbondi == 0 || d->nbondi == -1) B(zt3we1, jj, jh, jk, jf) = P(iihom + jh, jj);
                    }
        }
        switch (d->nbondi) {                                                    /* :158-174 */
        case -1:
            mppsend(d, 2, zt3we1, imigr, d->noea);
            mpprecv(d, 1, zt3ew1, imigr, d->noea);
            break;
        case 0:
            mppsend(d, 1, zt3ew1, imigr, d->nowe);
            mppsend(d, 2, zt3we1, imigr, d->noea);
            mpprecv(d, 1, zt3ew2, imigr, d->noea);
            mpprecv(d, 2, zt3we2, imigr, d->nowe);
            break;
        case 1:
            mppsend(d, 1, zt3ew1, imigr, d->nowe);
            mpprecv(d, 2, zt3we1, imigr, d->nowe);
            break;
        }
        iihom = nlci - NN_HLS;                                                  /* :179-213 */
        for (jf = 0; jf < nfld; ++jf) {
            double *a = ptab[jf];
            for (jk = 1; jk <= ipk; ++jk)
                for (jh = 1; jh <= NN_HLS; ++jh)
                    for (jj = 1; jj <= jpj; ++jj) {
                        switch (d->nbondi) {
                        case -1: P(iihom + jh, jj) = B(zt3ew1, jj, jh, jk, jf); break;
                        case 0:  P(jh, jj) = B(zt3we2, jj, jh, jk, jf); P(iihom + jh, jj) = B(zt3ew2, jj, jh, jk, jf); break;
                        case 1:  P(jh, jj) = B(zt3we1, jj, jh, jk, jf); break;
                        }
                    }
        }
#undef B
        free(zt3ew1); free(zt3we1); free(zt3ew2); free(zt3we2);
    }

    /* 3. north fold treatment (:217-228) */
    if (d->npolj != 0) {
        if (d->jpni == 1) lbc_nfd_generic(d, jpi, jpj, nlcj, nfld, ptab, cd_nat, psgn, ipk);
        else              mpp_nfd(d, nfld, ptab, cd_nat, psgn, ipk);
    }

    /* 4. North and south directions (:230-336).  zt3ns/zt3sn(jpi,nn_hls,ipk,ipl,ipf,{1,2}) */
    imigr = (size_t)NN_HLS * jpi * ipk * nfld;
    if (d->nbondj != 2) {
        double *zt3ns1 = (double *)malloc(imigr * sizeof(double)), *zt3sn1 = (double *)malloc(imigr * sizeof(double));
        double *zt3ns2 = (double *)malloc(imigr * sizeof(double)), *zt3sn2 = (double *)malloc(imigr * sizeof(double));
#define B(buf, ji_, jh_, jk_, jf_) (buf)[(((size_t)(jf_) * ipk + ((jk_) - 1)) * NN_HLS + ((jh_) - 1)) * jpi + ((ji_) - 1)]
        ijhom = nlcj - nrecj;
        for (jf = 0; jf < nfld; ++jf) {
            double *a = ptab[jf];
            for (jk = 1; jk <= ipk; ++jk)
                for (jh = 1; jh <= NN_HLS; ++jh)
                    for (ji = 1; ji <= jpi; ++ji) {
                        if (d->nbondj == 0 || d->nbondj == -1) B(zt3sn1, ji, jh, jk, jf) = P(ji, ijhom + jh);
                        if (d->nbondj == 0 || d->nbondj == 1)  B(zt3ns1, ji, jh, jk, jf) = P(ji, NN_HLS + jh);
                    }
        }
        switch (d->nbondj) {                                                    /* :279-295 */
        case -1:
            mppsend(d, 4, zt3sn1, imigr, d->nono);
            mpprecv(d, 3, zt3ns1, imigr, d->nono);
            break;
        case 0:
            mppsend(d, 3, zt3ns1, imigr, d->noso);
            mppsend(d, 4, zt3sn1, imigr, d->nono);
            mpprecv(d, 3, zt3ns2, imigr, d->nono);
            mpprecv(d, 4, zt3sn2, imigr, d->noso);
            break;
        case 1:
            mppsend(d, 3, zt3ns1, imigr, d->noso);
            mpprecv(d, 4, zt3sn1, imigr, d->noso);
            break;
        }
        ijhom = nlcj - NN_HLS;                                                  /* :299-334 */
        for (jf = 0; jf < nfld; ++jf) {
            double *a = ptab[jf];
            for (jk = 1; jk <= ipk; ++jk)
                for (jh = 1; jh <= NN_HLS; ++jh)
                    for (ji = 1; ji <= jpi; ++ji) {
                        switch (d->nbondj) {
                        case -1: P(ji, ijhom + jh) = B(zt3ns1, ji, jh, jk, jf); break;
                        case 0:  P(ji, jh) = B(zt3sn2, ji, jh, jk, jf); P(ji, ijhom + jh) = B(zt3ns2, ji, jh, jk, jf); break;
                        case 1:  P(ji, jh) = B(zt3sn1, ji, jh, jk, jf); break;
                        }
                    }
        }
#undef B
        free(zt3ns1); free(zt3sn1); free(zt3ns2); free(zt3sn2);
    }
#undef P
}

/* lbc_lnk_multi -> lbc_lnk_ptr (lbc_lnk_multi_generic.h90:52; lbclnk.F90:32-37 with MPI, :84-86 without) */
void lbc_lnk_multi(oce_dom *d, const char *cdname, int nfld, double **ptab, const char *cd_nat,
                   const double *psgn, int ipk, int has_pval, double pval)
{
    (void)cdname;
    if (d->world) mpp_lnk_generic(d, nfld, ptab, cd_nat, psgn, ipk, has_pval, pval);
    else          lbc_lnk_generic(d, nfld, ptab, cd_nat, psgn, ipk, has_pval, pval);
}

typedef struct { int nfld; double ***ptabs; const char *cd_nat; const double *psgn; int ipk; } lnk_arg;
static void lnk_thr(oce_dom *d, void *p)
{
    lnk_arg *a = (lnk_arg *)p;
    lbc_lnk_multi(d, "oracle", a->nfld, a->ptabs[d->nproc], a->cd_nat, a->psgn, a->ipk, 0, 0.0);
}
void oce_world_lbc_lnk(oce_world *w, int nfld, double ***ptabs, const char *cd_nat, const double *psgn, int ipk)
{
    lnk_arg a = { nfld, ptabs, cd_nat, psgn, ipk };
    oce_world_run(w, lnk_thr, &a);
}
