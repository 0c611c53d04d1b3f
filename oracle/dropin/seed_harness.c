/*
 * oracle/dropin/seed_harness.c -- TEST INFRASTRUCTURE.
 *
 * Plain-buffer entry points around the reference's OWN single-end seeding + locate
 * (alnse_seed_overlap, alnse.c:199-312; alnse_locate_alt, alnse.c:633-731) and its own index loader
 * (alnse_index_reload, indexio.c:23-50), compiled with the reference's unmodified objects into
 * oracle/_ref/libsaltref_seed.so.  tests/ use it as the oracle of row f1: the device must emit the identical
 * sorted candidate lists.  Nothing of the reference is copied here; this file only calls its functions.
 */
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "aln.h"          /* pulls in indexio.h (no include guard there) */

void alnse_seed_overlap(index_t *index, uint32_t l_seq, const uint8_t *seq, aln_opt_t *opt, aux_t *aux_data);
void alnse_locate_alt(index_t *index, uint32_t l_seq, uint32_t max_locate, aux_t *aux_data);
void alnse_locate(index_t *index, uint32_t l_seq, uint32_t max_locate, aux_t *aux_data);

/* 0: alnse_locate_alt (what alnse_overlap_alt, the single-end program, calls); 1: alnse_locate (alnse_overlap, paired-end) */
static int g_locate_mode;
void seedref_set_locate_mode(int m) { g_locate_mode = m; }

void *seedref_open(const char *prefix) { return alnse_index_reload(prefix); }
void seedref_close(void *ix) { if (ix) alnse_index_destroy((index_t *)ix); }

/* the in-memory index as the reference's loaders left it: what salt_fm_index_t (include/salt_b200.h) describes */
typedef struct {
    const uint32_t *c_bwt; size_t c_bwt_words;
    uint32_t c_primary, c_seq_len, c_L2[5];
    const uint32_t *c_sa; uint32_t c_n_sa, c_sa_intv;
    const uint32_t *lkt; uint32_t lkt_len;
    const uint32_t *r_bwt; size_t r_bwt_words;
    const uint32_t *r_occ; size_t r_occ_words;
    const uint32_t *r_occ_major; size_t r_occ_major_words;
    const uint32_t *r_sa_sharp; size_t r_n_sa_sharp;
    uint32_t r_cum[6], r_inv_sa0, r_text_len;
} seedref_view_t;

void seedref_view(void *p, seedref_view_t *v)
{
    index_t *ix = (index_t *)p;
    int i;
    memset(v, 0, sizeof *v);
    v->c_bwt = ix->cbwt->bwt; v->c_bwt_words = ix->cbwt->bwt_size;
    v->c_primary = ix->cbwt->primary; v->c_seq_len = ix->cbwt->seq_len;
    for (i = 0; i < 5; ++i) v->c_L2[i] = ix->cbwt->L2[i];
    v->c_sa = ix->cbwt->sa; v->c_n_sa = ix->cbwt->n_sa; v->c_sa_intv = (uint32_t)ix->cbwt->sa_intv;
    v->lkt = ix->lkt->item; v->lkt_len = ix->lkt->maxLookupLen;
    rbwt_t *r = ix->rbwt2->rbwt1;
    v->r_bwt = r->bwtCode; v->r_bwt_words = r->bwtSizeInWord;
    v->r_occ = r->occValue; v->r_occ_words = r->occSizeInWord;
    v->r_occ_major = r->occValueMajor; v->r_occ_major_words = r->occMajorSizeInWord;
    v->r_sa_sharp = r->saValueSharp; v->r_n_sa_sharp = r->saValueSizeSharp;
    for (i = 0; i < 6; ++i) v->r_cum[i] = r->cumulativeFreq[i];
    v->r_inv_sa0 = r->inverseSa0; v->r_text_len = r->textLength;
}

uint32_t seedref_mixref_len(void *p) { return ((index_t *)p)->mixRef->l; }
const uint32_t *seedref_mixref(void *p) { return ((index_t *)p)->mixRef->seq; }

/* Both strands of every read: offs0/offs1 (n_reads + 1) and the lists, in the order aux->loci holds them after
 * alnse_locate_alt.  Returns 0, or -1 when a list buffer is too small. */
int seedref_run(void *p, const uint8_t *codes, const uint32_t *roffs, uint32_t n_reads, int l_seed, int l_overlap,
                int max_seed, int max_locate, int seed_only_ref, uint32_t *offs0, uint32_t *loci0, size_t cap0,
                uint32_t *offs1, uint32_t *loci1, size_t cap1)
{
    index_t *ix = (index_t *)p;
    aln_opt_t opt;
    memset(&opt, 0, sizeof opt);
    opt.l_seed = l_seed; opt.l_overlap = l_overlap; opt.max_seed = max_seed; opt.max_locate = (uint32_t)max_locate;
    opt.seed_only_ref = seed_only_ref;
    uint32_t r, l_max = 1;
    for (r = 0; r < n_reads; ++r) if (roffs[r + 1] - roffs[r] > l_max) l_max = roffs[r + 1] - roffs[r];
    aux_t *aux = aux_init((int)l_max + l_seed, l_seed);
    uint8_t *rseq = malloc(l_max + 1);
    size_t n0 = 0, n1 = 0;
    int rc = 0;
    offs0[0] = offs1[0] = 0;
    for (r = 0; r < n_reads && rc == 0; ++r) {
        const uint8_t *seq = codes + roffs[r];
        const uint32_t L = roffs[r + 1] - roffs[r];
        uint32_t i;
        int s;
        for (i = 0; i < L; ++i) { uint8_t c = seq[L - 1 - i]; rseq[i] = c < 4 ? 3 - c : c; }      /* query.c:46-64 */
        for (s = 0; s < 2; ++s) {
            aux_reset(aux);
            if ((int)L >= l_seed) {
                alnse_seed_overlap(ix, L, s ? rseq : seq, &opt, aux);
                if (g_locate_mode) alnse_locate(ix, L, (uint32_t)max_locate, aux); else alnse_locate_alt(ix, L, (uint32_t)max_locate, aux);
            }
            size_t *n = s ? &n1 : &n0;
            uint32_t *dst = s ? loci1 : loci0;
            const size_t cap = s ? cap1 : cap0;
            if (*n + aux->loci.n > cap) { rc = -1; break; }
            memcpy(dst + *n, aux->loci.a, aux->loci.n * 4);
            *n += aux->loci.n;
            (s ? offs1 : offs0)[r + 1] = (uint32_t)*n;
        }
    }
    free(rseq);
    aux_destroy(aux);
    return rc;
}

/* The same on n_threads pthreads (contiguous blocks of reads, one aux_t per thread, as alnse_core_thread keeps one per
 * worker): the CPU side of the seeding throughput comparison.  Returns the wall time in seconds through *sec. */
typedef struct {
    index_t *ix; const uint8_t *codes; const uint32_t *roffs; uint32_t first, upto; aln_opt_t opt; uint32_t l_max;
    uint32_t *cnt0, *cnt1;            /* per read list lengths (shared arrays, disjoint ranges) */
    uint32_t *buf[2]; size_t n[2], cap[2];
} seed_job_t;

static void *seed_worker(void *arg)
{
    seed_job_t *J = (seed_job_t *)arg;
    aux_t *aux = aux_init((int)J->l_max + J->opt.l_seed, J->opt.l_seed);
    uint8_t *rseq = malloc(J->l_max + 1);
    uint32_t r;
    for (r = J->first; r < J->upto; ++r) {
        const uint8_t *seq = J->codes + J->roffs[r];
        const uint32_t L = J->roffs[r + 1] - J->roffs[r];
        uint32_t i; int s;
        for (i = 0; i < L; ++i) { uint8_t c = seq[L - 1 - i]; rseq[i] = c < 4 ? 3 - c : c; }
        for (s = 0; s < 2; ++s) {
            aux_reset(aux);
            if ((int)L >= J->opt.l_seed) {
                alnse_seed_overlap(J->ix, L, s ? rseq : seq, &J->opt, aux);
                if (g_locate_mode) alnse_locate(J->ix, L, J->opt.max_locate, aux); else alnse_locate_alt(J->ix, L, J->opt.max_locate, aux);
            }
            if (J->n[s] + aux->loci.n > J->cap[s]) {
                J->cap[s] = (J->n[s] + aux->loci.n) * 2 + 1024;
                J->buf[s] = realloc(J->buf[s], J->cap[s] * 4);
            }
            memcpy(J->buf[s] + J->n[s], aux->loci.a, aux->loci.n * 4);
            J->n[s] += aux->loci.n;
            (s ? J->cnt1 : J->cnt0)[r] = (uint32_t)aux->loci.n;
        }
    }
    free(rseq);
    aux_destroy(aux);
    return NULL;
}

int seedref_run_mt(void *p, const uint8_t *codes, const uint32_t *roffs, uint32_t n_reads, int l_seed, int l_overlap,
                   int max_seed, int max_locate, int seed_only_ref, int n_threads, uint32_t *offs0, uint32_t *loci0, size_t cap0,
                   uint32_t *offs1, uint32_t *loci1, size_t cap1, double *sec)
{
    if (n_threads < 1) n_threads = 1;
    seed_job_t *J = calloc((size_t)n_threads, sizeof *J);
    pthread_t *th = calloc((size_t)n_threads, sizeof *th);
    uint32_t *cnt0 = calloc((size_t)n_reads + 1, 4), *cnt1 = calloc((size_t)n_reads + 1, 4);
    uint32_t r, l_max = 1;
    int t, rc = 0;
    for (r = 0; r < n_reads; ++r) if (roffs[r + 1] - roffs[r] > l_max) l_max = roffs[r + 1] - roffs[r];
    struct timespec t0, t1;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    for (t = 0; t < n_threads; ++t) {
        memset(&J[t].opt, 0, sizeof J[t].opt);
        J[t].opt.l_seed = l_seed; J[t].opt.l_overlap = l_overlap; J[t].opt.max_seed = max_seed;
        J[t].opt.max_locate = (uint32_t)max_locate; J[t].opt.seed_only_ref = seed_only_ref;
        J[t].ix = (index_t *)p; J[t].codes = codes; J[t].roffs = roffs; J[t].l_max = l_max; J[t].cnt0 = cnt0; J[t].cnt1 = cnt1;
        J[t].first = (uint32_t)((uint64_t)n_reads * (uint64_t)t / (uint64_t)n_threads);
        J[t].upto = (uint32_t)((uint64_t)n_reads * (uint64_t)(t + 1) / (uint64_t)n_threads);
        pthread_create(&th[t], NULL, seed_worker, &J[t]);
    }
    for (t = 0; t < n_threads; ++t) pthread_join(th[t], NULL);
    clock_gettime(CLOCK_MONOTONIC, &t1);
    if (sec) *sec = (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
    size_t n0 = 0, n1 = 0;
    for (t = 0; t < n_threads; ++t) {
        if (n0 + J[t].n[0] > cap0 || n1 + J[t].n[1] > cap1) rc = -1;
        else {
            if (J[t].n[0]) memcpy(loci0 + n0, J[t].buf[0], J[t].n[0] * 4);
            if (J[t].n[1]) memcpy(loci1 + n1, J[t].buf[1], J[t].n[1] * 4);
        }
        n0 += J[t].n[0]; n1 += J[t].n[1];
        free(J[t].buf[0]); free(J[t].buf[1]);
    }
    offs0[0] = offs1[0] = 0;
    for (r = 0; r < n_reads; ++r) { offs0[r + 1] = offs0[r] + cnt0[r]; offs1[r + 1] = offs1[r] + cnt1[r]; }
    free(cnt0); free(cnt1); free(J); free(th);
    return rc;
}

/* The reference's own FASTQ reader (query_open / query_read_seq, query.c:66-239): codes, lengths, ambiguity counts, names,
 * comments and quality strings of up to max_reads records -- the oracle of the input side (row f4). */
int seedref_read_fastq(const char *fn, uint32_t max_reads, uint8_t *codes, size_t codes_cap, uint32_t *roffs, uint16_t *n_amb,
                       char *names, char *comments, char *quals, int text_stride)
{
    queryio_t *qs = query_open(fn);
    query_t q;
    uint32_t n = 0;
    size_t at = 0;
    roffs[0] = 0;
    memset(&q, 0, sizeof q);
    while (n < max_reads && query_read_seq(qs, &q) > 0) {
        if (at + (size_t)q.l_seq > codes_cap) break;
        memcpy(codes + at, q.seq, (size_t)q.l_seq);
        at += (size_t)q.l_seq;
        roffs[n + 1] = (uint32_t)at;
        n_amb[n] = (uint16_t)q.n_ambiguous;
        strncpy(names + (size_t)n * text_stride, q.name ? q.name : "", (size_t)text_stride - 1);
        strncpy(comments + (size_t)n * text_stride, q.comment ? q.comment : "", (size_t)text_stride - 1);
        strncpy(quals + (size_t)n * text_stride, q.qual ? (const char *)q.qual : "", (size_t)text_stride - 1);
        query_destroy(&q);
        memset(&q, 0, sizeof q);
        ++n;
    }
    query_close(qs);
    return (int)n;
}
