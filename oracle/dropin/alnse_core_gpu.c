/*
 * oracle/dropin/alnse_core_gpu.c -- TEST INFRASTRUCTURE / integration proof.
 *
 * A replacement for ONE function of the reference, alnse_core (alnse.c:1353-1480), written against
 * the reference's own headers and linked with the reference's own unmodified objects into
 * oracle/_ref/salt_dropin (see oracle/Makefile; alnse.c is compiled with
 * -Dalnse_core=alnse_core_reference so that aln.c's aln_main reaches this one instead).
 * Everything outside the verification stage stays the reference's code: option parsing, index
 * loading, FASTQ reading, seeding + locate (alnse_seed_overlap / alnse_locate_alt), SAM formatting
 * (aln_samse with its XA / MD / NM tags).  The verification stage -- alnse_check_nogap,
 * alnse_check_withgap, query_set_hits, query_gen_cigar -- runs on the GPU through
 * include/salt_host.h exactly as INTEGRATION.md describes.  tests/test_dropin.py runs the
 * reference binary and this one on the same index and reads and requires identical SAM.
 *
 * Nothing from the reference is copied here: this file only CALLS its functions.
 */
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "kvec.h"
#include "aln.h"
#include "sam.h"
#include "salt_host.h"

#define DROPIN_MAX_N_PERSEQ 200          /* alnse.c:1281 */
#define DROPIN_CHUNK_READS 20000u
#define DROPIN_CHUNK_CANDS (20000u * 256u)

static void die(const char *what)
{
    fprintf(stderr, "[salt_dropin] %s: %s\n", what, salt_b200_last_error());
    exit(1);
}

void dropin_tail_prepare(salt_b200_t *gpu, int slot, const query_t *multi_seqs, const int *slot_of, int first, int upto);
void dropin_tail_report(void);
void dropin_tail_begin_read(int j);
int dropin_tail_get(int j, const char **md, unsigned *nm, const uint16_t **xv, int *n_xv);
const char *dropin_xa_row(int j, int k);

/* alnse_core1's per-read work after verification: results -> query_t (alnse.c:1306 / :1342-1344); the SAM line
 * follows once the chunk's MD/NM/XV tags are back from the GPU */
static void finish_read(index_t *index, query_t *query, const aln_opt_t *aln_opt, const salt_chunk_t *ck, uint32_t i)
{
    salt_read_result_t r;
    int s, j;
    if (salt_chunk_result(ck, i, aln_opt->max_hits, &r) != SALT_OK) die("salt_chunk_result");
    query->pos = r.pos; query->strand = r.strand; query->n_diff = r.n_diff; query->is_gap = r.is_gap;
    query->b0 = r.b0; query->b1 = r.b1; query->mapq = (uint8_t)r.mapq;
    for (s = 0; s < 2; ++s)
        for (j = 0; j < r.n_alt[s]; ++j) {
            hit_t h;
            h.pos = r.alt[s][j].pos; h.n_diff = r.alt[s][j].n_diff; h.is_gap = r.alt[s][j].is_gap; h.strand = r.alt[s][j].strand;
            kv_push(hit_t, query->hits[s], h);
        }
    /* query_gen_cigar (query.c:282-295) */
    query->seq_start = 0; query->seq_end = query->l_seq - 1;
    if (query->pos != 0xFFFFFFFF) {
        if (query->is_gap) strncpy(query->cigar->s, r.cigar, query->cigar->m - 1);
        else ksprintf(query->cigar, "%dM", query->l_seq);
    }
}

/* ---- seeding on the host: the reference's own alnse_seed_overlap / alnse_locate_alt on -t worker threads, as
 * alnse_core_thread runs them (alnse.c:1282-1310, :1423-1428); every worker keeps its own aux_t pair ------------- */
typedef struct { uint32_t *a[2]; uint32_t n[2], m[2]; int skip; } seeded_read_t;
typedef struct {
    int tid, n_threads, n; index_t *index; query_t *queries; aln_opt_t *aln_opt; seeded_read_t *out; aux_t *aux[2];
} seed_thread_t;

static void *seed_worker(void *arg)
{
    seed_thread_t *T = (seed_thread_t *)arg;
    int i, s;
    for (i = T->tid; i < T->n; i += T->n_threads) {              /* static interleave, as alnse_core1 (alnse.c:1321) */
        query_t *query = T->queries + i;
        seeded_read_t *o = T->out + i;
        o->n[0] = o->n[1] = 0;
        o->skip = query->n_ambiguous > DROPIN_MAX_N_PERSEQ;      /* alnse.c:1296 / :1328 */
        if (o->skip) continue;
        if (query->l_seq - T->aln_opt->l_seed + 1 > T->aux[0]->n_sai_range) {
            int n_sai_range = query->l_seq - T->aln_opt->l_seed + 1;
            aux_resize(T->aux[0], n_sai_range);
            aux_resize(T->aux[1], n_sai_range);
        }
        aux_reset(T->aux[0]);
        aux_reset(T->aux[1]);
        /* the first half of alnse_overlap_alt (alnse.c:1065-1068) */
        alnse_seed_overlap(T->index, query->l_seq, query->seq, T->aln_opt, T->aux[0]);
        alnse_locate_alt(T->index, query->l_seq, T->aln_opt->max_locate, T->aux[0]);
        alnse_seed_overlap(T->index, query->l_seq, query->rseq, T->aln_opt, T->aux[1]);
        alnse_locate_alt(T->index, query->l_seq, T->aln_opt->max_locate, T->aux[1]);
        for (s = 0; s < 2; ++s) {
            const uint32_t k = (uint32_t)T->aux[s]->loci.n;
            if (k > o->m[s]) { o->m[s] = k + 16; o->a[s] = realloc(o->a[s], (size_t)o->m[s] * 4); }
            memcpy(o->a[s], T->aux[s]->loci.a, (size_t)k * 4);
            o->n[s] = k;
        }
    }
    return NULL;
}

/* hit selection and SAM text of a sub-chunk on the -t workers, as alnse_core_thread does both per read (alnse.c:1306-1307) */
typedef struct { int tid, n_threads, first, upto, phase; index_t *index; query_t *queries; aln_opt_t *aln_opt; const salt_chunk_t *ck; const int *slot_of; char **old_sam; } fin_thread_t;

/* SALT_DROPIN_SAM=native: the line is written by the host layer's own formatter (salt_sam_se, byte-identical to aln_samse:
 * tests/test_sam_text.py) from the same query_t fields and the tags / XA CIGARs the GPU prepared */
static int g_native_sam;
static salt_sam_refs_t g_sam_refs;
static void native_samse(int j, query_t *query, const aln_opt_t *aln_opt)
{
    salt_sam_read_t r;
    const char *xa[2 * SALT_MAX_HITS + 2];
    int s, i, k = 0;
    memset(&r, 0, sizeof r);
    r.name = query->name; r.seq = query->seq; r.qual = (const char *)query->qual; r.l_seq = (uint32_t)query->l_seq;
    r.pos = query->pos; r.strand = (uint8_t)query->strand; r.mapq = query->mapq; r.cigar = query->cigar->s;
    r.seq_start = query->seq_start; r.seq_end = query->seq_end;
    for (s = 0; s < 2; ++s) {
        r.n_alt[s] = (int)query->hits[s].n; r.alt[s] = (const salt_hit_t *)query->hits[s].a;       /* hit_t and salt_hit_t share their layout */
        if (aln_opt->print_xa_cigar)
            for (i = 0; i < r.n_alt[s]; ++i)
                if (query->hits[s].a[i].pos != query->pos && query->hits[s].a[i].is_gap && k < 2 * SALT_MAX_HITS) { xa[k] = dropin_xa_row(j, k); ++k; }
    }
    r.xa_cigars = xa;
    unsigned nm = 0; int n_xv = 0; const char *md = NULL; const uint16_t *xv = NULL;
    if (aln_opt->print_nm_md && query->pos != 0xFFFFFFFF && dropin_tail_get(j, &md, &nm, &xv, &n_xv)) { r.md = md; r.nm = nm; r.xv = xv; r.n_xv = n_xv; }
    kstring_t *ks = query->sam;
    for (;;) {
        const int n = salt_sam_se(&g_sam_refs, &r, aln_opt->print_xa_cigar, aln_opt->rg_id, ks->s, ks->m);
        if (n >= 0) { ks->l = (size_t)n; return; }
        if (n != SALT_ERR_NOMEM) { fprintf(stderr, "[salt_dropin] salt_sam_se failed on %s (%d)\n", query->name, n); exit(1); }
        ks->m = ks->m ? 2 * ks->m : 1024; ks->s = realloc(ks->s, ks->m);
    }
}

static void *finish_worker(void *arg)
{
    fin_thread_t *F = (fin_thread_t *)arg;
    int j;
    /* a contiguous share per worker, not the reference's interleave (alnse.c:1321): neighbouring reads' kstring_t headers share
       a cache line, and aln_samse updates its header a few hundred times per read */
    const int span = F->upto - F->first, lo = F->first + (int)((long)span * F->tid / F->n_threads),
              hi = F->first + (int)((long)span * (F->tid + 1) / F->n_threads);
    for (j = lo; j < hi; ++j) {
        if (F->slot_of[j] < 0) continue;
        if (F->phase == 0) finish_read(F->index, F->queries + j, F->aln_opt, F->ck, (uint32_t)F->slot_of[j]);
        else {
            /* The reader allocated sam->s (128 bytes, query.c:215) in ITS malloc arena; growing it from a worker makes realloc
               lock that one arena for every read of every worker.  Give the line a buffer from this thread's arena instead, large
               enough that aln_samse rarely grows it; the old one is freed by the thread that allocated it. */
            kstring_t *ks = F->queries[j].sam;
            if (ks && ks->m < 1024) { F->old_sam[j] = ks->s; ks->s = malloc(1024); ks->s[0] = '\0'; ks->m = 1024; }
            if (g_native_sam) native_samse(j, F->queries + j, F->aln_opt);
            else {
                dropin_tail_begin_read(j);
                aln_samse(F->index, F->queries + j, F->aln_opt);                                  /* alnse.c:1307 / :1345 */
            }
        }
    }
    return NULL;
}
static void finish_parallel(int n_threads, pthread_t *th, fin_thread_t *F, int phase, int first, int upto, index_t *index, query_t *queries,
                            aln_opt_t *aln_opt, const salt_chunk_t *ck, const int *slot_of, char **old_sam)
{
    int t;
    for (t = 0; t < n_threads; ++t) {
        F[t].tid = t; F[t].n_threads = n_threads; F[t].first = first; F[t].upto = upto; F[t].phase = phase;
        F[t].index = index; F[t].queries = queries; F[t].aln_opt = aln_opt; F[t].ck = ck; F[t].slot_of = slot_of; F[t].old_sam = old_sam;
    }
    if (n_threads == 1) { finish_worker(&F[0]); return; }
    for (t = 0; t < n_threads; ++t) pthread_create(&th[t], NULL, finish_worker, &F[t]);
    for (t = 0; t < n_threads; ++t) pthread_join(th[t], NULL);
}

static double now_s(void)
{
    struct timespec t;
    clock_gettime(CLOCK_MONOTONIC, &t);
    return (double)t.tv_sec + 1e-9 * (double)t.tv_nsec;
}

/* ---- the reference's FASTQ reader and its SAM printing on their own threads (the reference runs both on the main thread,
 * between its batches, alnse.c:1414-1440): batch k lives in buffer k % 3; the reader fills it, the main thread aligns it, the
 * writer prints and frees it.  Output order is the input order: batches are written in sequence. */
#define DROPIN_BATCH_BUFS 3
typedef struct {
    pthread_mutex_t mu; pthread_cond_t cv;
    queryio_t *qs;
    query_t *buf[DROPIN_BATCH_BUFS]; int cnt[DROPIN_BATCH_BUFS];
    char **old_sam[DROPIN_BATCH_BUFS];                /* per read: the reader's first sam buffer, replaced by a worker's */
    int n_read, n_done, n_written, n_destroyed, all_done;     /* batches read / aligned / written / freed so far */
    double t_read, t_print;
} batch_pipe_t;

static void batch_destroy(batch_pipe_t *P, int k)
{
    query_t *q = P->buf[k % DROPIN_BATCH_BUFS];
    char **old = P->old_sam[k % DROPIN_BATCH_BUFS];
    const int n = P->cnt[k % DROPIN_BATCH_BUFS];
    int i;
    for (i = 0; i < n; ++i) {
        free(old[i]); old[i] = NULL;
        query_destroy(q + i);
    }
    memset(q, 0, N_SEQS * sizeof(query_t));
    P->n_destroyed = k + 1;
}

static void *reader_thread(void *arg)
{
    batch_pipe_t *P = (batch_pipe_t *)arg;
    int k;
    for (k = 0;; ++k) {
        pthread_mutex_lock(&P->mu);
        while (k - P->n_written >= DROPIN_BATCH_BUFS) pthread_cond_wait(&P->cv, &P->mu);
        pthread_mutex_unlock(&P->mu);
        double t0 = now_s();
        if (k >= DROPIN_BATCH_BUFS) batch_destroy(P, k - DROPIN_BATCH_BUFS);      /* written by now; freed where it was allocated */
        int n = query_read_multiSeqs(P->qs, N_SEQS, P->buf[k % DROPIN_BATCH_BUFS]);
        P->t_read += now_s() - t0;
        pthread_mutex_lock(&P->mu);
        P->cnt[k % DROPIN_BATCH_BUFS] = n; P->n_read = k + 1;
        pthread_cond_broadcast(&P->cv);
        pthread_mutex_unlock(&P->mu);
        if (n <= 0) break;
    }
    return NULL;
}

static void *writer_thread(void *arg)
{
    batch_pipe_t *P = (batch_pipe_t *)arg;
    int k, i, n_tot = 0;
    for (k = 0;; ++k) {
        pthread_mutex_lock(&P->mu);
        while (P->n_done <= k && !P->all_done) pthread_cond_wait(&P->cv, &P->mu);
        const int have = P->n_done > k;
        pthread_mutex_unlock(&P->mu);
        if (!have) break;
        double t0 = now_s();
        query_t *q = P->buf[k % DROPIN_BATCH_BUFS];
        const int n = P->cnt[k % DROPIN_BATCH_BUFS];
        for (i = 0; i < n; ++i) puts(q[i].sam->s);     /* alnse.c:1433-1439 */
        n_tot += n;
        fprintf(stderr, "%d reads have been aligned!\n", n_tot);
        P->t_print += now_s() - t0;
        pthread_mutex_lock(&P->mu);
        P->n_written = k + 1;
        pthread_cond_broadcast(&P->cv);
        pthread_mutex_unlock(&P->mu);
    }
    return NULL;
}

int alnse_core(const opt_t *opt)
{
    const char *seed_env = getenv("SALT_DROPIN_SEED");
    const int gpu_seed = seed_env && !strcmp(seed_env, "gpu");
    fprintf(stderr, "[alnse_core/gpu]:  Start single end alignment (verification on libsalt_b200%s)\n",
            gpu_seed ? ", seeding + locate on libsalt_b200" : "");
    aln_opt_t *aln_opt = aln_opt_init(opt);
    double t_start = now_s();
    index_t *index = alnse_index_reload(opt->fn_index);
    const double t_load = now_s() - t_start;
    if (aln_opt->extend_algo == EXTEND_SW || aln_opt->l_overlap <= 0) {
        fprintf(stderr, "[salt_dropin] only the overlap/LV path (the reference's default) is served\n");
        exit(1);
    }
    double t_init = now_s();
    salt_b200_t *gpu = salt_b200_init(index->mixRef->seq, index->mixRef->l, index->pac, index->bntseq->l_pac, 0);
    if (!gpu) die("salt_b200_init");
    salt_seed_opt_t sopt;
    memset(&sopt, 0, sizeof sopt);                      /* locate_mode 0: alnse_locate_alt, the single-end program's */
    sopt.l_seed = aln_opt->l_seed; sopt.l_overlap = aln_opt->l_overlap; sopt.max_seed = aln_opt->max_seed;
    sopt.max_locate = (int)aln_opt->max_locate; sopt.seed_only_ref = aln_opt->seed_only_ref;
    if (gpu_seed) {
        /* the FM-indexes exactly as the reference's loaders left them in memory (indexio.c:23-50) */
        salt_fm_index_t fx;
        int i;
        rbwt_t *r = index->rbwt2->rbwt1;
        memset(&fx, 0, sizeof fx);
        fx.c_bwt = index->cbwt->bwt; fx.c_bwt_words = index->cbwt->bwt_size;
        fx.c_primary = index->cbwt->primary; fx.c_seq_len = index->cbwt->seq_len;
        for (i = 0; i < 5; ++i) fx.c_L2[i] = index->cbwt->L2[i];
        fx.c_sa = index->cbwt->sa; fx.c_n_sa = index->cbwt->n_sa; fx.c_sa_intv = (uint32_t)index->cbwt->sa_intv;
        fx.lkt = index->lkt->item; fx.lkt_len = index->lkt->maxLookupLen;
        fx.r_bwt = r->bwtCode; fx.r_bwt_words = r->bwtSizeInWord;
        fx.r_occ = r->occValue; fx.r_occ_words = r->occSizeInWord;
        fx.r_occ_major = r->occValueMajor; fx.r_occ_major_words = r->occMajorSizeInWord;
        fx.r_sa_sharp = r->saValueSharp; fx.r_n_sa_sharp = r->saValueSizeSharp;
        for (i = 0; i < 6; ++i) fx.r_cum[i] = r->cumulativeFreq[i];
        fx.r_inv_sa0 = r->inverseSa0; fx.r_text_len = r->textLength;
        if (salt_b200_set_index(gpu, &fx) != SALT_OK) die("salt_b200_set_index");
    }
    /* two chunk queues: while the GPU verifies one, the host finishes (hit selection, SAM text) the other */
    salt_chunk_t *ck[2];
    int c;
    for (c = 0; c < 2; ++c) {
        ck[c] = salt_chunk_new(DROPIN_CHUNK_READS, (size_t)DROPIN_CHUNK_READS * 1024, DROPIN_CHUNK_CANDS);
        if (!ck[c]) die("salt_chunk_new");
    }
    t_init = now_s() - t_init;
    const int n_threads = opt->n_threads > 1 ? opt->n_threads : 1;
    seed_thread_t *T = calloc((size_t)n_threads, sizeof *T);
    fin_thread_t *Fin = calloc((size_t)n_threads, sizeof *Fin);
    pthread_t *th = calloc((size_t)n_threads, sizeof *th);
    int t;
    for (t = 0; t < n_threads; ++t) {
        T[t].tid = t; T[t].n_threads = n_threads; T[t].index = index; T[t].aln_opt = aln_opt;
        T[t].aux[0] = aux_init(opt->l_read, opt->l_seed);
        T[t].aux[1] = aux_init(opt->l_read, opt->l_seed);
    }
    {
        const char *e = getenv("SALT_DROPIN_SAM");
        g_native_sam = e && !strcmp(e, "native");
        if (g_native_sam) {
            bntseq_t *bns = index->bntseq;
            const char **nm = calloc((size_t)bns->n_seqs, sizeof *nm); int64_t *of = calloc((size_t)bns->n_seqs, sizeof *of);
            for (int i = 0; i < bns->n_seqs; ++i) { nm[i] = bns->anns[i].name; of[i] = bns->anns[i].offset; }
            g_sam_refs.n_seqs = bns->n_seqs; g_sam_refs.names = nm; g_sam_refs.offsets = of; g_sam_refs.l_pac = bns->l_pac;
            fprintf(stderr, "[salt_dropin] SAM lines by salt_sam_se\n");
        }
    }
    queryio_t *qs = query_open(opt->fn_read1);
    batch_pipe_t P;
    memset(&P, 0, sizeof P);
    pthread_mutex_init(&P.mu, NULL); pthread_cond_init(&P.cv, NULL);
    P.qs = qs;
    for (t = 0; t < DROPIN_BATCH_BUFS; ++t) { P.buf[t] = calloc(N_SEQS, sizeof(query_t)); P.old_sam[t] = calloc(N_SEQS, sizeof(char *)); }
    seeded_read_t *seeds = calloc(N_SEQS, sizeof(seeded_read_t));
    int *slot_of = calloc(N_SEQS, sizeof(int));        /* index of read i in the GPU chunk that holds it */
    aln_samhead(opt, index->bntseq);
    fflush(stdout);
    pthread_t th_reader, th_writer;
    pthread_create(&th_reader, NULL, reader_thread, &P);
    pthread_create(&th_writer, NULL, writer_thread, &P);
    int n, i, n_tot = 0, n_chunks = 0, batch;
    double t_seed = 0, t_gpu_wait = 0, t_finish = 0, t_select = 0, t_tail = 0, t_starved = 0;
    for (batch = 0;; ++batch) {
        double tw = now_s();
        pthread_mutex_lock(&P.mu);
        while (P.n_read <= batch) pthread_cond_wait(&P.cv, &P.mu);
        n = P.cnt[batch % DROPIN_BATCH_BUFS];
        pthread_mutex_unlock(&P.mu);
        t_starved += now_s() - tw;
        if (n <= 0) break;
        query_t *multiSeqs = P.buf[batch % DROPIN_BATCH_BUFS];
        char **old_sam = P.old_sam[batch % DROPIN_BATCH_BUFS];
        n_tot += n;
        double t0 = now_s();
        if (!gpu_seed) {
            for (t = 0; t < n_threads; ++t) { T[t].n = n; T[t].queries = multiSeqs; T[t].out = seeds; }
            if (n_threads == 1) seed_worker(&T[0]);
            else {
                for (t = 0; t < n_threads; ++t) pthread_create(&th[t], NULL, seed_worker, &T[t]);
                for (t = 0; t < n_threads; ++t) pthread_join(th[t], NULL);
            }
        } else {
            for (i = 0; i < n; ++i) { seeds[i].skip = multiSeqs[i].n_ambiguous > DROPIN_MAX_N_PERSEQ; seeds[i].n[0] = seeds[i].n[1] = 0; }
        }
        t_seed += now_s() - t0;
        /* sub-chunks [first, upto) go to slots 0 / 1 alternately; chunk k-1 is finished on the host while k is on the GPU */
        int first = 0, k = 0;
        int pend_first = -1, pend_upto = -1, pend_c = -1;
        while (first < n || pend_c >= 0) {
            int upto = first, cur = -1;
            if (first < n) {
                cur = k & 1;
                salt_chunk_reset(ck[cur]);
                for (i = first; i < n; ++i) {
                    slot_of[i] = -1;
                    if (seeds[i].skip) continue;
                    query_t *query = multiSeqs + i;
                    int at = salt_chunk_add_read(ck[cur], query->seq, query->l_seq, seeds[i].a[0], seeds[i].n[0], seeds[i].a[1], seeds[i].n[1]);
                    if (at == SALT_ERR_NOMEM && salt_chunk_n_reads(ck[cur]) > 0) break;       /* queue full: verify what is queued */
                    if (at < 0) die("salt_chunk_add_read");
                    slot_of[i] = at;
                    if (gpu_seed && salt_chunk_n_reads(ck[cur]) * (size_t)sopt.max_locate >= DROPIN_CHUNK_CANDS) { ++i; break; }
                }
                upto = i;
                double tg = now_s();
                if (gpu_seed) {
                    /* SE thresholds: nogap 3 (alnse.c:1079), gapped l_seq/10 (alnse.c:1090) */
                    if (salt_chunk_seed_verify(gpu, ck[cur], &sopt, 3, -1) != SALT_OK) die("salt_chunk_seed_verify");
                } else if (salt_chunk_n_reads(ck[cur]) > 0) {
                    if (salt_chunk_submit(gpu, cur, ck[cur], 3, -1) != SALT_OK) die("salt_chunk_submit");
                }
                t_gpu_wait += now_s() - tg;
                ++n_chunks;
            }
            /* host seeding: finish the previous sub-chunk while the GPU verifies this one.  Device seeding is synchronous
               on slot 0 (the slot's resident reads feed the SAM tail), so there the sub-chunk just verified is finished at once */
            if (gpu_seed && cur >= 0) { pend_c = cur; pend_first = first; pend_upto = upto; first = upto; ++k; cur = -1; }
            if (pend_c >= 0) {
                int j;
                double tg = now_s();
                if (!gpu_seed && salt_chunk_wait(gpu, pend_c, ck[pend_c]) != SALT_OK) die("salt_chunk_wait");
                t_gpu_wait += now_s() - tg;
                double tf = now_s();
                (void)j;
                finish_parallel(n_threads, th, Fin, 0, pend_first, pend_upto, index, multiSeqs, aln_opt, ck[pend_c], slot_of, old_sam);
                t_select += now_s() - tf;
                double tt = now_s();
                if (aln_opt->print_nm_md || aln_opt->print_xa_cigar)
                    dropin_tail_prepare(gpu, gpu_seed ? 0 : pend_c, multiSeqs, slot_of, pend_first, pend_upto);
                t_tail += now_s() - tt;
                finish_parallel(n_threads, th, Fin, 1, pend_first, pend_upto, index, multiSeqs, aln_opt, ck[pend_c], slot_of, old_sam);
                t_finish += now_s() - tf;
                pend_c = -1;
            }
            if (cur >= 0) { pend_c = cur; pend_first = first; pend_upto = upto; first = upto; ++k; }
        }
        pthread_mutex_lock(&P.mu);
        P.n_done = batch + 1;
        pthread_cond_broadcast(&P.cv);
        pthread_mutex_unlock(&P.mu);
    }
    pthread_mutex_lock(&P.mu);
    P.all_done = 1;
    pthread_cond_broadcast(&P.cv);
    pthread_mutex_unlock(&P.mu);
    pthread_join(th_reader, NULL);
    pthread_join(th_writer, NULL);
    while (P.n_destroyed < batch) batch_destroy(&P, P.n_destroyed);
    fprintf(stderr, "[salt_dropin] %d reads, %d GPU chunks, %d seeding threads: seeding %s %.3f s, GPU calls (exposed wait) %.3f s, "
                    "hit selection + SAM text %.3f s\n", n_tot, n_chunks, n_threads, gpu_seed ? "(on the GPU, inside the GPU calls)" : "on the host",
            t_seed, t_gpu_wait, t_finish);
    fprintf(stderr, "[salt_dropin] wall %.3f s: index load (reference loaders) %.3f, GPU init + uploads %.3f; main thread: waiting for the reader %.3f, "
                    "hit selection %.3f, GPU tail calls %.3f, SAM text %.3f; beside it: FASTQ reader + free thread %.3f, SAM print thread %.3f\n",
            now_s() - t_start, t_load, t_init, t_starved, t_select, t_tail, t_finish - t_select - t_tail, P.t_read, P.t_print);
    for (t = 0; t < n_threads; ++t) { aux_destroy(T[t].aux[0]); aux_destroy(T[t].aux[1]); }
    for (i = 0; i < N_SEQS; ++i) { free(seeds[i].a[0]); free(seeds[i].a[1]); }
    free(T); free(th); free(seeds); free(Fin);
    for (t = 0; t < DROPIN_BATCH_BUFS; ++t) { free(P.buf[t]); free(P.old_sam[t]); }
    free(slot_of);
    query_close(qs);
    salt_chunk_free(ck[0]); salt_chunk_free(ck[1]);
    dropin_tail_report();
    salt_b200_destroy(gpu);
    alnse_index_destroy(index);
    aln_opt_destroy(aln_opt);
    return EXIT_SUCCESS;
}
