/*
 * salt_aln.c -- the aligner as a program of its own, built from this repository's libraries only (libsalt_host.so over
 * libsalt_b200.so): no reference code in the loop.
 *
 *     salt_aln [-p] [-a N] [-b N] [-r N] [-m N] [-s N] [-c] [-d] [-v] [-g RG] [-t N] PREFIX reads.fq[.gz] [mates.fq[.gz]] > out.sam
 *
 * The counterpart of the reference's `salt` (opt_parse, aln.c:138-226; alnse_core, alnse.c:1353-1480; alnpe_core,
 * alnpe.c:530-660) on an index written by its salt-idx, with the same options and the same SAM (except the @PG line):
 *
 *   single-end   FASTQ text --salt_fastq_pack--> reads --salt_chunk_seed_verify (seeding, locate, verification on the GPU)-->
 *                --salt_chunk_results (hit selection, mapq, CIGAR)--> --salt_b200_tail_primaries / salt_b200_lv_cigar (MD NM XV, XA CIGARs)-->
 *                --salt_sam_se--> SAM
 *   paired-end   two FASTQ texts --salt_fastq_pack--> mates 2i, 2i+1 --salt_b200_seed_locate (paired-end flavour)-->
 *                --salt_chunk_submit / _wait (thresholds 3 / 3)--> --salt_chunk_pair (pairing plans, mate rescue, apply)-->
 *                --salt_b200_md_nm / salt_b200_lv_cigar--> --salt_sam_pe--> SAM
 *
 * Reads go through in batches of N_SEQS (aln.h:27); the per-read host work of a batch (unpacking, SAM text) runs on -t
 * threads.  tests/test_native_pipeline.py runs the program on the SIMT emulator and on the GPU against the reference program.
 *
 * One deliberate difference (paired-end only): where an SNP-context interval is wider than -m the reference locates a random
 * subset of its rows (srand(time(0)) / rand(), alnse.c:538-552), so two runs of the reference itself differ there; the
 * device leaves that interval out and this program reports the number of mates concerned.
 */
#define _GNU_SOURCE
#include <errno.h>
#include <getopt.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <unistd.h>
#include <zlib.h>
#include <pthread.h>
#include "salt_host.h"

#define N_SEQS 100000                 /* aln.h:27 */
#define MAX_N_PERSEQ_SE 200           /* alnse.c:1281 */
#define MAX_N_PERSEQ_PE 5             /* alnpe.c:477 */
#define MAX_HITS 5                    /* aln.h:139 */
#define PE_LIST_CAP 1024              /* room per candidate list on the device to start with: a batch whose lists fill it is located
                                         again with four times the room, up to the engine's 16384 (the reference allows 262144,
                                         alnse.c:42); the scratch is n_mates x 2 x this x 4 bytes */
#define XA_STRIDE 256                 /* sam.c:216 */
#define MD_STRIDE 512
#define XV_STRIDE 64                  /* sam.c:242 */
#define UNMAPPED 0xFFFFFFFFu

typedef struct {
    int paired, n_threads, l_overlap, max_seed, max_locate, seed_only_ref, print_xa_cigar, print_nm_md, device;
    uint32_t min_tlen, max_tlen;
    const char *rg_id, *prefix, *fn[2];
    uint32_t batch; int list_cap;
} opts_t;

static double now(void) { struct timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return (double)t.tv_sec + 1e-9 * (double)t.tv_nsec; }

static void die(const char *what, int rc)
{
    fprintf(stderr, "[salt_aln] %s failed (%d): %s\n", what, rc, salt_b200_last_error());
    exit(1);
}
static void *xmalloc(size_t n) { void *p = malloc(n ? n : 1); if (!p) { fprintf(stderr, "[salt_aln] out of memory (%zu bytes)\n", n); exit(1); } return p; }
static void *xcalloc(size_t n, size_t s) { void *p = calloc(n ? n : 1, s); if (!p) { fprintf(stderr, "[salt_aln] out of memory\n"); exit(1); } return p; }
static void *xrealloc(void *q, size_t n) { void *p = realloc(q, n ? n : 1); if (!p) { fprintf(stderr, "[salt_aln] out of memory (%zu bytes)\n", n); exit(1); } return p; }

/* The per-read host work of a batch runs as T contiguous shares, share 0 on the calling thread. */
typedef void (*share_fn)(int t, int T, void *arg);
typedef struct { share_fn fn; int t, T; void *arg; } share_t;
static void *share_main(void *p) { share_t *s = (share_t *)p; s->fn(s->t, s->T, s->arg); return NULL; }
static void run_shares(int T, share_fn fn, void *arg)
{
    if (T > 256) T = 256;
    pthread_t th[256]; share_t sh[256];
    int started = 1;
    for (int t = 1; t < T; ++t) {
        sh[t].fn = fn; sh[t].t = t; sh[t].T = T; sh[t].arg = arg;
        if (pthread_create(&th[t], NULL, share_main, &sh[t]) != 0) break;
        ++started;
    }
    /* threads that could not be started: their shares run here after share 0 */
    fn(0, T, arg);
    for (int t = started; t < T; ++t) fn(t, T, arg);
    for (int t = 1; t < started; ++t) pthread_join(th[t], NULL);
}

/* ---- the index files as salt-idx writes them (what alnse_index_reload reads, indexio.c:23-50) ---------------------------- */
typedef struct {
    salt_fm_index_t fm;
    int l_seed;
    uint32_t l; const uint32_t *mixref;           /* PREFIX.ref */
    const uint8_t *pac;                           /* PREFIX.C.pac */
    salt_sam_refs_t refs; int *rec_len;           /* PREFIX.C.ann */
} index_files_t;

static void *slurp(const char *prefix, const char *ext, size_t *bytes)
{
    char fn[4096];
    snprintf(fn, sizeof fn, "%s%s", prefix, ext);
    FILE *f = fopen(fn, "rb");
    if (!f) { fprintf(stderr, "[salt_aln] cannot open %s: %s\n", fn, strerror(errno)); exit(1); }
    fseek(f, 0, SEEK_END);
    const long n = ftell(f);
    fseek(f, 0, SEEK_SET);
    void *p = xmalloc((size_t)n + 16);
    if (fread(p, 1, (size_t)n, f) != (size_t)n) { fprintf(stderr, "[salt_aln] short read on %s\n", fn); exit(1); }
    fclose(f);
    *bytes = (size_t)n;
    return p;
}
static void bad_index(const char *what) { fprintf(stderr, "[salt_aln] index file inconsistent: %s\n", what); exit(1); }

static void load_index(const char *prefix, index_files_t *ix)
{
    size_t nb;
    memset(ix, 0, sizeof *ix);
    salt_fm_index_t *fm = &ix->fm;
    uint32_t *raw = (uint32_t *)slurp(prefix, ".C.bwt", &nb);               /* primary, L2[1..4], BWT words (bwtio.c:55-74) */
    if (nb < 20) bad_index(".C.bwt");
    fm->c_primary = raw[0]; fm->c_L2[0] = 0;
    for (int i = 1; i < 5; ++i) fm->c_L2[i] = raw[i];
    fm->c_bwt = raw + 5; fm->c_bwt_words = nb / 4 - 5; fm->c_seq_len = fm->c_L2[4];
    raw = (uint32_t *)slurp(prefix, ".C.sa", &nb);                          /* primary, 4 skipped, sa_intv, seq_len, sa[1..] (bwtio.c:30-53) */
    if (nb < 28 || raw[0] != fm->c_primary || raw[6] != fm->c_seq_len || raw[5] == 0) bad_index(".C.sa against .C.bwt");
    fm->c_sa_intv = raw[5];
    fm->c_n_sa = (fm->c_seq_len + fm->c_sa_intv) / fm->c_sa_intv;
    if (nb / 4 < 7 + (size_t)fm->c_n_sa - 1) bad_index(".C.sa length");
    raw[6] = 0xFFFFFFFFu;                                                   /* sa[0] = (uint32_t)-1 (bwtio.c:46) sits right before sa[1] */
    fm->c_sa = raw + 6;
    raw = (uint32_t *)slurp(prefix, ".C.lkt", &nb);                         /* maxLookupLen, 4^len + 1 cumulative counts (lookup.c:47-65) */
    if (nb < 4 || raw[0] > 15 || nb / 4 < 1 + ((size_t)1 << (2 * raw[0])) + 1) bad_index(".C.lkt");
    fm->lkt_len = raw[0]; fm->lkt = raw + 1;
    raw = (uint32_t *)slurp(prefix, ".R.backward.bwt", &nb);                /* rbwt.c:250-270 */
    if (nb < 32 || nb / 4 < 8 + (size_t)raw[7]) bad_index(".R.backward.bwt");
    fm->r_text_len = raw[0]; fm->r_inv_sa0 = raw[1]; fm->r_cum[0] = 0;
    for (int i = 1; i < 6; ++i) fm->r_cum[i] = raw[1 + i];
    fm->r_bwt_words = raw[7]; fm->r_bwt = raw + 8;
    raw = (uint32_t *)slurp(prefix, ".R.backward.occ", &nb);                /* rbwt.c:271-288 */
    if (nb < 8 || nb / 4 < 2 + (size_t)raw[0] || nb / 4 < 2 + (size_t)raw[0] + raw[1 + raw[0]]) bad_index(".R.backward.occ");
    fm->r_occ_words = raw[0]; fm->r_occ = raw + 1;
    fm->r_occ_major_words = raw[1 + raw[0]]; fm->r_occ_major = raw + 2 + raw[0];
    raw = (uint32_t *)slurp(prefix, ".R.backward.sa", &nb);                 /* rbwt.c:558-574 */
    if (nb < 4 || nb / 4 < 1 + (size_t)raw[0]) bad_index(".R.backward.sa");
    fm->r_n_sa_sharp = raw[0]; fm->r_sa_sharp = raw + 1;
    raw = (uint32_t *)slurp(prefix, ".R.seedLen", &nb);                     /* aln.c:216-224 */
    if (nb < 4) bad_index(".R.seedLen");
    ix->l_seed = (int)raw[0];
    free(raw);
    raw = (uint32_t *)slurp(prefix, ".ref", &nb);                           /* uint32 l, then the 4-bit reference (metaref.c:61-93) */
    if (nb < 4 || nb / 4 < 1 + ((size_t)raw[0] + 7) / 8) bad_index(".ref");
    ix->l = raw[0]; ix->mixref = raw + 1;
    ix->pac = (const uint8_t *)slurp(prefix, ".C.pac", &nb);
    if (nb < ((size_t)ix->l + 3) / 4) bad_index(".C.pac shorter than the reference");
    /* PREFIX.C.ann as bns_dump writes it (bntseq.c): "l_pac n_seqs seed", then per record "gi name anno" / "offset len n_ambs" */
    char *ann = (char *)slurp(prefix, ".C.ann", &nb);
    ann[nb] = '\0';
    char *p = ann, *e;
    long long l_pac = strtoll(p, &e, 10); p = e;
    long n_seqs = strtol(p, &e, 10); p = e;
    if (l_pac != (long long)ix->l || n_seqs < 1) bad_index(".C.ann against .ref");
    p = strchr(p, '\n');
    const char **names = (const char **)xcalloc((size_t)n_seqs, sizeof *names);
    int64_t *offsets = (int64_t *)xcalloc((size_t)n_seqs, sizeof *offsets);
    ix->rec_len = (int *)xcalloc((size_t)n_seqs, sizeof(int));
    for (long i = 0; i < n_seqs; ++i) {
        if (!p) bad_index(".C.ann truncated");
        ++p;
        char *sp = strchr(p, ' ');                       /* gi */
        if (!sp) bad_index(".C.ann record line");
        char *nm = sp + 1;
        char *nl = strchr(nm, '\n');
        if (!nl) bad_index(".C.ann record line");
        char *sp2 = memchr(nm, ' ', (size_t)(nl - nm));
        *(sp2 ? sp2 : nl) = '\0';
        names[i] = nm;
        p = nl + 1;
        offsets[i] = strtoll(p, &e, 10); p = e;
        ix->rec_len[i] = (int)strtol(p, &e, 10); p = e;
        p = strchr(p, '\n');
    }
    ix->refs.n_seqs = (int)n_seqs; ix->refs.names = names; ix->refs.offsets = offsets; ix->refs.l_pac = l_pac;
}

/* ---- FASTQ input: blocks of text through salt_fastq_pack ------------------------------------------------------------------ */
typedef struct {
    gzFile f; int eof;                            /* plain or gzip-compressed text, as the reference's reader takes it (query.c:112) */
    char *text; size_t cap, len, used;            /* text[used, len) has not been parsed yet */
    salt_fastq_t fq;                              /* arrays for up to max_reads records */
    uint32_t max_reads;
    uint32_t *roffs;                              /* n + 1 base offsets of the batch */
} reader_t;

static void reader_alloc(reader_t *r, size_t cap)
{
    r->text = (char *)xrealloc(r->text, cap + 16); r->cap = cap;
    salt_fastq_t *q = &r->fq;
    q->bases_cap = cap; q->bases = (uint8_t *)xrealloc(q->bases, cap / 4 + 16);
    q->n_pos_cap = cap / 4 + 1024; q->n_pos = (uint32_t *)xrealloc(q->n_pos, q->n_pos_cap * 4);
}
static void reader_init(reader_t *r, gzFile f, uint32_t max_reads)
{
    memset(r, 0, sizeof *r);
    r->f = f;
    r->max_reads = max_reads;
    salt_fastq_t *q = &r->fq;
    q->lens = (uint16_t *)xmalloc((size_t)max_reads * 2); q->n_ambiguous = (uint16_t *)xmalloc((size_t)max_reads * 2);
    q->name_off = (uint32_t *)xmalloc((size_t)max_reads * 4); q->name_len = (uint16_t *)xmalloc((size_t)max_reads * 2);
    q->comment_off = (uint32_t *)xmalloc((size_t)max_reads * 4); q->comment_len = (uint16_t *)xmalloc((size_t)max_reads * 2);
    q->qual_off = (uint32_t *)xmalloc((size_t)max_reads * 4);
    r->roffs = (uint32_t *)xmalloc(((size_t)max_reads + 1) * 4);
    reader_alloc(r, (size_t)max_reads * 320 < ((size_t)1 << 20) ? ((size_t)1 << 20) : (size_t)max_reads * 320);
}
/* the next batch of up to `want` records; names and quality strings become NUL-terminated strings inside r->text */
static uint32_t reader_next(reader_t *r, uint32_t want)
{
    if (want > r->max_reads) want = r->max_reads;
    for (;;) {
        if (r->used) { memmove(r->text, r->text + r->used, r->len - r->used); r->len -= r->used; r->used = 0; }
        while (!r->eof && r->len < r->cap) {
            const size_t room = r->cap - r->len;
            const int got = gzread(r->f, r->text + r->len, (unsigned)(room < ((size_t)1 << 30) ? room : ((size_t)1 << 30)));
            if (got < 0) { int e = 0; fprintf(stderr, "[salt_aln] read error: %s\n", gzerror(r->f, &e)); exit(1); }
            if (got == 0) r->eof = 1;
            r->len += (size_t)got;
        }
        size_t consumed = 0;
        const int n = salt_fastq_pack(r->text, r->len, r->eof, want, &r->fq, &consumed);
        if (n < 0) die("salt_fastq_pack (malformed FASTQ)", n);
        if ((uint32_t)n < want && !r->eof) {      /* the block (or an array sized after it) ended before the batch did: a longer block */
            if (r->cap >= ((size_t)1 << 31)) { fprintf(stderr, "[salt_aln] a batch of FASTQ records longer than 2 GiB\n"); exit(1); }
            reader_alloc(r, r->cap * 2);
            continue;
        }
        r->used = n == 0 ? r->len : consumed;
        const salt_fastq_t *q = &r->fq;
        r->roffs[0] = 0;
        for (int i = 0; i < n; ++i) {
            r->roffs[i + 1] = r->roffs[i] + q->lens[i];
            r->text[q->name_off[i] + q->name_len[i]] = '\0';
            if (q->qual_off[i] != 0xFFFFFFFFu) {
                const size_t e = (size_t)q->qual_off[i] + q->lens[i];
                if (e > r->len || (e < r->len && r->text[e] != '\n' && r->text[e] != '\r')) {
                    fprintf(stderr, "[salt_aln] %s: quality strings over several lines are not supported\n", r->text + q->name_off[i]);
                    exit(1);
                }
                r->text[e] = '\0';
            }
        }
        return (uint32_t)n;
    }
}
/* Two frames per input file and a thread that fills the one the program is not looking at: while a batch is aligned and
 * printed, the next one is read and parsed.  The unparsed tail of a frame's text opens the next frame. */
typedef struct {
    reader_t fr[2]; int cur;                      /* fr[cur]: the batch the program holds */
    uint32_t want, n_ready;
    pthread_t th; pthread_mutex_t mu; pthread_cond_t cv;
    int kick, ready, quit;
    double secs;
} stream_t;
static void *stream_main(void *p)
{
    stream_t *s = (stream_t *)p;
    pthread_mutex_lock(&s->mu);
    for (;;) {
        while (!s->kick && !s->quit) pthread_cond_wait(&s->cv, &s->mu);
        if (s->quit) break;
        s->kick = 0;
        const reader_t *a = &s->fr[s->cur]; reader_t *b = &s->fr[s->cur ^ 1];
        const uint32_t want = s->want;
        pthread_mutex_unlock(&s->mu);
        const double t0 = now();
        const size_t left = a->len - a->used;
        if (b->cap < left || b->cap < a->cap) reader_alloc(b, a->cap > left ? a->cap : left);
        memcpy(b->text, a->text + a->used, left);
        b->len = left; b->used = 0; b->eof = a->eof;
        const uint32_t n = reader_next(b, want);
        const double dt = now() - t0;
        pthread_mutex_lock(&s->mu);
        s->n_ready = n; s->ready = 1; s->secs += dt;
        pthread_cond_broadcast(&s->cv);
    }
    pthread_mutex_unlock(&s->mu);
    return NULL;
}
static void stream_kick(stream_t *s, uint32_t want)
{
    pthread_mutex_lock(&s->mu);
    s->want = want; s->kick = 1;
    pthread_cond_broadcast(&s->cv);
    pthread_mutex_unlock(&s->mu);
}
static void stream_open(stream_t *s, const char *fn, uint32_t max_reads)
{
    memset(s, 0, sizeof *s);
    gzFile f = strcmp(fn, "-") ? gzopen(fn, "rb") : gzdopen(0, "rb");
    if (!f) { fprintf(stderr, "[salt_aln] cannot open %s: %s\n", fn, strerror(errno)); exit(1); }
    gzbuffer(f, 1u << 20);
    reader_init(&s->fr[0], f, max_reads); reader_init(&s->fr[1], f, max_reads);
    pthread_mutex_init(&s->mu, NULL); pthread_cond_init(&s->cv, NULL);
    if (pthread_create(&s->th, NULL, stream_main, s) != 0) { fprintf(stderr, "[salt_aln] cannot start the reader thread\n"); exit(1); }
    stream_kick(s, max_reads);
}
/* the batch that was asked for with the last stream_kick: its frame, *n records (0 at the end of the input) */
static const reader_t *stream_get(stream_t *s, uint32_t *n)
{
    pthread_mutex_lock(&s->mu);
    while (!s->ready) pthread_cond_wait(&s->cv, &s->mu);
    s->ready = 0; s->cur ^= 1; *n = s->n_ready;
    pthread_mutex_unlock(&s->mu);
    return &s->fr[s->cur];
}
static double stream_close(stream_t *s)
{
    pthread_mutex_lock(&s->mu);
    s->quit = 1;
    pthread_cond_broadcast(&s->cv);
    pthread_mutex_unlock(&s->mu);
    pthread_join(s->th, NULL);
    return s->secs;
}

/* read i's bases as one code per byte (query->seq, query.c:177-181): four bases per table look-up */
static uint32_t unpack4[256];
static void unpack_init(void)
{
    for (uint32_t b = 0; b < 256; ++b) unpack4[b] = (b & 3) | ((b >> 2 & 3) << 8) | ((b >> 4 & 3) << 16) | ((b >> 6 & 3) << 24);
}
static void unpack_read(const salt_fastq_t *q, uint32_t base0, uint32_t L, uint8_t *out)
{
    uint32_t k = 0, p = base0;
    for (; k < L && (p & 3); ++k, ++p) out[k] = (uint8_t)((q->bases[p >> 2] >> (2 * (p & 3))) & 3);
    for (; k + 4 <= L; k += 4, p += 4) memcpy(out + k, &unpack4[q->bases[p >> 2]], 4);
    for (; k < L; ++k, ++p) out[k] = (uint8_t)((q->bases[p >> 2] >> (2 * (p & 3))) & 3);
}
/* codes of reads [0, n) of the batch into dst[dst_off[i * stride] ...]; the N positions of the stream are ascending */
typedef struct { const reader_t *r; uint32_t n; uint8_t *dst; const uint32_t *dst_off; int stride; } unpack_arg_t;
static void unpack_share(int t, int T, void *arg)
{
    const unpack_arg_t *a = (const unpack_arg_t *)arg;
    const salt_fastq_t *q = &a->r->fq;
    const uint32_t lo = (uint32_t)((uint64_t)a->n * t / T), hi = (uint32_t)((uint64_t)a->n * (t + 1) / T);
    for (uint32_t i = lo; i < hi; ++i) unpack_read(q, a->r->roffs[i], q->lens[i], a->dst + a->dst_off[(size_t)i * a->stride]);
}
static void unpack_batch(const reader_t *r, uint32_t n, uint8_t *dst, const uint32_t *dst_off, int stride, int n_thr)
{
    const salt_fastq_t *q = &r->fq;
    unpack_arg_t a = {r, n, dst, dst_off, stride};
    run_shares(n < 4096 ? 1 : n_thr, unpack_share, &a);
    size_t i = 0;
    for (size_t j = 0; j < q->n_n; ++j) {
        const uint32_t p = q->n_pos[j];
        while (r->roffs[i + 1] <= p) ++i;
        dst[dst_off[i * stride] + (p - r->roffs[i])] = 4;
    }
}

/* ---- output: every thread formats a contiguous share of the batch into its own buffer, written in order ------------------- */
typedef struct { char *s; size_t len, cap; } outbuf_t;
static inline char *out_room(outbuf_t *o, size_t need)
{
    if (o->len + need > o->cap) { o->cap = (o->len + need) * 2; o->s = (char *)xrealloc(o->s, o->cap); }
    return o->s + o->len;
}

typedef struct {
    double index, gpu_init, parse, parse_beside, gpu, select, tail, text, write, write_wait;
    size_t reads, flagged, cut, relocated, md_tags, xa_cigars;
    salt_pe_stats_t pe;
} stats_t;

/* CIGARs of the gapped alternates that get printed (sam_add_xa, sam.c:205-225), one salt_b200_lv_cigar call for a batch:
 * alternates of read `row` whose position differs from primary[row]; xa_first[i] .. xa_first[i + 1] index xa_ptr. */
typedef struct {
    salt_pair_t *pairs; uint8_t *k; int8_t *e; char *cig; const char **ptr; uint32_t *first; size_t cap, n;
} xa_t;
static void xa_collect(xa_t *x, salt_b200_t *h, const salt_read_result_t *res, const int32_t *row_of, const uint32_t *primary, uint32_t n)
{
    x->n = 0;
    for (uint32_t i = 0; i < n; ++i) {
        x->first[i] = (uint32_t)x->n;
        if (row_of[i] < 0) continue;
        const salt_read_result_t *r = res + row_of[i];
        for (int s = 0; s < 2; ++s)
            for (int j = 0; j < r->n_alt[s]; ++j) {
                const salt_hit_t *a = &r->alt[s][j];
                if (a->pos == primary[i] || !a->is_gap) continue;
                if (x->n == x->cap) {
                    x->cap = x->cap ? x->cap * 2 : 4096;
                    x->pairs = (salt_pair_t *)xrealloc(x->pairs, x->cap * sizeof *x->pairs); x->k = (uint8_t *)xrealloc(x->k, x->cap);
                    x->e = (int8_t *)xrealloc(x->e, x->cap); x->cig = (char *)xrealloc(x->cig, x->cap * XA_STRIDE);
                    x->ptr = (const char **)xrealloc((void *)x->ptr, x->cap * sizeof *x->ptr);
                }
                x->pairs[x->n].rs = ((uint32_t)row_of[i] << 1) | (uint32_t)s; x->pairs[x->n].pos = a->pos; x->k[x->n] = a->n_diff;
                ++x->n;
            }
    }
    x->first[n] = (uint32_t)x->n;
    if (!x->n) return;
    const int rc = salt_b200_lv_cigar(h, x->pairs, x->k, x->n, x->cig, XA_STRIDE, x->e);
    if (rc != SALT_OK) die("salt_b200_lv_cigar (XA)", rc);
    for (size_t j = 0; j < x->n; ++j) {
        if (x->e[j] != (int8_t)x->k[j]) { fprintf(stderr, "[salt_aln] XA CIGAR: edit distance changed (%d != %d)\n", x->e[j], x->k[j]); exit(1); }  /* sam.c:219-223 */
        x->ptr[j] = x->cig + j * XA_STRIDE;
    }
}

/* The shares of a batch are written by a thread of their own while the next batch is under way: two sets of buffers, the
 * main thread formats into one while the writer drains the other. */
typedef struct {
    pthread_t th; pthread_mutex_t mu; pthread_cond_t cv;
    outbuf_t *set[2]; int n_thr;
    int job, busy, quit, failed;          /* job: the set being written while busy */
    double secs;
} writer_t;
static void *writer_main(void *p)
{
    writer_t *w = (writer_t *)p;
    pthread_mutex_lock(&w->mu);
    for (;;) {
        while (!w->busy && !w->quit) pthread_cond_wait(&w->cv, &w->mu);
        if (!w->busy) break;
        outbuf_t *ob = w->set[w->job];
        pthread_mutex_unlock(&w->mu);
        const double t0 = now();
        int bad = 0;
        for (int t = 0; t < w->n_thr; ++t) {
            if (ob[t].len && fwrite(ob[t].s, 1, ob[t].len, stdout) != ob[t].len) bad = 1;
            ob[t].len = 0;
        }
        const double dt = now() - t0;
        pthread_mutex_lock(&w->mu);
        w->secs += dt; w->failed |= bad; w->busy = 0;
        pthread_cond_broadcast(&w->cv);
    }
    pthread_mutex_unlock(&w->mu);
    return NULL;
}
static void writer_start(writer_t *w, int n_thr)
{
    memset(w, 0, sizeof *w);
    pthread_mutex_init(&w->mu, NULL); pthread_cond_init(&w->cv, NULL);
    w->n_thr = n_thr;
    for (int k = 0; k < 2; ++k) w->set[k] = (outbuf_t *)xcalloc((size_t)n_thr, sizeof(outbuf_t));
    if (pthread_create(&w->th, NULL, writer_main, w) != 0) { fprintf(stderr, "[salt_aln] cannot start the writer thread\n"); exit(1); }
}
/* hand set `k` to the writer (after it has finished the other one) and return the set to format into next */
static int writer_submit(writer_t *w, int k)
{
    pthread_mutex_lock(&w->mu);
    while (w->busy) pthread_cond_wait(&w->cv, &w->mu);
    if (w->failed) { fprintf(stderr, "[salt_aln] write error\n"); exit(1); }
    w->job = k; w->busy = 1;
    pthread_cond_broadcast(&w->cv);
    pthread_mutex_unlock(&w->mu);
    return k ^ 1;
}
static void writer_finish(writer_t *w, stats_t *st)
{
    pthread_mutex_lock(&w->mu);
    while (w->busy) pthread_cond_wait(&w->cv, &w->mu);
    w->quit = 1;
    pthread_cond_broadcast(&w->cv);
    pthread_mutex_unlock(&w->mu);
    pthread_join(w->th, NULL);
    if (w->failed) { fprintf(stderr, "[salt_aln] write error\n"); exit(1); }
    st->write += w->secs;
}

typedef struct {
    const opts_t *o; const index_files_t *ix; const reader_t *rd; const salt_chunk_t *ck; const uint8_t *codes; const int32_t *row_of;
    const salt_read_result_t *res; const xa_t *xa; outbuf_t *ob; uint32_t n; size_t md_tags; int failed;
    /* tags of the chunk's primaries as salt_b200_tail_primaries returns them (row = the read's row in the chunk); NULL: salt_chunk_md */
    const salt_mdnm_out_t *t_out; const uint32_t *t_offs; const char *t_md; const uint16_t *t_xv;
} se_text_t;
static void se_text_share(int t, int T, void *arg)
{
    se_text_t *a = (se_text_t *)arg;
    const opts_t *o = a->o; const reader_t *rd = a->rd; const salt_fastq_t *q = &rd->fq;
    const uint32_t lo = (uint32_t)((uint64_t)a->n * t / T), hi = (uint32_t)((uint64_t)a->n * (t + 1) / T);
    outbuf_t *b = &a->ob[t];
    size_t md_tags = 0; int failed = 0;
    for (uint32_t i = lo; i < hi && !failed; ++i) {
        if (a->row_of[i] < 0) { *out_room(b, 1) = '\n'; b->len += 1; continue; }
        const salt_read_result_t *r = a->res + a->row_of[i];
        salt_sam_read_t s;
        memset(&s, 0, sizeof s);
        s.name = rd->text + q->name_off[i]; s.seq = a->codes + rd->roffs[i];
        s.qual = q->qual_off[i] != 0xFFFFFFFFu ? rd->text + q->qual_off[i] : NULL;
        s.l_seq = q->lens[i]; s.pos = r->pos; s.strand = r->strand; s.mapq = r->mapq & 255; s.cigar = r->cigar;
        s.seq_start = 0; s.seq_end = s.l_seq ? s.l_seq - 1 : 0;
        for (int k = 0; k < 2; ++k) { s.n_alt[k] = r->n_alt[k]; s.alt[k] = r->alt[k]; }
        s.xa_cigars = a->xa->ptr ? a->xa->ptr + a->xa->first[i] : NULL;
        if (o->print_nm_md && r->pos != UNMAPPED) {
            int nm = 0, n_xv = 0; const uint16_t *xv = NULL;
            if (a->t_out) {
                const salt_mdnm_out_t *to = &a->t_out[a->row_of[i]];
                if (to->md_len < 0) { failed = 1; break; }                  /* -2 / -3: see salt_mdnm_out_t */
                s.md = a->t_md + a->t_offs[a->row_of[i]]; nm = to->nm; n_xv = to->n_xv;
                xv = n_xv ? a->t_xv + (size_t)a->row_of[i] * XV_STRIDE : NULL;
            } else s.md = salt_chunk_md(a->ck, (uint32_t)a->row_of[i], &nm, &xv, &n_xv);
            if (!s.md) { failed = 1; break; }
            s.nm = (uint32_t)nm; s.xv = xv; s.n_xv = n_xv; ++md_tags;
        }
        size_t room = (size_t)s.l_seq * 2 + 4096;
        for (;;) {
            char *w = out_room(b, room + 1);
            const int len = salt_sam_se(&a->ix->refs, &s, o->print_xa_cigar, o->rg_id, w, room);
            if (len == SALT_ERR_NOMEM && room < ((size_t)1 << 26)) { room *= 4; continue; }
            if (len < 0) { failed = 1; break; }
            w[len] = '\n'; b->len += (size_t)len + 1;                        /* the reference prints the line with puts */
            break;
        }
    }
    __atomic_fetch_add(&a->md_tags, md_tags, __ATOMIC_RELAXED);
    if (failed) __atomic_store_n(&a->failed, 1, __ATOMIC_RELAXED);
}

/* ---- single-end ----------------------------------------------------------------------------------------------------------- */
static void run_se(const opts_t *o, const index_files_t *ix, salt_b200_t *h, stats_t *st)
{
    stream_t in;
    stream_open(&in, o->fn[0], o->batch);
    salt_seed_opt_t so;
    memset(&so, 0, sizeof so);
    so.l_seed = ix->l_seed; so.l_overlap = o->l_overlap > 0 ? o->l_overlap : ix->l_seed; so.max_seed = o->max_seed;
    so.max_locate = o->max_locate; so.seed_only_ref = o->seed_only_ref;
    const uint32_t B = o->batch;
    size_t cap_bases = (size_t)B * 160, cap_cands = (size_t)B * 64;
    salt_chunk_t *ck = salt_chunk_new(B + 8, cap_bases, cap_cands);
    if (!ck) die("salt_chunk_new", SALT_ERR_NOMEM);
    uint8_t *codes = NULL; size_t codes_cap = 0;
    uint8_t *kcodes = NULL; uint32_t *kroffs = (uint32_t *)xmalloc(((size_t)B + 1) * 4);
    uint32_t *zeros = (uint32_t *)xcalloc((size_t)B + 1, 4);
    int32_t *row_of = (int32_t *)xmalloc((size_t)B * 4);
    uint32_t *primary = (uint32_t *)xmalloc((size_t)B * 4);
    salt_read_result_t *res = (salt_read_result_t *)xmalloc((size_t)B * sizeof *res);
    xa_t xa; memset(&xa, 0, sizeof xa); xa.first = (uint32_t *)xmalloc(((size_t)B + 1) * 4);
    salt_mdnm_out_t *t_out = NULL; uint32_t *t_offs = NULL; uint16_t *t_xv = NULL; char *t_md = NULL; size_t t_md_cap = 0;   /* pinned */
    int tail_fast_ok = 1;
    const int n_thr = o->n_threads;
    writer_t wr; writer_start(&wr, n_thr);
    int cur = 0;
    for (;;) {
        outbuf_t *ob = wr.set[cur];
        double t0 = now();
        uint32_t n = 0;
        const reader_t *rd = stream_get(&in, &n);
        if (!n) break;
        stream_kick(&in, B);                      /* the next batch is read and parsed while this one is aligned */
        const salt_fastq_t *q = &rd->fq;
        const size_t nb = rd->roffs[n];
        if (nb + 16 > codes_cap) { codes_cap = nb * 2 + 16; codes = (uint8_t *)xrealloc(codes, codes_cap); kcodes = (uint8_t *)xrealloc(kcodes, codes_cap); }
        unpack_batch(rd, n, codes, rd->roffs, 1, o->n_threads);
        /* reads with more than MAX_N_PERSEQ ambiguous bases are not aligned; their SAM line stays empty (alnse.c:1296) */
        uint32_t nk = 0; int all = 1;
        for (uint32_t i = 0; i < n; ++i) { row_of[i] = q->n_ambiguous[i] <= MAX_N_PERSEQ_SE ? (int32_t)nk++ : -1; all &= row_of[i] >= 0; }
        const uint8_t *cc = codes; const uint32_t *rr = rd->roffs;
        if (!all) {
            kroffs[0] = 0; nk = 0;
            for (uint32_t i = 0; i < n; ++i) if (row_of[i] >= 0) {
                memcpy(kcodes + kroffs[nk], codes + rd->roffs[i], q->lens[i]);
                kroffs[nk + 1] = kroffs[nk] + q->lens[i]; ++nk;
            }
            cc = kcodes; rr = kroffs;
        }
        st->parse += now() - t0; t0 = now();
        for (;;) {
            if (rr[nk] + 1024 > cap_bases) {
                cap_bases = (size_t)rr[nk] * 2 + 1024; salt_chunk_free(ck); ck = salt_chunk_new(B + 8, cap_bases, cap_cands);
                if (!ck) die("salt_chunk_new", SALT_ERR_NOMEM);
            }
            salt_chunk_reset(ck);
            int rc = nk ? salt_chunk_add_reads(ck, cc, rr, nk, zeros, NULL, zeros, NULL) : 0;
            if (rc < 0) die("salt_chunk_add_reads", rc);
            rc = salt_chunk_seed_verify(h, ck, &so, 3, -1);                          /* alnse.c:1079 / :1090 thresholds */
            if (rc == SALT_ERR_NOMEM && cap_cands < ((size_t)1 << 31)) {             /* longer candidate lists than the chunk has room for */
                cap_cands *= 4; salt_chunk_free(ck); ck = salt_chunk_new(B + 8, cap_bases, cap_cands);
                if (!ck) die("salt_chunk_new", SALT_ERR_NOMEM);
                continue;
            }
            if (rc != SALT_OK) die("salt_chunk_seed_verify", rc);
            break;
        }
        st->gpu += now() - t0; t0 = now();
        int fast_tail = 0;
        if (o->print_nm_md && nk) {
            /* sam_add_md_nm for every primary of the chunk (sam.c:246-328).  The records, strands and CIGARs of the chunk just
               verified are still on the device: salt_b200_tail_primaries makes the tags from them and sends the MD strings back
               packed.  Should it decline, salt_chunk_tail builds the same tags item by item. */
            if (!t_out) {
                t_out = (salt_mdnm_out_t *)salt_b200_host_alloc((size_t)B * sizeof *t_out); t_offs = (uint32_t *)salt_b200_host_alloc(((size_t)B + 1) * 4);
                t_xv = (uint16_t *)salt_b200_host_alloc((size_t)B * XV_STRIDE * 2);
                t_md_cap = (size_t)B * 48; t_md = (char *)salt_b200_host_alloc(t_md_cap);
            }
            if (tail_fast_ok && t_out && t_offs && t_xv && t_md) {
                size_t bytes = 0;
                int rc = salt_b200_tail_primaries(h, 0, t_out, t_offs, t_md, t_md_cap, &bytes, t_xv, XV_STRIDE);
                if (rc == SALT_ERR_NOMEM && bytes > t_md_cap) {              /* longer MD strings than the buffer: once more with room */
                    salt_b200_host_free(t_md);
                    t_md_cap = bytes + bytes / 4 + 4096; t_md = (char *)salt_b200_host_alloc(t_md_cap);
                    rc = t_md ? salt_b200_tail_primaries(h, 0, t_out, t_offs, t_md, t_md_cap, &bytes, t_xv, XV_STRIDE) : SALT_ERR_NOMEM;
                }
                if (rc == SALT_OK) fast_tail = 1;
                else { tail_fast_ok = 0; fprintf(stderr, "[salt_aln] salt_b200_tail_primaries declined (%d: %s): tags through salt_chunk_tail\n", rc, salt_b200_last_error()); }
            }
            if (!fast_tail) { const int rc = salt_chunk_tail(h, 0, ck); if (rc != SALT_OK) die("salt_chunk_tail", rc); }
        }
        st->tail += now() - t0; t0 = now();
        if (nk) { const int rc = salt_chunk_results(ck, MAX_HITS, res); if (rc != SALT_OK) die("salt_chunk_results", rc); }
        for (uint32_t i = 0; i < n; ++i) primary[i] = row_of[i] >= 0 ? res[row_of[i]].pos : UNMAPPED;
        st->select += now() - t0; t0 = now();
        if (o->print_xa_cigar) {
            /* an unmapped read has no alternates; for the others sam_add_xa skips the hit at the primary's position */
            xa_collect(&xa, h, res, row_of, primary, n);
            st->xa_cigars += xa.n;
        } else memset(xa.first, 0, ((size_t)n + 1) * 4);
        st->tail += now() - t0; t0 = now();
        se_text_t ta = {o, ix, rd, ck, codes, row_of, res, &xa, ob, n, 0, 0, fast_tail ? t_out : NULL, t_offs, t_md, t_xv};
        run_shares(n_thr, se_text_share, &ta);
        st->md_tags += ta.md_tags;
        const int failed = ta.failed;
        if (failed) die("salt_sam_se / salt_chunk_md", SALT_ERR_ARG);
        st->text += now() - t0; t0 = now();
        cur = writer_submit(&wr, cur);
        st->write_wait += now() - t0;
        st->reads += n;
    }
    writer_finish(&wr, st);
    st->parse_beside += stream_close(&in);
    salt_chunk_free(ck);
}

typedef struct {
    const opts_t *o; const index_files_t *ix; const reader_t *const *rd; const uint8_t *codes; const uint32_t *roffs; const salt_pair_final_t *fin;
    const salt_read_result_t *res; const xa_t *xa; const int32_t *tag_row; const char *tmd; const uint16_t *txv; const salt_mdnm_out_t *tout;
    outbuf_t *ob; uint32_t np; int failed;
} pe_text_t;
static void pe_text_share(int t, int T, void *arg)
{
    pe_text_t *a = (pe_text_t *)arg;
    const opts_t *o = a->o;
    const uint32_t lo = (uint32_t)((uint64_t)a->np * t / T), hi = (uint32_t)((uint64_t)a->np * (t + 1) / T);
    outbuf_t *b = &a->ob[t];
    int failed = 0;
    for (uint32_t p = lo; p < hi && !failed; ++p) {
        salt_sam_read_t s[2];
        memset(s, 0, sizeof s);
        for (int m = 0; m < 2; ++m) {
            const uint32_t i = 2 * p + (uint32_t)m;
            const reader_t *rd = a->rd[m]; const salt_fastq_t *q = &rd->fq;
            const salt_mate_final_t *f = &a->fin[p].mate[m];
            const salt_read_result_t *r = a->res + i;
            s[m].name = rd->text + q->name_off[p]; s[m].seq = a->codes + a->roffs[i];
            s[m].qual = q->qual_off[p] != 0xFFFFFFFFu ? rd->text + q->qual_off[p] : "";
            s[m].l_seq = q->lens[p]; s[m].pos = f->pos; s[m].strand = f->strand; s[m].mapq = f->mapq & 255; s[m].cigar = f->cigar;
            s[m].seq_start = f->seq_start; s[m].seq_end = f->seq_end;
            for (int k = 0; k < 2; ++k) { s[m].n_alt[k] = r->n_alt[k]; s[m].alt[k] = r->alt[k]; }
            s[m].xa_cigars = a->xa->ptr ? a->xa->ptr + a->xa->first[i] : NULL;
            if (a->tag_row[i] >= 0) {
                const int32_t j = a->tag_row[i];
                s[m].md = a->tmd + (size_t)j * MD_STRIDE; s[m].nm = (uint32_t)a->tout[j].nm;
                s[m].n_xv = a->tout[j].n_xv; s[m].xv = a->tout[j].n_xv ? a->txv + (size_t)j * XV_STRIDE : NULL;
            }
        }
        size_t room = (size_t)(s[0].l_seq + s[1].l_seq) * 2 + 4096;
        for (;;) {
            /* two lines back to back: each ends in a newline as alnpe_sam leaves it, and the reference prints it with
               "%s\n" (alnpe.c:620), so a blank line follows each */
            char *w = out_room(b, 2 * room + 4);
            int len[2] = {0, 0};
            const int rc2 = salt_sam_pe(&a->ix->refs, s, o->min_tlen, o->max_tlen, o->print_xa_cigar, o->rg_id, w, room, w + room + 2, room, len);
            if (rc2 == SALT_ERR_NOMEM && room < ((size_t)1 << 26)) { room *= 4; continue; }
            if (rc2 < 0) { failed = 1; break; }
            w[len[0]] = '\n';
            memmove(w + len[0] + 1, w + room + 2, (size_t)len[1]);
            w[len[0] + 1 + len[1]] = '\n';
            b->len += (size_t)len[0] + (size_t)len[1] + 2;
            break;
        }
    }
    if (failed) __atomic_store_n(&a->failed, 1, __ATOMIC_RELAXED);
}

/* ---- paired-end ----------------------------------------------------------------------------------------------------------- */
static void run_pe(const opts_t *o, const index_files_t *ix, salt_b200_t *h, stats_t *st)
{
    const uint32_t P = o->batch / 2 ? o->batch / 2 : 1;          /* pairs per batch: N_SEQS reads (query_read_multiPairedSeqs) */
    stream_t in[2];
    stream_open(&in[0], o->fn[0], P); stream_open(&in[1], o->fn[1], P);
    int last = 0;
    salt_seed_opt_t so;
    memset(&so, 0, sizeof so);
    so.l_seed = ix->l_seed; so.l_overlap = o->l_overlap > 0 ? o->l_overlap : ix->l_seed; so.max_seed = o->max_seed;
    so.max_locate = o->max_locate; so.seed_only_ref = o->seed_only_ref; so.locate_mode = 1; so.list_cap = o->list_cap;
    int8_t mat16[256], mat5[25];
    /* salt's two Smith-Waterman matrices by their rule (alnpe.c:52-73): SNP-aware 16 x 16, rows 1 2 4 8 score +1 where the
       column shares the row's bit, everything else -3; 2-bit 5 x 5, +1 / -3, N scores -1 */
    for (int r = 0; r < 16; ++r) for (int c = 0; c < 16; ++c) mat16[r * 16 + c] = (int8_t)(((r == 1 || r == 2 || r == 4 || r == 8) && (c & r)) ? 1 : -3);
    for (int r = 0; r < 5; ++r) for (int c = 0; c < 5; ++c) mat5[r * 5 + c] = (int8_t)((r == 4 || c == 4) ? -1 : r == c ? 1 : -3);
    const uint32_t M = 2 * P;
    size_t cap_bases = (size_t)M * 160, cap_cands = (size_t)M * 64;
    salt_chunk_t *ck = salt_chunk_new(M + 8, cap_bases, cap_cands);
    if (!ck) die("salt_chunk_new", SALT_ERR_NOMEM);
    uint8_t *codes = NULL; size_t codes_cap = 0;
    uint32_t *roffs = (uint32_t *)xmalloc(((size_t)M + 1) * 4);
    uint32_t *offs[2] = {(uint32_t *)xmalloc(((size_t)M + 1) * 4), (uint32_t *)xmalloc(((size_t)M + 1) * 4)};
    size_t loci_cap = cap_cands;
    uint32_t *loci[2] = {(uint32_t *)xmalloc(loci_cap * 4), (uint32_t *)xmalloc(loci_cap * 4)};
    uint8_t *stt[2] = {(uint8_t *)xmalloc(M), (uint8_t *)xmalloc(M)};
    int32_t *row_of = (int32_t *)xmalloc((size_t)M * 4);
    uint32_t *primary = (uint32_t *)xmalloc((size_t)M * 4);
    salt_read_result_t *res = (salt_read_result_t *)xmalloc((size_t)M * sizeof *res);
    salt_pair_final_t *fin = (salt_pair_final_t *)xmalloc((size_t)P * sizeof *fin);
    salt_mdnm_in_t *items = (salt_mdnm_in_t *)xmalloc((size_t)M * sizeof *items);
    salt_mdnm_out_t *tout = (salt_mdnm_out_t *)xmalloc((size_t)M * sizeof *tout);
    int32_t *tag_row = (int32_t *)xmalloc((size_t)M * 4);
    char *tcig = NULL, *tmd = NULL; uint16_t *txv = NULL;
    if (o->print_nm_md) { tcig = (char *)xmalloc((size_t)M * 256); tmd = (char *)xmalloc((size_t)M * MD_STRIDE); txv = (uint16_t *)xmalloc((size_t)M * XV_STRIDE * 2); }
    xa_t xa; memset(&xa, 0, sizeof xa); xa.first = (uint32_t *)xmalloc(((size_t)M + 1) * 4);
    const int n_thr = o->n_threads;
    writer_t wr; writer_start(&wr, n_thr);
    int cur = 0;
    for (uint32_t i = 0; i < M; ++i) row_of[i] = (int32_t)i;
    for (;;) {
        outbuf_t *ob = wr.set[cur];
        double t0 = now();
        if (last) break;
        uint32_t n0 = 0, n1 = 0;
        const reader_t *rd[2];
        rd[0] = stream_get(&in[0], &n0); rd[1] = stream_get(&in[1], &n1);
        if (!n0 || !n1) break;
        if (n0 != n1) {                             /* one file ended first: the surplus of the other is not aligned */
            fprintf(stderr, "[salt_aln] %s and %s hold different numbers of records: stopping after the last complete pair\n", o->fn[0], o->fn[1]);
            n0 = n0 < n1 ? n0 : n1; last = 1;
        } else { stream_kick(&in[0], P); stream_kick(&in[1], P); }
        const uint32_t np = n0, n = 2 * np;
        roffs[0] = 0;
        for (uint32_t p = 0; p < np; ++p) {
            roffs[2 * p + 1] = roffs[2 * p] + rd[0]->fq.lens[p];
            roffs[2 * p + 2] = roffs[2 * p + 1] + rd[1]->fq.lens[p];
        }
        const size_t nb = roffs[n];
        if (nb + 16 > codes_cap) { codes_cap = nb * 2 + 16; codes = (uint8_t *)xrealloc(codes, codes_cap); }
        unpack_batch(rd[0], np, codes, roffs, 2, o->n_threads); unpack_batch(rd[1], np, codes, roffs + 1, 2, o->n_threads);
        st->parse += now() - t0; t0 = now();
        /* phase 1: candidate lists of every mate, both strands (alnse_seed_overlap + alnse_locate, alnse.c:1010-1013) */
        salt_reads_t rs; rs.codes = codes; rs.offs = roffs; rs.n_reads = n;
        int rc = salt_b200_set_reads(h, &rs);
        if (rc != SALT_OK) die("salt_b200_set_reads", rc);
        size_t c0 = 0, c1 = 0;
        for (;;) {
            rc = salt_b200_seed_locate(h, 0, &so, offs[0], offs[1], loci[0], loci_cap, loci[1], loci_cap, &c0, &c1);
            if (rc == SALT_ERR_NOMEM) {
                loci_cap = (c0 > c1 ? c0 : c1) + 1024;
                loci[0] = (uint32_t *)xrealloc(loci[0], loci_cap * 4); loci[1] = (uint32_t *)xrealloc(loci[1], loci_cap * 4);
                rc = salt_b200_seed_locate(h, 0, &so, offs[0], offs[1], loci[0], loci_cap, loci[1], loci_cap, &c0, &c1);
            }
            if (rc != SALT_OK) die("salt_b200_seed_locate", rc);
            rc = salt_b200_seed_status(h, 0, stt[0], stt[1]);
            if (rc != SALT_OK) die("salt_b200_seed_status", rc);
            int cut = 0;                            /* bit 1: a list filled the room it was given before the reference's own limit */
            for (uint32_t i = 0; i < n; ++i) cut |= (stt[0][i] | stt[1][i]) & 2;
            if (!cut || so.list_cap >= 16384) break;
            so.list_cap = so.list_cap * 4 > 16384 ? 16384 : so.list_cap * 4;       /* kept for the batches that follow */
            ++st->relocated;
        }
        for (uint32_t i = 0; i < n; ++i) { st->flagged += ((stt[0][i] | stt[1][i]) & 1) != 0; st->cut += ((stt[0][i] | stt[1][i]) & 2) != 0; }
        /* a mate with more than MAX_N_PERSEQ ambiguous bases is not aligned (alnpe.c:491); it can still be rescued */
        int any_skip = 0;
        for (uint32_t i = 0; i < n; ++i) any_skip |= rd[i & 1]->fq.n_ambiguous[i >> 1] > MAX_N_PERSEQ_PE;
        if (any_skip)
            for (int s = 0; s < 2; ++s) {
                uint32_t w = 0, a = 0;                  /* w: end of the compacted lists so far; a: start of mate i's original list */
                for (uint32_t i = 0; i < n; ++i) {
                    const uint32_t b = offs[s][i + 1];
                    if (rd[i & 1]->fq.n_ambiguous[i >> 1] <= MAX_N_PERSEQ_PE) { memmove(loci[s] + w, loci[s] + a, (size_t)(b - a) * 4); w += b - a; }
                    a = b; offs[s][i + 1] = w;
                }
                if (s) c1 = w; else c0 = w;
            }
        /* phase 2: verification with the paired-end thresholds (alnse.c:1016, :1027), then the pair stage of the whole chunk */
        if (roffs[n] + 1024 > cap_bases || (c0 > c1 ? c0 : c1) + 64 > cap_cands) {
            if (roffs[n] + 1024 > cap_bases) cap_bases = (size_t)roffs[n] * 2 + 1024;
            if ((c0 > c1 ? c0 : c1) + 64 > cap_cands) cap_cands = (c0 > c1 ? c0 : c1) * 2 + 64;
            salt_chunk_free(ck); ck = salt_chunk_new(M + 8, cap_bases, cap_cands);
            if (!ck) die("salt_chunk_new", SALT_ERR_NOMEM);
        }
        salt_chunk_reset(ck);
        rc = salt_chunk_add_reads(ck, codes, roffs, n, offs[0], loci[0], offs[1], loci[1]);
        if (rc < 0) die("salt_chunk_add_reads", rc);
        rc = salt_chunk_submit(h, 0, ck, 3, 3);
        if (rc != SALT_OK) die("salt_chunk_submit", rc);
        rc = salt_chunk_wait(h, 0, ck);
        if (rc != SALT_OK) die("salt_chunk_wait", rc);
        st->gpu += now() - t0; t0 = now();
        salt_pe_stats_t ps; memset(&ps, 0, sizeof ps);
        rc = salt_chunk_pair(h, 0, ck, o->min_tlen, o->max_tlen, ix->l, MAX_HITS, mat16, mat5, 3, 1, 0, 20, 0, fin, NULL, NULL, 0, &ps);
        if (rc != SALT_OK) die("salt_chunk_pair", rc);
        st->pe.pairs += ps.pairs; st->pe.proper += ps.proper; st->pe.windows16 += ps.windows16; st->pe.windows5 += ps.windows5;
        st->pe.rescued += ps.rescued; st->pe.promoted += ps.promoted; st->pe.declined += ps.declined;
        rc = salt_chunk_results(ck, MAX_HITS, res);                                  /* query->hits of every mate: pairing does not change them */
        if (rc != SALT_OK) die("salt_chunk_results", rc);
        st->select += now() - t0; t0 = now();
        /* tags of every mapped mate as it stands after pairing (sam_add_md_nm, sam.c:246-328), one call for the chunk */
        uint32_t n_items = 0;
        for (uint32_t i = 0; i < n; ++i) {
            const salt_mate_final_t *f = &fin[i >> 1].mate[i & 1];
            primary[i] = f->pos; tag_row[i] = -1;
            if (!o->print_nm_md || f->pos == UNMAPPED) continue;
            items[n_items].rs = (i << 1) | (uint32_t)(f->strand & 1); items[n_items].pos = f->pos; items[n_items].seq_start = f->seq_start;
            memcpy(tcig + (size_t)n_items * 256, f->cigar, 256);
            tag_row[i] = (int32_t)n_items++;
        }
        if (n_items) {
            rc = salt_b200_md_nm(h, 0, items, n_items, tcig, 256, tmd, MD_STRIDE, txv, XV_STRIDE, tout);
            if (rc != SALT_OK) die("salt_b200_md_nm", rc);
            for (uint32_t j = 0; j < n_items; ++j)
                if (tout[j].md_len < 0) { fprintf(stderr, "[salt_aln] MD string of mate %u: code %d\n", items[j].rs >> 1, tout[j].md_len); exit(1); }
            st->md_tags += n_items;
        }
        if (o->print_xa_cigar) { xa_collect(&xa, h, res, row_of, primary, n); st->xa_cigars += xa.n; }
        else memset(xa.first, 0, ((size_t)n + 1) * 4);
        st->tail += now() - t0; t0 = now();
        pe_text_t ta = {o, ix, rd, codes, roffs, fin, res, &xa, tag_row, tmd, txv, tout, ob, np, 0};
        run_shares(n_thr, pe_text_share, &ta);
        const int failed = ta.failed;
        if (failed) die("salt_sam_pe", SALT_ERR_ARG);
        st->text += now() - t0; t0 = now();
        cur = writer_submit(&wr, cur);
        st->write_wait += now() - t0;
        st->reads += n;
    }
    writer_finish(&wr, st);
    st->parse_beside += stream_close(&in[0]) + stream_close(&in[1]);
    salt_chunk_free(ck);
}

static int usage(void)
{
    fprintf(stderr, "Usage: salt_aln [-p] [-a min_tlen] [-b max_tlen] [-r seed_step] [-m max_locate] [-s max_seed] [-c] [-d] [-v]\n"
                    "                [-g read_group] [-t threads] [-D device] <index prefix> <reads.fq> [mates.fq]\n"
                    "       (the options of the reference's `salt`, aln.c:138-226; -n -l -e -M -O -E are accepted and, as there, unused; -X 1 is refused)\n");
    return 1;
}

int main(int argc, char **argv)
{
    opts_t o;
    memset(&o, 0, sizeof o);
    o.n_threads = 1; o.l_overlap = -1; o.min_tlen = 250; o.max_tlen = 550; o.max_seed = 50; o.max_locate = 1000; o.batch = N_SEQS; o.list_cap = PE_LIST_CAP;   /* opt_init, aln.c:28-57 */
    int c;
    while ((c = getopt(argc, argv, "t:n:hpa:b:g:es:m:l:cdvr:M:O:E:X:D:B:L:")) >= 0) {
        switch (c) {
        case 't': o.n_threads = atoi(optarg); break;
        case 'p': o.paired = 1; break;
        case 'a': o.min_tlen = (uint32_t)atoi(optarg); break;
        case 'b': o.max_tlen = (uint32_t)atoi(optarg); break;
        case 'g': o.rg_id = optarg; break;
        case 's': o.max_seed = atoi(optarg); break;
        case 'm': o.max_locate = atoi(optarg); break;
        case 'c': o.print_xa_cigar = 1; break;
        case 'd': o.print_nm_md = 1; break;
        case 'v': o.seed_only_ref = 1; break;
        case 'r': o.l_overlap = atoi(optarg); break;
        case 'D': o.device = atoi(optarg); break;
        case 'B': o.batch = (uint32_t)atoi(optarg); break;                       /* reads per batch (tests) */
        case 'L': o.list_cap = atoi(optarg); break;                              /* paired-end: first list room on the device (tests) */
        case 'X':
            if (atoi(optarg) != 0) {              /* EXTEND_SW (aln.h:29): alnse_overlap_sw, alnse.c:1338 */
                fprintf(stderr, "[salt_aln] -X %s: the Smith-Waterman extension of single-end reads is not served -- the reference itself aborts on "
                                "that path (LandauVishkin.c:183)\n", optarg);
                return 1;
            }
            break;
        case 'n': case 'l': case 'e': case 'M': case 'O': case 'E': break;
        default: return usage();
        }
    }
    if (argc - optind < (o.paired ? 3 : 2)) return usage();
    o.prefix = argv[optind]; o.fn[0] = argv[optind + 1]; o.fn[1] = o.paired ? argv[optind + 2] : NULL;
    if (o.n_threads < 1) o.n_threads = 1;
    if (o.batch < 2) o.batch = 2;
    if (o.list_cap < 64) o.list_cap = 64;
    if (o.list_cap > 16384) o.list_cap = 16384;
    if (o.paired && o.max_tlen == 0) { fprintf(stderr, "infer isize func haven't been implemented\n"); return 1; }          /* alnpe.c:583 */
    salt_host_set_threads(o.n_threads);
    unpack_init();
    stats_t st;
    memset(&st, 0, sizeof st);
    const double t_start = now();
    index_files_t ix;
    load_index(o.prefix, &ix);
    st.index = now() - t_start;
    double t0 = now();
    salt_b200_t *h = salt_b200_init(ix.mixref, ix.l, ix.pac, (int64_t)ix.l, o.device);
    if (!h) die("salt_b200_init", SALT_ERR_CUDA);
    int rc = salt_b200_set_index(h, &ix.fm);
    if (rc != SALT_OK) die("salt_b200_set_index", rc);
    st.gpu_init = now() - t0;
    /* aln_samhead (sam.c:55-84) */
    static char obuf[1 << 20];
    setvbuf(stdout, obuf, _IOFBF, sizeof obuf);
    printf("@HD\tVN:ec1fec2\tSO:unsorted\n");
    for (int i = 0; i < ix.refs.n_seqs; ++i) printf("@SQ\tSN:%s\tLN:%d\n", ix.refs.names[i], ix.rec_len[i]);
    printf("@RG\tID:%s\n", o.rg_id ? o.rg_id : "(null)");
    printf("@PG\tID:salt_b200\tPN:salt_aln\tCL:\"");
    for (int i = 0; i < argc; ++i) printf("%s%s", i ? " " : "", argv[i]);
    printf("\"\n");
    if (o.paired) run_pe(&o, &ix, h, &st); else run_se(&o, &ix, h, &st);
    if (fflush(stdout) != 0 || ferror(stdout)) { fprintf(stderr, "[salt_aln] write error\n"); return 1; }
    int rc_end = salt_b200_sync(h);
    if (rc_end != SALT_OK) die("salt_b200_sync", rc_end);
    const int teardown = getenv("SALT_ALN_TEARDOWN") != NULL;     /* orderly release of every device and pinned allocation (leak checkers) */
    if (teardown) salt_b200_destroy(h);
    const double wall = now() - t_start;
    fprintf(stderr, "[salt_aln] %zu reads in %.3f s (%.0f reads/s): index files %.3f, GPU init + uploads %.3f; FASTQ -> codes %.3f (waiting for the reader thread and unpacking; it parsed for %.3f beside), "
                    "seeding + locate + verification %.3f, %s %.3f, tags + XA CIGARs %.3f, SAM text %.3f, waiting for the writer thread %.3f (it wrote for %.3f)  (%d host threads)\n",
            st.reads, wall, (double)st.reads / wall, st.index, st.gpu_init, st.parse, st.parse_beside, st.gpu,
            o.paired ? "pair stage + hit selection" : "hit selection", st.select, st.tail, st.text, st.write_wait, st.write, o.n_threads);
    fprintf(stderr, "[salt_aln] MD/NM/XV tags from the GPU: %zu, XA CIGARs from the GPU: %zu\n", st.md_tags, st.xa_cigars);
    if (o.paired)
        fprintf(stderr, "[salt_aln] pairs %zu: proper without rescue %zu, rescue windows %zu SNP-aware + %zu plain, mates rescued %zu, alternates promoted %zu, "
                        "windows declined %zu, mates with an SNP-context interval wider than -m (left out, the reference draws at random) %zu, "
                        "batches located again with more list room %zu, mates with a list cut at 16384 loci %zu\n",
                st.pe.pairs, st.pe.proper, st.pe.windows16, st.pe.windows5, st.pe.rescued, st.pe.promoted, st.pe.declined, st.flagged,
                st.relocated, st.cut);
    if (o.paired && (st.pe.declined || st.cut || st.flagged))
        fprintf(stderr, "[salt_aln] WARNING: %zu rescue windows beyond what the engine serves (traceback band over 512), %zu candidate lists cut, %zu "
                        "intervals the reference samples at random: the mates concerned may be placed differently from the reference program's output\n",
                st.pe.declined, st.cut, st.flagged);
    if (teardown) return 0;
    /* Everything is written.  The process ends here without freeing the device and pinned allocations one by one and without
       the CUDA runtime's exit handlers: the driver reclaims all of it with the process, in a fraction of the time. */
    fflush(NULL);
    _exit(0);
}
