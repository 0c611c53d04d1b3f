/*
 * oracle/dropin/alnpe_core_gpu.c -- TEST INFRASTRUCTURE / integration proof (paired-end).
 *
 * Replaces ONE function of the reference, alnpe_core (alnpe.c:530-661; alnpe.c is compiled with
 * -Dalnpe_core=alnpe_core_reference).  Per chunk of pairs:
 *   phase 1  seeding + locate of every mate with the reference's own alnse_seed_overlap /
 *            alnse_locate (the first half of alnse_overlap, alnse.c:1001-1004)
 *   phase 2  the verification stage of all mates on the GPU with the paired-end thresholds
 *            (nogap 3, gapped 3: alnse.c:1016,1027) + query_set_hits -- salt_host.h
 *   phase 3  the reference's own pairing2 / pairing_singleton and alnpe_sam, with the mate-rescue
 *            Smith-Waterman of the whole chunk done in ONE salt_b200_ssw call per flavour:
 *            alnpe.c is also compiled with its calls to ssw_init / ssw_align / init_destroy /
 *            align_destroy renamed to the hooks below.  Pass 1 runs pairing with hooks that only RECORD
 *            which (mate, strand, window) a rescue asks for and answer "not found" (the queries are
 *            snapshotted and restored around it); the recorded windows go to the GPU; pass 2 runs pairing
 *            again with hooks that REPLAY the GPU's s_align records, so every decision and every field
 *            update after a rescue is still the reference's own code.  Window coordinates are not visible
 *            to the hooks; they are recomputed from the anchor mate with pairing2's / pairing_singleton's
 *            arithmetic (alnpe.c:206-252, :413-466) and checked byte for byte against the window the
 *            reference unpacked -- a request that does not check, or that the engine declines (window
 *            touching the end of the reference, CIGAR longer than the stride), runs the reference's own
 *            ssw_align instead and is counted.
 */
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "kvec.h"
#include "aln.h"
#include "sam.h"
#include "ssw.h"
#include "salt_host.h"

#define DROPIN_PE_MAX_N_PERSEQ 5          /* alnpe.c:481 */
#define DROPIN_CHUNK_READS 20000u
#define DROPIN_CHUNK_CANDS (20000u * 256u)
#define DROPIN_CIG_STRIDE 64

int pairing2(index_t *index, query_t *q0, query_t *q1, const aln_opt_t *aln_opt);
int pairing_singleton(index_t *index, query_t *q0, query_t *q1, aln_opt_t *aln_opt);
void alnpe_sam(index_t *index, query_t *q, const aln_opt_t *opt);
void dropin_tail_prepare(salt_b200_t *gpu, int slot, const query_t *multi_seqs, const int *slot_of, int first, int upto);
void dropin_tail_report(void);
extern int8_t score_mat[25], score_mat2[256];       /* alnpe.c:52-73 */

static void die(const char *what)
{
    fprintf(stderr, "[salt_dropin/pe] %s: %s\n", what, salt_b200_last_error());
    exit(1);
}

/* ------------------------------------------------------------------ rescue hooks */
typedef struct { const int8_t *read; int32_t len; const int8_t *mat; int32_t n; s_profile *real; } hook_profile_t;
typedef struct {
    int pair, mate, strand, flavour;       /* flavour 16 = mixRef masks (snpaln_sw_snpaware), 5 = pac (snpaln_sw) */
    uint32_t start, end;
    int on_gpu;                             /* 0 = the reference's own ssw_align serves it */
} rescue_req_t;

enum { MODE_RECORD = 1, MODE_REPLAY = 2, MODE_DIRECT = 3 };     /* DIRECT: the hooks pass straight through to the reference's ssw_align */
static struct {
    int mode, func;                         /* func 2 = pairing2, 1 = pairing_singleton */
    index_t *index; const aln_opt_t *opt;
    query_t *q[2]; int pair; int chunk_idx[2];
    rescue_req_t *req; size_t n_req, m_req, cursor;
    salt_ssw_out_t *res; uint32_t *cig; size_t *res_of;     /* GPU results, indexed through res_of[request] */
    size_t n_cpu;
} H;

static int identify(const hook_profile_t *p, int *mate, int *strand)
{
    int m, s, i;
    for (m = 0; m < 2; ++m)
        for (s = 0; s < 2; ++s) {
            const uint8_t *seq = s ? H.q[m]->rseq : H.q[m]->seq;
            if (H.q[m]->l_seq != p->len) continue;
            if (p->n == 5) { if ((const uint8_t *)p->read == seq) { *mate = m; *strand = s; return 1; } continue; }
            for (i = 0; i < p->len; ++i) if ((uint8_t)p->read[i] != (uint8_t)(1u << seq[i])) break;
            if (i == p->len) { *mate = m; *strand = s; return 1; }
        }
    return 0;
}

/* the window pairing2 (func 2) / pairing_singleton (func 1) hands to its rescue call when `a` anchors and `t` is rescued */
static int window_of(const query_t *a, const query_t *t, int t_strand, uint32_t *start, uint32_t *end)
{
    const uint32_t l2 = a->l_seq + t->l_seq, l_pac = H.index->bntseq->l_pac;
    const uint32_t min_isize = H.opt->min_tlen > l2 ? H.opt->min_tlen - l2 : 0;
    const uint32_t max_isize = H.opt->max_tlen > l2 ? H.opt->max_tlen - l2 : 0;
    uint32_t s, e;
    if (a->pos == 0xFFFFFFFF) return 0;
    if (a->strand == 0) {                   /* anchor forward: mate expected downstream, on the other strand */
        if (t_strand != 1) return 0;
        s = a->pos + min_isize + a->l_seq;
        e = a->pos + max_isize + a->l_seq + t->l_seq;
    } else {
        if (t_strand != 0) return 0;
        s = a->pos > max_isize + t->l_seq ? a->pos - max_isize - t->l_seq : 0;
        e = a->pos > min_isize ? a->pos - min_isize : 0;
    }
    if (H.func == 2) { e = e >= l_pac ? l_pac : e; }
    else { s = s < l_pac - 1 ? s : l_pac - 1; e = e < l_pac - 1 ? e : l_pac - 1; }
    *start = s; *end = e;
    return 1;
}

static int window_matches(int flavour, uint32_t start, uint32_t end, const int8_t *ref, int32_t refLen)
{
    uint32_t i;
    if (end < start || (int64_t)end - start + 1 != refLen) return 0;
    if (end >= H.index->mixRef->l) return 0;             /* the engine does not read past the reference */
    for (i = 0; i < (uint32_t)refLen; ++i) {
        const uint32_t p = start + i;
        const int sym = flavour == 16 ? (int)((H.index->mixRef->seq[p >> 3] >> (4 * (p & 7))) & 15)
                                      : (int)((H.index->pac[p >> 2] >> ((~p & 3) << 1)) & 3);
        if (sym != ref[i]) return 0;
    }
    return 1;
}

s_profile *dropin_ssw_init(const int8_t *read, const int32_t readLen, const int8_t *mat, const int32_t n, const int8_t score_size)
{
    hook_profile_t *p = calloc(1, sizeof *p);
    p->read = read; p->len = readLen; p->mat = mat; p->n = n;
    p->real = ssw_init(read, readLen, mat, n, score_size);          /* cheap; only used when a request falls back */
    return (s_profile *)p;
}

void dropin_init_destroy(s_profile *pp)
{
    hook_profile_t *p = (hook_profile_t *)pp;
    init_destroy(p->real);
    free(p);
}

void dropin_align_destroy(s_align *a) { if (a) { free(a->cigar); free(a); } }

s_align *dropin_ssw_align(const s_profile *pp, const int8_t *ref, int32_t refLen, const uint8_t gapO, const uint8_t gapE,
                          const uint8_t flag, const uint16_t filters, const int32_t filterd, const int32_t maskLen)
{
    const hook_profile_t *p = (const hook_profile_t *)pp;
    int mate = 0, strand = 0;
    uint32_t start = 0, end = 0;
    const int known = identify(p, &mate, &strand) && window_of(H.q[1 - mate], H.q[mate], strand, &start, &end) &&
                      window_matches(p->n, start, end, ref, refLen) && H.chunk_idx[mate] >= 0;
    if (H.mode == MODE_DIRECT) return ssw_align(p->real, ref, refLen, gapO, gapE, flag, filters, filterd, maskLen);
    if (H.mode == MODE_RECORD) {
        if (H.n_req == H.m_req) { H.m_req = H.m_req ? H.m_req * 2 : 1024; H.req = realloc(H.req, H.m_req * sizeof *H.req); }
        rescue_req_t *r = &H.req[H.n_req++];
        r->pair = H.pair; r->mate = mate; r->strand = strand; r->flavour = p->n; r->start = start; r->end = end; r->on_gpu = known;
        s_align *a = calloc(1, sizeof *a);              /* "nothing found": read span 1 < filterd, so the caller moves on */
        return a;
    }
    /* replay: the k-th rescue call of this pair gets the k-th recorded request's answer */
    if (H.cursor >= H.n_req || H.req[H.cursor].pair != H.pair) { fprintf(stderr, "[salt_dropin/pe] replay ran out of requests\n"); exit(1); }
    const rescue_req_t *r = &H.req[H.cursor];
    const size_t ri = H.res_of[H.cursor];
    ++H.cursor;
    if (r->on_gpu && (!known || r->mate != mate || r->strand != strand || r->start != start || r->end != end)) {
        fprintf(stderr, "[salt_dropin/pe] replay diverged from the recorded run\n"); exit(1);
    }
    if (!r->on_gpu || H.res[ri].cigarLen < 0 || H.res[ri].cigarLen > DROPIN_CIG_STRIDE) {
        ++H.n_cpu;
        s_align *a = ssw_align(p->real, ref, refLen, gapO, gapE, flag, filters, filterd, maskLen);
        /* the caller frees through dropin_align_destroy: same layout, same free()s (ssw.c:858-862) */
        return a;
    }
    s_align *a = calloc(1, sizeof *a);
    const salt_ssw_out_t *o = &H.res[ri];
    a->score1 = o->score1; a->score2 = o->score2; a->ref_begin1 = o->ref_begin1; a->ref_end1 = o->ref_end1;
    a->read_begin1 = o->read_begin1; a->read_end1 = o->read_end1; a->ref_end2 = o->ref_end2; a->cigarLen = o->cigarLen;
    if (o->cigarLen > 0) {
        a->cigar = malloc((size_t)o->cigarLen * 4);
        memcpy(a->cigar, H.cig + ri * DROPIN_CIG_STRIDE, (size_t)o->cigarLen * 4);
    }
    return a;
}

/* ------------------------------------------------------------------ query snapshots around the recording pass */
typedef struct {
    uint32_t pos, seq_start, seq_end; int strand, b0, b1; uint8_t n_diff, is_gap, mapq;
    size_t cig_l; char cig[256]; size_t nh[2];
} qsnap_t;

static void snap(const query_t *q, qsnap_t *s)
{
    s->pos = q->pos; s->seq_start = q->seq_start; s->seq_end = q->seq_end; s->strand = q->strand; s->b0 = q->b0; s->b1 = q->b1;
    s->n_diff = q->n_diff; s->is_gap = q->is_gap; s->mapq = q->mapq; s->cig_l = q->cigar->l;
    memcpy(s->cig, q->cigar->s, q->cigar->m < sizeof s->cig ? q->cigar->m : sizeof s->cig);
    s->nh[0] = q->hits[0].n; s->nh[1] = q->hits[1].n;
}

static void unsnap(query_t *q, const qsnap_t *s)
{
    q->pos = s->pos; q->seq_start = s->seq_start; q->seq_end = s->seq_end; q->strand = s->strand; q->b0 = s->b0; q->b1 = s->b1;
    q->n_diff = s->n_diff; q->is_gap = s->is_gap; q->mapq = s->mapq; q->cigar->l = s->cig_l;
    memcpy(q->cigar->s, s->cig, q->cigar->m < sizeof s->cig ? q->cigar->m : sizeof s->cig);
    q->hits[0].n = s->nh[0]; q->hits[1].n = s->nh[1];
}

/* what alnse_overlap leaves in query_t after its checks and query_set_hits (alnse.c:1014-1036) */
static void apply_result(query_t *query, const aln_opt_t *aln_opt, const salt_chunk_t *ck, uint32_t i)
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
}

static void run_pairing(index_t *index, query_t *q0, query_t *q1, aln_opt_t *aln_opt)      /* alnpe.c:507-517 */
{
    if (q0->pos != 0xFFFFFFFF && q1->pos != 0xFFFFFFFF) { H.func = 2; pairing2(index, q0, q1, aln_opt); }
    else if (q0->pos != 0xFFFFFFFF || q1->pos != 0xFFFFFFFF) { H.func = 1; pairing_singleton(index, q0, q1, aln_opt); }
}

/* ---- salt_pair_plan (include/salt_host.h) checked against the reference's own pairing, pair by pair ---- */
static size_t plan_checked, plan_paired, plan_windows, plan_mismatch, plan_used;
static size_t apply_checked, apply_skipped, apply_mismatch, apply_rescued, apply_swapped, apply_final, apply_fallback;
static const salt_chunk_t *H_ck;          /* the chunk whose results the pairs under way came from */
static salt_b200_t *H_gpu;

static void result_of_query(const query_t *q, salt_read_result_t *r)       /* the query_t fields the verification stage set */
{
    int s; size_t i;
    memset(r, 0, sizeof *r);
    r->pos = q->pos; r->strand = (uint8_t)q->strand; r->n_diff = q->n_diff; r->is_gap = q->is_gap;
    r->b0 = q->b0; r->b1 = q->b1; r->mapq = q->mapq;
    for (s = 0; s < 2; ++s) {
        r->n_alt[s] = (int)q->hits[s].n;
        for (i = 0; i < q->hits[s].n && i < SALT_MAX_HITS; ++i) {
            const hit_t *h = q->hits[s].a + i;
            r->alt[s][i].pos = h->pos; r->alt[s][i].n_diff = h->n_diff; r->alt[s][i].is_gap = h->is_gap; r->alt[s][i].strand = h->strand;
        }
    }
}

static int plan_of(const query_t *q0, const query_t *q1, const aln_opt_t *opt, uint32_t l_pac, salt_pair_plan_t *plan)
{
    salt_read_result_t r0, r1;
    result_of_query(q0, &r0); result_of_query(q1, &r1);
    return salt_pair_plan(&r0, (uint32_t)q0->l_seq, &r1, (uint32_t)q1->l_seq, (uint32_t)opt->min_tlen, (uint32_t)opt->max_tlen, l_pac, plan);
}

static void check_plan(const salt_pair_plan_t *plan, int plan_rc, const query_t *q0, const query_t *q1, size_t req0)
{
    const size_t n_req = H.n_req - req0;
    int ok = plan_rc == SALT_OK;
    int w;
    ++plan_checked;
    if (plan->paired) {
        const query_t *q[2] = {q0, q1};
        ++plan_paired;
        ok = ok && n_req == 0;
        for (w = 0; w < 2; ++w)
            ok = ok && q[w]->pos == plan->hit[w].pos && q[w]->strand == (int)plan->hit[w].strand &&
                 q[w]->n_diff == plan->hit[w].n_diff && q[w]->is_gap == plan->hit[w].is_gap;
    } else {
        plan_windows += (size_t)plan->n_win;
        ok = ok && n_req == (size_t)plan->n_win;
        for (w = 0; ok && w < plan->n_win; ++w) {
            const rescue_req_t *r = &H.req[req0 + w];
            ok = r->mate == plan->win[w].mate && r->strand == plan->win[w].strand && r->flavour == plan->win[w].flavour &&
                 r->start == plan->win[w].start && r->end == plan->win[w].end;
        }
    }
    if (!ok) {
        if (plan_mismatch < 5) fprintf(stderr, "[salt_dropin/pe] salt_pair_plan disagrees with the reference on %s (rc %d, paired %d, windows %d, recorded %zu)\n",
                                       q0->name, plan_rc, plan->paired, plan->n_win, n_req);
        ++plan_mismatch;
    }
}

/* ---- salt_pair_apply checked against where the reference's pairing ends, pair by pair ---- */
static int results_for_apply(const query_t *q0, const query_t *q1, int ci0, int ci1, salt_read_result_t *r0, salt_read_result_t *r1)
{
    const query_t *q[2] = {q0, q1};
    salt_read_result_t *r[2] = {r0, r1};
    const int ci[2] = {ci0, ci1};
    int m;
    for (m = 0; m < 2; ++m) {
        result_of_query(q[m], r[m]);
        if (q[m]->pos != 0xFFFFFFFF && q[m]->is_gap) {          /* the verification stage's CIGAR of a gapped primary */
            salt_read_result_t full;
            if (ci[m] < 0 || salt_chunk_result(H_ck, (uint32_t)ci[m], H.opt->max_hits, &full) != SALT_OK) return 0;
            memcpy(r[m]->cigar, full.cigar, sizeof full.cigar);
        }
    }
    return 1;
}

static void check_apply(const salt_pair_plan_t *plan, const salt_read_result_t *r0, const salt_read_result_t *r1,
                        const query_t *q0, const query_t *q1, size_t k0, const aln_opt_t *opt)
{
    const query_t *q[2] = {q0, q1};
    salt_ssw_out_t ssw[2];
    uint32_t cig[2 * DROPIN_CIG_STRIDE];
    salt_mate_final_t fin[2];
    int w, m, ok = 1;
    for (w = 0; w < plan->n_win; ++w) {                           /* the GPU answers of this pair's windows, in plan order */
        const size_t k = k0 + (size_t)w;
        if (k >= H.n_req || H.req[k].pair != H.pair || !H.req[k].on_gpu) { ++apply_skipped; return; }
        const size_t ri = H.res_of[k];
        if (H.res[ri].cigarLen < 0 || H.res[ri].cigarLen > DROPIN_CIG_STRIDE) { ++apply_skipped; return; }
        ssw[w] = H.res[ri];
        memcpy(cig + (size_t)w * DROPIN_CIG_STRIDE, H.cig + ri * DROPIN_CIG_STRIDE, DROPIN_CIG_STRIDE * 4);
    }
    const int rc = salt_pair_apply(plan, r0, (uint32_t)q0->l_seq, r1, (uint32_t)q1->l_seq, ssw, cig, DROPIN_CIG_STRIDE,
                                   opt->filters, opt->filterd, fin);
    ++apply_checked;
    ok = rc >= 0;
    for (m = 0; ok && m < 2; ++m) {
        const salt_mate_final_t *f = &fin[m];
        ok = q[m]->pos == f->pos && (f->pos == 0xFFFFFFFF || (q[m]->strand == (int)f->strand && q[m]->n_diff == f->n_diff &&
             q[m]->is_gap == f->is_gap && q[m]->seq_start == f->seq_start && q[m]->seq_end == f->seq_end)) &&
             q[m]->b0 == f->b0 && q[m]->b1 == f->b1 && q[m]->mapq == (uint8_t)f->mapq;
        if (!ok || f->pos == 0xFFFFFFFF) continue;
        if (f->cigar_kind == 3) ++apply_rescued;
        if (f->cigar_kind == 4) {                                 /* a gapped alternate became the primary: one LV+CIGAR item */
            salt_pair_t p; uint8_t k = f->n_diff; int8_t e; char buf[256];
            ++apply_swapped;
            p.rs = ((uint32_t)H.chunk_idx[m] << 1) | f->strand; p.pos = f->pos;
            memset(buf, 0, sizeof buf);
            if (H.chunk_idx[m] < 0 || salt_b200_lv_cigar(H_gpu, &p, &k, 1, buf, sizeof buf, &e) != SALT_OK) { ok = 0; continue; }
            ok = strcmp(buf, q[m]->cigar->s) == 0;
        } else ok = strcmp(f->cigar, q[m]->cigar->s) == 0;
    }
    if (!ok) {
        if (apply_mismatch < 5) fprintf(stderr, "[salt_dropin/pe] salt_pair_apply disagrees with the reference on %s\n", q0->name);
        ++apply_mismatch;
    }
}

/* SALT_DROPIN_PLAN=2: write what salt_pair_apply decided into the two query_t; 0 = this pair needs the reference's pairing */
static int apply_to_queries(const salt_pair_plan_t *plan, const salt_read_result_t *r0, const salt_read_result_t *r1,
                            query_t *q0, query_t *q1, size_t k0, const aln_opt_t *opt)
{
    query_t *q[2] = {q0, q1};
    salt_ssw_out_t ssw[2];
    uint32_t cig[2 * DROPIN_CIG_STRIDE];
    salt_mate_final_t fin[2];
    char fresh[2][256];
    int w, m;
    for (w = 0; w < plan->n_win; ++w) {
        const size_t k = k0 + (size_t)w;
        if (k >= H.n_req || H.req[k].pair != H.pair || !H.req[k].on_gpu) return 0;
        const size_t ri = H.res_of[k];
        if (H.res[ri].cigarLen < 0 || H.res[ri].cigarLen > DROPIN_CIG_STRIDE) return 0;
        ssw[w] = H.res[ri];
        memcpy(cig + (size_t)w * DROPIN_CIG_STRIDE, H.cig + ri * DROPIN_CIG_STRIDE, DROPIN_CIG_STRIDE * 4);
    }
    if (salt_pair_apply(plan, r0, (uint32_t)q0->l_seq, r1, (uint32_t)q1->l_seq, ssw, cig, DROPIN_CIG_STRIDE,
                        opt->filters, opt->filterd, fin) < 0) return 0;
    for (m = 0; m < 2; ++m)                                       /* a gapped alternate promoted to primary: its CIGAR */
        if (fin[m].pos != 0xFFFFFFFF && fin[m].cigar_kind == 4) {
            salt_pair_t p; uint8_t k = fin[m].n_diff; int8_t e;
            if (H.chunk_idx[m] < 0) return 0;
            p.rs = ((uint32_t)H.chunk_idx[m] << 1) | fin[m].strand; p.pos = fin[m].pos;
            memset(fresh[m], 0, sizeof fresh[m]);
            if (salt_b200_lv_cigar(H_gpu, &p, &k, 1, fresh[m], sizeof fresh[m], &e) != SALT_OK || e != (int8_t)k) return 0;
        }
    for (m = 0; m < 2; ++m) {
        const salt_mate_final_t *f = &fin[m];
        if (f->pos == 0xFFFFFFFF) continue;                      /* an unmapped mate is left as the verification stage left it */
        q[m]->pos = f->pos; q[m]->strand = f->strand; q[m]->n_diff = f->n_diff; q[m]->is_gap = f->is_gap;
        q[m]->seq_start = f->seq_start; q[m]->seq_end = f->seq_end;
        q[m]->b0 = f->b0; q[m]->b1 = f->b1; q[m]->mapq = (uint8_t)f->mapq;
        q[m]->cigar->l = 0;
        kputs(f->cigar_kind == 4 ? fresh[m] : f->cigar, q[m]->cigar);
    }
    return 1;
}

/* pairs [first, upto) of the chunk: record, one GPU batch per flavour, replay + SAM */
static void pair_and_emit(salt_b200_t *gpu, index_t *index, aln_opt_t *aln_opt, query_t *multi_seqs, const int *slot_of,
                          int first, int upto)
{
    int j;
    size_t k;
    H.index = index; H.opt = aln_opt; H.n_req = 0;
    H.mode = MODE_RECORD;
    /* SALT_DROPIN_PLAN=1: the windows come from salt_pair_plan alone -- the re-staged flow (plan, one GPU batch,
       apply) -- instead of from a recording run of the reference's pairing; the replay below still refuses to
       continue if the reference then asks for anything else */
    const int plan_mode = getenv("SALT_DROPIN_PLAN") ? atoi(getenv("SALT_DROPIN_PLAN")) : 0;
    const int use_plan = plan_mode != 0;
    for (j = first; use_plan && j < upto; j += 2) {
        query_t *q0 = multi_seqs + j, *q1 = multi_seqs + j + 1;
        salt_pair_plan_t plan;
        int w;
        if (plan_of(q0, q1, aln_opt, index->bntseq->l_pac, &plan) != SALT_OK) die("salt_pair_plan");
        ++plan_used;
        for (w = 0; w < plan.n_win; ++w) {
            if (H.n_req == H.m_req) { H.m_req = H.m_req ? H.m_req * 2 : 1024; H.req = realloc(H.req, H.m_req * sizeof *H.req); }
            rescue_req_t *r = &H.req[H.n_req++];
            r->pair = j; r->mate = plan.win[w].mate; r->strand = plan.win[w].strand; r->flavour = plan.win[w].flavour;
            r->start = plan.win[w].start; r->end = plan.win[w].end;
            r->on_gpu = r->end >= r->start && r->end <= index->mixRef->l && slot_of[j + r->mate] >= 0;
        }
    }
    for (j = first; !use_plan && j < upto; j += 2) {
        query_t *q0 = multi_seqs + j, *q1 = multi_seqs + j + 1;
        qsnap_t s0, s1;
        H.q[0] = q0; H.q[1] = q1; H.pair = j; H.chunk_idx[0] = slot_of[j]; H.chunk_idx[1] = slot_of[j + 1];
        snap(q0, &s0); snap(q1, &s1);
        /* the host layer's re-staged pairing (salt_pair_plan) must predict what the reference's pairing2 /
           pairing_singleton is about to do: no rescue and these primaries, or exactly these windows in this order */
        salt_pair_plan_t plan;
        const size_t req0 = H.n_req;
        const int plan_rc = plan_of(q0, q1, aln_opt, index->bntseq->l_pac, &plan);
        run_pairing(index, q0, q1, aln_opt);
        check_plan(&plan, plan_rc, q0, q1, req0);
        unsnap(q0, &s0); unsnap(q1, &s1);
    }
    /* the windows the recording pass asked for, grouped by flavour */
    H.res = realloc(H.res, (H.n_req + 1) * sizeof *H.res);
    H.cig = realloc(H.cig, (H.n_req + 1) * DROPIN_CIG_STRIDE * 4);
    H.res_of = realloc(H.res_of, (H.n_req + 1) * sizeof *H.res_of);
    salt_win_t *wins = malloc((H.n_req + 1) * sizeof *wins);
    size_t n_out = 0;
    int flavour;
    for (flavour = 16; flavour >= 5; flavour -= 11) {
        const size_t base = n_out;
        size_t m = 0;
        for (k = 0; k < H.n_req; ++k) {
            const rescue_req_t *r = &H.req[k];
            if (r->flavour != flavour) continue;
            H.res_of[k] = base + m;
            if (!r->on_gpu) { H.res[base + m].cigarLen = -1; ++m; continue; }
            /* a window the engine may still decline comes back with cigarLen < 0: it gets its own call below */
            ++m;
        }
        /* compact GPU list */
        size_t g = 0, *map = malloc((m + 1) * sizeof *map);
        for (k = 0; k < H.n_req; ++k) {
            const rescue_req_t *r = &H.req[k];
            if (r->flavour != flavour || !r->on_gpu) continue;
            const int ci = slot_of[r->pair + r->mate];
            wins[g].rs = ((uint32_t)ci << 1) | (uint32_t)r->strand; wins[g].start = r->start; wins[g].end = r->end;
            map[g++] = H.res_of[k];
        }
        if (g) {
            salt_ssw_out_t *o = malloc(g * sizeof *o);
            uint32_t *cg = calloc(g * DROPIN_CIG_STRIDE, 4);
            const int rc = salt_b200_ssw(gpu, wins, g, flavour == 5, flavour == 5 ? score_mat : score_mat2, flavour, aln_opt->gap_op,
                                         aln_opt->gap_ex, 2, aln_opt->filters, aln_opt->filterd, -1, o, cg, DROPIN_CIG_STRIDE);
            if (rc != SALT_OK && rc != SALT_ERR_UNSUPPORTED) die("salt_b200_ssw");
            for (k = 0; k < g; ++k) {
                H.res[map[k]] = o[k];
                memcpy(H.cig + map[k] * DROPIN_CIG_STRIDE, cg + k * DROPIN_CIG_STRIDE, DROPIN_CIG_STRIDE * 4);
            }
            free(o); free(cg);
        }
        free(map);
        n_out = base + m;
    }
    free(wins);
    H.mode = MODE_REPLAY; H.cursor = 0;
    for (j = first; j < upto; j += 2) {
        query_t *q0 = multi_seqs + j, *q1 = multi_seqs + j + 1;
        H.q[0] = q0; H.q[1] = q1; H.pair = j; H.chunk_idx[0] = slot_of[j]; H.chunk_idx[1] = slot_of[j + 1];
        /* salt_pair_apply (include/salt_host.h) must end where the reference's pairing ends */
        salt_read_result_t pr0, pr1;
        salt_pair_plan_t plan;
        const size_t k0 = H.cursor;
        const int have_plan = results_for_apply(q0, q1, slot_of[j], slot_of[j + 1], &pr0, &pr1) &&
                              salt_pair_plan(&pr0, (uint32_t)q0->l_seq, &pr1, (uint32_t)q1->l_seq, (uint32_t)aln_opt->min_tlen,
                                             (uint32_t)aln_opt->max_tlen, index->bntseq->l_pac, &plan) == SALT_OK;
        /* SALT_DROPIN_PLAN=2: the reference's pairing is out of the loop -- salt_pair_apply writes the mates' final
           fields from the plan and the GPU's answers; a pair with a window the GPU did not serve goes through the
           reference's pairing with its own Smith-Waterman */
        if (plan_mode == 2) {
            if (have_plan && apply_to_queries(&plan, &pr0, &pr1, q0, q1, k0, aln_opt)) ++apply_final;
            else { H.mode = MODE_DIRECT; run_pairing(index, q0, q1, aln_opt); H.mode = MODE_REPLAY; ++apply_fallback; }
            while (H.cursor < H.n_req && H.req[H.cursor].pair == j) ++H.cursor;
            continue;
        }
        run_pairing(index, q0, q1, aln_opt);
        if (have_plan) check_apply(&plan, &pr0, &pr1, q0, q1, k0, aln_opt);
        else ++apply_skipped;
        if (H.cursor < H.n_req && H.req[H.cursor].pair == j) {
            /* a rescue that succeeds ends pairing early: skip the requests recorded after it */
            while (H.cursor < H.n_req && H.req[H.cursor].pair == j) ++H.cursor;
        }
    }
    /* every query_t of the chunk is final: its MD/NM/XV tags in one GPU call, then the SAM records */
    if (aln_opt->print_nm_md || aln_opt->print_xa_cigar) dropin_tail_prepare(gpu, 0, multi_seqs, slot_of, first, upto);
    for (j = first; j < upto; j += 2) alnpe_sam(index, multi_seqs + j, aln_opt);
}

/* phase 1 on -t worker threads, as alnpe_core runs alnpe_core1 on them (alnpe.c:482-528, :596-606): every worker its own
 * aux_t pair, mates interleaved statically (alnpe.c:487) */
typedef struct { uint32_t *a[2]; uint32_t n[2], m[2]; int skip; } pe_seeded_t;
typedef struct { int tid, n_threads, n; index_t *index; query_t *queries; aln_opt_t *aln_opt; pe_seeded_t *out; aux_t *aux[2]; } pe_seed_thread_t;

static void *pe_seed_worker(void *arg)
{
    pe_seed_thread_t *T = (pe_seed_thread_t *)arg;
    int i, s;
    for (i = T->tid; i < T->n; i += T->n_threads) {
        query_t *query = T->queries + i;
        pe_seeded_t *o = T->out + i;
        o->n[0] = o->n[1] = 0;
        o->skip = query->n_ambiguous > DROPIN_PE_MAX_N_PERSEQ;       /* alnpe.c:495 */
        if (o->skip) continue;
        if (query->l_seq - T->aln_opt->l_seed + 1 > T->aux[0]->n_sai_range) {
            int n_sai_range = query->l_seq - T->aln_opt->l_seed + 1;
            aux_resize(T->aux[0], n_sai_range);
            aux_resize(T->aux[1], n_sai_range);
        }
        aux_reset(T->aux[0]);
        aux_reset(T->aux[1]);
        alnse_seed_overlap(T->index, query->l_seq, query->seq, T->aln_opt, T->aux[0]);
        alnse_locate(T->index, query->l_seq, T->aln_opt->max_locate, T->aux[0]);
        alnse_seed_overlap(T->index, query->l_seq, query->rseq, T->aln_opt, T->aux[1]);
        alnse_locate(T->index, query->l_seq, T->aln_opt->max_locate, T->aux[1]);
        for (s = 0; s < 2; ++s) {
            const uint32_t k = (uint32_t)T->aux[s]->loci.n;
            if (k > o->m[s]) { o->m[s] = k + 16; o->a[s] = realloc(o->a[s], (size_t)o->m[s] * 4); }
            memcpy(o->a[s], T->aux[s]->loci.a, (size_t)k * 4);
            o->n[s] = k;
        }
    }
    return NULL;
}

/* phase 1 on the device (SALT_DROPIN_SEED=gpu): alnse_seed_overlap + alnse_locate of every mate through
 * salt_b200_seed_locate(locate_mode 1), in sub-batches; a strand the engine flags (an SNP-context interval the reference
 * subsamples with rand(), or a list beyond list_cap) gets the reference's own functions on the host instead. */
#define DROPIN_PE_LIST_CAP 4096
static size_t gpu_seed_flagged, gpu_seed_reads;
static void pe_seed_on_gpu(salt_b200_t *gpu, index_t *index, aln_opt_t *aln_opt, query_t *queries, int n, pe_seeded_t *out, aux_t *aux[2])
{
    salt_seed_opt_t so;
    memset(&so, 0, sizeof so);
    so.l_seed = aln_opt->l_seed; so.l_overlap = aln_opt->l_overlap; so.max_seed = aln_opt->max_seed;
    so.max_locate = (int)aln_opt->max_locate; so.seed_only_ref = aln_opt->seed_only_ref;
    so.locate_mode = 1; so.list_cap = DROPIN_PE_LIST_CAP;
    const int step = 20000;
    size_t cap = (size_t)step * 64;
    uint32_t *loci[2] = {malloc(cap * 4), malloc(cap * 4)};
    uint32_t *offs[2] = {malloc(((size_t)step + 1) * 4), malloc(((size_t)step + 1) * 4)};
    uint8_t *st[2] = {malloc((size_t)step), malloc((size_t)step)};
    uint32_t *roffs = malloc(((size_t)step + 1) * 4);
    int *who = malloc((size_t)step * sizeof *who);
    int b, i, s;
    for (b = 0; b < n; b += step) {
        const int e = b + step < n ? b + step : n;
        size_t nb = 0; uint32_t m = 0;
        for (i = b; i < e; ++i) {
            out[i].n[0] = out[i].n[1] = 0;
            out[i].skip = queries[i].n_ambiguous > DROPIN_PE_MAX_N_PERSEQ;
            if (!out[i].skip) nb += (size_t)queries[i].l_seq;
        }
        uint8_t *codes = malloc(nb + 16);
        roffs[0] = 0;
        for (i = b; i < e; ++i) {
            if (out[i].skip) continue;
            memcpy(codes + roffs[m], queries[i].seq, (size_t)queries[i].l_seq);
            roffs[m + 1] = roffs[m] + (uint32_t)queries[i].l_seq;
            who[m++] = i;
        }
        if (m) {
            salt_reads_t r; r.codes = codes; r.offs = roffs; r.n_reads = m;
            if (salt_b200_set_reads(gpu, &r) != SALT_OK) die("salt_b200_set_reads");
            size_t n0 = 0, n1 = 0;
            int rc = salt_b200_seed_locate(gpu, 0, &so, offs[0], offs[1], loci[0], cap, loci[1], cap, &n0, &n1);
            if (rc == SALT_ERR_NOMEM) {
                cap = (n0 > n1 ? n0 : n1) + 1024;
                loci[0] = realloc(loci[0], cap * 4); loci[1] = realloc(loci[1], cap * 4);
                rc = salt_b200_seed_locate(gpu, 0, &so, offs[0], offs[1], loci[0], cap, loci[1], cap, &n0, &n1);
            }
            if (rc != SALT_OK) die("salt_b200_seed_locate");
            if (salt_b200_seed_status(gpu, 0, st[0], st[1]) != SALT_OK) die("salt_b200_seed_status");
            for (uint32_t k = 0; k < m; ++k) {
                pe_seeded_t *o = out + who[k];
                query_t *query = queries + who[k];
                ++gpu_seed_reads;
                if (st[0][k] | st[1][k]) {               /* no single right answer on the device: the reference's own functions */
                    ++gpu_seed_flagged;
                    if (query->l_seq - aln_opt->l_seed + 1 > aux[0]->n_sai_range) {
                        aux_resize(aux[0], query->l_seq - aln_opt->l_seed + 1); aux_resize(aux[1], query->l_seq - aln_opt->l_seed + 1);
                    }
                    aux_reset(aux[0]); aux_reset(aux[1]);
                    alnse_seed_overlap(index, query->l_seq, query->seq, aln_opt, aux[0]);
                    alnse_locate(index, query->l_seq, aln_opt->max_locate, aux[0]);
                    alnse_seed_overlap(index, query->l_seq, query->rseq, aln_opt, aux[1]);
                    alnse_locate(index, query->l_seq, aln_opt->max_locate, aux[1]);
                }
                for (s = 0; s < 2; ++s) {
                    const int flagged = (st[0][k] | st[1][k]) != 0;
                    const uint32_t cnt = flagged ? (uint32_t)aux[s]->loci.n : offs[s][k + 1] - offs[s][k];
                    const uint32_t *src = flagged ? aux[s]->loci.a : loci[s] + offs[s][k];
                    if (cnt > o->m[s]) { o->m[s] = cnt + 16; o->a[s] = realloc(o->a[s], (size_t)o->m[s] * 4); }
                    memcpy(o->a[s], src, (size_t)cnt * 4);
                    o->n[s] = cnt;
                }
            }
        }
        free(codes);
    }
    free(loci[0]); free(loci[1]); free(offs[0]); free(offs[1]); free(st[0]); free(st[1]); free(roffs); free(who);
}

int alnpe_core(const opt_t *opt)
{
    int i;
    aln_opt_t *aln_opt = aln_opt_init(opt);
    const char *seed_env = getenv("SALT_DROPIN_SEED");
    const int gpu_seed = seed_env && !strcmp(seed_env, "gpu");
    fprintf(stderr, "[alnpe_core/gpu]:  Start paired end alignment (verification on libsalt_b200%s)\n",
            gpu_seed ? ", seeding + locate on libsalt_b200" : "");
    index_t *index = alnpe_index_reload(opt->fn_index);
    if (aln_opt->l_overlap <= 0) { fprintf(stderr, "[salt_dropin/pe] only the overlap path is served\n"); exit(1); }
    salt_b200_t *gpu = salt_b200_init(index->mixRef->seq, index->mixRef->l, index->pac, index->bntseq->l_pac, 0);
    if (!gpu) die("salt_b200_init");
    if (gpu_seed) {
        /* the FM-indexes exactly as the reference's loaders left them in memory (indexio.c:52-91) */
        salt_fm_index_t fx;
        int k;
        rbwt_t *r = index->rbwt2->rbwt1;
        memset(&fx, 0, sizeof fx);
        fx.c_bwt = index->cbwt->bwt; fx.c_bwt_words = index->cbwt->bwt_size;
        fx.c_primary = index->cbwt->primary; fx.c_seq_len = index->cbwt->seq_len;
        for (k = 0; k < 5; ++k) fx.c_L2[k] = index->cbwt->L2[k];
        fx.c_sa = index->cbwt->sa; fx.c_n_sa = index->cbwt->n_sa; fx.c_sa_intv = (uint32_t)index->cbwt->sa_intv;
        fx.lkt = index->lkt->item; fx.lkt_len = index->lkt->maxLookupLen;
        fx.r_bwt = r->bwtCode; fx.r_bwt_words = r->bwtSizeInWord;
        fx.r_occ = r->occValue; fx.r_occ_words = r->occSizeInWord;
        fx.r_occ_major = r->occValueMajor; fx.r_occ_major_words = r->occMajorSizeInWord;
        fx.r_sa_sharp = r->saValueSharp; fx.r_n_sa_sharp = r->saValueSizeSharp;
        for (k = 0; k < 6; ++k) fx.r_cum[k] = r->cumulativeFreq[k];
        fx.r_inv_sa0 = r->inverseSa0; fx.r_text_len = r->textLength;
        if (salt_b200_set_index(gpu, &fx) != SALT_OK) die("salt_b200_set_index");
    }
    salt_chunk_t *ck = salt_chunk_new(DROPIN_CHUNK_READS, (size_t)DROPIN_CHUNK_READS * 1024, DROPIN_CHUNK_CANDS);
    if (!ck) die("salt_chunk_new");
    queryio_t *qs[2];
    qs[0] = query_open(opt->fn_read1);
    qs[1] = query_open(opt->fn_read2);
    const int n_threads = opt->n_threads > 1 ? opt->n_threads : 1;
    pe_seed_thread_t *T = calloc((size_t)n_threads, sizeof *T);
    pthread_t *th = calloc((size_t)n_threads, sizeof *th);
    pe_seeded_t *seeds = calloc(N_SEQS, sizeof *seeds);
    int t;
    for (t = 0; t < n_threads; ++t) {
        T[t].tid = t; T[t].n_threads = n_threads; T[t].index = index; T[t].aln_opt = aln_opt;
        T[t].aux[0] = aux_init(opt->l_read, opt->l_seed);
        T[t].aux[1] = aux_init(opt->l_read, opt->l_seed);
    }
    aln_samhead(opt, index->bntseq);

    int n, tot = 0;
    size_t n_rescue = 0;
    query_t *multi_seqs = calloc(N_SEQS, sizeof(query_t));
    int *slot_of = calloc(N_SEQS, sizeof(int));
    char *verified = calloc(N_SEQS, 1);
    while ((n = query_read_multiPairedSeqs(qs, N_SEQS, multi_seqs)) > 0) {
        if (opt->max_tlen == 0) { fprintf(stderr, "infer isize func haven't been implemented\n"); break; }
        int first = 0;                               /* mates [first, i) are queued; first is always even */
        for (t = 0; t < n_threads; ++t) { T[t].n = n; T[t].queries = multi_seqs; T[t].out = seeds; }
        if (gpu_seed) pe_seed_on_gpu(gpu, index, aln_opt, multi_seqs, n, seeds, T[0].aux);
        else if (n_threads == 1) pe_seed_worker(&T[0]);
        else {
            for (t = 0; t < n_threads; ++t) pthread_create(&th[t], NULL, pe_seed_worker, &T[t]);
            for (t = 0; t < n_threads; ++t) pthread_join(th[t], NULL);
        }
        salt_chunk_reset(ck);
        for (i = 0; i <= n; ++i) {
            int flush = (i == n);
            if (!flush) {
                query_t *query = multi_seqs + i;
                int at;
                slot_of[i] = -1; verified[i] = 0;
                if (seeds[i].skip) {
                    /* alnpe.c:495 skips its verification; the mate still has to be on the device for a possible rescue */
                    at = salt_chunk_add_read(ck, query->seq, query->l_seq, NULL, 0, NULL, 0);
                } else {
                    at = salt_chunk_add_read(ck, query->seq, query->l_seq, seeds[i].a[0], seeds[i].n[0], seeds[i].a[1], seeds[i].n[1]);
                    verified[i] = 1;
                }
                if (at == SALT_ERR_NOMEM && salt_chunk_n_reads(ck) > 0) flush = 2;       /* queue full */
                else if (at < 0) die("salt_chunk_add_read");
                else slot_of[i] = at;
            }
            /* verify what is queued once a whole number of pairs is in (or the queue is full) */
            if (flush) {
                int upto = flush == 2 ? (i & ~1) : n, j;
                if (flush == 2 && upto <= first) die("one pair does not fit the queue");
                if (salt_chunk_n_reads(ck) > 0) {
                    if (salt_chunk_submit(gpu, 0, ck, 3, 3) != SALT_OK) die("salt_chunk_submit");   /* alnse.c:1016,1027 */
                    if (salt_chunk_wait(gpu, 0, ck) != SALT_OK) die("salt_chunk_wait");
                }
                for (j = first; j < upto; ++j)
                    if (verified[j]) apply_result(multi_seqs + j, aln_opt, ck, (uint32_t)slot_of[j]);
                H_ck = ck; H_gpu = gpu;
                pair_and_emit(gpu, index, aln_opt, multi_seqs, slot_of, first, upto);
                n_rescue += H.n_req;
                first = upto;
                salt_chunk_reset(ck);
                if (flush == 2) i = upto - 1;        /* re-queue from the first mate not yet verified */
            }
        }
        tot += n;
        for (i = 0; i < n / 2; ++i) {                /* alnpe.c:611-620 */
            query_t *q0 = multi_seqs + i * 2, *q1 = multi_seqs + i * 2 + 1;
            printf("%s\n", q0->sam->s);
            printf("%s\n", q1->sam->s);
            query_destroy(q0);
            query_destroy(q1);
        }
        memset(multi_seqs, '\0', sizeof(query_t) * N_SEQS);
        fprintf(stderr, "alned %d reads!\n", tot);
    }
    fprintf(stderr, "[salt_dropin/pe] rescue windows recorded: %zu, served by the reference's own ssw_align: %zu\n", n_rescue, H.n_cpu);
    fprintf(stderr, "[salt_dropin/pe] salt_pair_plan: %zu pairs checked against pairing2/pairing_singleton, %zu proper without rescue, "
                    "%zu windows planned, %zu mismatches; %zu pairs scheduled from the plan alone\n", plan_checked, plan_paired, plan_windows,
            plan_mismatch, plan_used);
    fprintf(stderr, "[salt_dropin/pe] salt_pair_apply: %zu pairs checked against the reference's final query_t (%zu rescued mates, "
                    "%zu alternates promoted with a fresh CIGAR), %zu not checked (rescue not on the GPU), %zu mismatches\n",
            apply_checked, apply_rescued, apply_swapped, apply_skipped, apply_mismatch);
    fprintf(stderr, "[salt_dropin/pe] pairs finished by salt_pair_apply alone: %zu, handed back to the reference's pairing: %zu\n",
            apply_final, apply_fallback);
    fprintf(stderr, "[salt_dropin/pe] %d seeding threads\n", n_threads);
    if (gpu_seed) fprintf(stderr, "[salt_dropin/pe] seeded on the GPU: %zu mates, %zu of them handed to the reference's own seeding (flagged lists)\n",
                          gpu_seed_reads, gpu_seed_flagged);
    for (t = 0; t < n_threads; ++t) { aux_destroy(T[t].aux[0]); aux_destroy(T[t].aux[1]); }
    for (i = 0; i < N_SEQS; ++i) { free(seeds[i].a[0]); free(seeds[i].a[1]); }
    free(T); free(th); free(seeds);
    query_close(qs[0]);
    query_close(qs[1]);
    free(multi_seqs); free(slot_of); free(verified);
    salt_chunk_free(ck);
    dropin_tail_report();
    salt_b200_destroy(gpu);
    alnpe_index_destroy(index);
    aln_opt_destroy(aln_opt);
    return 0;
}
