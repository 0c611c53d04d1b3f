/* salt_host.c -- see include/salt_host.h.  Host-side C over the C ABI of libsalt_b200.so. */
#include "../../include/salt_host.h"

#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

struct salt_chunk {
    uint32_t max_reads; size_t max_bases, max_cands;
    uint32_t n_reads;
    /* pinned inputs */
    uint8_t *codes; uint32_t *offs_all; uint32_t *roffs;
    uint32_t *offs[2]; uint32_t *loci[2];
    /* pinned outputs */
    salt_verify_out_t *rec; int8_t *acc[2]; char *cigars;
    int lv_T0; int done;          /* done: results of the queued reads are in the output arrays */
    int submitted, slot;          /* in flight on pipeline slot `slot` since salt_chunk_submit */
    /* SAM tail (salt_chunk_tail): per read, rows of the arrays below; -1 = unmapped */
    int tail_done; int32_t *tail_row; salt_mdnm_out_t *tail_out; char *tail_md; uint16_t *tail_xv;
};

#define TAIL_MD_STRIDE 256
#define TAIL_XV_STRIDE 64

static void *pinned(size_t bytes) { return salt_b200_host_alloc(bytes ? bytes : 1); }

salt_chunk_t *salt_chunk_new(uint32_t max_reads, size_t max_bases, size_t max_cands)
{
    salt_chunk_t *c = (salt_chunk_t *)calloc(1, sizeof *c);
    if (!c) return NULL;
    c->max_reads = max_reads; c->max_bases = max_bases; c->max_cands = max_cands;
    c->codes = (uint8_t *)pinned(max_bases + 16);
    /* read offsets and the two candidate-offset arrays share one allocation; salt_chunk_submit lays them back to
       back for the chunk's actual read count so that the engine uploads them in one copy */
    c->offs_all = (uint32_t *)pinned(6 * ((size_t)max_reads + 1) * 4);
    c->roffs = c->offs_all;
    c->rec = (salt_verify_out_t *)pinned((size_t)max_reads * sizeof(salt_verify_out_t));
    c->cigars = (char *)pinned((size_t)max_reads * 128);
    for (int s = 0; s < 2; ++s) {
        c->offs[s] = c->offs_all ? c->offs_all + (size_t)(s + 1) * ((size_t)max_reads + 1) : NULL;
        c->loci[s] = (uint32_t *)pinned(max_cands * 4 + 4);
        c->acc[s] = (int8_t *)pinned(max_cands + 1);
    }
    if (!c->codes || !c->roffs || !c->rec || !c->cigars || !c->offs[0] || !c->offs[1] || !c->loci[0] ||
        !c->loci[1] || !c->acc[0] || !c->acc[1]) { salt_chunk_free(c); return NULL; }
    salt_chunk_reset(c);
    return c;
}

void salt_chunk_free(salt_chunk_t *c)
{
    if (!c) return;
    salt_b200_host_free(c->codes); salt_b200_host_free(c->offs_all); salt_b200_host_free(c->rec);
    salt_b200_host_free(c->cigars);
    free(c->tail_row); free(c->tail_out); free(c->tail_md); free(c->tail_xv);
    for (int s = 0; s < 2; ++s) { salt_b200_host_free(c->loci[s]); salt_b200_host_free(c->acc[s]); }
    free(c);
}

void salt_chunk_reset(salt_chunk_t *c)
{
    c->n_reads = 0; c->done = 0; c->submitted = 0; c->tail_done = 0;
    c->roffs[0] = 0; c->offs[0][0] = 0; c->offs[1][0] = 0;
}

uint32_t salt_chunk_n_reads(const salt_chunk_t *c) { return c->n_reads; }

int salt_chunk_add_read(salt_chunk_t *c, const uint8_t *seq, uint32_t l_seq,
                        const uint32_t *loci0, uint32_t n0, const uint32_t *loci1, uint32_t n1)
{
    const uint32_t i = c->n_reads;
    if (c->done || c->submitted) return SALT_ERR_ARG;          /* in flight, or holding results: salt_chunk_reset first */
    if (i >= c->max_reads) return SALT_ERR_NOMEM;
    if ((size_t)c->roffs[i] + l_seq > c->max_bases) return SALT_ERR_NOMEM;
    if ((size_t)c->offs[0][i] + n0 > c->max_cands || (size_t)c->offs[1][i] + n1 > c->max_cands) return SALT_ERR_NOMEM;
    if (l_seq && !seq) return SALT_ERR_ARG;
    memcpy(c->codes + c->roffs[i], seq, l_seq);
    c->roffs[i + 1] = c->roffs[i] + l_seq;
    if (n0) memcpy(c->loci[0] + c->offs[0][i], loci0, (size_t)n0 * 4);
    if (n1) memcpy(c->loci[1] + c->offs[1][i], loci1, (size_t)n1 * 4);
    c->offs[0][i + 1] = c->offs[0][i] + n0;
    c->offs[1][i + 1] = c->offs[1][i] + n1;
    c->n_reads = i + 1;
    return (int)i;
}

int salt_chunk_submit(salt_b200_t *h, int slot, salt_chunk_t *c, int nogap_T0, int lv_T0)
{
    salt_reads_t r; salt_cands_t k;
    if (!h || !c || c->submitted) return SALT_ERR_ARG;         /* one submit per fill; salt_chunk_wait ends it */
    /* compact copy of the three offset arrays, back to back, in the second half of the allocation */
    const size_t m1 = (size_t)c->n_reads + 1;
    uint32_t *pk = c->offs_all + 3 * ((size_t)c->max_reads + 1);
    memcpy(pk, c->roffs, m1 * 4); memcpy(pk + m1, c->offs[0], m1 * 4); memcpy(pk + 2 * m1, c->offs[1], m1 * 4);
    r.codes = c->codes; r.offs = pk; r.n_reads = c->n_reads;
    k.offs[0] = pk + m1; k.offs[1] = pk + 2 * m1; k.loci[0] = c->loci[0]; k.loci[1] = c->loci[1];
    c->lv_T0 = lv_T0; c->done = 0;
    if (!c->n_reads) { c->submitted = 1; c->slot = slot; return SALT_OK; }
    /* query->cigar starts out empty (kstring, query.c:208-209): gapped primaries get theirs from the GPU */
    for (uint32_t i = 0; i < c->n_reads; ++i) c->cigars[(size_t)i * 128] = '\0';
    const int rc = salt_b200_verify_submit(h, slot, &r, &k, nogap_T0, lv_T0, c->rec, c->acc[0], c->acc[1], c->cigars, 128);
    if (rc == SALT_OK) { c->submitted = 1; c->slot = slot; }
    return rc;
}

int salt_chunk_wait(salt_b200_t *h, int slot, salt_chunk_t *c)
{
    if (!h || !c) return SALT_ERR_ARG;
    if (!c->submitted) return c->done ? SALT_OK : SALT_ERR_ARG;        /* waiting twice is harmless; waiting for nothing is a bug */
    if (slot != c->slot) return SALT_ERR_ARG;                          /* the chunk went to another slot */
    int rc = c->n_reads ? salt_b200_verify_wait(h, slot) : SALT_OK;
    c->submitted = 0;
    if (rc == SALT_OK) c->done = 1;
    return rc;
}

int salt_chunk_seed_verify(salt_b200_t *h, salt_chunk_t *c, const salt_seed_opt_t *opt, int nogap_T0, int lv_T0)
{
    if (!h || !c || !opt || c->submitted) return SALT_ERR_ARG;
    c->lv_T0 = lv_T0; c->done = 0;
    if (!c->n_reads) { c->done = 1; return SALT_OK; }
    salt_reads_t r; r.codes = c->codes; r.offs = c->roffs; r.n_reads = c->n_reads;
    int rc = salt_b200_set_reads(h, &r);
    if (rc != SALT_OK) return rc;
    size_t n0 = 0, n1 = 0;
    rc = salt_b200_seed_locate(h, 0, opt, c->offs[0], c->offs[1], c->loci[0], c->max_cands, c->loci[1], c->max_cands, &n0, &n1);
    if (rc != SALT_OK) return rc;
    for (uint32_t i = 0; i < c->n_reads; ++i) c->cigars[(size_t)i * 128] = '\0';
    rc = salt_b200_verify_seeded(h, 0, nogap_T0, lv_T0, c->rec, c->acc[0], c->acc[1], c->cigars, 128);
    if (rc == SALT_OK) c->done = 1;
    return rc;
}

int salt_chunk_hits(const salt_chunk_t *c, uint32_t i, int strand, salt_hit_t *out, int cap)
{
    if (!c->done || i >= c->n_reads || strand < 0 || strand > 1) return SALT_ERR_ARG;
    const salt_verify_out_t *q = &c->rec[i];
    const uint8_t gapped = q->lv_ran ? 1 : 0;        /* all hits of a read come from one stage (alnse.c:1089) */
    int n = 0;
    for (uint32_t j = c->offs[strand][i]; j < c->offs[strand][i + 1] && n < cap; ++j) {
        if (c->acc[strand][j] < 0) continue;
        out[n].pos = c->loci[strand][j]; out[n].n_diff = (uint8_t)c->acc[strand][j];
        out[n].is_gap = gapped; out[n].strand = (uint16_t)strand;
        ++n;
    }
    return n;
}

static uint32_t gen_mapq(uint32_t b0, uint32_t b1)            /* query.c:270-281 */
{
    if (b0 == 0) return 0;
    uint32_t mapq = (uint32_t)(255.0 * ((double)abs((int)(b0 - b1)) / (double)b0));
    return mapq < 254 ? mapq : 254;
}

int salt_chunk_result(const salt_chunk_t *c, uint32_t i, int max_hits, salt_read_result_t *out)
{
    /* max_hits < 1 would never meet the reference's `tot_hits == max_hits` stop test (query.c:327) and run past alt[] */
    if (!c->done || i >= c->n_reads || !out || max_hits < 1 || max_hits > SALT_MAX_HITS) return SALT_ERR_ARG;
    const salt_verify_out_t *q = &c->rec[i];
    /* no memset of the 700-byte record: alt[] beyond n_alt[] and cigar[] beyond its terminator are not part of the result */
    out->n_alt[0] = out->n_alt[1] = 0; out->cigar[0] = '\0';
    out->pos = q->pos; out->strand = q->strand; out->n_diff = q->n_diff; out->is_gap = q->is_gap;
    /* query_set_hits (query.c:297-333).  The reference compares a->n_diff -- element 0 of the strand's
     * hit vector -- for every j (:317-318); kept.  tot_hits == max_hits is tested after every visited hit. */
    const uint8_t gapped = q->lv_ran ? 1 : 0;
    int tot = 0;
    out->b0 = q->n_diff; out->b1 = 100000;
    for (int s = 0; s < 2; ++s) {
        int first_nd = -1;
        for (uint32_t j = c->offs[s][i]; j < c->offs[s][i + 1]; ++j) {
            const int nd = c->acc[s][j];
            if (nd < 0) continue;                                 /* not in aux->hits */
            if (first_nd < 0) first_nd = nd;
            const uint32_t pos = c->loci[s][j];
            if (pos != 0xFFFFFFFFu && pos != q->pos) {            /* last_pos stays (uint32_t)-1 in the reference */
                if (first_nd <= (int)q->n_diff) {
                    if (first_nd <= out->b1) out->b1 = first_nd;
                    salt_hit_t *hh = &out->alt[s][out->n_alt[s]++];
                    hh->pos = pos; hh->n_diff = (uint8_t)nd; hh->is_gap = gapped; hh->strand = (uint16_t)s;
                    ++tot;
                }
                if (tot == max_hits) goto done;
            }
        }
    }
done:
    out->mapq = gen_mapq((uint32_t)out->b0, (uint32_t)out->b1);
    /* query_gen_cigar (query.c:282-295) */
    if (q->pos != 0xFFFFFFFFu) {
        if (q->is_gap) { strncpy(out->cigar, c->cigars + (size_t)i * 128, 127); out->cigar[127] = '\0'; }
        else snprintf(out->cigar, sizeof out->cigar, "%dM", (int)(c->roffs[i + 1] - c->roffs[i]));
    }
    return SALT_OK;
}

/* sam_add_md_nm for the whole chunk (sam.c:246-328): one GPU call on the slot's resident reads */
int salt_chunk_tail(salt_b200_t *h, int slot, salt_chunk_t *c)
{
    if (!c || !c->done) return SALT_ERR_ARG;
    const uint32_t n = c->n_reads;
    uint32_t m = 0, i;
    c->tail_row = realloc(c->tail_row, ((size_t)n + 1) * sizeof *c->tail_row);
    if (!c->tail_row) return SALT_ERR_NOMEM;
    for (i = 0; i < n; ++i) c->tail_row[i] = c->rec[i].pos != 0xFFFFFFFFu ? (int32_t)m++ : -1;
    c->tail_done = 1;
    if (!m) return SALT_OK;
    salt_mdnm_in_t *in = malloc((size_t)m * sizeof *in);
    char *cg = calloc((size_t)m, 128);
    c->tail_out = realloc(c->tail_out, (size_t)m * sizeof *c->tail_out);
    c->tail_md = realloc(c->tail_md, (size_t)m * TAIL_MD_STRIDE);
    c->tail_xv = realloc(c->tail_xv, (size_t)m * TAIL_XV_STRIDE * sizeof *c->tail_xv);
    if (!in || !cg || !c->tail_out || !c->tail_md || !c->tail_xv) { free(in); free(cg); c->tail_done = 0; return SALT_ERR_NOMEM; }
    for (i = 0; i < n; ++i) {
        const int32_t k = c->tail_row[i];
        if (k < 0) continue;
        const salt_verify_out_t *q = &c->rec[i];
        in[k].rs = (i << 1) | (uint32_t)(q->strand & 1); in[k].pos = q->pos; in[k].seq_start = 0;     /* query.c:284 */
        if (q->is_gap) { strncpy(cg + (size_t)k * 128, c->cigars + (size_t)i * 128, 127); }          /* query.c:288 */
        else snprintf(cg + (size_t)k * 128, 128, "%dM", (int)(c->roffs[i + 1] - c->roffs[i]));       /* query.c:291 */
    }
    const int rc = salt_b200_md_nm(h, slot, in, m, cg, 128, c->tail_md, TAIL_MD_STRIDE, c->tail_xv, TAIL_XV_STRIDE, c->tail_out);
    free(in); free(cg);
    if (rc != SALT_OK) c->tail_done = 0;
    return rc;
}

const char *salt_chunk_md(const salt_chunk_t *c, uint32_t i, int *nm, const uint16_t **xv, int *n_xv)
{
    if (!c || !c->tail_done || i >= c->n_reads) return NULL;
    const int32_t k = c->tail_row[i];
    if (nm) *nm = 0;
    if (xv) *xv = NULL;
    if (n_xv) *n_xv = 0;
    if (k < 0) return "";
    if (c->tail_out[k].md_len < 0) return NULL;                     /* -2 / -3: see salt_mdnm_out_t */
    if (nm) *nm = c->tail_out[k].nm;
    if (n_xv) *n_xv = c->tail_out[k].n_xv;
    if (xv && c->tail_out[k].n_xv) *xv = c->tail_xv + (size_t)k * TAIL_XV_STRIDE;
    return c->tail_md + (size_t)k * TAIL_MD_STRIDE;
}

/* ------------------------------------------------------------------ paired-end plan (alnpe.c:94-257, :395-473) */
enum { PP_IN_RANGE, PP_LOW, PP_HIGH };
static int pp_in_range(uint32_t a, uint32_t b, uint32_t small, uint32_t large)     /* CHECK_IN_RANGE, alnpe.c:76-81 */
{
    const uint32_t r = a < b ? b - a : a - b;
    if (a > b || r < small) return PP_LOW;
    if (r > large) return PP_HIGH;
    return PP_IN_RANGE;
}

/* the hit x hit scan of pairing2 (alnpe.c:131-191): forward hits of one mate against backward hits of the other.
 * The reference never advances its lower bound (`j == jj;` is a comparison, alnpe.c:154): every forward hit scans
 * the backward hits from the first one and stops at the first that lies beyond the insert range. */
static void pp_scan(const salt_read_result_t *f, uint32_t lf, const salt_read_result_t *b, uint32_t min_isize,
                    uint32_t max_isize, uint32_t *min_errors, salt_hit_t *bf, salt_hit_t *bb)
{
    for (int i = 0; i < f->n_alt[0]; ++i) {
        const uint32_t pos0 = f->alt[0][i].pos;
        for (int j = 0; j < b->n_alt[1]; ++j) {
            const int range = pp_in_range(pos0 + lf, b->alt[1][j].pos, min_isize, max_isize);
            if (range == PP_IN_RANGE) {
                const uint32_t e = (uint32_t)f->alt[0][i].n_diff + b->alt[1][j].n_diff;
                if (e < *min_errors) { *min_errors = e; *bf = f->alt[0][i]; *bb = b->alt[1][j]; }
            } else if (range == PP_HIGH) break;
        }
    }
}

/* the window in which mate t is looked for when mate a anchors (alnpe.c:206-252 / :413-466) */
static int pp_window(const salt_read_result_t *a, uint32_t la, uint32_t lt, int t_mate, uint32_t min_isize,
                     uint32_t max_isize, uint32_t l_pac, int singleton, salt_rescue_t *w)
{
    uint32_t s, e;
    if (a->strand == 0) {                    /* anchor forward: the mate lies downstream on the other strand */
        s = a->pos + min_isize + la;
        e = a->pos + max_isize + la + lt;
        w->strand = 1;
    } else {
        s = a->pos > max_isize + lt ? a->pos - max_isize - lt : 0;
        e = a->pos > min_isize ? a->pos - min_isize : 0;
        w->strand = 0;
    }
    if (singleton) { s = s < l_pac - 1 ? s : l_pac - 1; e = e < l_pac - 1 ? e : l_pac - 1; }   /* __min(.., l_pac-1), alnpe.c:417-420 */
    else e = e >= l_pac ? l_pac : e;                                                          /* alnpe.c:212 */
    w->mate = t_mate; w->flavour = singleton ? 5 : 16; w->start = s; w->end = e;
    return s >= l_pac ? SALT_ERR_ARG : SALT_OK;       /* alnpe.c:265-268 / :334-337: the reference exits */
}

int salt_pair_plan(const salt_read_result_t *r0, uint32_t l0, const salt_read_result_t *r1, uint32_t l1,
                   uint32_t min_tlen, uint32_t max_tlen, uint32_t l_pac, salt_pair_plan_t *out)
{
    if (!r0 || !r1 || !out) return SALT_ERR_ARG;
    memset(out, 0, sizeof *out);
    const uint32_t l2 = l0 + l1;
    const uint32_t min_isize = min_tlen > l2 ? min_tlen - l2 : 0, max_isize = max_tlen > l2 ? max_tlen - l2 : 0;
    const int m0 = r0->pos != 0xFFFFFFFFu, m1 = r1->pos != 0xFFFFFFFFu;
    if (!m0 && !m1) return SALT_OK;                                   /* alnpe.c:513-517: nothing to do */
    int rc = SALT_OK;
    if (m0 && m1) {                                                    /* pairing2 */
        const salt_hit_t p0 = {r0->pos, r0->n_diff, r0->is_gap, r0->strand}, p1 = {r1->pos, r1->n_diff, r1->is_gap, r1->strand};
        if ((r0->strand == 0 && r1->strand == 1 && r0->pos < r1->pos &&
             pp_in_range(r0->pos + l0, r1->pos, min_isize, max_isize) == PP_IN_RANGE) ||
            (r1->strand == 0 && r0->strand == 1 && r1->pos < r0->pos &&
             pp_in_range(r1->pos + l1, r0->pos, min_isize, max_isize) == PP_IN_RANGE)) {
            out->paired = 1; out->hit[0] = p0; out->hit[1] = p1;      /* the primaries pair as they are (alnpe.c:109-127) */
            return SALT_OK;
        }
        uint32_t min_errors = 0xFFFFFFFFu;
        salt_hit_t b0 = p0, b1 = p1;
        pp_scan(r0, l0, r1, min_isize, max_isize, &min_errors, &b0, &b1);     /* mate 0 forward, mate 1 backward */
        pp_scan(r1, l1, r0, min_isize, max_isize, &min_errors, &b1, &b0);     /* mate 1 forward, mate 0 backward */
        if (min_errors != 0xFFFFFFFFu) { out->paired = 1; out->hit[0] = b0; out->hit[1] = b1; return SALT_OK; }
        /* rescue: mate 1 around mate 0, then mate 0 around mate 1 (alnpe.c:206-252) */
        int rc0 = pp_window(r0, l0, l1, 1, min_isize, max_isize, l_pac, 0, &out->win[0]);
        int rc1 = pp_window(r1, l1, l0, 0, min_isize, max_isize, l_pac, 0, &out->win[1]);
        out->n_win = 2;
        rc = rc0 != SALT_OK ? rc0 : rc1;
    } else if (m0) {                                                   /* pairing_singleton, mate 0 anchors */
        rc = pp_window(r0, l0, l1, 1, min_isize, max_isize, l_pac, 1, &out->win[0]);
        out->n_win = 1;
    } else {
        rc = pp_window(r1, l1, l0, 0, min_isize, max_isize, l_pac, 1, &out->win[0]);
        out->n_win = 1;
    }
    return rc;
}

/* query_gen_cigar (query.c:282-295) for a mate whose primary is `h` */
static void pp_keep(const salt_read_result_t *r, uint32_t l, const salt_hit_t *h, int swapped, salt_mate_final_t *o)
{
    memset(o, 0, sizeof *o);
    o->pos = h->pos; o->strand = (uint8_t)h->strand; o->n_diff = h->n_diff; o->is_gap = h->is_gap;
    o->b0 = r->b0; o->b1 = r->b1; o->mapq = r->mapq;
    o->seq_start = 0; o->seq_end = l - 1;
    if (o->pos == 0xFFFFFFFFu) { o->cigar_kind = 0; return; }
    if (!o->is_gap) { o->cigar_kind = 1; snprintf(o->cigar, sizeof o->cigar, "%dM", (int)l); }
    else if (!swapped) { o->cigar_kind = 2; strncpy(o->cigar, r->cigar, sizeof o->cigar - 1); }
    else o->cigar_kind = 4;
}

int salt_pair_apply(const salt_pair_plan_t *plan, const salt_read_result_t *r0, uint32_t l0,
                    const salt_read_result_t *r1, uint32_t l1, const salt_ssw_out_t *ssw, const uint32_t *ssw_cigars,
                    int cigar_stride, int filters, int filterd, salt_mate_final_t out[2])
{
    if (!plan || !r0 || !r1 || !out || (plan->n_win > 0 && (!ssw || !ssw_cigars))) return SALT_ERR_ARG;
    const salt_read_result_t *r[2] = {r0, r1};
    const uint32_t l[2] = {l0, l1};
    const salt_hit_t prim[2] = {{r0->pos, r0->n_diff, r0->is_gap, r0->strand}, {r1->pos, r1->n_diff, r1->is_gap, r1->strand}};
    if (plan->paired) {
        for (int m = 0; m < 2; ++m) {
            const int swapped = plan->hit[m].pos != prim[m].pos || plan->hit[m].strand != prim[m].strand;
            pp_keep(r[m], l[m], &plan->hit[m], swapped, &out[m]);
        }
        return 1;
    }
    pp_keep(r0, l0, &prim[0], 0, &out[0]);
    pp_keep(r1, l1, &prim[1], 0, &out[1]);
    for (int w = 0; w < plan->n_win; ++w) {
        const salt_ssw_out_t *a = &ssw[w];
        if (!((int)a->score1 >= filters && a->read_end1 - a->read_begin1 + 1 >= filterd)) continue;     /* alnpe.c:295 / :362 */
        salt_mate_final_t *o = &out[plan->win[w].mate];
        o->b0 = a->score1; o->b1 = a->score2; o->mapq = gen_mapq((uint32_t)o->b0, (uint32_t)o->b1);
        o->pos = (uint32_t)a->ref_begin1 + plan->win[w].start; o->strand = (uint8_t)plan->win[w].strand;
        o->seq_start = (uint32_t)a->read_begin1; o->seq_end = (uint32_t)a->read_end1;
        o->cigar_kind = 3;
        size_t at = 0;
        o->cigar[0] = 0;
        /* a rescued mate without its whole CIGAR is an error, not a shorter string: the engine declined the window
           (cigarLen < 0), the row was too short for it (cigarLen > cigar_stride) or it does not fit 256 bytes */
        if (a->cigarLen < 0 || a->cigarLen > cigar_stride) return SALT_ERR_UNSUPPORTED;
        for (int j = 0; j < a->cigarLen; ++j) {
            const uint32_t c = ssw_cigars[(size_t)w * (size_t)cigar_stride + (size_t)j];
            if (at + 16 >= sizeof o->cigar) return SALT_ERR_UNSUPPORTED;
            at += (size_t)snprintf(o->cigar + at, sizeof o->cigar - at, "%u%c", c >> 4, "MID"[c & 15]);
        }
        return 1;
    }
    return 0;
}

/* host threads for the per-pair loops (hit selection, plans, apply): the reference runs them on its -t workers */
static int g_host_threads = 1;
void salt_host_set_threads(int n) { g_host_threads = n < 1 ? 1 : (n > 256 ? 256 : n); }

/* A parallel-for over [0, n) on the calling thread plus a pool of workers that belongs to the calling thread (two driver
 * threads, INTEGRATION 2a, each get their own).  The workers are persistent: a chunk runs four or five of these loops, and
 * creating sixteen threads for each of them cost 0.3-0.8 ms a time.  The pool follows salt_host_set_threads and is torn
 * down when its owner exits. */
typedef struct pfor_pool pfor_pool_t;
typedef struct { pfor_pool_t *p; int idx; pthread_t th; } pfor_worker_t;
struct pfor_pool {
    int n_workers;                                   /* threads besides the owner */
    pthread_mutex_t mu; pthread_cond_t cv_go, cv_done;
    unsigned long gen;                               /* bumped per job */
    int pending, stop;
    void (*fn)(void *ctx, uint32_t first, uint32_t upto); void *ctx; uint32_t n; int shares;
    pfor_worker_t *w;
};

static void *pfor_worker(void *a)
{
    pfor_worker_t *W = (pfor_worker_t *)a;
    pfor_pool_t *p = W->p;
    unsigned long seen = 0;
    pthread_mutex_lock(&p->mu);
    for (;;) {
        while (!p->stop && p->gen == seen) pthread_cond_wait(&p->cv_go, &p->mu);
        if (p->stop) break;
        seen = p->gen;
        void (*fn)(void *, uint32_t, uint32_t) = p->fn; void *ctx = p->ctx;
        const uint32_t n = p->n; const int T = p->shares, t = W->idx + 1;          /* share 0 is the owner's */
        pthread_mutex_unlock(&p->mu);
        if (t < T) fn(ctx, (uint32_t)((uint64_t)n * (uint64_t)t / (uint64_t)T), (uint32_t)((uint64_t)n * (uint64_t)(t + 1) / (uint64_t)T));
        pthread_mutex_lock(&p->mu);
        if (--p->pending == 0) pthread_cond_signal(&p->cv_done);
    }
    pthread_mutex_unlock(&p->mu);
    return NULL;
}

static void pfor_pool_destroy(void *a)
{
    pfor_pool_t *p = (pfor_pool_t *)a;
    if (!p) return;
    pthread_mutex_lock(&p->mu); p->stop = 1; pthread_cond_broadcast(&p->cv_go); pthread_mutex_unlock(&p->mu);
    for (int i = 0; i < p->n_workers; ++i) pthread_join(p->w[i].th, NULL);
    pthread_mutex_destroy(&p->mu); pthread_cond_destroy(&p->cv_go); pthread_cond_destroy(&p->cv_done);
    free(p->w); free(p);
}

static pthread_key_t g_pool_key;
static pthread_once_t g_pool_once = PTHREAD_ONCE_INIT;
static void pfor_key_init(void) { pthread_key_create(&g_pool_key, pfor_pool_destroy); }

static pfor_pool_t *pfor_pool_get(int n_workers)
{
    pthread_once(&g_pool_once, pfor_key_init);
    pfor_pool_t *p = (pfor_pool_t *)pthread_getspecific(g_pool_key);
    if (p && p->n_workers == n_workers) return p;
    if (p) { pthread_setspecific(g_pool_key, NULL); pfor_pool_destroy(p); }
    p = (pfor_pool_t *)calloc(1, sizeof *p);
    if (!p) return NULL;
    p->w = (pfor_worker_t *)calloc((size_t)n_workers, sizeof *p->w);
    if (!p->w) { free(p); return NULL; }
    pthread_mutex_init(&p->mu, NULL); pthread_cond_init(&p->cv_go, NULL); pthread_cond_init(&p->cv_done, NULL);
    for (int i = 0; i < n_workers; ++i) {
        p->w[i].p = p; p->w[i].idx = i;
        if (pthread_create(&p->w[i].th, NULL, pfor_worker, &p->w[i]) != 0) { p->n_workers = i; pfor_pool_destroy(p); return NULL; }
        p->n_workers = i + 1;
    }
    pthread_setspecific(g_pool_key, p);
    return p;
}

static void pfor_g(uint32_t n, uint32_t grain, void (*fn)(void *, uint32_t, uint32_t), void *ctx)
{
    int T = g_host_threads;
    if ((uint32_t)T > n / grain + 1) T = (int)(n / grain + 1);
    pfor_pool_t *p = T > 1 ? pfor_pool_get(g_host_threads - 1) : NULL;
    if (!p) { fn(ctx, 0, n); return; }
    pthread_mutex_lock(&p->mu);
    p->fn = fn; p->ctx = ctx; p->n = n; p->shares = T; p->pending = p->n_workers; ++p->gen;
    pthread_cond_broadcast(&p->cv_go);
    pthread_mutex_unlock(&p->mu);
    fn(ctx, 0, (uint32_t)((uint64_t)n / (uint64_t)T));
    pthread_mutex_lock(&p->mu);
    while (p->pending) pthread_cond_wait(&p->cv_done, &p->mu);
    pthread_mutex_unlock(&p->mu);
}
static uint32_t g_host_grain = 256;
void salt_host_set_grain(uint32_t items) { g_host_grain = items < 1 ? 1 : items; }
static void pfor(uint32_t n, void (*fn)(void *, uint32_t, uint32_t), void *ctx) { pfor_g(n, g_host_grain, fn, ctx); }


#define COPY_PIECES 64
typedef struct { uint8_t *dst[3]; const uint8_t *src[3]; size_t bytes[3]; } copy3_t;
static void copy3_range(void *a, uint32_t first, uint32_t upto)
{
    copy3_t *X = (copy3_t *)a;
    for (uint32_t i = first; i < upto; ++i) {
        const int k = (int)(i / COPY_PIECES); const size_t piece = i % COPY_PIECES;
        const size_t lo = X->bytes[k] * piece / COPY_PIECES, hi = X->bytes[k] * (piece + 1) / COPY_PIECES;
        if (hi > lo) memcpy(X->dst[k] + lo, X->src[k] + lo, hi - lo);
    }
}

/* ------------------------------------------------------------------ bulk queueing */
int salt_chunk_add_reads(salt_chunk_t *c, const uint8_t *codes, const uint32_t *roffs, uint32_t n,
                         const uint32_t *offs0, const uint32_t *loci0, const uint32_t *offs1, const uint32_t *loci1)
{
    if (!c || !roffs || !offs0 || !offs1 || (n && !codes)) return SALT_ERR_ARG;
    if (c->done || c->submitted) return SALT_ERR_ARG;          /* in flight, or holding results: salt_chunk_reset first */
    const uint32_t at = c->n_reads;
    const size_t nb = (size_t)roffs[n] - roffs[0], n0 = (size_t)offs0[n] - offs0[0], n1 = (size_t)offs1[n] - offs1[0];
    if ((size_t)at + n > c->max_reads || (size_t)c->roffs[at] + nb > c->max_bases ||
        (size_t)c->offs[0][at] + n0 > c->max_cands || (size_t)c->offs[1][at] + n1 > c->max_cands) return SALT_ERR_NOMEM;
    if ((n0 && !loci0) || (n1 && !loci1)) return SALT_ERR_ARG;
    /* three large copies into pinned memory: shared out over the host threads (one thread moves ~5 GB/s) */
    copy3_t cp = {{c->codes + c->roffs[at], (uint8_t *)(c->loci[0] + c->offs[0][at]), (uint8_t *)(c->loci[1] + c->offs[1][at])},
                  {codes + roffs[0], (const uint8_t *)(loci0 ? loci0 + offs0[0] : NULL), (const uint8_t *)(loci1 ? loci1 + offs1[0] : NULL)},
                  {nb, n0 * 4, n1 * 4}};
    if (nb + n0 * 4 + n1 * 4 < ((size_t)1 << 20)) copy3_range(&cp, 0, 3 * COPY_PIECES);       /* small: not worth the threads */
    else pfor_g(3 * COPY_PIECES, 4, copy3_range, &cp);
    const uint32_t rb = c->roffs[at] - roffs[0], b0 = c->offs[0][at] - offs0[0], b1 = c->offs[1][at] - offs1[0];
    for (uint32_t i = 1; i <= n; ++i) {
        c->roffs[at + i] = roffs[i] + rb; c->offs[0][at + i] = offs0[i] + b0; c->offs[1][at + i] = offs1[i] + b1;
    }
    c->n_reads = at + n;
    return (int)at;
}

/* ------------------------------------------------------------------ paired-end stage of a chunk */
static double ms_now(void)
{
    struct timespec t;
    clock_gettime(CLOCK_MONOTONIC, &t);
    return 1e3 * (double)t.tv_sec + 1e-6 * (double)t.tv_nsec;
}

#define PE_CIG_STRIDE 64
#define PE_TAIL_CIG 64               /* bytes per mate for the M/I/D string sent to salt_b200_md_nm (longer ones: 256) */

typedef struct {
    salt_chunk_t *c; int max_hits; uint32_t min_tlen, max_tlen, l_pac; int filters, filterd;
    salt_read_result_t *res; salt_pair_plan_t *plan; uint32_t *win_at;
    salt_ssw_out_t *so[2]; uint32_t *sc[2];
    salt_pair_final_t *out; salt_mdnm_in_t *tin; char *tcg; int tcs;
    int rc;                                  /* first error of any thread (benign race: any error fails the call) */
    size_t declined, rescued, k4, long_cigar;
    pthread_mutex_t mu;
} pe_ctx_t;

static void pe_plan_range(void *a, uint32_t first, uint32_t upto)
{
    pe_ctx_t *X = (pe_ctx_t *)a;
    salt_chunk_t *c = X->c;
    for (uint32_t p = first; p < upto; ++p) {
        int rc = salt_chunk_result(c, 2 * p, X->max_hits, &X->res[2 * p]);
        if (rc == SALT_OK) rc = salt_chunk_result(c, 2 * p + 1, X->max_hits, &X->res[2 * p + 1]);
        const uint32_t l0 = c->roffs[2 * p + 1] - c->roffs[2 * p], l1 = c->roffs[2 * p + 2] - c->roffs[2 * p + 1];
        if (rc == SALT_OK) rc = salt_pair_plan(&X->res[2 * p], l0, &X->res[2 * p + 1], l1, X->min_tlen, X->max_tlen, X->l_pac, &X->plan[p]);
        if (rc != SALT_OK) { X->rc = rc; return; }
    }
}

static void pe_apply_range(void *a, uint32_t first, uint32_t upto)
{
    pe_ctx_t *X = (pe_ctx_t *)a;
    salt_chunk_t *c = X->c;
    size_t declined = 0, rescued = 0, k4 = 0, long_cigar = 0;
    for (uint32_t p = first; p < upto; ++p) {
        const uint32_t l0 = c->roffs[2 * p + 1] - c->roffs[2 * p], l1 = c->roffs[2 * p + 2] - c->roffs[2 * p + 1];
        salt_ssw_out_t ws[2];
        uint32_t wc[2 * PE_CIG_STRIDE];
        const salt_pair_plan_t *pl = &X->plan[p];
        for (int w = 0; w < pl->n_win; ++w) {
            const int f = pl->win[w].flavour == 5;
            const uint32_t at = X->win_at[2 * p + (uint32_t)w];
            ws[w] = X->so[f][at];
            memcpy(wc + w * PE_CIG_STRIDE, X->sc[f] + (size_t)at * PE_CIG_STRIDE, PE_CIG_STRIDE * 4);
            if (ws[w].cigarLen < 0) ++declined;
        }
        const int r = salt_pair_apply(pl, &X->res[2 * p], l0, &X->res[2 * p + 1], l1, ws, wc, PE_CIG_STRIDE, X->filters, X->filterd,
                                      X->out[p].mate);
        if (r < 0) { X->rc = r; return; }
        X->out[p].paired = r;
        for (int m = 0; m < 2; ++m) {
            const salt_mate_final_t *mf = &X->out[p].mate[m];
            if (mf->cigar_kind == 3) ++rescued;
            if (mf->cigar_kind == 4) ++k4;
            if (X->tin) {                                /* the mate's row of the SAM-tail batch */
                const uint32_t i = 2 * p + (uint32_t)m;
                X->tin[i].rs = (i << 1) | (uint32_t)(mf->strand & 1); X->tin[i].pos = mf->pos; X->tin[i].seq_start = mf->seq_start;
                const size_t len = strnlen(mf->cigar, sizeof mf->cigar);
                if (len >= (size_t)X->tcs) ++long_cigar;
                else memcpy(X->tcg + (size_t)i * (size_t)X->tcs, mf->cigar, len + 1);
            }
        }
    }
    pthread_mutex_lock(&X->mu);
    X->declined += declined; X->rescued += rescued; X->k4 += k4; X->long_cigar += long_cigar;
    pthread_mutex_unlock(&X->mu);
}

typedef struct { const salt_chunk_t *c; int max_hits; salt_read_result_t *out; int rc; } results_ctx_t;
static void results_range(void *a, uint32_t first, uint32_t upto)
{
    results_ctx_t *X = (results_ctx_t *)a;
    for (uint32_t i = first; i < upto; ++i) {
        const int rc = salt_chunk_result(X->c, i, X->max_hits, &X->out[i]);
        if (rc != SALT_OK) { X->rc = rc; return; }
    }
}

int salt_chunk_results(const salt_chunk_t *c, int max_hits, salt_read_result_t *out)
{
    if (!c || !c->done || !out) return SALT_ERR_ARG;
    results_ctx_t X = {c, max_hits, out, SALT_OK};
    pfor(c->n_reads, results_range, &X);
    return X.rc;
}

int salt_chunk_pair(salt_b200_t *h, int slot, salt_chunk_t *c, uint32_t min_tlen, uint32_t max_tlen, uint32_t l_pac,
                    int max_hits, const int8_t *mat16, const int8_t *mat5, int gapO, int gapE, int filters, int filterd,
                    int with_tail, salt_pair_final_t *out, salt_mdnm_out_t *tail_out, char *tail_md, int md_stride,
                    salt_pe_stats_t *stats)
{
    if (!h || !c || !c->done || !out || !mat16 || !mat5 || (c->n_reads & 1)) return SALT_ERR_ARG;
    if (with_tail && (!tail_out || !tail_md || md_stride < 2)) return SALT_ERR_ARG;
    const uint32_t np = c->n_reads / 2;
    salt_pe_stats_t st;
    memset(&st, 0, sizeof st);
    st.pairs = np;
    int rc = SALT_OK;
    pe_ctx_t X;
    memset(&X, 0, sizeof X);
    pthread_mutex_init(&X.mu, NULL);
    X.c = c; X.max_hits = max_hits; X.min_tlen = min_tlen; X.max_tlen = max_tlen; X.l_pac = l_pac; X.filters = filters; X.filterd = filterd;
    X.out = out; X.rc = SALT_OK;
    X.res = malloc((size_t)c->n_reads * sizeof *X.res + 1);
    X.plan = malloc((size_t)np * sizeof *X.plan + 1);
    /* windows by flavour; win_at[2 * pair + w] = row of the plan's w-th window in its flavour's batch */
    salt_win_t *win[2] = {malloc(((size_t)np * 2 + 1) * sizeof(salt_win_t)), malloc(((size_t)np * 2 + 1) * sizeof(salt_win_t))};
    X.win_at = malloc(((size_t)np * 2 + 1) * 4);
    size_t nw[2] = {0, 0};
    if (!X.res || !X.plan || !win[0] || !win[1] || !X.win_at) { rc = SALT_ERR_NOMEM; goto done; }

    double t0 = ms_now();
    pfor(np, pe_plan_range, &X);                          /* query_set_hits of both mates + the pair's plan */
    if ((rc = X.rc) != SALT_OK) goto done;
    for (uint32_t p = 0; p < np; ++p) {
        if (X.plan[p].paired) ++st.proper;
        for (int w = 0; w < X.plan[p].n_win; ++w) {
            const salt_rescue_t *rw = &X.plan[p].win[w];
            const int f = rw->flavour == 5;
            salt_win_t *o = &win[f][nw[f]];
            o->rs = ((2 * p + (uint32_t)rw->mate) << 1) | (uint32_t)rw->strand; o->start = rw->start; o->end = rw->end;
            X.win_at[2 * p + (uint32_t)w] = (uint32_t)nw[f]++;
        }
    }
    st.windows16 = nw[0]; st.windows5 = nw[1];
    st.ms_plan = ms_now() - t0;

    /* one Smith-Waterman batch per flavour on the chunk's resident reads (alnpe.c:261 / :330) */
    t0 = ms_now();
    if ((rc = salt_b200_use_slot(h, slot)) != SALT_OK) goto done;
    for (int f = 0; f < 2 && rc == SALT_OK; ++f) {
        if (!nw[f]) continue;
        X.so[f] = malloc(nw[f] * sizeof(salt_ssw_out_t));
        X.sc[f] = calloc(nw[f] * PE_CIG_STRIDE, 4);
        if (!X.so[f] || !X.sc[f]) { rc = SALT_ERR_NOMEM; break; }
        rc = salt_b200_ssw(h, win[f], nw[f], f, f ? mat5 : mat16, f ? 5 : 16, gapO, gapE, 2, filters, filterd, -1, X.so[f], X.sc[f], PE_CIG_STRIDE);
    }
    st.ms_ssw = ms_now() - t0;
    if (rc != SALT_OK) { salt_b200_use_slot(h, 0); goto done; }

    t0 = ms_now();
    if (with_tail) {
        X.tcs = PE_TAIL_CIG;
        X.tin = malloc((size_t)c->n_reads * sizeof *X.tin + 1);
        X.tcg = malloc((size_t)c->n_reads * (size_t)X.tcs + 1);
        if (!X.tin || !X.tcg) { rc = SALT_ERR_NOMEM; salt_b200_use_slot(h, 0); goto done; }
    }
    pfor(np, pe_apply_range, &X);
    rc = X.rc;
    st.declined = X.declined; st.rescued = X.rescued;
    if (rc == SALT_OK && X.k4) {
        /* gapped alternates that became primaries: their CIGARs in one salt_b200_lv_cigar call (query.c:288) */
        const size_t n_k4 = X.k4;
        salt_pair_t *pp = malloc(n_k4 * sizeof *pp);
        uint8_t *kk = malloc(n_k4);
        char *cg = calloc(n_k4, 128);
        int8_t *eo = malloc(n_k4);
        size_t q = 0;
        if (!pp || !kk || !cg || !eo) rc = SALT_ERR_NOMEM;
        for (uint32_t p = 0; p < np && rc == SALT_OK; ++p)
            for (int m = 0; m < 2; ++m)
                if (out[p].mate[m].cigar_kind == 4) {
                    pp[q].rs = ((2 * p + (uint32_t)m) << 1) | (uint32_t)(out[p].mate[m].strand & 1); pp[q].pos = out[p].mate[m].pos;
                    kk[q] = out[p].mate[m].n_diff > 30 ? 30 : out[p].mate[m].n_diff; ++q;
                }
        if (rc == SALT_OK) rc = salt_b200_lv_cigar(h, pp, kk, n_k4, cg, 128, eo);
        q = 0;
        for (uint32_t p = 0; p < np && rc == SALT_OK; ++p)
            for (int m = 0; m < 2; ++m)
                if (out[p].mate[m].cigar_kind == 4) {
                    strncpy(out[p].mate[m].cigar, cg + q * 128, 127);
                    if (X.tcg) { strncpy(X.tcg + (size_t)(2 * p + (uint32_t)m) * (size_t)X.tcs, cg + q * 128, (size_t)X.tcs - 1);
                                 X.tcg[(size_t)(2 * p + (uint32_t)m) * (size_t)X.tcs + (size_t)X.tcs - 1] = 0;
                                 if (strlen(cg + q * 128) >= (size_t)X.tcs) ++X.long_cigar; }
                    ++q; ++st.promoted;
                }
        free(pp); free(kk); free(cg); free(eo);
    }
    salt_b200_use_slot(h, 0);
    st.ms_apply = ms_now() - t0;
    if (rc != SALT_OK) goto done;

    if (with_tail) {
        /* sam_add_md_nm of every mapped mate (sam.c:246-328) in one call on the slot's reads */
        t0 = ms_now();
        const uint32_t n = c->n_reads;
        if (X.long_cigar) {                               /* rare: a CIGAR of 64+ characters -- redo the batch's strings at full width */
            free(X.tcg);
            X.tcs = 256;
            X.tcg = malloc((size_t)n * 256 + 1);
            if (!X.tcg) rc = SALT_ERR_NOMEM;
            for (uint32_t i = 0; i < n && rc == SALT_OK; ++i) {
                memcpy(X.tcg + (size_t)i * 256, out[i / 2].mate[i & 1].cigar, 256);
                X.tcg[(size_t)i * 256 + 255] = 0;
            }
        }
        if (rc == SALT_OK) rc = salt_b200_md_nm(h, slot, X.tin, n, X.tcg, X.tcs, tail_md, md_stride, NULL, 0, tail_out);
        st.ms_tail = ms_now() - t0;
    }
done:
    free(X.res); free(X.plan); free(win[0]); free(win[1]); free(X.win_at);
    free(X.so[0]); free(X.so[1]); free(X.sc[0]); free(X.sc[1]); free(X.tin); free(X.tcg);
    pthread_mutex_destroy(&X.mu);
    if (stats) *stats = st;
    return rc;
}

/* ------------------------------------------------------------------ several GPUs in one process */
typedef struct {
    salt_b200_t *h; salt_packed_chunk_t view; uint32_t chunk_reads; int nogap_T0, lv_T0;
    salt_verify_out_t *rec; int8_t *acc0, *acc1; char *cigars; int cigar_stride; int rc;
} multi_job_t;

/* one persistent host thread per device: a thread's first CUDA call binds it to its device, which is not free, so the
   threads live as long as the handle set and are woken per batch */
typedef struct {
    pthread_t th; pthread_mutex_t mu; pthread_cond_t cv;
    int state;                       /* 0 idle, 1 job posted, 2 job done, 3 quit */
    multi_job_t job;
} multi_worker_t;

struct salt_multi { int n; salt_b200_t **h; multi_worker_t *w; };

static void *multi_worker(void *arg)
{
    multi_worker_t *W = (multi_worker_t *)arg;
    pthread_mutex_lock(&W->mu);
    for (;;) {
        while (W->state != 1 && W->state != 3) pthread_cond_wait(&W->cv, &W->mu);
        if (W->state == 3) break;
        pthread_mutex_unlock(&W->mu);
        multi_job_t *J = &W->job;
        J->rc = J->view.n_reads ? salt_b200_verify_batch_packed(J->h, &J->view, J->chunk_reads, J->nogap_T0, J->lv_T0, J->rec, J->acc0,
                                                               J->acc1, J->cigars, J->cigar_stride) : SALT_OK;
        pthread_mutex_lock(&W->mu);
        W->state = 2;
        pthread_cond_broadcast(&W->cv);
    }
    pthread_mutex_unlock(&W->mu);
    return NULL;
}

salt_multi_t *salt_multi_init(const uint32_t *mixref, uint32_t l, const uint8_t *pac, int64_t l_pac, const int *devices, int n_devices)
{
    if (n_devices < 1 || !devices) return NULL;
    salt_multi_t *m = calloc(1, sizeof *m);
    if (!m) return NULL;
    m->h = calloc((size_t)n_devices, sizeof *m->h);
    m->w = calloc((size_t)n_devices, sizeof *m->w);
    if (!m->h || !m->w) { free(m->h); free(m->w); free(m); return NULL; }
    m->n = n_devices;
    for (int i = 0; i < n_devices; ++i) {
        m->h[i] = salt_b200_init(mixref, l, pac, l_pac, devices[i]);
        if (!m->h[i]) { m->n = i; salt_multi_destroy(m); return NULL; }
    }
    for (int i = 0; i < n_devices; ++i) {
        pthread_mutex_init(&m->w[i].mu, NULL); pthread_cond_init(&m->w[i].cv, NULL);
        m->w[i].state = 0;
        pthread_create(&m->w[i].th, NULL, multi_worker, &m->w[i]);
    }
    return m;
}

void salt_multi_destroy(salt_multi_t *m)
{
    if (!m) return;
    for (int i = 0; i < m->n; ++i) {
        if (m->w && m->w[i].th) {
            pthread_mutex_lock(&m->w[i].mu); m->w[i].state = 3; pthread_cond_broadcast(&m->w[i].cv); pthread_mutex_unlock(&m->w[i].mu);
            pthread_join(m->w[i].th, NULL);
            pthread_mutex_destroy(&m->w[i].mu); pthread_cond_destroy(&m->w[i].cv);
        }
        if (m->h[i]) salt_b200_destroy(m->h[i]);
    }
    free(m->h); free(m->w); free(m);
}

int salt_multi_n(const salt_multi_t *m) { return m ? m->n : 0; }
salt_b200_t *salt_multi_handle(salt_multi_t *m, int i) { return (m && i >= 0 && i < m->n) ? m->h[i] : NULL; }

static uint32_t pk_count(const salt_packed_chunk_t *pc, int s, uint32_t i)
{
    return pc->count_bits == 16 ? ((const uint16_t *)pc->n_cand[s])[i] : ((const uint32_t *)pc->n_cand[s])[i];
}

int salt_multi_verify_batch_packed(salt_multi_t *m, const salt_packed_chunk_t *pc, uint32_t chunk_reads, int nogap_T0, int lv_T0,
                                   salt_verify_out_t *rec, int8_t *acc0, int8_t *acc1, char *cigars, int cigar_stride)
{
    if (!m || !pc || !rec || (pc->count_bits != 16 && pc->count_bits != 32) || (pc->n_reads && (!pc->n_cand[0] || !pc->n_cand[1])))
        return SALT_ERR_ARG;
    if (chunk_reads == 0) chunk_reads = 100000;
    const int G = m->n;
    /* contiguous shares on chunk boundaries; one pass over the lengths and counts finds where each share starts */
    const uint32_t n_chunks = (pc->n_reads + chunk_reads - 1) / chunk_reads;
    uint64_t base = pc->base_start; size_t c0 = 0, c1 = 0, np = 0;
    uint32_t r = 0;
    const size_t cb = (size_t)pc->count_bits / 8;
    for (int g = 0; g < G; ++g) {
        const uint32_t first_chunk = (uint32_t)((uint64_t)n_chunks * (uint64_t)g / (uint64_t)G);
        const uint32_t upto_chunk = (uint32_t)((uint64_t)n_chunks * (uint64_t)(g + 1) / (uint64_t)G);
        const uint32_t first = first_chunk * chunk_reads;
        uint32_t upto = upto_chunk * chunk_reads; if (upto > pc->n_reads) upto = pc->n_reads;
        multi_job_t *j = &m->w[g].job;
        j->h = m->h[g]; j->chunk_reads = chunk_reads; j->nogap_T0 = nogap_T0; j->lv_T0 = lv_T0; j->cigar_stride = cigar_stride;
        j->view = *pc;
        j->view.n_reads = upto - first;
        j->view.base_start = (uint32_t)base;
        j->view.lens = pc->lens ? pc->lens + first : NULL;
        while (np < pc->n_n && pc->n_pos[np] < base) ++np;
        j->view.n_pos = pc->n_pos ? pc->n_pos + np : NULL; j->view.n_n = pc->n_n - np;
        j->view.n_cand[0] = (const uint8_t *)pc->n_cand[0] + (size_t)first * cb;
        j->view.n_cand[1] = (const uint8_t *)pc->n_cand[1] + (size_t)first * cb;
        j->view.loci[0] = pc->loci[0] ? pc->loci[0] + c0 : NULL; j->view.loci[1] = pc->loci[1] ? pc->loci[1] + c1 : NULL;
        j->rec = rec + first; j->acc0 = acc0 ? acc0 + c0 : NULL; j->acc1 = acc1 ? acc1 + c1 : NULL;
        j->cigars = cigars ? cigars + (size_t)first * cigar_stride : NULL;
        for (; r < upto; ++r) {
            base += pc->lens ? pc->lens[r] : pc->l_seq;
            c0 += pk_count(pc, 0, r); c1 += pk_count(pc, 1, r);
        }
        pthread_mutex_lock(&m->w[g].mu); m->w[g].state = 1; pthread_cond_broadcast(&m->w[g].cv); pthread_mutex_unlock(&m->w[g].mu);
    }
    int rc = SALT_OK;
    for (int g = 0; g < G; ++g) {
        pthread_mutex_lock(&m->w[g].mu);
        while (m->w[g].state != 2) pthread_cond_wait(&m->w[g].cv, &m->w[g].mu);
        m->w[g].state = 0;
        pthread_mutex_unlock(&m->w[g].mu);
        if (rc == SALT_OK) rc = m->w[g].job.rc;
    }
    return rc;
}

/* ------------------------------------------------------------------ FASTQ -> compact transport (query.c:146-239) */
static inline int fq_is_blank(char c) { return c == ' ' || c == '\t' || c == '\n' || c == '\r' || c == '\v' || c == '\f'; }

/* The 2-bit base stream is written through a small bit accumulator (no read-modify-write per base); bases arrive eight at a
 * time when a sequence line holds nothing but A/C/G/T in either case, which eight bytes are tested for at once. */
typedef struct { uint8_t *out; size_t byte; uint64_t acc; int bits; } fq_bitw_t;
static inline void fq_put(fq_bitw_t *w, uint64_t v, int nbits)
{
    w->acc |= v << w->bits; w->bits += nbits;
    while (w->bits >= 8) { w->out[w->byte++] = (uint8_t)w->acc; w->acc >>= 8; w->bits -= 8; }
}
#define FQ_ONES 0x0101010101010101ULL
static inline uint64_t fq_eq_bytes(uint64_t x, unsigned c)        /* 0x80 in every byte of x that equals c */
{
    const uint64_t y = x ^ (FQ_ONES * c);
    const uint64_t t = ((y & (FQ_ONES * 0x7F)) + (FQ_ONES * 0x7F)) | y;          /* bit 7 set iff the byte is non-zero */
    return ~t & (FQ_ONES * 0x80);
}
static inline int fq_has_blank(uint64_t x) { return ((x - FQ_ONES * 0x21) & ~x & (FQ_ONES * 0x80)) != 0; }   /* some byte <= ' ' */
#if defined(__SSE2__)
#include <emmintrin.h>
/* sixteen bases at once: 1 and *packed = their 32 code bits when all sixteen are A/C/G/T in either case, else 0 */
static inline int fq_pack16(const char *src, uint32_t *packed)
{
    const __m128i x = _mm_loadu_si128((const __m128i *)src);
    const __m128i u = _mm_and_si128(x, _mm_set1_epi8((char)0xDF));
    const __m128i ok = _mm_or_si128(_mm_or_si128(_mm_cmpeq_epi8(u, _mm_set1_epi8('A')), _mm_cmpeq_epi8(u, _mm_set1_epi8('C'))),
                                    _mm_or_si128(_mm_cmpeq_epi8(u, _mm_set1_epi8('G')), _mm_cmpeq_epi8(u, _mm_set1_epi8('T'))));
    if (_mm_movemask_epi8(ok) != 0xFFFF) return 0;
    __m128i t = _mm_and_si128(_mm_xor_si128(_mm_srli_epi16(x, 1), _mm_srli_epi16(x, 2)), _mm_set1_epi8(3));
    t = _mm_and_si128(_mm_or_si128(t, _mm_srli_epi16(t, 6)), _mm_set1_epi16(0x000F));
    t = _mm_and_si128(_mm_or_si128(t, _mm_srli_epi32(t, 12)), _mm_set1_epi32(0x000000FF));
    t = _mm_packus_epi16(_mm_packs_epi32(t, t), _mm_setzero_si128());
    *packed = (uint32_t)_mm_cvtsi128_si32(t);
    return 1;
}
static inline int fq_blank16(const char *src)          /* some byte <= ' ' (or >= 0x80: left to the byte loop) among sixteen */
{
    return _mm_movemask_epi8(_mm_cmplt_epi8(_mm_loadu_si128((const __m128i *)src), _mm_set1_epi8(0x21))) != 0;
}
#endif

int salt_fastq_pack(const char *text, size_t len, int final, uint32_t max_reads, salt_fastq_t *o, size_t *consumed)
{
    if (!text || !o || !o->bases || !o->lens || !o->n_ambiguous || !o->name_off || !o->name_len || !consumed) return SALT_ERR_ARG;
    /* nst_nt4_table (bntseq.c): A/a C/c G/g T/t -> 0..3, '-' and everything else -> 4 (5 for '-' is > 3 all the same) */
    static uint8_t nt4[256]; static int nt4_ready;
    if (!nt4_ready) {
        memset(nt4, 4, sizeof nt4);
        nt4['A'] = nt4['a'] = 0; nt4['C'] = nt4['c'] = 1; nt4['G'] = nt4['g'] = 2; nt4['T'] = nt4['t'] = 3;
        nt4_ready = 1;
    }
    uint32_t n = 0;
    size_t nb = o->n_bases = 0, nn = o->n_n = 0;
    size_t p = 0;
    *consumed = 0;
    o->n_reads = 0;
    fq_bitw_t w = {o->bases, 0, 0, 0};
    size_t dirty = 0;                      /* bytes a given-back record may have written */
    while (n < max_reads) {
        /* kseq: skip to the next header character */
        while (p < len && text[p] != '@' && text[p] != '>') ++p;
        if (p >= len) break;
        const size_t rec = p++;
        size_t q = p;
        while (q < len && !fq_is_blank(text[q])) ++q;
        if (q >= len && !final) break;
        size_t name_a = p, name_b = q;
        if (name_b - name_a > 2 && text[name_b - 2] == '/' && text[name_b - 1] >= '0' && text[name_b - 1] <= '9') name_b -= 2;   /* trim_readno */
        size_t com_a = q, com_b = q;
        if (q < len && text[q] != '\n') {                                   /* comment: the rest of the header line */
            com_a = q + 1;
            const char *e = (const char *)memchr(text + q, '\n', len - q);
            q = e ? (size_t)(e - text) : len;
            com_b = q;
            while (com_b > com_a && text[com_b - 1] == '\r') --com_b;
        }
        if (q >= len && !final) break;
        p = q < len ? q + 1 : q;
        /* sequence lines until a line starting with '+', '>' or '@' */
        const size_t base0 = nb, nn0 = nn; size_t amb = 0;
        const fq_bitw_t w0 = w;
        int complete = 0, has_plus = 0;
        while (1) {
            if (p >= len) { complete = final; break; }
            const char c0 = text[p];
            if (c0 == '+') { has_plus = 1; complete = 1; break; }
            if (c0 == '>' || c0 == '@') { complete = 1; break; }
            const char *e = (const char *)memchr(text + p, '\n', len - p);
            const size_t le = e ? (size_t)(e - text) : len;                  /* end of this line (or of the block) */
            if (nb + (le - p) > o->bases_cap) goto full;                     /* conservative: blanks inside the line count too */
#if defined(__SSE2__)
            while (p + 16 <= le) {
                uint32_t v;
                if (!fq_pack16(text + p, &v)) break;
                fq_put(&w, v, 32);
                nb += 16; p += 16;
            }
#endif
            while (p + 8 <= le) {
                uint64_t x; memcpy(&x, text + p, 8);
                const uint64_t u = x & (FQ_ONES * 0xDF);                     /* upper case */
                const uint64_t ok = fq_eq_bytes(u, 'A') | fq_eq_bytes(u, 'C') | fq_eq_bytes(u, 'G') | fq_eq_bytes(u, 'T');
                if (ok != FQ_ONES * 0x80) break;                             /* something else in these eight: one at a time below */
                uint64_t t = ((x >> 1) ^ (x >> 2)) & (FQ_ONES * 3);          /* A C G T (either case) -> 0 1 2 3 */
                t = (t | (t >> 6)) & 0x000F000F000F000FULL;
                t = (t | (t >> 12)) & 0x000000FF000000FFULL;
                t = (t | (t >> 24)) & 0xFFFFULL;
                fq_put(&w, t, 16);
                nb += 8; p += 8;
            }
            while (p < le) {
#if defined(__SSE2__)
                if (p + 16 <= le) {                                          /* back to sixteen at a time after the odd byte(s) */
                    uint32_t v;
                    if (fq_pack16(text + p, &v)) { fq_put(&w, v, 32); nb += 16; p += 16; continue; }
                }
#endif
                if (p + 8 <= le) {
                    uint64_t x; memcpy(&x, text + p, 8);
                    const uint64_t u = x & (FQ_ONES * 0xDF);
                    const uint64_t ok = fq_eq_bytes(u, 'A') | fq_eq_bytes(u, 'C') | fq_eq_bytes(u, 'G') | fq_eq_bytes(u, 'T');
                    if (ok == FQ_ONES * 0x80) {
                        uint64_t t = ((x >> 1) ^ (x >> 2)) & (FQ_ONES * 3);
                        t = (t | (t >> 6)) & 0x000F000F000F000FULL;
                        t = (t | (t >> 12)) & 0x000000FF000000FFULL;
                        t = (t | (t >> 24)) & 0xFFFFULL;
                        fq_put(&w, t, 16);
                        nb += 8; p += 8;
                        continue;
                    }
                }
                const unsigned char ch = (unsigned char)text[p++];
                if (ch <= ' ') continue;                                     /* kseq keeps graphic characters only */
                const unsigned v = nt4[ch];
                if (v > 3) {
                    ++amb;
                    if (o->n_pos) { if (nn >= o->n_pos_cap) goto full; o->n_pos[nn] = (uint32_t)nb; }
                    ++nn;
                    fq_put(&w, 0, 2);
                } else fq_put(&w, v, 2);
                ++nb;
            }
            if (p < len) ++p; else { complete = final; break; }
        }
        if (!complete) { nb = base0; nn = nn0; dirty = w.byte + 1; w = w0; break; }
        const size_t L = nb - base0;
        if (L > 65535) return SALT_ERR_ARG;
        size_t qual_at = (size_t)-1;
        if (has_plus) {
            const char *e = (const char *)memchr(text + p, '\n', len - p);   /* the rest of the '+' line */
            p = e ? (size_t)(e - text) : len;
            if (p >= len) { if (!final) { nb = base0; nn = nn0; dirty = w.byte + 1; w = w0; break; } }
            else ++p;
            qual_at = p;
            size_t ql = 0;
            if (p + L <= len) {                                              /* the usual case: L graphic characters in a row */
                size_t k = 0; int blank = 0;
#if defined(__SSE2__)
                for (; k + 16 <= L; k += 16) if (fq_blank16(text + p + k)) { blank = 1; break; }
#endif
                for (; !blank && k + 8 <= L; k += 8) { uint64_t x; memcpy(&x, text + p + k, 8); if (fq_has_blank(x)) { blank = 1; break; } }
                if (!blank) for (; k < L; ++k) if ((unsigned char)text[p + k] <= ' ') { blank = 1; break; }
                if (!blank) { ql = L; p += L; }
            }
            while (ql < L && p < len) { const unsigned char ch = (unsigned char)text[p++]; if (ch > ' ') ++ql; }
            if (ql < L) { if (final) return SALT_ERR_ARG; nb = base0; nn = nn0; dirty = w.byte + 1; w = w0; break; }
            e = (const char *)memchr(text + p, '\n', len - p);               /* kseq reads whole lines */
            p = e ? (size_t)(e - text) : len;
            if (p < len) ++p; else if (!final) { nb = base0; nn = nn0; dirty = w.byte + 1; w = w0; break; }
        }
        if (L == 0) { /* query_read_seq stops at an empty record (l_seq <= 0, query.c:155) */ *consumed = p; break; }
        o->lens[n] = (uint16_t)L; o->n_ambiguous[n] = (uint16_t)(amb > 65535 ? 65535 : amb);
        o->name_off[n] = (uint32_t)name_a; o->name_len[n] = (uint16_t)(name_b - name_a);
        if (o->comment_off) { o->comment_off[n] = (uint32_t)com_a; o->comment_len[n] = (uint16_t)(com_b - com_a); }
        if (o->qual_off) o->qual_off[n] = qual_at == (size_t)-1 ? 0xFFFFFFFFu : (uint32_t)qual_at;
        ++n;
        *consumed = p;
        (void)rec;
        continue;
full:
        /* an array filled up inside this record: give back what belongs to it and stop before it */
        nb = base0; nn = nn0; dirty = w.byte + 1; w = w0;
        break;
    }
    /* the last, partly filled byte; whatever a given-back record left behind it is cleared */
    {
        size_t end = w.byte;
        if (w.bits > 0) o->bases[end++] = (uint8_t)w.acc;
        const size_t room = o->bases_cap / 4 + 1;
        if (dirty > room) dirty = room;
        if (end < dirty) memset(o->bases + end, 0, dirty - end);
    }
    o->n_reads = n; o->n_bases = nb; o->n_n = nn;
    return (int)n;
}

/* Cut points for parsing one FASTQ text on several threads (salt_fastq_pack is re-entrant, and every part can travel as
 * chunks of its own: the compact transport does not care where a chunk's bases start).  A cut is the offset of a line that
 * starts with '@' and whose next-but-one line starts with '+': a record header of the usual four-line layout.  A quality line
 * may start with '@' too, but the line two below it is then a sequence line. */
int salt_fastq_split(const char *text, size_t len, int n_parts, size_t *cuts)
{
    if (!text || !cuts || n_parts < 1) return SALT_ERR_ARG;
    int made = 0;
    cuts[0] = 0;
    for (int k = 1; k < n_parts; ++k) {
        size_t s = (size_t)((unsigned long long)len * (unsigned long long)k / (unsigned long long)n_parts);
        if (s <= cuts[made]) continue;
        const char *e = (const char *)memchr(text + s, '\n', len - s);
        if (!e) break;                                          /* the rest is one line: no further cut */
        s = (size_t)(e - text) + 1;
        int found = 0;
        for (int tries = 0; tries < 4096 && s < len; ++tries) {
            const char *e1 = (const char *)memchr(text + s, '\n', len - s);
            if (!e1) break;
            const size_t l1 = (size_t)(e1 - text) + 1;
            if (text[s] == '@' && l1 < len) {
                const char *e2 = (const char *)memchr(text + l1, '\n', len - l1);
                if (e2) {
                    const size_t l2 = (size_t)(e2 - text) + 1;
                    if (l2 < len && text[l2] == '+') { found = 1; break; }
                }
            }
            s = l1;
        }
        if (!found) { if (s >= len) break; return SALT_ERR_UNSUPPORTED; }       /* records span several lines: parse on one thread */
        if (s > cuts[made]) cuts[++made] = s;
    }
    cuts[++made] = len;
    return made;
}

/* ------------------------------------------------------------------ SAM text (sam.c:86-180, :186-240, :331-455) */
typedef struct { char *s; size_t l, cap; int ovf; } sam_buf_t;
static inline void sb_putc(sam_buf_t *b, char c) { if (b->l + 1 < b->cap) b->s[b->l] = c; else b->ovf = 1; ++b->l; }
static inline void sb_puts(sam_buf_t *b, const char *t) { while (*t) sb_putc(b, *t++); }
static inline void sb_putu(sam_buf_t *b, uint64_t v)
{
    char tmp[24]; int n = 0;
    do { tmp[n++] = (char)('0' + v % 10); v /= 10; } while (v);
    while (n) sb_putc(b, tmp[--n]);
}
static inline void sb_puti(sam_buf_t *b, int64_t v) { if (v < 0) { sb_putc(b, '-'); sb_putu(b, (uint64_t)(-v)); } else sb_putu(b, (uint64_t)v); }

/* bns_coor_pac2real (bntseq.c:269-284): the record a coordinate lies in */
static int sam_rid(const salt_sam_refs_t *r, int64_t pac_coor)
{
    int left = 0, mid = 0, right = r->n_seqs;
    while (left < right) {
        mid = (left + right) >> 1;
        if (pac_coor >= r->offsets[mid]) {
            if (mid == r->n_seqs - 1) break;
            if (pac_coor < r->offsets[mid + 1]) break;
            left = mid + 1;
        } else right = mid;
    }
    return mid;
}

static void sam_seq_qual(sam_buf_t *b, const salt_sam_read_t *q)
{
    const uint32_t L = q->l_seq;
    const int has_qual = q->qual && q->qual[0];
    if (q->strand == 1) {                                    /* query->rseq: reversed, A<->T C<->G, N stays (query.c:46-64) */
        for (uint32_t j = 0; j < L; ++j) { const uint8_t c = q->seq[L - 1 - j]; sb_putc(b, "ACGTN"[c < 4 ? 3 - c : 4]); }
        sb_putc(b, '\t');
        if (has_qual) for (uint32_t j = L; j-- > 0;) sb_putc(b, q->qual[j]);
        else sb_putc(b, '*');
    } else {
        for (uint32_t j = 0; j < L; ++j) sb_putc(b, "ACGTN"[q->seq[j] < 4 ? q->seq[j] : 4]);
        sb_putc(b, '\t');
        if (has_qual) sb_puts(b, q->qual); else sb_putc(b, '*');
    }
}

/* sam_add_xa (sam.c:186-240) */
static int sam_xa(sam_buf_t *b, const salt_sam_refs_t *r, const salt_sam_read_t *q, int is_cigar)
{
    int first = 1, k_gap = 0;
    for (int strand = 0; strand < 2; ++strand)
        for (int i = 0; i < q->n_alt[strand]; ++i) {
            const salt_hit_t *h = q->alt[strand] + i;
            if (h->pos == q->pos) continue;
            if ((int64_t)h->pos >= r->l_pac) return SALT_ERR_ARG;
            if (first) { sb_puts(b, "\tXA:Z:"); first = 0; }
            const int rid = sam_rid(r, (int64_t)h->pos);
            sb_puts(b, r->names[rid]); sb_putc(b, ',');
            sb_putc(b, "+-"[strand]); sb_puti(b, (int64_t)h->pos - r->offsets[rid] + 1); sb_putc(b, ',');
            if (is_cigar) {
                if (h->is_gap) {
                    if (!q->xa_cigars || !q->xa_cigars[k_gap]) return SALT_ERR_ARG;
                    sb_puts(b, q->xa_cigars[k_gap++]); sb_putc(b, ',');
                } else { sb_putu(b, q->l_seq); sb_puts(b, "M,"); }
            } else sb_puts(b, "*,");
            sb_putu(b, h->n_diff); sb_putc(b, ';');
        }
    return SALT_OK;
}

static void sam_tags(sam_buf_t *b, const salt_sam_read_t *q, const char *rg_id)
{
    if (q->md && q->pos != 0xFFFFFFFFu) {                    /* sam_add_md_nm (sam.c:246-328) */
        sb_puts(b, "\tMD:Z:"); sb_puts(b, q->md);
        sb_puts(b, "\tNM:i:"); sb_putu(b, q->nm);
        if (q->n_xv > 0) {
            sb_puts(b, "\tXV:i:");
            for (int i = 0; i < q->n_xv; ++i) { if (i) sb_putc(b, ','); sb_putu(b, q->xv[i]); }
        }
    }
    if (rg_id) { sb_puts(b, "\tRG:Z:"); sb_puts(b, rg_id); }
}

int salt_sam_se(const salt_sam_refs_t *refs, const salt_sam_read_t *q, int print_xa_cigar, const char *rg_id, char *out, size_t cap)
{
    if (!refs || !q || !out || !cap || !q->name || (q->l_seq && !q->seq)) return SALT_ERR_ARG;
    sam_buf_t b = {out, 0, cap, 0};
    sb_puts(&b, q->name); sb_putc(&b, '\t');
    if (q->pos == 0xFFFFFFFFu) {                             /* sam.c:104-122: no tags, the read as it was sequenced */
        sb_puts(&b, "4\t*\t0\t0\t*\t*\t0\t0\t");
        for (uint32_t j = 0; j < q->l_seq; ++j) sb_putc(&b, "ACGTN"[q->seq[j] < 4 ? q->seq[j] : 4]);
        sb_putc(&b, '\t');
        if (q->qual) sb_puts(&b, q->qual); else sb_putc(&b, '*');
    } else {
        if ((int64_t)q->pos >= refs->l_pac || !q->cigar) return SALT_ERR_ARG;
        const int rid = sam_rid(refs, (int64_t)q->pos);
        sb_putu(&b, q->strand ? 16u : 0u); sb_putc(&b, '\t');
        sb_puts(&b, refs->names[rid]); sb_putc(&b, '\t');
        sb_puti(&b, (int64_t)q->pos - refs->offsets[rid] + 1); sb_putc(&b, '\t');
        sb_putu(&b, q->mapq); sb_putc(&b, '\t');
        sb_puts(&b, q->cigar);
        sb_puts(&b, "\t*\t0\t0\t");
        sam_seq_qual(&b, q);
        const int rc = sam_xa(&b, refs, q, print_xa_cigar);
        if (rc != SALT_OK) return rc;
        sam_tags(&b, q, rg_id);
    }
    if (b.ovf || b.l >= cap) return SALT_ERR_NOMEM;
    out[b.l] = '\0';
    return (int)b.l;
}

int salt_sam_pe(const salt_sam_refs_t *refs, const salt_sam_read_t q[2], uint32_t min_tlen, uint32_t max_tlen, int print_xa_cigar,
                const char *rg_id, char *out0, size_t cap0, char *out1, size_t cap1, int len[2])
{
    if (!refs || !q || !out0 || !out1 || !cap0 || !cap1 || !len) return SALT_ERR_ARG;
    int rid[2] = {-1, -1}, is_map[2] = {0, 0};
    uint32_t pos[2] = {0, 0};
    for (int i = 0; i < 2; ++i) {
        if (!q[i].name || (q[i].l_seq && !q[i].seq)) return SALT_ERR_ARG;
        if (q[i].pos != 0xFFFFFFFFu) {
            if ((int64_t)q[i].pos >= refs->l_pac || !q[i].cigar) return SALT_ERR_ARG;
            is_map[i] = 1;
            rid[i] = sam_rid(refs, (int64_t)q[i].pos);
            pos[i] = (uint32_t)((int64_t)q[i].pos - refs->offsets[rid[i]] + 1);
        }
    }
    int tlen = 0;                                            /* sam.c:352-358, its second branch reads q[1].seq_start as written there */
    if (is_map[0] && is_map[1]) {
        if (rid[0] != rid[1]) tlen = 0;
        else if (pos[0] < pos[1]) tlen = (int)(pos[1] + q[1].seq_end - q[1].seq_start + 1 - pos[0]);
        else tlen = (int)(pos[0] + q[0].seq_end - q[1].seq_start + 1 - pos[1]);
        if ((uint32_t)tlen > max_tlen || (uint32_t)tlen < min_tlen) tlen = 0;
    }
    char *outs[2] = {out0, out1}; const size_t caps[2] = {cap0, cap1};
    for (int i = 0; i < 2; ++i) {
        sam_buf_t b = {outs[i], 0, caps[i], 0};
        sb_puts(&b, q[i].name); sb_putc(&b, '\t');
        unsigned flag = 0x1;
        if (!is_map[i]) flag |= 0x4;
        if (!is_map[1 - i]) flag |= 0x8;
        if (q[i].strand == 1) flag |= 0x10;
        if (q[1 - i].strand == 1) flag |= 0x20;
        if (tlen != 0) flag |= 0x2;
        flag |= i == 0 ? 0x40 : 0x80;
        sb_putu(&b, flag); sb_putc(&b, '\t');
        if (is_map[i]) {
            sb_puts(&b, refs->names[rid[i]]); sb_putc(&b, '\t');
            sb_putu(&b, pos[i]); sb_putc(&b, '\t');
            sb_putu(&b, q[i].mapq); sb_putc(&b, '\t');
            if (q[i].seq_start != 0) { sb_puti(&b, (int)q[i].seq_start); sb_putc(&b, 'S'); }
            sb_puts(&b, q[i].cigar);
            if (q[i].seq_end != q[i].l_seq - 1) { sb_puti(&b, (int)(q[i].l_seq - q[i].seq_end - 1)); sb_putc(&b, 'S'); }
            sb_putc(&b, '\t');
        } else if (is_map[1 - i]) {
            sb_puts(&b, refs->names[rid[1 - i]]); sb_putc(&b, '\t');
            sb_putu(&b, pos[1 - i]); sb_putc(&b, '\t');
            sb_puts(&b, "255\t*\t");
        } else sb_puts(&b, "*\t0\t255\t*\t");
        if (is_map[1 - i]) {                                 /* Rnext, Pnext */
            if (rid[i] == rid[1 - i] || !is_map[i]) sb_puts(&b, "=\t");
            else { sb_puts(&b, refs->names[rid[1 - i]]); sb_putc(&b, '\t'); }
            sb_putu(&b, pos[1 - i]); sb_putc(&b, '\t');
        } else sb_puts(&b, "*\t0\t");
        if (tlen != 0) {
            if (q[i].pos >= q[1 - i].pos) sb_putc(&b, '-');
            sb_puti(&b, tlen); sb_putc(&b, '\t');
        } else sb_puts(&b, "0\t");
        sam_seq_qual(&b, &q[i]);
        const int rc = sam_xa(&b, refs, &q[i], print_xa_cigar);
        if (rc != SALT_OK) return rc;
        sam_tags(&b, &q[i], rg_id);
        sb_putc(&b, '\n');
        if (b.ovf || b.l >= caps[i]) return SALT_ERR_NOMEM;
        outs[i][b.l] = '\0';
        len[i] = (int)b.l;
    }
    return SALT_OK;
}
