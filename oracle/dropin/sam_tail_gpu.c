/* oracle/dropin/sam_tail_gpu.c -- TEST INFRASTRUCTURE ONLY (part of oracle/_ref/salt_dropin).
 *
 * The MD / NM / XV tags of a whole chunk from ONE salt_b200_md_nm call.  sam.c is compiled
 * unmodified (position independent, its sam_add_md_nm weakened with objcopy, see oracle/Makefile);
 * the definition below takes its place at link time, so aln_samse / alnpe_sam (sam.c:180, :448)
 * print the engine's tags.  sam.c's one call of ed_diff_withcigar (sam.c:218, the CIGAR of each
 * gapped XA alternate under -c) is renamed to dropin_xa_cigar below, which hands out the strings of
 * ONE salt_b200_lv_cigar call per chunk.  The chunk loops call dropin_tail_prepare once all query_t
 * fields of the chunk are final and before the first SAM line is formatted. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
#include "aln.h"
#include "query.h"
#include "kstring.h"
#include "salt_b200.h"

#define TAIL_MD_STRIDE 256
#define TAIL_XV_STRIDE 64

#define TAIL_XA_STRIDE 256          /* sam.c:216 callocs 256 bytes */

static struct {
    const query_t *base; int n_q;
    int *idx;                       /* per query: row in the result arrays, -1 = none */
    salt_mdnm_out_t *out; char *md; uint16_t *xv;
    /* gapped XA alternates in the order sam_add_xa visits them (sam.c:195-197) */
    size_t n_xa, xa_cursor;
    size_t *xa_first; int n_first;  /* per query: its first row among the XA alternates */
    size_t tot_md, tot_xa, used_md, used_xa;       /* prepared on the GPU / printed into SAM lines */
    const uint8_t **xa_seq; uint32_t *xa_pos; int8_t *xa_e; char *xa_cig;
} T;

static __thread size_t t_xa_cursor;

static void tail_die(const char *what)
{
    fprintf(stderr, "[salt_dropin] %s: %s\n", what, salt_b200_last_error());
    exit(1);
}

/* CIGARs of the chunk's gapped XA alternates (sam.c:195-231): one salt_b200_lv_cigar call */
static void dropin_xa_prepare(salt_b200_t *gpu, int slot, const query_t *multi_seqs, const int *slot_of, int first, int upto)
{
    int j, strand;
    size_t i, n = 0, cap = 0;
    salt_pair_t *pairs = NULL; uint8_t *k_each = NULL;
    T.n_xa = 0; T.xa_cursor = 0; t_xa_cursor = 0;
    T.xa_first = realloc(T.xa_first, (size_t)(upto + 1) * sizeof *T.xa_first); T.n_first = upto;
    for (j = 0; j < upto; ++j) T.xa_first[j] = 0;
    if (salt_b200_use_slot(gpu, slot) != SALT_OK) tail_die("salt_b200_use_slot");      /* the per-pair entry points follow the chunk's slot */
    for (j = first; j < upto; ++j) {
        const query_t *q = multi_seqs + j;
        T.xa_first[j] = n;
        if (slot_of[j] < 0) continue;
        for (strand = 0; strand < 2; ++strand)
            for (i = 0; i < q->hits[strand].n; ++i) {
                const hit_t *hit = q->hits[strand].a + i;
                if (hit->pos == q->pos || !hit->is_gap) continue;          /* sam.c:205, :215 */
                if (n == cap) {
                    cap = cap ? cap * 2 : 1024;
                    pairs = realloc(pairs, cap * sizeof *pairs); k_each = realloc(k_each, cap);
                    T.xa_seq = realloc(T.xa_seq, cap * sizeof *T.xa_seq); T.xa_pos = realloc(T.xa_pos, cap * sizeof *T.xa_pos);
                }
                pairs[n].rs = ((uint32_t)slot_of[j] << 1) | (uint32_t)strand; pairs[n].pos = hit->pos;
                k_each[n] = hit->n_diff;
                T.xa_seq[n] = strand == 0 ? q->seq : q->rseq; T.xa_pos[n] = hit->pos;
                ++n;
            }
    }
    if (n) {
        T.xa_e = realloc(T.xa_e, n);
        T.xa_cig = realloc(T.xa_cig, n * TAIL_XA_STRIDE);
        memset(T.xa_cig, 0, n * TAIL_XA_STRIDE);
        if (salt_b200_lv_cigar(gpu, pairs, k_each, n, T.xa_cig, TAIL_XA_STRIDE, T.xa_e) != SALT_OK) tail_die("salt_b200_lv_cigar");
    }
    salt_b200_use_slot(gpu, 0);
    T.n_xa = n; T.tot_xa += n;
    free(pairs); free(k_each);
}

/* The SAM text of a chunk may be written by several threads (as the reference's workers do, alnse.c:1307): each thread
 * announces the read it is about to print, and the XA hook then walks that read's own alternates. */
void dropin_tail_begin_read(int j) { t_xa_cursor = (T.xa_first && j >= 0 && j < T.n_first) ? T.xa_first[j] : 0; }

/* replaces the ed_diff_withcigar call at sam.c:218 (renamed with -D on sam.c only) */
int dropin_xa_cigar(const uint32_t *mixRef, uint32_t ref_st, uint32_t l_ref, const uint8_t *seq, uint32_t l_seq,
                    int max_k_diff, char *cigarBuf, int cigarLen, int useM, int cigarFormat)
{
    (void)mixRef; (void)l_ref; (void)l_seq; (void)max_k_diff; (void)useM; (void)cigarFormat;
    while (t_xa_cursor < T.n_xa && !(T.xa_seq[t_xa_cursor] == seq && T.xa_pos[t_xa_cursor] == ref_st)) ++t_xa_cursor;
    if (t_xa_cursor >= T.n_xa) { fprintf(stderr, "[salt_dropin] no XA CIGAR prepared for pos %u\n", ref_st); exit(1); }
    const size_t k = t_xa_cursor++;
    __atomic_fetch_add(&T.used_xa, 1, __ATOMIC_RELAXED);
    strncpy(cigarBuf, T.xa_cig + k * TAIL_XA_STRIDE, (size_t)cigarLen - 1);
    return T.xa_e[k];
}

/* queries [first, upto) of multi_seqs; slot_of[j] = index of read j among the reads resident in `slot` */
void dropin_tail_prepare(salt_b200_t *gpu, int slot, const query_t *multi_seqs, const int *slot_of, int first, int upto)
{
    int j, n = 0;
    size_t cs = 2;
    T.base = multi_seqs; T.n_q = upto;
    dropin_xa_prepare(gpu, slot, multi_seqs, slot_of, first, upto);
    T.idx = realloc(T.idx, (size_t)(upto + 1) * sizeof *T.idx);
    for (j = 0; j < upto; ++j) T.idx[j] = -1;
    for (j = first; j < upto; ++j) {
        const query_t *q = multi_seqs + j;
        if (slot_of[j] < 0 || q->pos == 0xFFFFFFFF) continue;
        if (strlen(q->cigar->s) + 1 > cs) cs = strlen(q->cigar->s) + 1;
        T.idx[j] = n++;
    }
    if (!n) return;
    salt_mdnm_in_t *in = malloc((size_t)n * sizeof *in);
    char *cg = calloc((size_t)n, cs);
    T.out = realloc(T.out, (size_t)n * sizeof *T.out);
    T.md = realloc(T.md, (size_t)n * TAIL_MD_STRIDE);
    T.xv = realloc(T.xv, (size_t)n * TAIL_XV_STRIDE * sizeof *T.xv);
    for (j = first; j < upto; ++j) {
        const query_t *q = multi_seqs + j;
        const int k = T.idx[j];
        if (k < 0) continue;
        in[k].rs = ((uint32_t)slot_of[j] << 1) | (uint32_t)(q->strand & 1);
        in[k].pos = q->pos; in[k].seq_start = q->seq_start;
        strcpy(cg + (size_t)k * cs, q->cigar->s);
    }
    if (salt_b200_md_nm(gpu, slot, in, (size_t)n, cg, (int)cs, T.md, TAIL_MD_STRIDE, T.xv, TAIL_XV_STRIDE, T.out) != SALT_OK)
        tail_die("salt_b200_md_nm");
    T.tot_md += (size_t)n;
    free(in); free(cg);
}

/* replaces sam.c:246 */
void sam_add_md_nm(kstring_t *s, index_t *index, query_t *q)
{
    (void)index;
    if (q->pos == 0xFFFFFFFF) return;                               /* sam.c:248 */
    const long j = q - T.base;
    if (j < 0 || j >= T.n_q || T.idx[j] < 0) { fprintf(stderr, "[salt_dropin] no SAM tail prepared for %s\n", q->name); exit(1); }
    const int k = T.idx[j];
    if (T.out[k].md_len < 0) { fprintf(stderr, "[salt_dropin] MD of %s: engine code %d\n", q->name, T.out[k].md_len); exit(1); }
    __atomic_fetch_add(&T.used_md, 1, __ATOMIC_RELAXED);
    ksprintf(s, "\tMD:Z:%s", T.md + (size_t)k * TAIL_MD_STRIDE);
    ksprintf(s, "\tNM:i:%u", (unsigned)T.out[k].nm);
    if (T.out[k].n_xv > 0) {
        int i;
        ksprintf(s, "\tXV:i:");
        for (i = 0; i < T.out[k].n_xv; ++i) {
            if (i != 0) ksprintf(s, ",");
            ksprintf(s, "%d", (int)T.xv[(size_t)k * TAIL_XV_STRIDE + i]);
        }
    }
}

/* For the host layer's own SAM formatter (salt_sam_se, SALT_DROPIN_SAM=native): the tags of read j and the CIGAR of its k-th
 * printed gapped alternate, as prepared by dropin_tail_prepare.  Returns 0 when read j has no tags (unmapped / not prepared). */
int dropin_tail_get(int j, const char **md, unsigned *nm, const uint16_t **xv, int *n_xv)
{
    if (j < 0 || j >= T.n_q || !T.idx || T.idx[j] < 0) return 0;
    const int k = T.idx[j];
    if (T.out[k].md_len < 0) { fprintf(stderr, "[salt_dropin] MD of read %d: engine code %d\n", j, T.out[k].md_len); exit(1); }
    *md = T.md + (size_t)k * TAIL_MD_STRIDE; *nm = (unsigned)T.out[k].nm;
    *xv = T.xv + (size_t)k * TAIL_XV_STRIDE; *n_xv = T.out[k].n_xv;
    __atomic_fetch_add(&T.used_md, 1, __ATOMIC_RELAXED);
    return 1;
}
const char *dropin_xa_row(int j, int k)
{
    if (!T.xa_first || j < 0 || j >= T.n_first) return NULL;
    const size_t row = T.xa_first[j] + (size_t)k;
    if (row >= T.n_xa) return NULL;
    __atomic_fetch_add(&T.used_xa, 1, __ATOMIC_RELAXED);
    return T.xa_cig + row * TAIL_XA_STRIDE;
}

void dropin_tail_report(void)
{
    fprintf(stderr, "[salt_dropin] SAM tail from the GPU: MD/NM/XV tags prepared %zu printed %zu, XA CIGARs prepared %zu printed %zu\n",
            T.tot_md, T.used_md, T.tot_xa, T.used_xa);
}
