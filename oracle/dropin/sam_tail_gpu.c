/* oracle/dropin/sam_tail_gpu.c -- TEST INFRASTRUCTURE ONLY (part of oracle/_ref/salt_dropin).
 *
 * The MD / NM / XV tags of a whole chunk from ONE salt_b200_md_nm call.  sam.c is compiled
 * unmodified (position independent, its sam_add_md_nm weakened with objcopy, see oracle/Makefile);
 * the definition below takes its place at link time, so aln_samse / alnpe_sam (sam.c:180, :448)
 * print the engine's tags.  The chunk loops call dropin_tail_prepare once all query_t fields of the
 * chunk are final and before the first SAM line is formatted. */
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

static struct {
    const query_t *base; int n_q;
    int *idx;                       /* per query: row in the result arrays, -1 = none */
    salt_mdnm_out_t *out; char *md; uint16_t *xv;
} T;

static void tail_die(const char *what)
{
    fprintf(stderr, "[salt_dropin] %s: %s\n", what, salt_b200_last_error());
    exit(1);
}

/* queries [first, upto) of multi_seqs; slot_of[j] = index of read j among the reads resident in `slot` */
void dropin_tail_prepare(salt_b200_t *gpu, int slot, const query_t *multi_seqs, const int *slot_of, int first, int upto)
{
    int j, n = 0;
    size_t cs = 2;
    T.base = multi_seqs; T.n_q = upto;
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
