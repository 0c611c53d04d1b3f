/* oracle/dropin/pair_harness.c -- TEST INFRASTRUCTURE ONLY.
 *
 * The reference's own pairing2 / pairing_singleton (alnpe.c:94-257, :395-473, compiled unmodified and position
 * independent into oracle/_ref/libsaltref_pair.so, see oracle/Makefile) behind a plain-buffer entry point, so that
 * tests can pin salt_pair_plan / salt_pair_apply (include/salt_host.h) against them on arbitrary hit lists.
 * alnpe.c's two Smith-Waterman rescue functions are weakened with objcopy and replaced here by recorders that note
 * the window they were asked for and report "nothing found". */
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
#include "aln.h"
#include "query.h"
#include "kstring.h"

int pairing2(index_t *index, query_t *q0, query_t *q1, const aln_opt_t *aln_opt);
int pairing_singleton(index_t *index, query_t *q0, query_t *q1, aln_opt_t *aln_opt);

static struct { query_t *q[2]; int n; int w[4][5]; } P;     /* windows: mate, strand, flavour, start, end */

static int record(uint32_t *start, uint32_t *end, query_t *q, int strand, int flavour)
{
    if (P.n < 4) {
        int *w = P.w[P.n++];
        w[0] = q == P.q[1]; w[1] = strand; w[2] = flavour; w[3] = (int)*start; w[4] = (int)*end;
    }
    return 0;                                                /* SW_UNPAIRED */
}
int snpaln_sw_snpaware(index_t *index, uint32_t *start, uint32_t *end, query_t *q, int strand, const aln_opt_t *opt)
{ (void)index; (void)opt; return record(start, end, q, strand, 16); }
int snpaln_sw(index_t *index, uint32_t *start, uint32_t *end, query_t *q, int strand, const aln_opt_t *opt)
{ (void)index; (void)opt; return record(start, end, q, strand, 5); }

/* qf: [2][4] pos, strand, n_diff, is_gap; hits: [2 mates][2 strands][max_hits][3] pos, n_diff, is_gap; n_hits [2][2].
 * out_q: [2][6] pos, strand, n_diff, is_gap, seq_start, seq_end after pairing; out_w: [4][5].
 * Returns (number of windows) | (the pairing function's return value << 8). */
int ref_pairing(uint32_t l_pac, int min_tlen, int max_tlen, const uint32_t *qf, const int *l_seq, const uint32_t *hits,
                const int *n_hits, int max_hits, uint32_t *out_q, int *out_w, char *out_cigar /* [2][64] */)
{
    bntseq_t bns; memset(&bns, 0, sizeof bns); bns.l_pac = l_pac;
    mixRef_t mr; mr.l = l_pac; mr.seq = calloc(l_pac / 8 + 64, 4);
    index_t index; memset(&index, 0, sizeof index); index.bntseq = &bns; index.mixRef = &mr;
    aln_opt_t opt; memset(&opt, 0, sizeof opt); opt.min_tlen = min_tlen; opt.max_tlen = max_tlen; opt.max_hits = 5;
    query_t q[2]; kstring_t cg[2];
    int m, s, i, rc;
    memset(q, 0, sizeof q); memset(cg, 0, sizeof cg);
    for (m = 0; m < 2; ++m) {
        q[m].l_seq = l_seq[m];
        q[m].seq = calloc((size_t)l_seq[m] + 16, 1); q[m].rseq = calloc((size_t)l_seq[m] + 16, 1);
        q[m].pos = qf[m * 4]; q[m].strand = (int)qf[m * 4 + 1]; q[m].n_diff = (uint8_t)qf[m * 4 + 2]; q[m].is_gap = (uint8_t)qf[m * 4 + 3];
        q[m].cigar = &cg[m]; cg[m].m = 128; cg[m].s = calloc(128, 1);    /* query.c:208-209 */
        q[m].name = (char *)"q";
        for (s = 0; s < 2; ++s)
            for (i = 0; i < n_hits[m * 2 + s]; ++i) {
                const uint32_t *h = hits + (((size_t)m * 2 + s) * max_hits + i) * 3;
                hit_t hh; hh.pos = h[0]; hh.n_diff = (uint8_t)h[1]; hh.is_gap = (uint8_t)h[2]; hh.strand = (uint16_t)s;
                kv_push(hit_t, q[m].hits[s], hh);
            }
    }
    P.q[0] = &q[0]; P.q[1] = &q[1]; P.n = 0;
    if (q[0].pos != 0xFFFFFFFF && q[1].pos != 0xFFFFFFFF) rc = pairing2(&index, &q[0], &q[1], &opt);          /* alnpe.c:507-517 */
    else if (q[0].pos != 0xFFFFFFFF || q[1].pos != 0xFFFFFFFF) rc = pairing_singleton(&index, &q[0], &q[1], &opt);
    else rc = 0;
    for (m = 0; m < 2; ++m) {
        out_q[m * 6] = q[m].pos; out_q[m * 6 + 1] = (uint32_t)q[m].strand; out_q[m * 6 + 2] = q[m].n_diff; out_q[m * 6 + 3] = q[m].is_gap;
        out_q[m * 6 + 4] = q[m].seq_start; out_q[m * 6 + 5] = q[m].seq_end;
        strncpy(out_cigar + m * 64, cg[m].s ? cg[m].s : "", 63);
        free(q[m].seq); free(q[m].rseq); free(cg[m].s); free(q[m].hits[0].a); free(q[m].hits[1].a);
    }
    memcpy(out_w, P.w, sizeof P.w);
    free(mr.seq);
    return P.n | (rc << 8);
}
