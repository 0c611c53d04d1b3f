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
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

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

int alnse_core(const opt_t *opt)
{
    fprintf(stderr, "[alnse_core/gpu]:  Start single end alignment (verification on libsalt_b200)\n");
    aln_opt_t *aln_opt = aln_opt_init(opt);
    index_t *index = alnse_index_reload(opt->fn_index);
    if (aln_opt->extend_algo == EXTEND_SW || aln_opt->l_overlap <= 0) {
        fprintf(stderr, "[salt_dropin] only the overlap/LV path (the reference's default) is served\n");
        exit(1);
    }
    salt_b200_t *gpu = salt_b200_init(index->mixRef->seq, index->mixRef->l, index->pac, index->bntseq->l_pac, 0);
    if (!gpu) die("salt_b200_init");
    salt_chunk_t *ck = salt_chunk_new(DROPIN_CHUNK_READS, (size_t)DROPIN_CHUNK_READS * 1024, DROPIN_CHUNK_CANDS);
    if (!ck) die("salt_chunk_new");
    aux_t *aux[2];
    aux[0] = aux_init(opt->l_read, opt->l_seed);
    aux[1] = aux_init(opt->l_read, opt->l_seed);

    queryio_t *qs = query_open(opt->fn_read1);
    query_t *multiSeqs = calloc(N_SEQS, sizeof(query_t));
    int *slot_of = calloc(N_SEQS, sizeof(int));        /* index of read i in the GPU chunk being filled */
    aln_samhead(opt, index->bntseq);
    int n, i, n_tot = 0;
    while ((n = query_read_multiSeqs(qs, N_SEQS, multiSeqs)) > 0) {
        n_tot += n;
        int first = 0;                                 /* reads [first, i) are queued in ck */
        salt_chunk_reset(ck);
        for (i = 0; i <= n; ++i) {
            int flush = (i == n);
            if (!flush) {
                query_t *query = multiSeqs + i;
                slot_of[i] = -1;
                if (query->n_ambiguous > DROPIN_MAX_N_PERSEQ) continue;          /* alnse.c:1328 */
                if (query->l_seq - aln_opt->l_seed + 1 > aux[0]->n_sai_range) {
                    int n_sai_range = query->l_seq - aln_opt->l_seed + 1;
                    aux_resize(aux[0], n_sai_range);
                    aux_resize(aux[1], n_sai_range);
                }
                aux_reset(aux[0]);
                aux_reset(aux[1]);
                /* seeding + locate: the reference's own code, the first half of alnse_overlap_alt (alnse.c:1065-1068) */
                alnse_seed_overlap(index, query->l_seq, query->seq, aln_opt, aux[0]);
                alnse_locate_alt(index, query->l_seq, aln_opt->max_locate, aux[0]);
                alnse_seed_overlap(index, query->l_seq, query->rseq, aln_opt, aux[1]);
                alnse_locate_alt(index, query->l_seq, aln_opt->max_locate, aux[1]);
                int at = salt_chunk_add_read(ck, query->seq, query->l_seq, aux[0]->loci.a, aux[0]->loci.n,
                                             aux[1]->loci.a, aux[1]->loci.n);
                if (at == SALT_ERR_NOMEM && salt_chunk_n_reads(ck) > 0) { flush = 1; --i; }   /* queue full: verify what is queued, retry */
                else if (at < 0) die("salt_chunk_add_read");
                else slot_of[i] = at;
            }
            if (flush) {
                int upto = (i < n) ? i + 1 : n, j;
                if (salt_chunk_n_reads(ck) > 0) {
                    /* SE thresholds: nogap 3 (alnse.c:1079), gapped l_seq/10 (alnse.c:1090) */
                    if (salt_chunk_submit(gpu, 0, ck, 3, -1) != SALT_OK) die("salt_chunk_submit");
                    if (salt_chunk_wait(gpu, 0, ck) != SALT_OK) die("salt_chunk_wait");
                }
                for (j = first; j < upto; ++j)
                    if (slot_of[j] >= 0) finish_read(index, multiSeqs + j, aln_opt, ck, (uint32_t)slot_of[j]);
                if (aln_opt->print_nm_md || aln_opt->print_xa_cigar) dropin_tail_prepare(gpu, 0, multiSeqs, slot_of, first, upto);
                for (j = first; j < upto; ++j)
                    if (slot_of[j] >= 0) aln_samse(index, multiSeqs + j, aln_opt);     /* alnse.c:1307 / :1345 */
                first = upto;
                salt_chunk_reset(ck);
            }
        }
        for (i = 0; i < n; ++i) {
            query_t *query = multiSeqs + i;
            puts(query->sam->s);                        /* alnse.c:1433-1439 */
            query_destroy(query);
        }
        memset(multiSeqs, 0, N_SEQS * sizeof(query_t));
        fprintf(stderr, "%d reads have been aligned!\n", n_tot);
    }
    aux_destroy(aux[0]);
    aux_destroy(aux[1]);
    free(multiSeqs); free(slot_of);
    query_close(qs);
    salt_chunk_free(ck);
    dropin_tail_report();
    salt_b200_destroy(gpu);
    alnse_index_destroy(index);
    aln_opt_destroy(aln_opt);
    return EXIT_SUCCESS;
}
