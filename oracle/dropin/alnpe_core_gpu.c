/*
 * oracle/dropin/alnpe_core_gpu.c -- TEST INFRASTRUCTURE / integration proof (paired-end).
 *
 * Replaces ONE function of the reference, alnpe_core (alnpe.c:530-661; alnpe.c is compiled with
 * -Dalnpe_core=alnpe_core_reference).  Per chunk of pairs:
 *   phase 1  seeding + locate of every mate with the reference's own alnse_seed_overlap /
 *            alnse_locate (the first half of alnse_overlap, alnse.c:1001-1004)
 *   phase 2  the verification stage of all mates on the GPU with the paired-end thresholds
 *            (nogap 3, gapped 3: alnse.c:1016,1027) + query_set_hits -- salt_host.h
 *   phase 3  the reference's own pairing2 / pairing_singleton (mate rescue included) and alnpe_sam.
 * Only phase 2 leaves the reference's code.  (Batching phase 3's Smith-Waterman rescues through
 * salt_b200_ssw needs pairing2 re-staged; the kernel itself is parity-tested
 * against ssw_align separately.)
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "kvec.h"
#include "aln.h"
#include "sam.h"
#include "salt_host.h"

#define DROPIN_PE_MAX_N_PERSEQ 5          /* alnpe.c:481 */
#define DROPIN_CHUNK_READS 20000u
#define DROPIN_CHUNK_CANDS (20000u * 256u)

int pairing2(index_t *index, query_t *q0, query_t *q1, const aln_opt_t *aln_opt);
int pairing_singleton(index_t *index, query_t *q0, query_t *q1, aln_opt_t *aln_opt);
void alnpe_sam(index_t *index, query_t *q, const aln_opt_t *opt);

static void die(const char *what)
{
    fprintf(stderr, "[salt_dropin/pe] %s: %s\n", what, salt_b200_last_error());
    exit(1);
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

int alnpe_core(const opt_t *opt)
{
    int i;
    aln_opt_t *aln_opt = aln_opt_init(opt);
    fprintf(stderr, "[alnpe_core/gpu]:  Start paired end alignment (verification on libsalt_b200)\n");
    index_t *index = alnpe_index_reload(opt->fn_index);
    if (aln_opt->l_overlap <= 0) { fprintf(stderr, "[salt_dropin/pe] only the overlap path is served\n"); exit(1); }
    salt_b200_t *gpu = salt_b200_init(index->mixRef->seq, index->mixRef->l, index->pac, index->bntseq->l_pac, 0);
    if (!gpu) die("salt_b200_init");
    salt_chunk_t *ck = salt_chunk_new(DROPIN_CHUNK_READS, (size_t)DROPIN_CHUNK_READS * 1024, DROPIN_CHUNK_CANDS);
    if (!ck) die("salt_chunk_new");
    queryio_t *qs[2];
    qs[0] = query_open(opt->fn_read1);
    qs[1] = query_open(opt->fn_read2);
    aux_t *aux[2];
    aux[0] = aux_init(opt->l_read, opt->l_seed);
    aux[1] = aux_init(opt->l_read, opt->l_seed);
    aln_samhead(opt, index->bntseq);

    int n, tot = 0;
    query_t *multi_seqs = calloc(N_SEQS, sizeof(query_t));
    int *slot_of = calloc(N_SEQS, sizeof(int));
    while ((n = query_read_multiPairedSeqs(qs, N_SEQS, multi_seqs)) > 0) {
        if (opt->max_tlen == 0) { fprintf(stderr, "infer isize func haven't been implemented\n"); break; }
        int first = 0;                               /* mates [first, i) are queued; first is always even */
        salt_chunk_reset(ck);
        for (i = 0; i <= n; ++i) {
            int flush = (i == n);
            if (!flush) {
                query_t *query = multi_seqs + i;
                slot_of[i] = -1;
                if (query->n_ambiguous > DROPIN_PE_MAX_N_PERSEQ) continue;               /* alnpe.c:495 */
                if (query->l_seq - aln_opt->l_seed + 1 > aux[0]->n_sai_range) {
                    int n_sai_range = query->l_seq - aln_opt->l_seed + 1;
                    aux_resize(aux[0], n_sai_range);
                    aux_resize(aux[1], n_sai_range);
                }
                aux_reset(aux[0]);
                aux_reset(aux[1]);
                alnse_seed_overlap(index, query->l_seq, query->seq, aln_opt, aux[0]);
                alnse_locate(index, query->l_seq, aln_opt->max_locate, aux[0]);
                alnse_seed_overlap(index, query->l_seq, query->rseq, aln_opt, aux[1]);
                alnse_locate(index, query->l_seq, aln_opt->max_locate, aux[1]);
                int at = salt_chunk_add_read(ck, query->seq, query->l_seq, aux[0]->loci.a, aux[0]->loci.n,
                                             aux[1]->loci.a, aux[1]->loci.n);
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
                    if (slot_of[j] >= 0) apply_result(multi_seqs + j, aln_opt, ck, (uint32_t)slot_of[j]);
                for (j = first; j + 1 < upto + 1 && j < upto; j += 2) {                  /* alnpe.c:507-519 */
                    query_t *q0 = multi_seqs + j, *q1 = multi_seqs + j + 1;
                    if (q0->pos != 0xFFFFFFFF && q1->pos != 0xFFFFFFFF) pairing2(index, q0, q1, aln_opt);
                    else if (q0->pos != 0xFFFFFFFF || q1->pos != 0xFFFFFFFF) pairing_singleton(index, q0, q1, aln_opt);
                    alnpe_sam(index, multi_seqs + j, aln_opt);
                }
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
    aux_destroy(aux[0]);
    aux_destroy(aux[1]);
    query_close(qs[0]);
    query_close(qs[1]);
    free(multi_seqs); free(slot_of);
    salt_chunk_free(ck);
    salt_b200_destroy(gpu);
    alnpe_index_destroy(index);
    aln_opt_destroy(aln_opt);
    return 0;
}
