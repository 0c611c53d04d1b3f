/*
 * salt_host.h -- host-side C layer over libsalt_b200.so: the part of salt's per-chunk loop that
 * sits between seeding and SAM output, re-staged for a batched engine.
 *
 * In the reference every worker thread runs, per read,
 *     seed + locate  ->  alnse_check_nogap / alnse_check_withgap  ->  query_set_hits  ->  query_gen_cigar
 * (alnse.c:1045-1100 alnse_overlap_alt, alnse.c:985-1040 alnse_overlap, query.c:282-333).
 * With the GPU engine the middle step is done for a whole chunk at once, so the loop becomes
 *     workers: seed + locate, push (read, sorted loci) into a pinned chunk queue      [salt_chunk_add_read]
 *     main   : submit the chunk to a pipeline slot, keep seeding the next chunk        [salt_chunk_submit]
 *     main   : wait; per read rebuild aux->hits, run query_set_hits / gen_mapq         [salt_chunk_wait, salt_chunk_result]
 * Plain C, no CUDA types.  Compiled by gcc into libsalt_host.so, which links libsalt_b200.so.
 */
#ifndef SALT_HOST_H
#define SALT_HOST_H
#include <stddef.h>
#include <stdint.h>
#include "salt_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

#define SALT_MAX_HITS 16            /* aln_opt->max_hits is 5 in the reference (aln.h:139) */

/* hit_t (query.h:28-33) */
typedef struct { uint32_t pos; uint8_t n_diff; uint8_t is_gap; uint16_t strand; } salt_hit_t;

/* What the verification stage leaves in query_t for one read (query.h:37-64):
 * pos/strand/n_diff/is_gap (alnse.c:348-393), hits[2] + b0/b1/mapq (query_set_hits, query.c:297-333;
 * gen_mapq, query.c:270-281), cigar (query_gen_cigar, query.c:282-295). */
typedef struct {
    uint32_t pos;                    /* 0xFFFFFFFF = unmapped */
    uint8_t strand, n_diff, is_gap;  /* 3 / 255 / 255 when unmapped */
    int b0, b1;
    uint32_t mapq;
    int n_alt[2];
    salt_hit_t alt[2][SALT_MAX_HITS];
    char cigar[128];                 /* "" when unmapped; "<l_seq>M" for an ungapped primary */
} salt_read_result_t;

typedef struct salt_chunk salt_chunk_t;

/* Pinned queues for one chunk: up to max_reads reads, max_bases read bases, max_cands candidate
 * loci per strand.  (N_SEQS = 100000 reads per chunk in the reference, aln.h:27.) */
salt_chunk_t *salt_chunk_new(uint32_t max_reads, size_t max_bases, size_t max_cands);
void salt_chunk_free(salt_chunk_t *c);
void salt_chunk_reset(salt_chunk_t *c);
uint32_t salt_chunk_n_reads(const salt_chunk_t *c);

/* Append read i = salt_chunk_n_reads() with its candidate loci per strand as alnse_locate[_alt]
 * leaves them in aux->loci (sorted ascending).  seq: codes 0..4 (query->seq).  Call in read order
 * from one thread (e.g. after the seeding workers have joined, alnse.c:1428).  Returns the read's
 * index in the chunk, or a negative SALT_ERR_* when a queue is full. */
int salt_chunk_add_read(salt_chunk_t *c, const uint8_t *seq, uint32_t l_seq,
                        const uint32_t *loci0, uint32_t n0, const uint32_t *loci1, uint32_t n1);

/* Send the chunk through pipeline slot `slot` (salt_b200_verify_submit) / wait for it.
 * nogap_T0 = 3 (alnse.c:1016,1079); lv_T0 < 0 = l_seq/10, the SE rule (alnse.c:1090), 3 = PE (alnse.c:1027). */
int salt_chunk_submit(salt_b200_t *h, int slot, salt_chunk_t *c, int nogap_T0, int lv_T0);
int salt_chunk_wait(salt_b200_t *h, int slot, salt_chunk_t *c);

/* Seeding on the device: the chunk's reads were added with n0 = n1 = 0; this uploads them,
 * runs alnse_seed_overlap + alnse_locate_alt and the verification stage on the GPU (salt_b200_set_index must have been
 * called), and brings the candidate lists back into the chunk so that salt_chunk_result / salt_chunk_hits work as
 * after salt_chunk_wait.  Synchronous, slot 0.  SALT_ERR_NOMEM when the lists exceed the chunk's max_cands. */
int salt_chunk_seed_verify(salt_b200_t *h, salt_chunk_t *c, const salt_seed_opt_t *opt, int nogap_T0, int lv_T0);

/* After salt_chunk_wait: the query_t fields of read i, with query_set_hits(max_hits) applied to
 * the accepted hits exactly as the reference does (including its use of element 0's n_diff,
 * query.c:317-318).  1 <= max_hits <= SALT_MAX_HITS. */
int salt_chunk_result(const salt_chunk_t *c, uint32_t i, int max_hits, salt_read_result_t *out);

/* salt_chunk_result for every read of the chunk, on the host threads of salt_host_set_threads (the reference does this
 * per read on its -t workers, alnse.c:1306): out[i] = the query_t fields of read i. */
int salt_chunk_results(const salt_chunk_t *c, int max_hits, salt_read_result_t *out);

/* The SAM tail of the chunk's primaries (sam_add_md_nm, sam.c:246-328; option -d): MD string, NM and
 * XV of every mapped read from ONE salt_b200_md_nm call on the reads still resident in `slot`.
 * Call after salt_chunk_wait and before the slot is submitted again.  Needs the 2-bit pac. */
int salt_chunk_tail(salt_b200_t *h, int slot, salt_chunk_t *c);

/* Read i after salt_chunk_tail: the "MD:Z:" value ("" for an unmapped read), *nm = the NM value,
 * *xv / *n_xv = the XV read offsets (NULL / 0 when there are none).  NULL on misuse. */
const char *salt_chunk_md(const salt_chunk_t *c, uint32_t i, int *nm, const uint16_t **xv, int *n_xv);

/* aux->hits of read i on one strand, in acceptance order.  Returns the number written (<= cap). */
int salt_chunk_hits(const salt_chunk_t *c, uint32_t i, int strand, salt_hit_t *out, int cap);

/* ---- paired-end: what happens to a pair after both mates went through the verification stage ----
 * pairing2 (alnpe.c:94-257, both mates mapped) and pairing_singleton (alnpe.c:395-473, one mapped)
 * re-staged as a PLAN: either the pair is proper as it stands or through one combination of the mates'
 * alternates (no rescue), or up to two mate-rescue windows have to be aligned with Smith-Waterman, in the
 * order the reference tries them (the second only if the first finds nothing).  A chunk's windows can then
 * go to salt_b200_ssw in one call per flavour, without a host round trip per pair. */
typedef struct {
    int mate;                        /* the mate aligned into the window (the other one anchors it) */
    int strand;                      /* that mate's strand there */
    int flavour;                     /* 16: SNP-aware, mixRef masks + score_mat2 (snpaln_sw_snpaware, alnpe.c:261);
                                        5: plain, 2-bit pac + score_mat (snpaln_sw, alnpe.c:330) */
    uint32_t start, end;             /* reference bases [start, end] inclusive */
} salt_rescue_t;

typedef struct {
    int paired;                      /* 1: proper pair without rescue; hit[m] = what mate m's primary becomes */
    salt_hit_t hit[2];
    int n_win;                       /* 0..2 rescue windows */
    salt_rescue_t win[2];
} salt_pair_plan_t;

/* r0 / r1: the two mates' results (salt_chunk_result, PE thresholds); l0 / l1: their lengths;
 * min_tlen / max_tlen: the -a / -b options (aln.h:127-128); l_pac: reference length.
 * Returns SALT_OK, or SALT_ERR_ARG where the reference would exit (a window starting past the reference). */
int salt_pair_plan(const salt_read_result_t *r0, uint32_t l0, const salt_read_result_t *r1, uint32_t l1,
                   uint32_t min_tlen, uint32_t max_tlen, uint32_t l_pac, salt_pair_plan_t *out);

/* What one mate looks like after pairing -- the query_t fields alnpe_sam prints (sam.c:331-455). */
typedef struct {
    uint32_t pos;                    /* 0xFFFFFFFF = unmapped */
    uint8_t strand, n_diff, is_gap;  /* n_diff / is_gap keep the verification stage's values after a rescue, as in the reference */
    uint32_t seq_start, seq_end;     /* aligned part of the read (soft clips around it, sam.c:392-394) */
    int b0, b1; uint32_t mapq;
    int cigar_kind;                  /* 0 none (unmapped); 1 "<l_seq>M"; 2 the verification stage's CIGAR (r->cigar);
                                        3 Smith-Waterman ops of the rescue; 4 NOT filled: a gapped alternate became the
                                        primary, its CIGAR is one salt_b200_lv_cigar item (pos, strand, k = n_diff) */
    char cigar[256];
} salt_mate_final_t;

/* Apply a plan: `ssw` / `ssw_cigars` (cigar_stride uint32 per window: len << 4 | op, op 0/1/2 = M/I/D) are the
 * salt_b200_ssw results of plan->win[0 .. n_win), in that order; filters / filterd the accept rule of
 * snpaln_sw[_snpaware] (alnpe.c:295 / :362: score1 >= filters and aligned read span >= filterd).  The first window
 * that passes rescues its mate; the other mate keeps its primary.  Returns 1 when the pair ends up aligned as a
 * pair (PAIRED_ALNED), 0 otherwise, negative SALT_ERR_* on misuse; SALT_ERR_UNSUPPORTED when the rescuing window has
 * no complete CIGAR (declined by the engine, longer than cigar_stride, or longer than the 256-byte string). */
int salt_pair_apply(const salt_pair_plan_t *plan, const salt_read_result_t *r0, uint32_t l0,
                    const salt_read_result_t *r1, uint32_t l1, const salt_ssw_out_t *ssw, const uint32_t *ssw_cigars,
                    int cigar_stride, int filters, int filterd, salt_mate_final_t out[2]);

/* Append n reads at once (the chunk's queues are plain arrays: this is three memcpys and three offset loops).
 * roffs / offs0 / offs1: n + 1 offsets each, any base (they are rebased).  Returns the index of the first read added. */
int salt_chunk_add_reads(salt_chunk_t *c, const uint8_t *codes, const uint32_t *roffs, uint32_t n,
                         const uint32_t *offs0, const uint32_t *loci0, const uint32_t *offs1, const uint32_t *loci1);

/* ---- the paired-end stage of one chunk: alnpe_core1 after seeding (alnpe.c:482-528) re-staged for batches ----------
 * The chunk holds mates 2i, 2i+1 of pair i and has been through salt_chunk_submit (PE thresholds 3 / 3) and
 * salt_chunk_wait on `slot`.  This runs, for the whole chunk: query_set_hits per mate (salt_chunk_result), pairing2 /
 * pairing_singleton as plans (salt_pair_plan), ONE salt_b200_ssw call per rescue flavour on the reads still resident
 * in the slot, salt_pair_apply, one salt_b200_lv_cigar call for gapped alternates that became primaries, and one
 * salt_b200_md_nm call for the MD / NM / XV tags of every mapped mate (with_tail != 0; needs the 2-bit pac).
 * out[i] = the two mates of pair i as alnpe_sam would print them; md / nm of mate m of pair i are at row 2i+m (tail_out[].n_xv counts
 * the XV entries; the offsets themselves are not returned here: a caller that prints XV asks salt_b200_md_nm for the finals, as
 * salt_aln does). */
typedef struct {
    salt_mate_final_t mate[2];
    int paired;                      /* PAIRED_ALNED */
} salt_pair_final_t;
typedef struct {
    size_t pairs, proper, windows16, windows5, rescued, promoted, declined;
    double ms_plan, ms_ssw, ms_apply, ms_tail;
} salt_pe_stats_t;
/* host threads for salt_chunk_pair's per-pair loops (the reference runs its pairing on the -t workers, alnpe.c:596-606) */
void salt_host_set_threads(int n);
/* per-read / per-pair loops shorter than `items` per thread stay on fewer threads (256 by default; tests lower it) */
void salt_host_set_grain(uint32_t items);
int salt_chunk_pair(salt_b200_t *h, int slot, salt_chunk_t *c, uint32_t min_tlen, uint32_t max_tlen, uint32_t l_pac,
                    int max_hits, const int8_t *mat16 /* score_mat2, 256 */, const int8_t *mat5 /* score_mat, 25 */,
                    int gapO, int gapE, int filters, int filterd, int with_tail,
                    salt_pair_final_t *out, salt_mdnm_out_t *tail_out, char *tail_md, int md_stride, salt_pe_stats_t *stats);

/* ---- input side: FASTQ text -> the compact transport, in one pass ---------------------------
 * What query_read_seq does per record (query.c:146-239) -- name up to the first blank with a trailing "/<digit>" trimmed
 * (:140-144), comment, bases through nst_nt4_table (A/C/G/T in either case -> 0..3, anything else -> 4 and counted in
 * n_ambiguous), quality string -- but emitting the bases directly at 2 bits per base with the N positions on the side,
 * i.e. a salt_packed_chunk_t ready for salt_b200_set_reads_packed / _verify_submit_packed / _align_batch_packed.  The
 * reverse complement (query->rseq, query.c:46-64) is made on the device.  kseq's grammar is kept: '@' or '>' records,
 * sequence and quality may span lines, the quality block ends when it is as long as the sequence.
 * All arrays are the caller's; name / comment / quality stay in `text` and are returned as offsets.  (A record without a
 * comment reports an empty one; the reference's query->comment holds the previous record's there, query.c:160 -- it is
 * never printed.) */
typedef struct {
    uint8_t *bases; size_t bases_cap;        /* in: room for bases_cap bases (bases_cap / 4 bytes, + 1); out: 2-bit stream */
    uint32_t *n_pos; size_t n_pos_cap;       /* stream positions of non-ACGT bases */
    uint16_t *lens; uint16_t *n_ambiguous;   /* per read (max_reads entries each) */
    uint32_t *name_off; uint16_t *name_len;  /* per read: the trimmed name inside `text` */
    uint32_t *comment_off; uint16_t *comment_len;
    uint32_t *qual_off;                      /* per read: first quality character inside `text` (0xFFFFFFFF: none; multi-line
                                                quality blocks are reported by their first line) */
    /* results */
    uint32_t n_reads; size_t n_bases, n_n;
} salt_fastq_t;

/* Parse complete records from text[0, len) until max_reads records, a full array, or the end of the text.
 * final != 0: the text ends the input (a last record without a trailing newline is complete).  *consumed = bytes of
 * text used by the records returned; the caller re-presents the rest together with the next block.
 * Returns the number of records, or a negative SALT_ERR_*: SALT_ERR_ARG on malformed input (a quality block shorter than
 * its sequence at the end of the input, a read longer than 65535 bases). */
int salt_fastq_pack(const char *text, size_t len, int final, uint32_t max_reads, salt_fastq_t *out, size_t *consumed);

/* Cut points for parsing one FASTQ text on several host threads: cuts[0] = 0 < cuts[1] < ... < cuts[parts] = len, every inner
 * cut the offset of a four-line record's header (a line starting with '@' whose next-but-one line starts with '+'; a quality
 * line may start with '@', the line two below it is then a sequence line).  Each part is a text of its own for
 * salt_fastq_pack(final = 1) and travels as chunks of its own.  cuts needs n_parts + 1 entries.  Returns the number of parts
 * made (fewer than asked when the text is short), SALT_ERR_UNSUPPORTED when no such header is found where a cut is due
 * (records spanning several lines, FASTA): parse on one thread. */
int salt_fastq_split(const char *text, size_t len, int n_parts, size_t *cuts);

/* ---- several GPUs in one process (reads shard, the reference is replicated, no collective) ----
 * salt is one process (alnse.c:1414-1440); this keeps it one: one handle per device, the batch split into contiguous
 * shares, every share through its device's own chunk pipeline on its own host thread, every result written at the
 * position of its read -- input order by construction. */
typedef struct salt_multi salt_multi_t;
salt_multi_t *salt_multi_init(const uint32_t *mixref, uint32_t l, const uint8_t *pac, int64_t l_pac, const int *devices, int n_devices);
void salt_multi_destroy(salt_multi_t *m);
int salt_multi_n(const salt_multi_t *m);
salt_b200_t *salt_multi_handle(salt_multi_t *m, int i);
int salt_multi_verify_batch_packed(salt_multi_t *m, const salt_packed_chunk_t *pc, uint32_t chunk_reads, int nogap_T0, int lv_T0,
                                   salt_verify_out_t *rec, int8_t *acc0, int8_t *acc1, char *cigars, int cigar_stride);

/* ---- SAM text (aln_samse sam.c:86-180, alnpe_sam sam.c:331-455, sam_add_xa :186-240,
 * bns_coor_pac2real bntseq.c:269-303) ----------------------------------------------------------------------------
 * One SAM line per read from the fields the stages above produce, byte for byte what the reference's formatter leaves in
 * query->sam: no per-field vsnprintf, no allocation -- the line goes into the caller's buffer. */
typedef struct {
    int n_seqs; const char *const *names; const int64_t *offsets;      /* bntseq->anns[i].name / .offset */
    int64_t l_pac;
} salt_sam_refs_t;

typedef struct {
    const char *name;                /* query->name (already trimmed) */
    const uint8_t *seq;              /* codes 0..4, forward strand; the reverse complement is made here (query.c:46-64) */
    const char *qual;                /* NULL or "" prints '*' */
    uint32_t l_seq;
    uint32_t pos;                    /* 0xFFFFFFFF = unmapped */
    uint8_t strand; uint32_t mapq;
    const char *cigar;               /* query->cigar->s */
    uint32_t seq_start, seq_end;     /* soft clips around the aligned part (paired-end lines only, sam.c:392-394) */
    int n_alt[2]; const salt_hit_t *alt[2];           /* query->hits[strand] */
    const char *const *xa_cigars;    /* CIGARs of the gapped alternates that get printed, in the order sam_add_xa visits them */
    const char *md; uint32_t nm; const uint16_t *xv; int n_xv;         /* tags of sam_add_md_nm; md == NULL: not printed */
} salt_sam_read_t;

/* aln_samse: the single-end line (no trailing newline; the reference prints it with puts).  print_xa_cigar: option -c.
 * Returns the line's length; SALT_ERR_NOMEM when it does not fit cap (nothing useful is left in out), SALT_ERR_ARG when a
 * position lies outside the reference (the reference exits there). */
int salt_sam_se(const salt_sam_refs_t *refs, const salt_sam_read_t *q, int print_xa_cigar, const char *rg_id, char *out, size_t cap);

/* alnpe_sam: the two lines of a pair, each with its trailing newline as the reference leaves them in query->sam.
 * len[0] / len[1] receive the lengths. */
int salt_sam_pe(const salt_sam_refs_t *refs, const salt_sam_read_t q[2], uint32_t min_tlen, uint32_t max_tlen, int print_xa_cigar,
                const char *rg_id, char *out0, size_t cap0, char *out1, size_t cap1, int len[2]);

#ifdef __cplusplus
}
#endif
#endif /* SALT_HOST_H */
