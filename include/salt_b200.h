/*
 * salt_b200.h -- C ABI of libsalt_b200.so, the B200 (sm_100a) engine for salt's
 * SNP-aware verification / extension hot path.
 *
 * Plain C, opaque handle, int return codes, no exit() inside the library.  Every entry
 * point names the reference interface it stands in for (paths relative to the reference
 * tree's Align_src/).  There is no CPU fallback: without a CUDA device every compute
 * entry point fails with SALT_ERR_NODEVICE.
 *
 * Data conventions (all from the reference):
 *   mixRef : uint32 words, 4 bit/base allele mask A=1 C=2 G=4 T=8, base p in bits
 *            4*(p%8).. of word p/8                       (metaref.h:2-5, metaref.c:54-56)
 *   pac    : 2 bit/base, base p in bits ((~p&3)<<1) of byte p>>2       (alnpe.c:47)
 *   reads  : codes A,C,G,T,N = 0..4, one byte per base                 (query.c:177-181)
 *   strand : 0 = read as given, 1 = reverse complement                 (query.c:46-64)
 */
#ifndef SALT_B200_H
#define SALT_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SALT_B200_ABI_VERSION 1

enum {
    SALT_OK = 0,
    SALT_ERR_ARG = -101,          /* bad argument */
    SALT_ERR_CUDA = -102,         /* a CUDA call failed; see salt_b200_last_error() */
    SALT_ERR_NOMEM = -103,
    SALT_ERR_UNSUPPORTED = -104,  /* parameters outside what the kernels reproduce exactly */
    SALT_ERR_NODEVICE = -105      /* no usable CUDA device: there is no CPU path */
};

typedef struct salt_b200 salt_b200_t;

/* (read, candidate locus) pair.  rs = (read_id << 1) | strand. */
typedef struct { uint32_t rs; uint32_t pos; } salt_pair_t;

/* One chunk of reads (alnse.c:1414 reads N_SEQS=100000 at a time; any size works). */
typedef struct {
    const uint8_t *codes;    /* concatenated codes 0..4 */
    const uint32_t *offs;    /* n_reads+1 offsets into codes */
    uint32_t n_reads;
} salt_reads_t;

/* Sorted candidate loci per read and strand, CSR (what alnse_locate[_alt] leaves in
 * aux->loci, alnse.c:501-731). */
typedef struct {
    const uint32_t *offs[2];  /* n_reads+1 each */
    const uint32_t *loci[2];
} salt_cands_t;

/* Mate-rescue window: read rs against reference bases [start, end] inclusive
 * (alnpe.c:213-252 compute start/end; alnpe.c:261 / :330 consume them). */
typedef struct { uint32_t rs; uint32_t start; uint32_t end; } salt_win_t;

/* s_align (ssw.h:37-47) without the heap pointer. */
typedef struct {
    uint16_t score1, score2;
    int32_t ref_begin1, ref_end1, read_begin1, read_end1, ref_end2;
    int32_t cigarLen;
} salt_ssw_out_t;

/* One alignment whose SAM tail is wanted: the read as aligned (strand bit of rs), its leftmost
 * reference base and the read offset the alignment starts at (query->seq_start: 0 in SE, the
 * soft-clip length after a mate rescue, alnpe.c:301).  pos = 0xFFFFFFFF: unmapped, no tags. */
typedef struct { uint32_t rs; uint32_t pos; uint32_t seq_start; } salt_mdnm_in_t;

/* nm: the NM value.  md_len: characters of the MD value (excluding the NUL); -2 = MD buffer too
 * small (string truncated), -3 = an M run reaches past the end of the reference (the reference
 * asserts there, sam.c:270).  n_xv: entries of the XV list (at most 64, sam.c:242). */
typedef struct { int32_t nm; int16_t md_len; uint16_t n_xv; } salt_mdnm_out_t;

/* Per-read outcome of the verification stage: the query_t fields that
 * alnse_check_nogap / alnse_check_withgap set (alnse.c:348-393), plus hit counts. */
typedef struct {
    uint32_t pos;        /* 0xFFFFFFFF = unmapped (query.c:202) */
    uint8_t strand;      /* 3 = unset (query.c:204) */
    uint8_t n_diff;      /* 255 = unset */
    uint8_t is_gap;      /* 255 = unset */
    uint8_t lv_ran;      /* 1 if the gapped stage ran for this read */
    int32_t n_hits[2];   /* accepted candidates per strand (aux->hits) */
} salt_verify_out_t;

const char *salt_b200_last_error(void);
int salt_b200_abi_version(void);
int salt_b200_device_count(void);

/* Upload the SNP-aware reference (and optionally the 2-bit pac) once.
 * Stands in for the in-memory mixRef/pac that alnse_index_reload builds (indexio.c:23-50). */
salt_b200_t *salt_b200_init(const uint32_t *mixref, uint32_t l, const uint8_t *pac, int64_t l_pac, int device);
void salt_b200_destroy(salt_b200_t *h);

/* A second handle on the same device that shares the first one's resident reference (and FM-indexes) but has its own
 * streams, slots and scratch: what each of the reference's -t worker threads takes (a handle is single-threaded).  The
 * parent must outlive it. */
salt_b200_t *salt_b200_attach(salt_b200_t *parent);

/* Replace the resident reference of a handle that owns one (re-using its allocation when the new one fits). */
int salt_b200_reload_ref(salt_b200_t *h, const uint32_t *mixref, uint32_t l, const uint8_t *pac, int64_t l_pac);

/* Build the SNP-aware reference on the device from raw bases and SNP rows instead of
 * uploading it: Index_src/mixRef.c:93-197 (build_mixRef) + Index_src/hapmap.c:92-158.
 * bases: ASCII, all records concatenated; snp_pos: 0-based global positions; snp_mask: allele
 * masks (low 4 bits used).  Returns a handle owning the device copy; salt_b200_get_mixref
 * downloads the words (file layout of PREFIX.ref minus the length word). */
salt_b200_t *salt_b200_init_from_bases(const char *bases, uint32_t l, const uint32_t *snp_pos,
                                       const uint8_t *snp_mask, size_t n_snp, int device);
int salt_b200_get_mixref(salt_b200_t *h, uint32_t *words_out, size_t n_words);

/* Use an existing CUDA stream (cudaStream_t) for the synchronous entry points and the *_dev
 * variants of this handle (slot 0 of the chunk pipeline); NULL = own stream. */
int salt_b200_set_stream(salt_b200_t *h, void *cuda_stream);
int salt_b200_sync(salt_b200_t *h);

/* Pinned host memory for the caller's queues. */
void *salt_b200_host_alloc(size_t bytes);
void salt_b200_host_free(void *p);

/* Upload + pack one chunk of reads (both strands) into HBM.  Pair/window `rs` fields
 * index into the current chunk. */
int salt_b200_set_reads(salt_b200_t *h, const salt_reads_t *reads);

/* Batched ed_mismatch (editdistance.c:88): out[i] = n if n <= max_err else -1.
 * Pairs with pos + l_seq > l give -1 (the reference leaves that to its callers). */
int salt_b200_mismatch(salt_b200_t *h, const salt_pair_t *pairs, size_t n, int max_err, int8_t *out);

/* Batched ed_diff -> computeEditDistance (editdistance.c:174, LandauVishkin.c:19) with
 * l_ref = l_seq + 4 (alnse.c:373).  k < 0 means l_seq/10 per read (alnse.c:1090).
 * out[i] = e <= min(k,30), or -1. */
int salt_b200_lv(salt_b200_t *h, const salt_pair_t *pairs, size_t n, int k, int8_t *out);

/* Batched ed_diff_withcigar -> computeEditDistanceWithCigar, useM=1, COMPACT_CIGAR_STRING
 * (editdistance.c:234, LandauVishkin.c:176; call sites query.c:288, sam.c:218).
 * k_each[i] must be < 31.  cigars: n buffers of `stride` bytes, written like the reference
 * writes its caller's buffer (NUL-terminated string; bytes after it untouched).
 * out[i] = e, -1 (not within k) or -2 (buffer too small). */
int salt_b200_lv_cigar(salt_b200_t *h, const salt_pair_t *pairs, const uint8_t *k_each, size_t n,
                       char *cigars, int stride, int8_t *out);

/* Batched ssw_init(score_size=1) + ssw_align (ssw.c:742, :771) on reference windows.
 * use_pac = 0: window symbols are mixRef masks, read symbol = 1<<code (N -> 16), the
 *              snpaln_sw_snpaware call (alnpe.c:261-293);
 * use_pac = 1: window symbols are 2-bit pac codes, read symbol = code, the snpaln_sw call
 *              (alnpe.c:330-351).
 * mat has n_sym*n_sym entries and is indexed mat[ref_sym*n_sym + read_sym] exactly as
 * ssw.c:361 does; an index of n_sym*n_sym (read N against mask 15, which the reference
 * reads one past its array) is scored as mat[n_sym*n_sym-1].
 * mask_len < 0 means l_seq/2 (alnpe.c:287).  Requires gapO > gapE.
 * cigars: n * cigar_stride uint32 (len<<4|op, op 0/1/2 = M/I/D); out[i].cigarLen is the
 * true length even if it exceeds cigar_stride (the row then holds only the last cigar_stride ops).
 * A window the engine declines does not fail the batch: out[i].cigarLen = -1 (read id outside the
 * chunk, start > end, end > l -- end == l is served, position l scoring as symbol 0 -- or wider than
 * the widest supported window), -2 (traceback band beyond 512, ssw.c:570-631 keeps doubling). */
int salt_b200_ssw(salt_b200_t *h, const salt_win_t *wins, size_t n, int use_pac,
                  const int8_t *mat, int n_sym, int gapO, int gapE, int flag,
                  int filters, int filterd, int mask_len,
                  salt_ssw_out_t *out, uint32_t *cigars, int cigar_stride);

/* MD / NM / XV of n alignments of the reads resident in `slot` (0 = the read set of the synchronous
 * entry points; a pipeline slot keeps its chunk's reads until it is submitted again) --
 * sam_add_md_nm (sam.c:246-328), the tags behind `-d`.  cigars: n NUL-terminated M/I/D strings (query->cigar->s, no soft clips),
 * cigar_stride bytes apart.  md: n strings of md_stride bytes ("MD:Z:" value only).  xv: n rows
 * of xv_stride uint16 read offsets (relative to seq_start); a row holds the first
 * min(n_xv, xv_stride) entries.  Needs the 2-bit pac given to salt_b200_init. */
int salt_b200_md_nm(salt_b200_t *h, int slot, const salt_mdnm_in_t *items, size_t n, const char *cigars, int cigar_stride,
                    char *md, int md_stride, uint16_t *xv, int xv_stride, salt_mdnm_out_t *out);

/* The same tags for the PRIMARIES of the chunk just verified in `slot` (salt_b200_verify / _verify_wait, with CIGARs):
 * positions, strands and CIGARs are already on the device, so nothing crosses the host link on the way in, and the MD
 * strings come back packed -- string i (NUL-terminated, "" for an unmapped read) starts at md_packed[md_offs[i]];
 * *md_bytes = md_offs[n_reads] bytes in all (SALT_ERR_NOMEM when md_cap is smaller).  out / xv as in salt_b200_md_nm,
 * one row per read of the chunk.  This is sam_add_md_nm for a whole chunk of aln_samse records (sam.c:87, :246-328). */
int salt_b200_tail_primaries(salt_b200_t *h, int slot, salt_mdnm_out_t *out, uint32_t *md_offs, char *md_packed, size_t md_cap,
                             size_t *md_bytes, uint16_t *xv, int xv_stride);

/* The same, asynchronous: _tail_submit queues the work on the slot's stream -- also right behind a
 * salt_b200_verify_submit[_packed] of the same slot, without waiting for it -- and _tail_wait completes both.  The buffers
 * must stay valid until _tail_wait returns. */
int salt_b200_tail_submit(salt_b200_t *h, int slot, salt_mdnm_out_t *out, uint32_t *md_offs, char *md_packed, size_t md_cap,
                          uint16_t *xv, int xv_stride);
int salt_b200_tail_wait(salt_b200_t *h, int slot, size_t *md_bytes);

/* The whole verification stage for a chunk, as alnse_overlap_alt (SE) / alnse_overlap (PE)
 * run it after seeding (alnse.c:1077-1097 / :1014-1036):
 *   nogap on strand 0 then 1 with threshold nogap_T0 (3) tightening as candidates are
 *   accepted; if neither strand matched, the gapped stage with lv_T0 (< 0: l_seq/10, the SE
 *   rule; 3 is the PE rule); then CIGARs for gapped primaries (query.c:282).
 * acc[s][i] (one per candidate, same order as cands->loci[s]) = accepted n_diff or -1;
 * together with rec[] that is exactly aux->hits + the primary.  cigars: n_reads buffers of
 * `cigar_stride` bytes (query->cigar, 128 in the reference), written only for gapped
 * primaries. acc / cigars may be NULL. */
int salt_b200_verify(salt_b200_t *h, const salt_cands_t *cands, int nogap_T0, int lv_T0,
                     salt_verify_out_t *rec, int8_t *acc0, int8_t *acc1,
                     char *cigars, int cigar_stride);

/* ---- asynchronous chunk pipeline ------------------------------------------------------------
 * alnse_core / alnpe_core work through the input in chunks of N_SEQS = 100000 reads
 * (aln.h:27, alnse.c:1414-1440): read a chunk, run the workers, print.  The engine keeps
 * salt_b200_n_slots() chunks in flight, each on its own CUDA stream with its own staging in HBM,
 * so the upload of chunk c+1, the kernels of chunk c and the download of chunk c-1 overlap (and the
 * host can seed chunk c+1 meanwhile).  Buffers passed to _submit (inputs and outputs; pinned memory
 * from salt_b200_host_alloc makes the copies truly asynchronous) must stay valid and untouched
 * until _wait(slot) returns.  Results are exactly those of salt_b200_set_reads + salt_b200_verify
 * on the same chunk; read ids inside a chunk are chunk-relative. */
int salt_b200_n_slots(void);
int salt_b200_verify_submit(salt_b200_t *h, int slot, const salt_reads_t *reads, const salt_cands_t *cands,
                            int nogap_T0, int lv_T0, salt_verify_out_t *rec, int8_t *acc0, int8_t *acc1,
                            char *cigars, int cigar_stride);
int salt_b200_verify_wait(salt_b200_t *h, int slot);
/* The loop above over a whole batch: reads/cands describe n_reads reads (global CSR offsets);
 * chunks of chunk_reads reads (0 = 100000) go through the slots round-robin; outputs are indexed
 * like the inputs.  Returns when every chunk is done. */
int salt_b200_verify_batch(salt_b200_t *h, const salt_reads_t *reads, const salt_cands_t *cands, uint32_t chunk_reads,
                           int nogap_T0, int lv_T0, salt_verify_out_t *rec, int8_t *acc0, int8_t *acc1,
                           char *cigars, int cigar_stride);

/* ---- compact transport ------------------------------------------------------------------------
 * The verification stage is bound by the host link, not by the kernels, so a caller
 * that controls its own queues can send a chunk in fewer bytes: bases at 2 or 4 bits instead of the
 * reference's one byte per base (query->seq, query.c:177-181 -- the FASTQ parser can emit either for
 * free), read lengths and per-read candidate counts instead of three 32-bit offset arrays.  The engine
 * rebuilds codes and offsets on the device; results are identical to salt_b200_verify_submit on the
 * same chunk.
 *   base_bits 2: codes 0..3, base p of the stream in bits 2*(p&3) of byte p>>2; an N is sent as any
 *                code and its stream position listed in n_pos (ascending).
 *   base_bits 4: codes 0..4, base p in bits 4*(p&1) of byte p>>1.
 * Reads lie back to back in the stream; read 0's first base is stream position base_start (so a view
 * into a longer stream needs no re-packing).  `bases` must be readable up to the byte holding the last
 * base.  lens == NULL: every read has l_seq bases.  n_cand[s][i] = candidates of read i on strand s
 * (count_bits 16 or 32); loci[s] = the lists themselves, reads back to back, as in salt_cands_t. */
typedef struct {
    uint32_t n_reads;
    int base_bits;
    const uint8_t *bases;
    uint32_t base_start;
    const uint16_t *lens;
    uint32_t l_seq;
    const uint32_t *n_pos;
    size_t n_n;
    int count_bits;
    const void *n_cand[2];
    const uint32_t *loci[2];
} salt_packed_chunk_t;

/* salt_b200_set_reads / _verify_submit / _verify_batch for the compact format (n_cand / loci are not
 * read by _set_reads_packed). */
int salt_b200_set_reads_packed(salt_b200_t *h, const salt_packed_chunk_t *pc);
int salt_b200_verify_submit_packed(salt_b200_t *h, int slot, const salt_packed_chunk_t *pc, int nogap_T0, int lv_T0,
                                   salt_verify_out_t *rec, int8_t *acc0, int8_t *acc1, char *cigars, int cigar_stride);
int salt_b200_verify_batch_packed(salt_b200_t *h, const salt_packed_chunk_t *pc, uint32_t chunk_reads, int nogap_T0, int lv_T0,
                                  salt_verify_out_t *rec, int8_t *acc0, int8_t *acc1, char *cigars, int cigar_stride);

/* ---- seeding + locate on the device ------------------------------------
 * The reference spends > 90 % of its time producing the candidate lists (alnse_seed_overlap, alnse.c:199-312;
 * alnse_locate_alt, alnse.c:633-731).  With its FM-indexes resident in HBM the engine produces the identical
 * sorted lists itself, so a chunk needs only its reads uploaded and the lists never cross the host link.
 *
 * salt_fm_index_t carries the arrays exactly as the reference's loaders leave them in memory
 * (indexio.c:23-50): the BWA-layout BWT and sampled SA of the primary reference (bwt.h:44-57: PREFIX.C.bwt,
 * PREFIX.C.sa; sa[0] = (uint32_t)-1, bwtio.c:46), the 12-mer lookup table (lookup.h:21-25: PREFIX.C.lkt) and the
 * backward SNP-context index (rbwt.h:60-80: PREFIX.R.backward.bwt / .occ / .sa).  cum[0] = 0 as Rbwt_restore_bwt
 * sets it (rbwt.c:262). */
typedef struct {
    const uint32_t *c_bwt; size_t c_bwt_words;           /* bwt_t::bwt, bwt_size */
    uint32_t c_primary, c_seq_len, c_L2[5];
    const uint32_t *c_sa; uint32_t c_n_sa, c_sa_intv;
    const uint32_t *lkt; uint32_t lkt_len;                /* lookupTable_t::item (4^lkt_len + 1 entries), maxLookupLen */
    const uint32_t *r_bwt; size_t r_bwt_words;            /* rbwt_t::bwtCode, bwtSizeInWord */
    const uint32_t *r_occ; size_t r_occ_words;            /* occValue, occSizeInWord */
    const uint32_t *r_occ_major; size_t r_occ_major_words;
    const uint32_t *r_sa_sharp; size_t r_n_sa_sharp;      /* saValueSharp, saValueSizeSharp */
    uint32_t r_cum[6], r_inv_sa0, r_text_len;             /* cumulativeFreq, inverseSa0, textLength */
} salt_fm_index_t;

/* aln_opt_t fields the seeding reads (aln.h:121-151): l_seed (persisted by the indexer in PREFIX.R.seedLen),
 * l_overlap (defaults to l_seed), max_seed (-n), max_locate (-m), seed_only_ref (-R). */
typedef struct {
    int l_seed, l_overlap, max_seed, max_locate, seed_only_ref;
    int locate_mode;     /* 0: alnse_locate_alt, the single-end program (alnse.c:633-731); 1: alnse_locate, the paired-end
                            program (alnse.c:501-631): up to max_locate + 1 rows of every primary-index interval, MAX_LOC_POS
                            loci in all, and SNP-context intervals wider than max_locate subsampled with rand() */
    int list_cap;        /* locate_mode 1: room per list on the device, 64..16384 (the reference allows 262144) */
} salt_seed_opt_t;

/* Upload the indexes once (the handle keeps its own copy). */
int salt_b200_set_index(salt_b200_t *h, const salt_fm_index_t *ix);

/* alnse_seed_overlap + alnse_locate_alt for both strands of every read resident in `slot` (uploaded by
 * salt_b200_set_reads[_packed]; slot 0 for the synchronous entry points).  The sorted lists stay on the device as the
 * slot's candidate lists -- salt_b200_verify_seeded consumes them in place -- and are optionally downloaded:
 * offs0 / offs1 (n_reads + 1 each) and loci0 / loci1 (cap0 / cap1 entries of room; SALT_ERR_NOMEM if a strand has
 * more) may be NULL.  *n0 / *n1 receive the totals.  Limits: at most 1024 seed starts per strand
 * ((l_seq - l_seed) / l_overlap + 1), max_locate <= 16384, l_seed >= the lookup length. */
int salt_b200_seed_locate(salt_b200_t *h, int slot, const salt_seed_opt_t *opt, uint32_t *offs0, uint32_t *offs1,
                          uint32_t *loci0, size_t cap0, uint32_t *loci1, size_t cap1, size_t *n0, size_t *n1);

/* locate_mode 1 only: per read and strand, bit 0 = an SNP-context interval was wider than max_locate -- the reference
 * picks a random subset of its rows there (srand(time(0)), alnse.c:538-552), so no list is "the" right one: the interval
 * is left out and the caller decides (the reference's own functions on the host, or accept the shorter list); bit 1 = the
 * list filled list_cap before the reference's own limit.  n_reads bytes per strand. */
int salt_b200_seed_status(salt_b200_t *h, int slot, uint8_t *st0, uint8_t *st1);

/* salt_b200_verify on the lists salt_b200_seed_locate left in the slot.  acc0 / acc1 need the totals it reported. */
int salt_b200_verify_seeded(salt_b200_t *h, int slot, int nogap_T0, int lv_T0, salt_verify_out_t *rec,
                            int8_t *acc0, int8_t *acc1, char *cigars, int cigar_stride);

/* The whole single-end stage of alnse_overlap_alt (alnse.c:1045-1100) for a batch in the compact transport format
 * (n_cand / loci of `pc` are not read): per chunk of chunk_reads reads, upload the reads, seed, locate, verify, and
 * download the per-read records (and CIGARs).  Chunks go through the pipeline slots like salt_b200_verify_batch. */
int salt_b200_align_batch_packed(salt_b200_t *h, const salt_packed_chunk_t *pc, const salt_seed_opt_t *opt, uint32_t chunk_reads,
                                 int nogap_T0, int lv_T0, salt_verify_out_t *rec, char *cigars, int cigar_stride);

/* The per-pair entry points (salt_b200_mismatch / _lv / _lv_cigar / _ssw and their _dev variants) work on the reads of
 * slot 0 unless told otherwise: after salt_b200_verify_wait(slot) a chunk's reads stay resident in that slot until it
 * is submitted again, so the paired-end stage can run its mate rescues on them (alnpe.c:206-252) without a re-upload. */
int salt_b200_use_slot(salt_b200_t *h, int slot);

/* Landau-Vishkin work mapping: 0 = automatic (one thread per pair with all diagonals in
 * registers for k <= 15 inside the verify stage, one warp per pair with lanes over diagonals
 * beyond that and on flat pair lists), 1 = always one warp (or sub-warp group) per pair,
 * 2 = one thread per pair whenever k <= 15.  Adding 16 runs the thread-per-pair kernel in one pass instead of two (three
 * levels for every survivor of the pre-filter, the full depth only for the pairs that need it).  Results are identical;
 * this exists for measurement. */
int salt_b200_set_lv_mapping(salt_b200_t *h, int mapping);

/* Pigeonhole pre-filter in front of Landau-Vishkin (default on): a pair can only be within k
 * differences if at least plen/8 - k of the read's full 8-base words match the window exactly on
 * some diagonal |d| <= k; pairs failing that get -1 without running Landau-Vishkin.  Results are
 * identical with the filter off; the switch exists for measurement and tests. */
int salt_b200_set_lv_filter(salt_b200_t *h, int enable);

/* Widest rescue window (in bases) the *_dev SSW entry point must handle; the host entry point
 * sets it from its arguments.  Default 1024. */
int salt_b200_set_max_window(salt_b200_t *h, int cols);

/* ---- device-resident variants (inputs already in HBM; asynchronous on the handle's
 * stream; used for kernel-level measurement and by callers that keep queues on the GPU) */
int salt_b200_mismatch_dev(salt_b200_t *h, const salt_pair_t *d_pairs, size_t n, int max_err, int8_t *d_out);
int salt_b200_lv_dev(salt_b200_t *h, const salt_pair_t *d_pairs, size_t n, int k, int8_t *d_out);
/* Device variant of salt_b200_verify.  CIGARs come back compact: d_cigars[i*stride] belongs to
 * read d_cig_reads[i], i < *d_cig_count (only gapped primaries have one).  Pass d_cigars = NULL
 * to skip them.  d_rec must be 16-byte aligned. */
int salt_b200_verify_dev(salt_b200_t *h, const uint32_t *d_offs0, const uint32_t *d_loci0, size_t n0,
                         const uint32_t *d_offs1, const uint32_t *d_loci1, size_t n1,
                         int nogap_T0, int lv_T0, salt_verify_out_t *d_rec, int8_t *d_acc0, int8_t *d_acc1,
                         char *d_cigars, int cigar_stride, uint32_t *d_cig_reads, uint32_t *d_cig_count);
int salt_b200_ssw_dev(salt_b200_t *h, const salt_win_t *d_wins, size_t n, int use_pac,
                      const int8_t *mat, int n_sym, int gapO, int gapE, int flag,
                      int filters, int filterd, int mask_len,
                      salt_ssw_out_t *d_out, uint32_t *d_cigars, int cigar_stride);

/* Per-stage device timing (CUDA events on the handle's stream around each kernel).
 * After a verify / ssw call with profiling enabled, salt_b200_profile_read synchronises and fills
 * ms[0..5]  = (unused), nogap_fused, (unused), lv (pre-filter + Landau-Vishkin), scan_gap, lv_cigar   (verify stage)
 * ms[6..11] = prep_fwd, dp_fwd, prep_rev, dp_rev, banded, banded_overflow (ssw)
 * with -1 for stages that did not run. */
int salt_b200_profile(salt_b200_t *h, int enable);
int salt_b200_profile_read(salt_b200_t *h, float *ms);

/* Counters for benchmarking: kernels launched by this handle since the last reset. */
uint64_t salt_b200_launch_count(salt_b200_t *h, int reset);

#ifdef __cplusplus
}
#endif
#endif /* SALT_B200_H */
