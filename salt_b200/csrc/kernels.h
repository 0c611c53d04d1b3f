// kernels.h -- launchers implemented in verify.cu / ssw.cu / samtail.cu / mixref.cu, used by engine.cu.
#pragma once
#if !defined(SALT_EMUL)
#include <cuda_runtime.h>
#endif

#include "common.cuh"

namespace salt {

cudaError_t launch_pack_reads(const uint8_t *codes, const uint32_t *offs, uint32_t n_reads, uint32_t W64,
                              uint64_t *rd4, uint16_t *rd_len, cudaStream_t st);
cudaError_t launch_mismatch(const DevCtx &c, const salt_pair_t *pairs, size_t n, int max_err, int8_t *out, cudaStream_t st);
struct LvFilterScratch {      // survivors of the pigeonhole filter: room for every input pair
    salt_pair_t *pairs; uint32_t *slots; uint32_t *count;
    salt_pair_t *pairs2 = nullptr; uint32_t *slots2 = nullptr; uint32_t *count2 = nullptr;   // second Landau-Vishkin pass (optional)
};
struct LvDefer { salt_pair_t *pairs; uint32_t *slots; uint32_t *count; };    // where a first pass queues the pairs it does not finish
cudaError_t launch_lv(const DevCtx &c, const salt_pair_t *pairs, size_t n, int k,
                      const uint32_t *worklist, const uint32_t *wl_count, size_t wl_cap,
                      int8_t *out, int sm_count, cudaStream_t st, int mapping = 0, const LvFilterScratch *f = nullptr);
cudaError_t launch_lv_cigar(const DevCtx &c, const salt_pair_t *pairs, const uint8_t *k_each, size_t n,
                            const uint32_t *worklist, const uint32_t *wl_count, size_t wl_cap,
                            const salt_verify_out_t *rec, int kmax, char *cigars, int stride, int8_t *out,
                            int sm_count, cudaStream_t st, int mapping);
cudaError_t launch_nogap_fused(const DevCtx &c, const uint32_t *offs0, const uint32_t *loci0,
                               const uint32_t *offs1, const uint32_t *loci1, size_t n0, int T0,
                               int8_t *acc, salt_verify_out_t *rec, salt_pair_t *lv_pairs, uint32_t *lv_slots,
                               uint32_t *lv_count /* [0] pairs, [1] reads */, uint32_t *lv_reads, cudaStream_t st);
cudaError_t launch_scan_gap(const DevCtx &c, const uint32_t *offs0, const uint32_t *loci0,
                            const uint32_t *offs1, const uint32_t *loci1, size_t n0, int lv_T0,
                            int8_t *acc, salt_verify_out_t *rec, const uint32_t *lv_reads, const uint32_t *lv_read_count,
                            uint32_t *cig_list, uint32_t *cig_count, int sm_count, cudaStream_t st);

// samtail.cu
cudaError_t launch_md_nm(const DevCtx &c, const uint8_t *codes, const uint32_t *roffs, const salt_mdnm_in_t *items, size_t n,
                         const char *cigars, int cstride, char *md, int mstride, uint16_t *xv, int xstride,
                         salt_mdnm_out_t *out, cudaStream_t st);

cudaError_t launch_tail_primaries(const DevCtx &c, const uint8_t *codes, const uint32_t *roffs, const salt_verify_out_t *rec,
                                  const uint32_t *cig_reads, const uint32_t *cig_count, const char *cigs, int cstride,
                                  int32_t *cig_row, char *md, int mstride, uint16_t *xv, int xstride, salt_mdnm_out_t *out,
                                  uint32_t *md_bytes, cudaStream_t st);
cudaError_t launch_tail_pack(const char *md, int mstride, const uint32_t *offs, uint32_t n, char *packed, cudaStream_t st);

// mixref.cu
cudaError_t launch_build_mixref(const char *bases, uint32_t l, const uint32_t *snp_pos, const uint8_t *snp_mask,
                                size_t n_snp, uint32_t *words, cudaStream_t st);

// transport.cu: the compact chunk format (salt_packed_chunk_t)
struct Scan3 {                // exclusive prefix sums: out[k][0..n] from n counts each
    const void *in[3];        // null = every count is uniform[k]
    int width[3];             // 16 or 32 bits per count
    uint32_t uniform[3];
    uint32_t *out[3];
    size_t n;
    uint32_t *partial;        // 3 * scan3_blocks(n) words of scratch
    uint32_t n_blocks;        // filled by the launcher
};
cudaError_t launch_cig_slim(const char *rows, int stride, const uint32_t *count, uint32_t eager, char *slim, cudaStream_t st);
uint32_t scan3_blocks(size_t n);
cudaError_t launch_scan3(Scan3 a, int n_arrays, cudaStream_t st);
cudaError_t launch_unpack_bases(const uint8_t *in, uint32_t phase, int bits, size_t n_bases, uint8_t *codes,
                                const uint32_t *n_pos, size_t n_n, uint32_t origin, cudaStream_t st);

// seed.cu: single-end seeding + locate on the reference's FM-indexes (row f1)
struct FmIndexDev {
    // primary reference ("C part"): BWA layout, bwt.h:44-57
    const uint32_t *cbwt; uint32_t c_primary, c_seq_len; uint32_t c_L2[5];
    const uint32_t *c_sa; uint32_t c_sa_intv, c_n_sa;
    const uint32_t *lkt; int l_lkt;                       // lookup.h:21-25
    const uint4 *c32;                                     // device layout: per 32 bases {4 counts}{2 words} (seed.cu)
    const uint4 *r64;                                     // device layout: per 64 characters {5 counts}{8 words}
    // SNP-context index, backward direction ("R part"): rbwt.h:60-80
    const uint32_t *r_bwt, *r_occ, *r_occ_major, *r_sa_sharp;
    uint32_t r_cum[6], r_inv_sa0, r_text_len, r_n_sa_sharp;
};
struct SeedOpt { int l_seed, l_overlap, max_seed, max_locate, seed_only_ref, mode /* 0 alnse_locate_alt, 1 alnse_locate */, list_cap /* list stride */; };
struct SeedSai { uint32_t sp, ep, offset; };              // sai_t, aln.h:91-95
size_t seed_sai_bytes(uint32_t n_reads, int max_seeds);
size_t fm_c32_entries(size_t c_bwt_words);
size_t fm_r64_entries(uint32_t r_text_len);
cudaError_t launch_build_dense_index(const FmIndexDev &ix, size_t c_bwt_words, size_t r_bwt_words_padded, uint4 *c32, uint4 *r64, cudaStream_t st);
cudaError_t launch_seed(const FmIndexDev &ix, const SeedOpt &opt, const uint8_t *codes, const uint32_t *roffs, uint32_t n_reads,
                        int max_seeds, SeedSai *sai, cudaStream_t st);
cudaError_t launch_locate(const FmIndexDev &ix, const SeedOpt &opt, const uint32_t *roffs, uint32_t n_reads, int max_seeds,
                          uint32_t ref_l, SeedSai *sai, uint32_t *counts, uint32_t *lists, uint32_t *long_list /* 2*n_reads */,
                          uint32_t *long_count, uint8_t *status /* [2][n_reads] or null */, int sm_count, cudaStream_t st);
cudaError_t launch_seed_gather(const uint32_t *lists, int max_locate, const uint32_t *offs0, const uint32_t *offs1,
                               uint32_t n_reads, uint32_t *loci0, uint32_t *loci1, cudaStream_t st);

// ssw.cu
struct SswParams {
    int use_pac, n_sym, gapO, gapE, flag, filters, filterd, mask_len;
    int8_t table[17 * 8];     // score[ref_sym][read code 0..4, 5 = pad row], row 16 = pad column
};
struct SswScratch {           // device buffers owned by the engine, sized for >= n tasks
    uint32_t *win4;           // per pair-of-tasks packed window symbols, interleaved (w0,w1) per 8 columns
    uint32_t *rsel;           // per task read-code selectors
    uint16_t *maxcol;         // per task, per column
    int32_t *fwd;             // per task forward results
    int8_t *dirs;             // banded traceback direction bytes
    size_t dirs_bytes;
    size_t cap_tasks; int cap_cols; int cap_rows;
};
size_t ssw_scratch_bytes(size_t n_tasks, int max_cols, int max_rows, size_t *layout /*[8]*/);
cudaError_t launch_ssw(const DevCtx &c, const salt_win_t *wins, size_t n, const SswParams &prm,
                       void *scratch, size_t scratch_bytes, int max_cols,
                       salt_ssw_out_t *out, uint32_t *cigars, int cigar_stride, int sm_count, cudaStream_t st,
                       uint64_t *launches, cudaEvent_t *ev /* null or [7]: boundaries of the 6 stages */,
                       uint8_t *ovf_dirs /* ssw_overflow_bytes(l_max) of scratch for bands wider than 16, or null */);
constexpr unsigned SSW_OVF_THREADS = 128;
size_t ssw_overflow_bytes(int max_rows);

}  // namespace salt
