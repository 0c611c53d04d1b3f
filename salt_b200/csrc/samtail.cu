// samtail.cu -- the SAM tail of a mapped read on the device: MD string, NM and XV.
// Restates sam_add_md_nm (sam.c:246-328).  One thread per alignment walks the M/I/D runs of the
// read's CIGAR (query->cigar->s; soft clips are printed elsewhere, sam.c:392-394) over the 2-bit
// reference (pac, MSB-first, sam.c:244) and the read as aligned (query->seq or its reverse
// complement, query.c:46-64), and the 4-bit mixRef for the XV test.
//   MD   match runs as decimal counts, a mismatch prints the reference base, a deletion prints
//        '^' and its bases after flushing the pending run.  The reference prints no "0" between
//        adjacent mismatches and none after a deletion; kept.
//   NM   mismatches + inserted + deleted bases.
//   XV   read offsets (relative to seq_start) of mismatching bases whose one-hot code is in the
//        mixRef allele mask, i.e. a known non-reference allele; at most 64 (MAX_RS, sam.c:242).
// Byte-serial string work, ~10 instructions per base: bound by nothing worth naming at chunk sizes
// (2 M reads ~ 0.1 ms); it is here so a chunk's SAM fields can leave the device in one pass.
#if !defined(SALT_EMUL)
#include <cuda_runtime.h>
#endif

#include "common.cuh"
#include "kernels.h"

namespace salt {

struct MdOut {
    char *buf; int cap; int len; bool ovf;
    __device__ void put(char ch)
    {
        if (len < cap - 1) buf[len] = ch; else ovf = true;
        ++len;
    }
    __device__ void put_num(int v)
    {
        char tmp[12]; int nd = 0;
        do { tmp[nd++] = (char)('0' + v % 10); v /= 10; } while (v > 0);
        while (nd > 0) put(tmp[--nd]);
    }
};

// One alignment.  cg / cstride: its M/I/D string; cg == nullptr: the ungapped "<L>M" of query_gen_cigar (query.c:291).
__device__ __forceinline__ salt_mdnm_out_t md_nm_one(const DevCtx &c, const uint8_t *__restrict__ codes, const uint32_t *__restrict__ roffs,
                                                     const salt_mdnm_in_t in, const char *__restrict__ cg, int cstride,
                                                     char *__restrict__ mdbuf, int mstride, uint16_t *__restrict__ xvrow, int xstride)
{
    salt_mdnm_out_t o; o.nm = 0; o.md_len = 0; o.n_xv = 0;
    MdOut m{mdbuf, mstride, 0, false};
    const uint32_t rid = in.rs >> 1;
    const bool rev = (in.rs & 1u) != 0;
    if (in.pos == 0xFFFFFFFFu || rid >= c.n_reads) {                    // sam.c:248: unmapped reads get no tags
        if (mstride > 0) m.buf[0] = '\0';
        return o;
    }
    const int L = (int)c.rd_len[rid];
    const uint8_t *__restrict__ rd = codes + roffs[rid];
    uint32_t ref_pos = in.pos;
    int si = (int)in.seq_start;
    int nm = 0, n_match = 0, n_rs = 0;
    bool past_end = false;
    int ci = 0;
    bool synth_done = false;
    while (!past_end) {
        int run = 0;
        char op;
        if (cg) {
            if (!(ci < cstride && cg[ci] != '\0')) break;
            while (ci < cstride && cg[ci] >= '0' && cg[ci] <= '9') { run = run * 10 + (cg[ci] - '0'); ++ci; }
            op = ci < cstride ? cg[ci] : '\0';
        } else {
            if (synth_done) break;
            run = L; op = 'M'; synth_done = true;
        }
        if (op == 'M') {
            for (int i = 0; i < run; ++i) {
                if ((int64_t)ref_pos >= c.l_pac) { past_end = true; break; }     // sam.c:270 asserts
                const uint32_t bt = (c.pac[ref_pos >> 2] >> ((~ref_pos & 3u) << 1)) & 3u;
                uint32_t b = 4u;                                                 // beyond the read: never equal
                if (si >= 0 && si < L) {
                    b = rev ? rd[L - 1 - si] : rd[si];
                    if (rev && b < 4u) b = 3u - b;
                }
                if (bt == b) ++n_match;
                else {
                    const uint32_t meta = (c.mixref[ref_pos >> 3] >> (4u * (ref_pos & 7u))) & 15u;
                    if ((meta & (1u << b)) != 0u && n_rs < 64) {                 // sam.c:281-286
                        if (n_rs < xstride) xvrow[n_rs] = (uint16_t)(si - (int)in.seq_start);
                        ++n_rs;
                    }
                    ++nm;
                    if (n_match != 0) m.put_num(n_match);
                    n_match = 0;
                    m.put("ACGT"[bt]);
                }
                ++ref_pos; ++si;
            }
        } else if (op == 'I') { nm += run; si += run; }
        else if (op == 'D') {
            if (n_match != 0) m.put_num(n_match);
            n_match = 0; nm += run;
            m.put('^');
            for (int i = 0; i < run; ++i) {
                const uint32_t p = ref_pos < (uint32_t)c.l_pac ? ref_pos : (uint32_t)c.l_pac - 1u;   // the reference reads on; stay inside
                m.put("ACGT"[(c.pac[p >> 2] >> ((~p & 3u) << 1)) & 3u]);
                ++ref_pos;
            }
        }
        if (cg && op != '\0') ++ci;
    }
    if (n_match != 0) m.put_num(n_match);
    if (mstride > 0) m.buf[m.len < mstride - 1 ? m.len : mstride - 1] = '\0';
    o.nm = nm; o.n_xv = (uint16_t)n_rs;
    o.md_len = past_end ? (int16_t)-3 : (m.ovf ? (int16_t)-2 : (int16_t)m.len);
    return o;
}

__global__ void __launch_bounds__(128)
md_nm_kernel(DevCtx c, const uint8_t *__restrict__ codes, const uint32_t *__restrict__ roffs,
             const salt_mdnm_in_t *__restrict__ items, size_t n, const char *__restrict__ cigars, int cstride,
             char *__restrict__ md, int mstride, uint16_t *__restrict__ xv, int xstride, salt_mdnm_out_t *__restrict__ out)
{
    const size_t it = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (it >= n) return;
    out[it] = md_nm_one(c, codes, roffs, items[it], cigars + it * (size_t)cstride, cstride, md + it * (size_t)mstride, mstride,
                        xv + it * (size_t)xstride, xstride);
}

// ---- the tags of a verified chunk's primaries, without anything crossing the host link on the way in ----
// cig_row[read] = row of the read's CIGAR in the verification stage's compact list (gapped primaries only)
__global__ void __launch_bounds__(256)
tail_cigrow_kernel(const uint32_t *__restrict__ cig_reads, const uint32_t *__restrict__ cig_count, uint32_t n_reads, int32_t *__restrict__ cig_row)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t n = min(*cig_count, n_reads);
    if (i < n && cig_reads[i] < n_reads) cig_row[cig_reads[i]] = (int32_t)i;
}

__global__ void __launch_bounds__(128)
tail_primaries_kernel(DevCtx c, const uint8_t *__restrict__ codes, const uint32_t *__restrict__ roffs,
                      const salt_verify_out_t *__restrict__ rec, const int32_t *__restrict__ cig_row, const char *__restrict__ cigs,
                      int cstride, char *__restrict__ md, int mstride, uint16_t *__restrict__ xv, int xstride,
                      salt_mdnm_out_t *__restrict__ out, uint32_t *__restrict__ md_bytes)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= c.n_reads) return;
    const salt_verify_out_t q = rec[i];
    salt_mdnm_in_t in; in.rs = (i << 1) | (uint32_t)(q.strand & 1u); in.pos = q.pos; in.seq_start = 0;       // query.c:284
    const char *cg = nullptr;
    if (q.pos != 0xFFFFFFFFu && q.is_gap == 1) {
        const int32_t row = cig_row[i];
        if (row < 0) { in.pos = 0xFFFFFFFFu; }               // a gapped primary without its CIGAR: no tags rather than wrong ones
        else cg = cigs + (size_t)row * (size_t)cstride;
    }
    const salt_mdnm_out_t o = md_nm_one(c, codes, roffs, in, cg, cstride, md + (size_t)i * (size_t)mstride, mstride,
                                        xv + (size_t)i * (size_t)xstride, xstride);
    out[i] = o;
    md_bytes[i] = (uint32_t)((o.md_len > 0 ? o.md_len : 0) + 1);       // string + NUL in the packed stream
}

__global__ void __launch_bounds__(256)
tail_pack_kernel(const char *__restrict__ md, int mstride, const uint32_t *__restrict__ offs, uint32_t n, char *__restrict__ packed)
{
    const uint32_t i = (blockIdx.x * blockDim.x + threadIdx.x) >> 2;          // four lanes per string
    const int lane = threadIdx.x & 3;
    if (i >= n) return;
    const uint32_t a = offs[i], len = offs[i + 1] - a;                        // len includes the NUL
    const char *__restrict__ src = md + (size_t)i * (size_t)mstride;
    for (uint32_t k = lane; k + 1 < len; k += 4) packed[a + k] = src[k];
    if (lane == 0) packed[a + len - 1] = '\0';
}

cudaError_t launch_tail_primaries(const DevCtx &c, const uint8_t *codes, const uint32_t *roffs, const salt_verify_out_t *rec,
                                  const uint32_t *cig_reads, const uint32_t *cig_count, const char *cigs, int cstride,
                                  int32_t *cig_row, char *md, int mstride, uint16_t *xv, int xstride, salt_mdnm_out_t *out,
                                  uint32_t *md_bytes, cudaStream_t st)
{
    const uint32_t n = c.n_reads;
    if (!n) return cudaSuccess;
    cudaError_t e = cudaMemsetAsync(cig_row, 0xFF, (size_t)n * 4, st);
    if (e != cudaSuccess) return e;
    if (cig_reads) SALT_LAUNCH(tail_cigrow_kernel, (n + 255) / 256, 256, 0, st, cig_reads, cig_count, n, cig_row);
    SALT_LAUNCH(tail_primaries_kernel, (n + 127) / 128, 128, 0, st, c, codes, roffs, rec, cig_row, cigs, cstride, md, mstride, xv, xstride,
                out, md_bytes);
    return cudaGetLastError();
}

cudaError_t launch_tail_pack(const char *md, int mstride, const uint32_t *offs, uint32_t n, char *packed, cudaStream_t st)
{
    if (!n) return cudaSuccess;
    SALT_LAUNCH(tail_pack_kernel, (unsigned)(((size_t)n * 4 + 255) / 256), 256, 0, st, md, mstride, offs, n, packed);
    return cudaGetLastError();
}

cudaError_t launch_md_nm(const DevCtx &c, const uint8_t *codes, const uint32_t *roffs, const salt_mdnm_in_t *items, size_t n,
                         const char *cigars, int cstride, char *md, int mstride, uint16_t *xv, int xstride,
                         salt_mdnm_out_t *out, cudaStream_t st)
{
    if (!n) return cudaSuccess;
    SALT_LAUNCH(md_nm_kernel, (unsigned)((n + 127) / 128), 128, 0, st, c, codes, roffs, items, n, cigars, cstride,
                md, mstride, xv, xstride, out);
    return cudaGetLastError();
}

}  // namespace salt
