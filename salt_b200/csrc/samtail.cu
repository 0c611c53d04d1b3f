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

__global__ void __launch_bounds__(128)
md_nm_kernel(DevCtx c, const uint8_t *__restrict__ codes, const uint32_t *__restrict__ roffs,
             const salt_mdnm_in_t *__restrict__ items, size_t n, const char *__restrict__ cigars, int cstride,
             char *__restrict__ md, int mstride, uint16_t *__restrict__ xv, int xstride, salt_mdnm_out_t *__restrict__ out)
{
    const size_t it = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (it >= n) return;
    const salt_mdnm_in_t in = items[it];
    salt_mdnm_out_t o; o.nm = 0; o.md_len = 0; o.n_xv = 0;
    MdOut m{md + it * (size_t)mstride, mstride, 0, false};
    const uint32_t rid = in.rs >> 1;
    const bool rev = (in.rs & 1u) != 0;
    if (in.pos == 0xFFFFFFFFu || rid >= c.n_reads) {                    // sam.c:248: unmapped reads get no tags
        if (mstride > 0) m.buf[0] = '\0';
        out[it] = o;
        return;
    }
    const int L = (int)c.rd_len[rid];
    const uint8_t *__restrict__ rd = codes + roffs[rid];
    const char *__restrict__ cg = cigars + it * (size_t)cstride;
    uint32_t ref_pos = in.pos;
    int si = (int)in.seq_start;
    int nm = 0, n_match = 0, n_rs = 0;
    bool past_end = false;
    int ci = 0;
    while (ci < cstride && cg[ci] != '\0' && !past_end) {
        int run = 0;
        while (ci < cstride && cg[ci] >= '0' && cg[ci] <= '9') { run = run * 10 + (cg[ci] - '0'); ++ci; }
        const char op = ci < cstride ? cg[ci] : '\0';
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
                        if (n_rs < xstride) xv[it * (size_t)xstride + n_rs] = (uint16_t)(si - (int)in.seq_start);
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
        if (op != '\0') ++ci;
    }
    if (n_match != 0) m.put_num(n_match);
    if (mstride > 0) m.buf[m.len < mstride - 1 ? m.len : mstride - 1] = '\0';
    o.nm = nm; o.n_xv = (uint16_t)n_rs;
    o.md_len = past_end ? (int16_t)-3 : (m.ovf ? (int16_t)-2 : (int16_t)m.len);
    out[it] = o;
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
