// seed.cu -- sm_100a kernels for salt's single-end seeding + locate:
//
//   alnse_seed_overlap   alnse.c:199-312   per seed start: 12-mer lookup (lookup.h:39) + backward search on the
//                                          primary-reference FM-index (bwt.c:281 bwt_match_exact_alt, :140 bwt_2occ),
//                                          the same on the SNP-context FM-index (rbwt.c:619, :159), both extended to the
//                                          left while the interval is wider than max_seed; intervals sorted by width
//   alnse_locate_alt     alnse.c:633-731   suffix-array walks (bwt.c:89 bwt_sa, rbwt.c:316 Rbwt_back_bwt_sa), position
//                                          minus seed offset in uint32 arithmetic, at most max_locate loci, sorted
//
// The index arrives exactly as the reference's loaders leave it in memory (indexio.c:23-50): PREFIX.C.bwt / .C.sa
// (BWA layout: 128-base blocks of 4 counts + 8 words), PREFIX.C.lkt, PREFIX.R.backward.{bwt,occ,sa}.
//
// Mapping: seed_kernel runs one thread per (read, strand, seed start) -- every LF step is a dependent random access,
// so parallelism across seeds is what hides HBM latency; locate_kernel runs one warp per (read, strand): lane 0 orders
// the intervals with the reference's own (unstable) introsort so that ties break identically, the lanes walk the
// suffix array for 32 rows at a time, a ballot + prefix count reproduces "stop at max_locate" exactly, and the list is
// sorted in shared memory (bitonic).  Results are the candidate lists of the slot: the verification stage consumes
// them in place, they never cross the host link.
#if !defined(SALT_EMUL)
#include <cuda_runtime.h>
#endif

#include "common.cuh"
#include "kernels.h"

namespace salt {

// ---------------------------------------------------------------- primary-reference FM-index (bwt.h layout)
__device__ __forceinline__ uint32_t c_count16(uint32_t x, int c, int m)
{
    // bases of a 16-base word sit 2 bits each, first base in the top bits (bwt.h:62 bwt_B0); count c among the first m
    uint32_t y = ((c & 2) ? x : ~x) >> 1 & ((c & 1) ? x : ~x) & 0x55555555u;
    if (m < 16) y &= ~((1u << (2 * (16 - m))) - 1u);
    return (uint32_t)__popc(y);
}

// The device keeps its own, denser copy of the reference's BWT (built once by build_c32_kernel from the BWA-layout
// words): per 32 bases one 32-byte entry = the four counts up to the entry's first base + its two 16-base words.  An
// occurrence count is then ONE aligned 32-byte read and at most two masked popcounts -- no loop, so the lanes of a warp
// stay together -- where the 128-base blocks of the file layout (bwt.h:44-57) need up to eight words.  HBM is what a
// B200 has plenty of: 1 byte per base (3.1 GB at GRCh38 size) instead of 0.375.
__device__ __forceinline__ uint32_t c_pick(const uint4 v, int c) { return c == 0 ? v.x : c == 1 ? v.y : c == 2 ? v.z : v.w; }

// bwt_occ (bwt.c:106-129): occurrences of c in B[0..k]
__device__ __forceinline__ uint32_t c_occ(const FmIndexDev &ix, uint32_t k, int c)
{
    if (k == ix.c_seq_len) return ix.c_L2[c + 1] - ix.c_L2[c];
    if (k == 0xFFFFFFFFu) return 0u;
    if (k >= ix.c_primary) --k;
    const uint4 *__restrict__ e = ix.c32 + (size_t)(k >> 5) * 2;
    const uint4 cnt = e[0], w = e[1];
    const int r = (int)(k & 31u) + 1;                   // bases of this entry to count: 1..32
    uint32_t n = c_pick(cnt, c) + c_count16(w.x, c, r < 16 ? r : 16);
    if (r > 16) n += c_count16(w.y, c, r - 16);
    return n;
}

__device__ __forceinline__ int c_B0(const FmIndexDev &ix, uint32_t k)
{
    const uint4 w = ix.c32[(size_t)(k >> 5) * 2 + 1];
    return (int)(((k & 16u) ? w.y : w.x) >> ((~k & 0xfu) << 1) & 3u);
}

// one entry of the dense table from the file layout: counts of block k/128 plus the words of the block before the entry
__global__ void __launch_bounds__(256)
build_c32_kernel(const uint32_t *__restrict__ cbwt, size_t n_words, size_t n_entries, uint4 *__restrict__ c32)
{
    const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n_entries) return;
    const size_t blk = e >> 2; const int sub = (int)(e & 3);
    const uint32_t *__restrict__ p = cbwt + blk * 12;
    auto word = [&](size_t i) { return blk * 12 + i < n_words ? p[i] : 0u; };
    uint32_t cnt[4];
    for (int c = 0; c < 4; ++c) {
        uint32_t n = word((size_t)c);
        for (int w = 0; w < 2 * sub; ++w) n += c_count16(word(4 + (size_t)w), c, 16);
        cnt[c] = n;
    }
    c32[e * 2] = make_uint4(cnt[0], cnt[1], cnt[2], cnt[3]);
    c32[e * 2 + 1] = make_uint4(word(4 + 2 * (size_t)sub), word(5 + 2 * (size_t)sub), 0u, 0u);
}

// bwt_invPsi (bwt.h:67-72)
__device__ __forceinline__ uint32_t c_inv_psi(const FmIndexDev &ix, uint32_t k)
{
    if (k == ix.c_primary) return 0u;
    const int c = k < ix.c_primary ? c_B0(ix, k) : c_B0(ix, k - 1);
    return ix.c_L2[c] + c_occ(ix, k, c);
}

// bwt_sa (bwt.c:89-104)
__device__ __forceinline__ uint32_t c_sa(const FmIndexDev &ix, uint32_t k)
{
    uint32_t sa = 0;
    while (k % ix.c_sa_intv != 0) { ++sa; k = c_inv_psi(ix, k); }
    return sa + ix.c_sa[k / ix.c_sa_intv];
}

// occurrences of c among the block-relative bases [a, b] of one 128-base block (p = its eight 16-base words)
__device__ __forceinline__ uint32_t c_count_range(const uint32_t *__restrict__ p, int a, int b, int c)
{
    uint32_t n = 0;
    for (int w = a >> 4; w <= (b >> 4); ++w) {
        const uint32_t x = p[w];
        uint32_t y = ((c & 2) ? x : ~x) >> 1 & ((c & 1) ? x : ~x) & 0x55555555u;
        const int lo = w << 4;
        if (a > lo) y &= 0xFFFFFFFFu >> (2 * (a - lo));                  // drop bases before a (top bits)
        if (b < lo + 15) y &= ~((1u << (2 * (lo + 15 - b))) - 1u);       // drop bases after b
        n += (uint32_t)__popc(y);
    }
    return n;
}

// one backward-search step (bwt.c:293-299): returns false when the interval empties.  bwt_2occ (bwt.c:140-176) is an
// optimised pair of bwt_occ calls; so is this: when both ends fall into one 128-base block -- the usual case once the
// interval has narrowed -- the second count is the first plus the occurrences in between.
__device__ __forceinline__ bool c_step(const FmIndexDev &ix, int c, uint32_t &k, uint32_t &l)
{
    const uint32_t km = k - 1;
    uint32_t ok, ol;
    const bool plain = km == 0xFFFFFFFFu || km == ix.c_seq_len || l == ix.c_seq_len || l == 0xFFFFFFFFu;
    const uint32_t kk = km >= ix.c_primary ? km - 1 : km, ll = l >= ix.c_primary ? l - 1 : l;
    if (!plain && (kk >> 5) == (ll >> 5) && kk <= ll) {
        const uint4 *__restrict__ e = ix.c32 + (size_t)(kk >> 5) * 2;
        const uint4 cnt = e[0], w = e[1];
        const uint32_t ww[2] = {w.x, w.y};
        const int rk = (int)(kk & 31u) + 1;
        ok = c_pick(cnt, c) + c_count16(w.x, c, rk < 16 ? rk : 16) + (rk > 16 ? c_count16(w.y, c, rk - 16) : 0u);
        ol = ok;
        if (ll > kk) ol += c_count_range(ww, (int)(kk & 31u) + 1, (int)(ll & 31u), c);
    } else { ok = c_occ(ix, km, c); ol = c_occ(ix, l, c); }
    k = ix.c_L2[c] + ok + 1;
    l = ix.c_L2[c] + ol;
    return k <= l;
}

// ---------------------------------------------------------------- SNP-context FM-index (rbwt.h layout)
// occurrences of c among BWT characters [a, b): 4 bits each, first character in the top nibble (rbwt.h:113-117)
__device__ __forceinline__ uint32_t r_match8(uint32_t word, uint32_t pat)
{
    uint32_t x = word ^ pat;                              // matching nibbles become 0
    x |= x >> 1; x |= x >> 2;
    return ~x & 0x11111111u;
}

__device__ __forceinline__ uint32_t r_count(const uint32_t *__restrict__ code, uint32_t a, uint32_t b, uint32_t c)
{
    const uint32_t pat = c * 0x11111111u;
    const uint32_t wa = a >> 3, wb = (b - 1) >> 3;        // b > a
    uint32_t first = r_match8(code[wa], pat);
    if (a & 7u) first &= 0xFFFFFFFFu >> (4 * (a & 7u));                  // drop the characters before a (top nibbles)
    if (wa == wb) {
        if (b & 7u) first &= ~((1u << (4 * (8 - (b & 7u)))) - 1u);       // ... and those from b on
        return (uint32_t)__popc(first);
    }
    uint32_t n = (uint32_t)__popc(first);
    for (uint32_t w = wa + 1; w < wb; ++w) n += (uint32_t)__popc(r_match8(code[w], pat));
    uint32_t last = r_match8(code[wb], pat);
    if (b & 7u) last &= ~((1u << (4 * (8 - (b & 7u)))) - 1u);
    return n + (uint32_t)__popc(last);
}

// occurrences of c among BWT characters [0, index) from the reference's own tables: Rbwt_BWTOccValue (rbwt.c:159-189)
// with BWTOccValueExplicit (:38-79) -- bidirectional explicit counts every 256 characters (16 bit, two per word) on top
// of major counts every 65536 -- without the inverseSa0 adjustment.  Used once, to build the dense table below.
__device__ __forceinline__ uint32_t r_occ_file(const FmIndexDev &ix, uint32_t index, uint32_t c)
{
    const uint32_t e = (index + 127u) >> 8, at = e << 8;
    const uint32_t minor = ix.r_occ[(size_t)(e >> 1) * 5 + c];
    uint32_t v = ix.r_occ_major[(size_t)(at >> 16) * 5 + c] + ((e & 1u) ? (minor & 0xFFFFu) : (minor >> 16));
    if (at < index) v += r_count(ix.r_bwt, at, index, c);
    else if (at > index) v -= r_count(ix.r_bwt, index, at, c);
    return v;
}

// The device's own layout of the SNP-context index: per 64 characters one 64-byte line = the five counts up to the
// line's first character (A, C, G, T, #) + its eight 4-bit words.  One aligned line per occurrence count, at most eight
// words to look at, no second and third table.
__global__ void __launch_bounds__(256)
build_r64_kernel(FmIndexDev ix, size_t n_entries, size_t n_words, uint4 *__restrict__ r64)
{
    const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n_entries) return;
    const uint32_t at = (uint32_t)(e << 6);
    uint32_t cnt[5];
    for (uint32_t c = 0; c < 5; ++c) cnt[c] = r_occ_file(ix, at, c);
    uint32_t w[8];
    for (int i = 0; i < 8; ++i) { const size_t wi = e * 8 + (size_t)i; w[i] = wi < n_words ? ix.r_bwt[wi] : 0u; }
    r64[e * 4] = make_uint4(cnt[0], cnt[1], cnt[2], cnt[3]);
    r64[e * 4 + 1] = make_uint4(cnt[4], 0u, 0u, 0u);
    r64[e * 4 + 2] = make_uint4(w[0], w[1], w[2], w[3]);
    r64[e * 4 + 3] = make_uint4(w[4], w[5], w[6], w[7]);
}

// Rbwt_BWTOccValue (rbwt.c:159-189): occurrences of c among the BWT characters before `index` ($ is not encoded)
__device__ __forceinline__ uint32_t r_occ(const FmIndexDev &ix, uint32_t index, uint32_t c)
{
    if (index > ix.r_inv_sa0) --index;
    const uint4 *__restrict__ e = ix.r64 + (size_t)(index >> 6) * 4;
    const uint4 c03 = e[0];
    uint32_t v = c == 4u ? e[1].x : c_pick(c03, (int)c);
    const uint32_t r = index & 63u;                       // characters of this line to count: 0..63
    if (r) {
        const uint32_t pat = c * 0x11111111u;
        const uint4 a = e[2];
        const uint32_t wa[4] = {a.x, a.y, a.z, a.w};
        const uint32_t full = r >> 3, part = r & 7u;      // whole words, characters of the next one
#pragma unroll
        for (uint32_t i = 0; i < 4; ++i) {
            uint32_t m = r_match8(wa[i], pat);
            if (i > full || (i == full && part == 0)) m = 0;
            else if (i == full) m &= ~((1u << (4 * (8 - part))) - 1u);
            v += (uint32_t)__popc(m);
        }
        if (r > 32) {
            const uint4 b = e[3];
            const uint32_t wb[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (uint32_t i = 0; i < 4; ++i) {
                uint32_t m = r_match8(wb[i], pat);
                if (i + 4 > full || (i + 4 == full && part == 0)) m = 0;
                else if (i + 4 == full) m &= ~((1u << (4 * (8 - part))) - 1u);
                v += (uint32_t)__popc(m);
            }
        }
    }
    return v;
}

__device__ __forceinline__ uint32_t r_nt(const FmIndexDev &ix, uint32_t pos)       // Rbwt_bwt2nt, rbwt.h:101-121
{
    if (pos == ix.r_inv_sa0) return 4u;
    if (pos > ix.r_inv_sa0) --pos;
    return (ix.r_bwt[pos >> 3] >> (4u * (7u - (pos & 7u)))) & 15u;
}

// Rbwt_back_bwt_sa (rbwt.c:316-333)
__device__ __forceinline__ uint32_t r_sa(const FmIndexDev &ix, uint32_t i)
{
    uint32_t step = 0;
    const uint32_t n_acgt = ix.r_cum[4];
    while (i <= n_acgt) {
        const uint32_t c = r_nt(ix, i);
        i = ix.r_cum[c] + r_occ(ix, i, c) + 1;
        ++step;
    }
    return ix.r_sa_sharp[i - n_acgt - 1] + step - 1;
}

// The pair Rbwt_BWTOccValue(k, c), Rbwt_BWTOccValue(l + 1, c) of one backward step (rbwt.c:641-642).  Both are plain
// occurrence counts, so when the two (adjusted) indexes are close the second is the first plus the occurrences in
// between -- identical values, about half the words read once the interval has narrowed.
__device__ __forceinline__ void r_occ2(const FmIndexDev &ix, uint32_t k, uint32_t l1, uint32_t c, uint32_t &ok, uint32_t &ol)
{
    ok = r_occ(ix, k, c);
    const uint32_t ka = k > ix.r_inv_sa0 ? k - 1 : k, la = l1 > ix.r_inv_sa0 ? l1 - 1 : l1;
    if (la >= ka && la - ka <= 64u) ol = la > ka ? ok + r_count(ix.r_bwt, ka, la, c) : ok;
    else ol = r_occ(ix, l1, c);
}

__device__ __forceinline__ bool r_step(const FmIndexDev &ix, uint32_t c, uint32_t &k, uint32_t &l)
{
    uint32_t ok, ol;
    r_occ2(ix, k, l + 1, c, ok, ol);
    k = ix.r_cum[c] + ok + 1;
    l = ix.r_cum[c] + ol;
    return k <= l;
}

// ---------------------------------------------------------------- reads
// code of base i of strand s of a read: query->seq / query->rseq (query.c:46-64: reverse, 3-c for c < 4)
__device__ __forceinline__ int seed_code(const uint8_t *__restrict__ codes, int L, int strand, int i)
{
    if (!strand) return codes[i];
    const int c = codes[L - 1 - i];
    return c < 4 ? 3 - c : c;
}

// ---------------------------------------------------------------- alnse_seed_overlap, one seed start per thread
__global__ void __launch_bounds__(128)
seed_kernel(FmIndexDev ix, SeedOpt opt, const uint8_t *__restrict__ codes, const uint32_t *__restrict__ roffs,
            uint32_t n_reads, int max_seeds, SeedSai *__restrict__ sai /* [rs][2][max_seeds] */)
{
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t rs = tid / (unsigned)max_seeds;
    const int si = (int)(tid % (unsigned)max_seeds);
    if (rs >= (size_t)n_reads * 2) return;
    const uint32_t r = (uint32_t)(rs >> 1);
    const int strand = (int)(rs & 1);
    const uint32_t o = roffs[r];
    const int L = (int)(roffs[r + 1] - o);
    const uint8_t *__restrict__ rd = codes + o;
    SeedSai *out = sai + rs * 2 * (size_t)max_seeds;
    SeedSai none; none.sp = 1; none.ep = 0; none.offset = 0;
    out[si] = none; out[max_seeds + si] = none;
    const int seed_start = si * opt.l_overlap;                       // alnse.c:226-228
    if (L < opt.l_seed || seed_start > L - opt.l_seed) return;
    const int seed_end = seed_start + opt.l_seed - 1;

    // ---- primary reference: 12-mer lookup, then the rest of the seed, then extension to the left
    {
        uint32_t k = 1, l = 0;
        uint32_t item = 0; bool has_n = false;
        for (int i = seed_end - ix.l_lkt + 1; i <= seed_end; ++i) {   // LKT_seq2LktItem, lookup.c:163-176
            const int c = seed_code(rd, L, strand, i);
            if (c > 3) has_n = true;
            item = (item << 2) | (uint32_t)(c & 3);
        }
        if (!has_n) { k = ix.lkt[item]; l = ix.lkt[item + 1] - 1; }  // lookup.h:39-52
        bool ok = k <= l;
        for (int i = opt.l_seed - ix.l_lkt - 1; ok && i >= 0; --i) { // bwt_match_exact_alt on seq[seed_start .. +l_seed-l_lkt)
            const int c = seed_code(rd, L, strand, seed_start + i);
            if (c > 3) { ok = false; break; }
            ok = c_step(ix, c, k, l);
        }
        if (ok) {
            int l_extend = 0;
            while (l - k > (uint32_t)opt.max_seed && l_extend < seed_start) {      // alnse.c:249-259
                const int c = seed_code(rd, L, strand, seed_start - l_extend - 1);
                if (c > 3) break;
                uint32_t nk = k, nl = l;
                if (!c_step(ix, c, nk, nl)) break;
                k = nk; l = nl;
                ++l_extend;
                if (l - k <= (uint32_t)opt.max_seed) break;
            }
            SeedSai s; s.sp = k; s.ep = l; s.offset = (uint32_t)(seed_start - l_extend);
            out[si] = s;
        }
    }
    if (opt.seed_only_ref) return;
    // ---- SNP-context index: the whole seed backwards (Rbwt_exact_match_backward, rbwt.c:619-649), then extension
    {
        uint32_t k = 0, l = ix.r_text_len;
        bool ok = true;
        int step = 0;
        while (k <= l && step < opt.l_seed) {
            const int c = seed_code(rd, L, strand, seed_start + opt.l_seed - step - 1);
            if (c > 3) { ok = false; break; }
            r_step(ix, (uint32_t)c, k, l);
            ++step;
        }
        if (ok && l >= k) {
            int l_extend = 0;
            while (l - k > (uint32_t)opt.max_seed && l_extend < seed_start) {      // alnse.c:280-291
                // the reference indexes cumulativeFreq / occ with the raw code here: an N (4) extends over '#'
                const uint32_t c = (uint32_t)seed_code(rd, L, strand, seed_start - l_extend - 1);
                uint32_t okk, oll;
                r_occ2(ix, k, l + 1, c, okk, oll);
                if (okk + 1 > oll) break;
                k = ix.r_cum[c] + okk + 1;
                l = ix.r_cum[c] + oll;
                ++l_extend;
                if (l - k <= (uint32_t)opt.max_seed) break;
            }
            SeedSai s; s.sp = k; s.ep = l; s.offset = (uint32_t)(seed_start - l_extend);
            out[max_seeds + si] = s;
        }
    }
}

// ---------------------------------------------------------------- ks_introsort_sai (ksort.h:180-226, alnse.c:35-38)
// The reference's sort is not stable, and which interval comes first decides whose loci survive the max_locate cut:
// the same algorithm (median-of-three pivot moved to the end, Hoare partition, ranges of <= 17 left to one final
// insertion sort, comb sort when the depth budget runs out) is run here on the same input order.
__device__ __forceinline__ bool sai_lt(const SeedSai &a, const SeedSai &b) { return a.ep - a.sp < b.ep - b.sp; }

__device__ void sai_insertsort(SeedSai *s, SeedSai *t)             // __ks_insertsort, ksort.h:148-155
{
    for (SeedSai *i = s + 1; i < t; ++i)
        for (SeedSai *j = i; j > s && sai_lt(*j, *(j - 1)); --j) { const SeedSai tmp = *j; *j = *(j - 1); *(j - 1) = tmp; }
}

__device__ void sai_combsort(int n, SeedSai *a)                     // ks_combsort, ksort.h:156-178
{
    const double shrink = 1.2473309501039786540366528676643;
    int do_swap;
    int gap = n;
    do {
        if (gap > 2) {
            gap = (int)(gap / shrink);
            if (gap == 9 || gap == 10) gap = 11;
        }
        do_swap = 0;
        for (SeedSai *i = a; i < a + n - gap; ++i) {
            SeedSai *j = i + gap;
            if (sai_lt(*j, *i)) { const SeedSai tmp = *i; *i = *j; *j = tmp; do_swap = 1; }
        }
    } while (do_swap || gap > 2);
    if (gap != 1) sai_insertsort(a, a + n);
}

__device__ void sai_introsort(int n, SeedSai *a)
{
    if (n < 1) return;
    if (n == 2) { if (sai_lt(a[1], a[0])) { const SeedSai tmp = a[0]; a[0] = a[1]; a[1] = tmp; } return; }
    int d;
    for (d = 2; (1 << d) < n; ++d);
    struct Frame { SeedSai *left, *right; int depth; };
    Frame stack[64];
    Frame *top = stack;
    SeedSai *s = a, *t = a + (n - 1);
    d <<= 1;
    for (;;) {
        if (s < t) {
            if (--d == 0) { sai_combsort((int)(t - s) + 1, s); t = s; continue; }
            SeedSai *i = s, *j = t, *k = i + ((j - i) >> 1) + 1;
            if (sai_lt(*k, *i)) { if (sai_lt(*k, *j)) k = j; }
            else k = sai_lt(*j, *i) ? i : j;
            const SeedSai rp = *k;
            if (k != t) { const SeedSai tmp = *k; *k = *t; *t = tmp; }
            for (;;) {
                do ++i; while (sai_lt(*i, rp));
                do --j; while (i <= j && sai_lt(rp, *j));
                if (j <= i) break;
                const SeedSai tmp = *i; *i = *j; *j = tmp;
            }
            { const SeedSai tmp = *i; *i = *t; *t = tmp; }
            if (i - s > t - i) {
                if (i - s > 16) { top->left = s; top->right = i - 1; top->depth = d; ++top; }
                s = t - i > 16 ? i + 1 : t;
            } else {
                if (t - i > 16) { top->left = i + 1; top->right = t; top->depth = d; ++top; }
                t = i - s > 16 ? i - 1 : s;
            }
        } else {
            if (top == stack) { sai_insertsort(a, a + n); return; }
            --top; s = top->left; t = top->right; d = top->depth;
        }
    }
}

// ---------------------------------------------------------------- alnse_locate[_alt], up to 64 (read, strand)s per CTA
// A strand's lists are short on most genomes (a handful of intervals of a few rows each) while a suffix-array walk is
// long (up to sa_intv - 1 LF steps on the primary index, up to a whole local pattern on the SNP-context index), so rows
// are the unit of parallel work, not strands.  One thread keeps a strand's books (interval order, the max_locate cut,
// the final sort of a short list); the walks of ALL strands of the CTA are laid end to end and shared out over its 128
// threads, one index at a time, several walks per thread, so that every lane walks, all lanes run the same code and
// long and short walks average out.
// Loci go straight to the strand's row of the fixed-stride list array; lists of up to LOC_SMALL entries are sorted by
// their thread in shared memory, longer ones are queued for sort_long_kernel.
// Shared memory: per strand max_seeds sai (12 B) and max_seeds + 1 row offsets (64 bit); per CTA the staging of one
// round of walks.
constexpr int LOC_STRANDS = 64;                // at most, per CTA of 128 threads (fewer when there are many seed starts)
constexpr int LOC_SMALL = 16;
constexpr int LOC_STAGE = LOC_STRANDS * LOC_SMALL;   // walks per round and CTA; also the short lists' sorting room

__host__ __device__ __forceinline__ size_t locate_strand_words(int max_seeds)
{
    return (((size_t)max_seeds * 3 + 1) & ~(size_t)1) + 2 * ((size_t)max_seeds + 2);
}

__global__ void __launch_bounds__(128)
locate_kernel(FmIndexDev ix, SeedOpt opt, const uint32_t *__restrict__ roffs, uint32_t n_reads, int max_seeds, int strands,
              uint32_t ref_l, SeedSai *__restrict__ sai, uint32_t *__restrict__ counts /* [2][n_reads] */,
              uint32_t *__restrict__ lists /* [rs][list_cap] */, uint32_t *__restrict__ long_list, uint32_t *__restrict__ long_count,
              uint8_t *__restrict__ status /* [2][n_reads] */)
{
    SALT_DYN_SMEM(uint32_t, s_mem);
    __shared__ uint32_t s_pos[LOC_STAGE];
    __shared__ uint8_t s_keep[LOC_STAGE];
    __shared__ uint32_t s_want[LOC_STRANDS + 1];                      // exclusive prefix of the strands' shares of a round
    __shared__ unsigned long long s_done[LOC_STRANDS];                // rows of the current part already walked
    __shared__ uint32_t s_n[LOC_STRANDS];                             // aux->loci.n
    __shared__ int s_m[LOC_STRANDS];                                  // valid intervals of the current part
    const int t = threadIdx.x;
    const size_t per_strand = locate_strand_words(max_seeds);
    const size_t row_off = ((size_t)max_seeds * 3 + 1) & ~(size_t)1;  // row offsets are 64 bit: an interval that could not be
                                                                      // narrowed may span the whole suffix array
    const size_t rs0 = (size_t)blockIdx.x * (size_t)strands;
    const size_t rs = rs0 + (size_t)t;
    const bool mine = t < strands && rs < (size_t)n_reads * 2;        // this thread keeps the books of strand rs
    SeedSai *my_sai = reinterpret_cast<SeedSai *>(s_mem + (size_t)(t < strands ? t : 0) * per_strand);
    unsigned long long *my_row = reinterpret_cast<unsigned long long *>(s_mem + (size_t)(t < strands ? t : 0) * per_strand + row_off);
    // mode 0 = alnse_locate_alt (single-end): at most max_locate loci in all.  mode 1 = alnse_locate (paired-end,
    // alnse.c:501-631): at most max_locate + 1 rows of each primary-index interval (:521), MAX_LOC_POS loci in all (:533),
    // SNP-context intervals wider than max_locate are subsampled with rand() by the reference (:577-596) -- such an
    // interval is left out here and the strand is flagged (bit 0), as is a list cut by list_cap < MAX_LOC_POS (bit 1).
    const bool pe = opt.mode == 1;
    const uint32_t stride = (uint32_t)opt.list_cap;
    const uint32_t max_locate = pe ? (stride < 0x40000u ? stride : 0x40000u) : (uint32_t)opt.max_locate;
    uint32_t flags = 0;
    if (t < LOC_STRANDS) { s_n[t] = 0; s_m[t] = 0; s_done[t] = 0; }
    for (int part = 0; part < 2; ++part) {                            // 0: sai_C, 1: sai_backwardR (sai_forwardR is empty here)
        // ---- plan: the valid intervals in seed order, sorted as the reference sorts them, their rows laid end to end
        if (t < strands) {
            int m = 0;
            if (mine) {
                const SeedSai *g = sai + rs * 2 * (size_t)max_seeds + (size_t)part * max_seeds;
                for (int i = 0; i < max_seeds; ++i) { const SeedSai v = g[i]; if (v.sp <= v.ep) my_sai[m++] = v; }
                sai_introsort(m, my_sai);
            }
            unsigned long long acc = 0;
            for (int i = 0; i < m; ++i) {
                const SeedSai v = my_sai[i];
                my_row[i] = acc;
                if (!pe) {
                    uint32_t skip = 1;
                    if (part == 1) { skip = (v.ep + 1 - v.sp) / 0x40000u; if ((int)skip <= 0) skip = 1; }    // alnse.c:702-703
                    acc += ((unsigned long long)(v.ep - v.sp)) / skip + 1;                                   // rows sp, sp+skip, .. <= ep
                } else if (part == 0) {
                    const unsigned long long w = (unsigned long long)(v.ep - v.sp) + 1;
                    acc += w < (unsigned long long)opt.max_locate + 1 ? w : (unsigned long long)opt.max_locate + 1;   // alnse.c:521
                } else if (v.ep - v.sp > (uint32_t)opt.max_locate) flags |= 1u;                              // alnse.c:577: rand()
                else acc += (unsigned long long)(v.ep - v.sp) + 1;
            }
            my_row[m] = acc;
            s_m[t] = m; s_done[t] = 0;
        }
        __syncthreads();
        // ---- rounds: every strand asks for the rows it may still push, all threads walk them, every strand takes its own
        for (;;) {
            if (t == 0) {
                uint32_t acc = 0;
                for (int gi = 0; gi < strands; ++gi) {
                    s_want[gi] = acc;
                    const unsigned long long *row_g = reinterpret_cast<const unsigned long long *>(s_mem + (size_t)gi * per_strand + row_off);
                    const unsigned long long left = s_m[gi] ? row_g[s_m[gi]] - s_done[gi] : 0ull;
                    const uint32_t room = max_locate - s_n[gi];          // the reference stops at max_locate pushes (alnse.c:678)
                    uint32_t w = (uint32_t)(left < (unsigned long long)room ? left : (unsigned long long)room);
                    if (w > (uint32_t)LOC_STAGE - acc) w = (uint32_t)LOC_STAGE - acc;
                    acc += w;
                }
                s_want[strands] = acc;
            }
            __syncthreads();
            const uint32_t total = s_want[strands];
            if (total == 0) break;
            for (uint32_t f = (uint32_t)t; f < total; f += blockDim.x) {
                int glo = 0, ghi = strands - 1;                        // the strand whose share holds row f
                while (glo < ghi) { const int mid = (glo + ghi + 1) >> 1; if (s_want[mid] <= f) glo = mid; else ghi = mid - 1; }
                const int gi = glo;
                const uint32_t *gb = s_mem + (size_t)gi * per_strand;
                const SeedSai *sai_g = reinterpret_cast<const SeedSai *>(gb);
                const unsigned long long *row_g = reinterpret_cast<const unsigned long long *>(gb + row_off);
                const unsigned long long flat = s_done[gi] + (unsigned long long)(f - s_want[gi]);
                int lo = 0, hi = s_m[gi] - 1;                          // last interval whose first row is <= flat
                while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (row_g[mid] <= flat) lo = mid; else hi = mid - 1; }
                const SeedSai v = sai_g[lo];
                uint32_t skip = 1;
                if (!pe && part == 1) { skip = (v.ep + 1 - v.sp) / 0x40000u; if ((int)skip <= 0) skip = 1; }
                const uint32_t j = v.sp + (uint32_t)(flat - row_g[lo]) * skip;
                const uint32_t pos = (part == 0 ? c_sa(ix, j) : r_sa(ix, j)) - v.offset;      // uint32 arithmetic, may wrap (alnse.c:669)
                const size_t rsg = rs0 + (size_t)gi;
                const uint32_t l_seq = roffs[(rsg >> 1) + 1] - roffs[rsg >> 1];
                bool keep = !(pos + l_seq > ref_l);
                if (part == 1 && pos > ref_l) keep = false;                                   // alnse.c:711
                s_pos[f] = pos; s_keep[f] = keep ? 1 : 0;
            }
            __syncthreads();
            if (t < strands) {                                         // each strand pushes its rows in order
                const uint32_t f0 = s_want[t], cnt = s_want[t + 1] - f0;
                if (cnt) {
                    uint32_t n = s_n[t];
                    uint32_t *dst = lists + rs * (size_t)stride;
                    for (uint32_t i = 0; i < cnt; ++i)
                        if (s_keep[f0 + i] && n < max_locate) dst[n++] = s_pos[f0 + i];
                    s_n[t] = n; s_done[t] += cnt;
                }
            }
            __syncthreads();
        }
        __syncthreads();
    }
    if (!mine) return;
    const uint32_t r = (uint32_t)(rs >> 1);
    const uint32_t n = s_n[t];
    uint32_t *dst = lists + rs * (size_t)stride;
    counts[(rs & 1) * (size_t)n_reads + r] = n;
    if (pe && stride < 0x40000u && n == stride) flags |= 2u;
    if (status) status[(rs & 1) * (size_t)n_reads + r] = (uint8_t)flags;
    if (n > LOC_SMALL) { long_list[atomicAdd(long_count, 1u)] = (uint32_t)rs; return; }     // ks_introsort(uint32_t) of a long list
    if (n < 2) return;
    // any correct sort gives the reference's array: insertion sort in this strand's corner of the staging area (no other
    // thread of the CTA touches the staging area after the last round)
    uint32_t *mine_sorted = s_pos + (size_t)t * LOC_SMALL;
    for (uint32_t i = 0; i < n; ++i) mine_sorted[i] = dst[i];
    for (uint32_t i = 1; i < n; ++i) {
        const uint32_t v = mine_sorted[i];
        uint32_t j = i;
        while (j > 0 && mine_sorted[j - 1] > v) { mine_sorted[j] = mine_sorted[j - 1]; --j; }
        mine_sorted[j] = v;
    }
    for (uint32_t i = 0; i < n; ++i) dst[i] = mine_sorted[i];
}

// lists longer than LOC_SMALL: one warp per queued (read, strand), bitonic in shared memory over cap2 >= max_locate
__global__ void __launch_bounds__(128)
sort_long_kernel(const uint32_t *__restrict__ long_list, const uint32_t *__restrict__ long_count, uint32_t n_reads, int max_locate /* list stride */,
                 int cap2, const uint32_t *__restrict__ counts, uint32_t *__restrict__ lists)
{
    constexpr unsigned FULL = 0xffffffffu;
    SALT_DYN_SMEM(uint32_t, s_mem);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint32_t *s_loci = s_mem + (size_t)warp * (size_t)cap2;
    const uint32_t total = *long_count;
    const uint32_t n_warps = gridDim.x * (blockDim.x >> 5);
    for (uint32_t e = blockIdx.x * (blockDim.x >> 5) + warp; e < total; e += n_warps) {
        const uint32_t rs = long_list[e];
        const uint32_t n = counts[(rs & 1) * (size_t)n_reads + (rs >> 1)];
        uint32_t *dst = lists + (size_t)rs * (size_t)max_locate;
        uint32_t n2 = 2;
        while (n2 < n) n2 <<= 1;
        for (uint32_t i = lane; i < n2; i += 32) s_loci[i] = i < n ? dst[i] : 0xFFFFFFFFu;
        __syncwarp(FULL);
        for (uint32_t k = 2; k <= n2; k <<= 1)
            for (uint32_t j = k >> 1; j > 0; j >>= 1) {
                for (uint32_t i = lane; i < n2; i += 32) {
                    const uint32_t p = i ^ j;
                    if (p > i) {
                        const uint32_t a = s_loci[i], b = s_loci[p];
                        const bool up = (i & k) == 0;
                        if ((a > b) == up) { s_loci[i] = b; s_loci[p] = a; }
                    }
                }
                __syncwarp(FULL);
            }
        for (uint32_t i = lane; i < n; i += 32) dst[i] = s_loci[i];
        __syncwarp(FULL);
    }
}

// CSR gather: strand s of read r takes counts[s][r] loci from its fixed-stride row
__global__ void __launch_bounds__(256)
seed_gather_kernel(const uint32_t *__restrict__ lists, int max_locate, const uint32_t *__restrict__ offs0,
                   const uint32_t *__restrict__ offs1, uint32_t n_reads, uint32_t *__restrict__ loci0, uint32_t *__restrict__ loci1)
{
    const size_t rs = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 3;         // eight lanes per list
    const int lane = threadIdx.x & 7;
    if (rs >= (size_t)n_reads * 2) return;
    const uint32_t r = (uint32_t)(rs >> 1);
    const uint32_t *__restrict__ offs = (rs & 1) ? offs1 : offs0;
    uint32_t *__restrict__ dst = ((rs & 1) ? loci1 : loci0) + offs[r];
    const uint32_t n = offs[r + 1] - offs[r];
    const uint32_t *__restrict__ src = lists + rs * (size_t)max_locate;
    for (uint32_t i = lane; i < n; i += 8) dst[i] = src[i];
}

// ---------------------------------------------------------------- launchers
size_t fm_c32_entries(size_t c_bwt_words) { return (c_bwt_words + 11) / 12 * 4; }
size_t fm_r64_entries(uint32_t r_text_len) { return ((size_t)r_text_len >> 6) + 2; }

cudaError_t launch_build_dense_index(const FmIndexDev &ix, size_t c_bwt_words, size_t r_bwt_words_padded, uint4 *c32, uint4 *r64, cudaStream_t st)
{
    const size_t nc = fm_c32_entries(c_bwt_words), nr = fm_r64_entries(ix.r_text_len);
    SALT_LAUNCH(build_c32_kernel, (unsigned)((nc + 255) / 256), 256, 0, st, ix.cbwt, c_bwt_words, nc, c32);
    SALT_LAUNCH(build_r64_kernel, (unsigned)((nr + 255) / 256), 256, 0, st, ix, nr, r_bwt_words_padded, r64);
    return cudaGetLastError();
}

size_t seed_sai_bytes(uint32_t n_reads, int max_seeds) { return (size_t)n_reads * 2 * 2 * (size_t)max_seeds * sizeof(SeedSai); }

cudaError_t launch_seed(const FmIndexDev &ix, const SeedOpt &opt, const uint8_t *codes, const uint32_t *roffs, uint32_t n_reads,
                        int max_seeds, SeedSai *sai, cudaStream_t st)
{
    const size_t total = (size_t)n_reads * 2 * (size_t)max_seeds;
    if (!total) return cudaSuccess;
    SALT_LAUNCH(seed_kernel, (unsigned)((total + 127) / 128), 128, 0, st, ix, opt, codes, roffs, n_reads, max_seeds, sai);
    return cudaGetLastError();
}

cudaError_t launch_locate(const FmIndexDev &ix, const SeedOpt &opt, const uint32_t *roffs, uint32_t n_reads, int max_seeds,
                          uint32_t ref_l, SeedSai *sai, uint32_t *counts, uint32_t *lists, uint32_t *long_list, uint32_t *long_count,
                          uint8_t *status, int sm_count, cudaStream_t st)
{
    if (!n_reads) return cudaSuccess;
    cudaError_t e = cudaMemsetAsync(long_count, 0, 4, st);
    if (e != cudaSuccess) return e;
    const size_t per_strand = locate_strand_words(max_seeds) * 4;
    int strands = LOC_STRANDS;                              // per CTA of 128 threads
    while (strands > 8 && per_strand * strands > 96 * 1024) strands >>= 1;
    const size_t smem = per_strand * strands;
    if (smem > 200 * 1024) return cudaErrorInvalidValue;
    {
        auto kern = locate_kernel;
        if (smem > 48 * 1024 && (e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return e;
        const size_t n_rs = (size_t)n_reads * 2;
        SALT_LAUNCH(kern, (unsigned)((n_rs + strands - 1) / strands), 128, smem, st, ix, opt, roffs, n_reads, max_seeds, strands, ref_l, sai,
                    counts, lists, long_list, long_count, status);
    }
    if (opt.list_cap > LOC_SMALL) {
        int cap2 = 2 * LOC_SMALL;
        while (cap2 < opt.list_cap) cap2 <<= 1;
        int warps = 4;
        while (warps > 1 && (size_t)cap2 * 4 * warps > 96 * 1024) warps >>= 1;
        const size_t smem2 = (size_t)cap2 * 4 * warps;
        auto kern = sort_long_kernel;
        if (smem2 > 48 * 1024 && (e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2)) != cudaSuccess) return e;
        const size_t want = ((size_t)n_reads * 2 + warps - 1) / warps;
        const size_t cap = (size_t)sm_count * 8;
        SALT_LAUNCH(kern, (unsigned)(want < cap ? want : cap), warps * 32, smem2, st, long_list, long_count, n_reads, opt.list_cap, cap2,
                    counts, lists);
    }
    return cudaGetLastError();
}

cudaError_t launch_seed_gather(const uint32_t *lists, int max_locate, const uint32_t *offs0, const uint32_t *offs1,
                               uint32_t n_reads, uint32_t *loci0, uint32_t *loci1, cudaStream_t st)
{
    const size_t total = (size_t)n_reads * 2 * 8;
    if (!total) return cudaSuccess;
    SALT_LAUNCH(seed_gather_kernel, (unsigned)((total + 255) / 256), 256, 0, st, lists, max_locate, offs0, offs1, n_reads, loci0, loci1);
    return cudaGetLastError();
}

}  // namespace salt
