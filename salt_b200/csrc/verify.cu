// verify.cu -- sm_100a kernels for the ungapped / gapped verification of (read, locus) pairs
// and the per-read acceptance scans that reproduce salt's sequential threshold logic.
//
//   pack_reads      query codes -> one-hot nibbles for both strands      (query.c:46-64, editdistance.c:215-218)
//   mismatch        ed_mismatch                                          (editdistance.c:88-163)
//   lv              ed_diff -> computeEditDistance                       (editdistance.c:174, LandauVishkin.c:19)
//   lv_cigar        ed_diff_withcigar -> computeEditDistanceWithCigar    (editdistance.c:234, LandauVishkin.c:176)
//   nogap_fused     alnse_check_nogap x2 + running threshold + primary     (alnse.c:734, :1014-1036, :1077-1097)
//   lv_filter       exact pigeonhole pre-filter in front of Landau-Vishkin
//   scan_gap        alnse_check_withgap's acceptance logic                 (alnse.c:871, :372-393)
#if !defined(SALT_EMUL)
#include <cuda_runtime.h>
#endif

#include "common.cuh"
#include "lv_core.cuh"
#include "kernels.h"

namespace salt {

// --------------------------------------------------------------------------------------
// pack_reads: eight lanes per read-strand, lane t writing 64-bit words t, t+8, .. (16 bases each).
// A code c becomes the nibble 1 << c, N (and anything above 3) becomes 15: one shift of a 20-bit
// table.  Strand 1 is the reverse complement (query.c:46-64): base i comes from L-1-i with 3 - c.
// --------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t pack_nib(uint32_t c, bool comp)
{
    c = c > 4u ? 4u : c;
    // forward: 0,1,2,3,N -> 1,2,4,8,15     complement: 0,1,2,3,N -> 8,4,2,1,15
    return ((comp ? 0xF1248u : 0xF8421u) >> (4u * c)) & 15u;
}

__global__ void __launch_bounds__(256)
pack_reads_kernel(const uint8_t *__restrict__ codes, const uint32_t *__restrict__ offs, uint32_t n_reads,
                  uint32_t W64, uint64_t *__restrict__ rd4, uint16_t *__restrict__ rd_len)
{
    const size_t g = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 3;      // read-strand
    const int lane = threadIdx.x & 7;
    if (g >= (size_t)n_reads * 2) return;
    const uint32_t rid = (uint32_t)(g >> 1);
    const bool rev = (g & 1) != 0;
    const uint32_t o = offs[rid];
    const int L = (int)(offs[rid + 1] - o);
    const uint8_t *__restrict__ src = codes + o;
    uint64_t *__restrict__ dst = rd4 + g * W64;
    for (uint32_t w = lane; w < W64; w += 8) {
        const int i0 = (int)w * 16;
        uint64_t word = 0;
        if (i0 < L) {
            const int nb = L - i0 < 16 ? L - i0 : 16;
            if (!rev) {
                if (nb == 16 && ((o + (uint32_t)i0) & 3u) == 0u) {              // four aligned 32-bit loads
                    const uint32_t *__restrict__ s4 = reinterpret_cast<const uint32_t *>(src + i0);
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const uint32_t v = s4[k];
#pragma unroll
                        for (int b = 0; b < 4; ++b) word |= (uint64_t)pack_nib((v >> (8 * b)) & 255u, false) << (4 * (4 * k + b));
                    }
                } else {
                    for (int b = 0; b < nb; ++b) word |= (uint64_t)pack_nib(src[i0 + b], false) << (4 * b);
                }
            } else {
                const int top = L - 1 - i0;                                     // base i0 of the strand comes from here
                if (nb == 16 && ((o + (uint32_t)(top - 15)) & 3u) == 0u) {
                    const uint32_t *__restrict__ s4 = reinterpret_cast<const uint32_t *>(src + top - 15);
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const uint32_t v = s4[k];                               // bytes top-15+4k .. top-12+4k
#pragma unroll
                        for (int b = 0; b < 4; ++b) word |= (uint64_t)pack_nib((v >> (8 * b)) & 255u, true) << (4 * (15 - 4 * k - b));
                    }
                } else {
                    for (int b = 0; b < nb; ++b) word |= (uint64_t)pack_nib(src[top - b], true) << (4 * b);
                }
            }
        }
        dst[w] = word;
    }
    if (lane == 0 && !rev) rd_len[rid] = (uint16_t)L;
}

// --------------------------------------------------------------------------------------
// mismatch: batched ed_mismatch on a flat pair list (editdistance.c:88-163).  G lanes per pair,
// lane t owning 64-bit words t, t+G, .. of the packed read and of the aligned window -- the
// counting scheme of nogap_fused below (one coalesced 8-byte-aligned window load per lane, the next
// lane's low word by shuffle, one funnel shift each way, popc of window AND read) without the
// read reuse, since neighbouring pairs may belong to different reads.  Four pairs are in flight per
// group; control flow is warp-uniform so the shuffles run on the full mask.
// G*WPL*16 >= l_max + 16 (launcher), so the last lane never needs a word from beyond the group.
// --------------------------------------------------------------------------------------
template <int G, int WPL>
__global__ void __launch_bounds__(256)
mismatch_kernel(DevCtx c, const salt_pair_t *__restrict__ pairs, size_t n, int max_err, int8_t *__restrict__ out)
{
    constexpr unsigned FULL = 0xffffffffu;
    constexpr int U = 4;                                 // pairs per group per pass
    const int lane = threadIdx.x % G;
    const size_t group = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) / G;
    const size_t ngroups = ((size_t)gridDim.x * blockDim.x) / G;
    const uint2 *__restrict__ mixl = reinterpret_cast<const uint2 *>(c.mixref) + lane;
    const size_t passes = (n + ngroups * U - 1) / (ngroups * U);          // identical for every group
    for (size_t ps = 0; ps < passes; ++ps) {
        const size_t first = (ps * ngroups + group) * U;
        uint2 q[U][WPL], rw[U][WPL], rwh[U][WPL];
        uint32_t ph[U]; int Ls[U]; bool good[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const size_t i = first + u;
            salt_pair_t p; p.rs = 0; p.pos = 0;
            if (i < n) p = pairs[i];
            const uint32_t rid = p.rs >> 1;
            const bool live = i < n && rid < c.n_reads;
            const int L = live ? (int)c.rd_len[rid] : 0;
            Ls[u] = L;
            good[u] = live && L > 0 && (uint64_t)p.pos + (uint64_t)L <= (uint64_t)c.l;
            const uint32_t cpos = good[u] ? p.pos : 0u;                  // a bad pair counts window 0 and is ignored
            ph[u] = ((cpos & 7u) << 2) | ((cpos & 8u) << 28);
            const uint2 *__restrict__ wp = mixl + (cpos >> 4);
            const uint2 *__restrict__ rrow = reinterpret_cast<const uint2 *>(c.rd4 + (size_t)(live ? p.rs : 0u) * c.W64);
#pragma unroll
            for (int w = 0; w < WPL; ++w) {
                q[u][w] = wp[w * G];
                rw[u][w] = (live && (uint32_t)(lane + w * G) < c.W64) ? rrow[lane + w * G] : make_uint2(0u, 0u);
            }
        }
        uint32_t packed[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int sh = (int)ph[u];
            const bool hi = (int)ph[u] < 0;
            uint32_t m = 0;
#pragma unroll
            for (int w = 0; w < WPL; ++w) {
                uint32_t up = __shfl_up_sync(FULL, rw[u][w].y, 1, G);     // the read moved up by one 32-bit word
                if (lane == 0) up = 0u;
                if (w > 0) { const uint32_t wrap = __shfl_sync(FULL, rw[u][w > 0 ? w - 1 : 0].y, G - 1, G); if (lane == 0) up = wrap; }
                rwh[u][w] = make_uint2(up, rw[u][w].x);
                uint32_t n0 = __shfl_down_sync(FULL, q[u][w].x, 1, G);
                if (w + 1 < WPL) {
                    const uint32_t w0 = __shfl_sync(FULL, q[u][w + 1 < WPL ? w + 1 : w].x, 0, G);
                    if (lane == G - 1) n0 = w0;
                }
                uint32_t x0 = __funnelshift_r(q[u][w].x, q[u][w].y, sh) & (hi ? rwh[u][w].x : rw[u][w].x);
                uint32_t x1 = __funnelshift_r(q[u][w].y, n0, sh) & (hi ? rwh[u][w].y : rw[u][w].y);
                // general form (reads may contain N = 15): a matching nibble is a non-zero nibble
                x0 |= x0 >> 1; x0 |= x0 >> 2; x0 &= 0x11111111u;
                x1 |= x1 >> 1; x1 |= x1 >> 2; x1 &= 0x11111111u;
                m += (uint32_t)(__popc(x0) + __popc(x1));
            }
            packed[u] = m;
        }
#pragma unroll
        for (int o = G / 2; o > 0; o >>= 1)
#pragma unroll
            for (int u = 0; u < U; ++u) packed[u] += __shfl_xor_sync(FULL, packed[u], o, G);
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const size_t i = first + u;
            if (lane == u && i < n) {
                const int nmis = Ls[u] - (int)packed[u];
                out[i] = (int8_t)((!good[u] || nmis > max_err) ? -1 : nmis);
            }
        }
    }
}

// --------------------------------------------------------------------------------------
// Staging of one pair's text window and pattern into shared memory (32-bit words).
// T: nibbles pos..pos+tlen-1 of mixRef, zero from tlen on.  P: the packed read, zero from plen on.
// --------------------------------------------------------------------------------------
template <int G>
__device__ __forceinline__ void lv_stage(const DevCtx &c, uint32_t rs, uint32_t pos, int plen, int tlen,
                                         uint32_t *T, uint32_t *P, int TW, int PW, int lane)
{
    const uint32_t *__restrict__ mix = c.mixref;
    for (int i = lane; i < TW; i += G) {
        const int valid = tlen - 8 * i;
        uint32_t x = 0;
        if (valid > 0) {
            const uint32_t o = pos + 8u * (uint32_t)i;
            const uint32_t w = o >> 3;
            const int sh = (int)(o & 7u) * 4;
            x = __funnelshift_r(mix[w], mix[w + 1], sh);
            if (valid < 8) x &= (1u << (4 * valid)) - 1u;
        }
        T[i] = x;
    }
    const uint32_t *__restrict__ r32 = reinterpret_cast<const uint32_t *>(c.rd4 + (size_t)rs * c.W64);
    const int have = (int)c.W64 * 2;
    for (int i = lane; i < PW; i += G) P[i] = i < have ? r32[i] : 0u;
    (void)plen;
}

// Cooperative level-0 extension: lane t looks at symbols [8t, 8t+8) of each 8G-symbol block.
template <int G>
__device__ __forceinline__ int lv_level0(const uint32_t *T, const uint32_t *P, int plen, int tlen,
                                         int lane, unsigned gmask, int gshift)
{
    const int e0 = imin(plen, tlen);
    for (int base = 0; base < e0; base += 8 * G) {
        const int off = base + 8 * lane;
        // lanes past the end report a stop at their first symbol; the min() below clamps it
        const uint32_t z = off < e0 ? zero_nibbles(nib8(P, off) & nib8(T, off)) : 1u;
        const unsigned bal = (__ballot_sync(gmask, z != 0) >> gshift) & (unsigned)((1ull << G) - 1ull);
        if (bal) {
            const int first = __ffs((int)bal) - 1;
            const int idx = __shfl_sync(gmask, z ? first_set_nibble(z) : 0, first, G);
            return imin(base + 8 * first + idx, e0);
        }
    }
    return e0;
}

// Shared-memory words needed per pair for a chunk whose longest read is l_max.
__host__ __device__ inline int lv_tw(int l_max) { return (l_max + 4 + 64) / 8 + 2; }
__host__ __device__ inline int lv_pw(int l_max) { return (l_max + 64) / 8 + 2; }
// thread-per-pair kernel: window kept as raw 16-byte-aligned reference words (up to 31 nibbles in front)
// The thread-per-pair kernels never look past nibble toff + textLen + 2 of the window or patternLen + 1 of the read
// (extensions stop at endl(d), lv_core.cuh), each plus the 8 nibbles and the next word nib8 touches.
__host__ __device__ inline int lv_twr(int l_max) { return ((31 + l_max + 4 + 10) / 8 + 2 + 3) & ~3; }
__host__ __device__ inline int lv_pwt(int l_max) { return (l_max + 10) / 8 + 2; }
__device__ __forceinline__ uint32_t lv_tailmask(int r)             // keep the low r nibbles (r <= 0: none, r >= 8: all)
{
    return r >= 8 ? 0xffffffffu : (r <= 0 ? 0u : (1u << (4 * r)) - 1u);
}

// --------------------------------------------------------------------------------------
// Group-cooperative longest common extension (same result as lv_extend, lv_core.cuh).
// Every lane first looks at the gate and at the first 8 symbols of ITS diagonal; almost all
// diagonals stop there.  The few that run on (the diagonal the read really lies on) are then
// finished one at a time by the whole group, lane t testing symbols [8t, 8t+8) of each block of 8G
// beyond the point reached, so a 100-base extension costs one ballot instead of a dozen dependent
// iterations of one lane while the others wait.
// --------------------------------------------------------------------------------------
template <int G>
__device__ __forceinline__ int lv_extend_group(const uint32_t *T, const uint32_t *P, int best, int d, bool active,
                                               int plen, int tlen, int lane, unsigned gmask, int gshift)
{
    constexpr unsigned gfull = (unsigned)((1ull << G) - 1ull);
    const int lim = imin(plen, tlen - d);
    bool more = false;
    if (active) {
        const uint32_t pc = nib8(P, best), tc = nib8(T, d + best);
        const uint32_t pb = pc & 15u, tb = tc & 15u;
        if (pb == tb) {                                             // equality gate (LandauVishkin.c:79)
            if (pb == 0) best = lim;                                // both exhausted: the 8-byte x == 0 shortcut
            else {
                const uint32_t z = zero_nibbles(pc & tc);
                if (z) best = imin(best + first_set_nibble(z), lim);
                else if (best + 8 >= lim) best = lim;
                else { best += 8; more = true; }
            }
        }
    }
    unsigned todo = (__ballot_sync(gmask, more) >> gshift) & gfull;
    while (todo) {
        const int src = __ffs((int)todo) - 1;
        todo &= todo - 1;
        const int b = __shfl_sync(gmask, best, src, G);
        const int dd = __shfl_sync(gmask, d, src, G);
        const int e = imin(plen, tlen - dd);
        int res = e;
        for (int base = b; base < e; base += 8 * G) {
            const int off = base + 8 * lane;
            // lanes past the end report a stop at their first symbol; the min() below clamps it
            const uint32_t z = off < e ? zero_nibbles(nib8(P, off) & nib8(T, dd + off)) : 1u;
            const unsigned bal = (__ballot_sync(gmask, z != 0) >> gshift) & gfull;
            if (bal) {
                const int first = __ffs((int)bal) - 1;
                const int idx = __shfl_sync(gmask, z ? first_set_nibble(z) : 0, first, G);
                res = imin(base + 8 * first + idx, e);
                break;
            }
        }
        if (lane == src) best = res;
    }
    return best;
}

// --------------------------------------------------------------------------------------
// lv: G lanes per pair, DPL diagonals per lane (diagonal d = lane*DPL + q - G*DPL/2).
// Furthest-reaching values of the previous level live in registers; neighbours are
// exchanged with __shfl_up/down inside the group.  The diagonal order of the reference
// (0,+1,-1,..) does not affect the returned level, so all diagonals of a level run at once.
// Work items: pairs[i] -> out[i] for i < n; with `slots` non-null the list length is read from
// *wl_count on the device and pairs[i] -> out[slots[i]] (persistent groups stride over the list).
// --------------------------------------------------------------------------------------
template <int G, int DPL>
__global__ void __launch_bounds__(128)
lv_kernel(DevCtx c, const salt_pair_t *__restrict__ pairs, size_t n, int k_fixed,
          const uint32_t *__restrict__ worklist, const uint32_t *__restrict__ wl_count,
          int8_t *__restrict__ out)
{
    SALT_DYN_SMEM(uint32_t, smem);
    constexpr int ND = G * DPL, C = ND / 2;
    const int TW = lv_tw((int)c.l_max), PW = lv_pw((int)c.l_max);
    const int gl = threadIdx.x / G, lane = threadIdx.x % G;
    const int gshift = (threadIdx.x & 31) / G * G;
    const unsigned gmask = (unsigned)(((1ull << G) - 1ull) << gshift);
    uint32_t *T = smem + (size_t)gl * (TW + PW);
    uint32_t *P = T + TW;
    const size_t groups = (size_t)gridDim.x * (blockDim.x / G);
    const size_t count = worklist ? (size_t)*wl_count : n;

    for (size_t it = (size_t)blockIdx.x * (blockDim.x / G) + gl; it < count; it += groups) {
        const size_t slot = worklist ? worklist[it] : it;
        const salt_pair_t p = pairs[it];
        const uint32_t rid = p.rs >> 1;
        const int plen = rid < c.n_reads ? (int)c.rd_len[rid] : 0;
        const int tlen = plen + 4;                                  // alnse.c:373
        int k = k_fixed >= 0 ? k_fixed : plen / 10;                 // alnse.c:1090
        k = imin(k, LV_MAXK - 1);                                   // LandauVishkin.c:31
        k = imin(k, C - 1);                                         // diagonals this instantiation holds
        int result = -1;
        // editdistance.c:178: window must lie inside the reference
        const bool ok = plen > 0 && (uint64_t)p.pos + (uint64_t)tlen <= (uint64_t)c.l;
        if (ok) {
            lv_stage<G>(c, p.rs, p.pos, plen, tlen, T, P, TW, PW, lane);
            __syncwarp(gmask);
            const int L0 = lv_level0<G>(T, P, plen, tlen, lane, gmask, gshift);
            if (L0 == plen) {
                result = 0;                                         // endl(0) == plen since tlen > plen
            } else {
                int Lp[DPL];
#pragma unroll
                for (int q = 0; q < DPL; ++q) Lp[q] = (lane * DPL + q - C == 0) ? L0 : -2;
                for (int e = 1; e <= k; ++e) {
                    int lft = __shfl_up_sync(gmask, Lp[DPL - 1], 1, G);
                    int rgt = __shfl_down_sync(gmask, Lp[0], 1, G);
                    if (lane == 0) lft = -2;
                    if (lane == G - 1) rgt = -2;
                    int Ln[DPL];
                    bool hit = false;
#pragma unroll
                    for (int q = 0; q < DPL; ++q) {
                        const int d = lane * DPL + q - C;
                        const int left = q == 0 ? lft : Lp[q - 1];
                        const int right = q == DPL - 1 ? rgt : Lp[q + 1];
                        int best = imax(imax(Lp[q] + 1, left), right + 1);
                        const bool act = d >= -e && d <= e;
                        best = lv_extend_group<G>(T, P, best, d, act, plen, tlen, lane, gmask, gshift);
                        hit = hit || (act && best == plen);
                        Ln[q] = act ? best : -2;
                    }
                    if (__any_sync(gmask, hit)) { result = e; break; }
#pragma unroll
                    for (int q = 0; q < DPL; ++q) Lp[q] = Ln[q];
                }
            }
            __syncwarp(gmask);
        }
        if (lane == 0) out[slot] = (int8_t)result;
    }
}

// Thread-per-pair staging: the lane fills its own row with the window as the raw reference words from the
// 16-byte boundary below pos (text position 0 is nibble `toff` of the row, returned; nibbles from
// toff + tlen on are cleared -- the text is "0 beyond textLen"), then the packed read.  Wide independent
// loads; with an odd row stride the 32 rows of a warp sit in 32 different banks.
__device__ __forceinline__ int lv_stage_own(const DevCtx &c, const salt_pair_t p, int tlen, uint32_t *T, uint32_t *P,
                                            int TW, int PW)
{
    const uint32_t w0 = (p.pos >> 3) & ~3u;
    const int toff = (int)(p.pos - 8u * w0);
    const int tend = toff + tlen;
    const uint4 *__restrict__ src = reinterpret_cast<const uint4 *>(c.mixref + w0);
#pragma unroll 4
    for (int v = 0; v < TW / 4; ++v) {
        const int rem = tend - 32 * v;                      // valid nibbles from this vector's first one
        uint4 x = make_uint4(0u, 0u, 0u, 0u);
        if (rem > 0) x = src[v];
        T[4 * v + 0] = x.x & lv_tailmask(rem);
        T[4 * v + 1] = x.y & lv_tailmask(rem - 8);
        T[4 * v + 2] = x.z & lv_tailmask(rem - 16);
        T[4 * v + 3] = x.w & lv_tailmask(rem - 24);
    }
    const uint2 *__restrict__ r2 = reinterpret_cast<const uint2 *>(c.rd4 + (size_t)p.rs * c.W64);
#pragma unroll 4
    for (int i = 0; i < PW; i += 2) {
        uint2 y = make_uint2(0u, 0u);
        if ((i >> 1) < (int)c.W64) y = r2[i >> 1];
        P[i] = y.x;
        if (i + 1 < PW) P[i + 1] = y.y;
    }
    return toff;
}

// --------------------------------------------------------------------------------------
// lv_tpp: one THREAD per pair, all 2K+1 diagonals of a level in registers (loops over the
// diagonals are fully unrolled, so L[e-1][d-1], L[e-1][d], L[e-1][d+1] are plain registers).
// The warp-per-pair kernel above keeps at most 2e+1 of its 32 lanes busy at level e; here every
// lane works on its own pair, which costs ~6x fewer issue slots per pair for k <= 15.  Windows
// and reads are staged per warp: for each of its 32 pairs the lanes copy the pair's text and
// pattern words to shared memory with coalesced loads (row stride odd => conflict-free columns).
// Work items as in lv_kernel.
// --------------------------------------------------------------------------------------
template <int K>
__global__ void __launch_bounds__(128, K <= 10 ? 10 : 4)
lv_tpp_kernel(DevCtx c, const salt_pair_t *__restrict__ pairs, size_t n, int k_fixed,
              const uint32_t *__restrict__ slots, const uint32_t *__restrict__ wl_count,
              int8_t *__restrict__ out, LvDefer df)
{
    SALT_DYN_SMEM(uint32_t, smem);
    const int TW = lv_twr((int)c.l_max), PW = lv_pwt((int)c.l_max);
    const int stride = (TW + PW) | 1;                  // odd: the 32 rows of a warp sit in 32 different banks
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t *wbase = smem + (size_t)warp * 32 * stride;
    uint32_t *T = wbase + (size_t)lane * stride;
    uint32_t *P = T + TW;
    const size_t count = slots ? (size_t)*wl_count : n;
    const size_t step = (size_t)gridDim.x * blockDim.x;

    for (size_t base = (size_t)blockIdx.x * blockDim.x + (size_t)warp * 32; base < count; base += step) {
        const size_t it = base + lane;
        const bool live = it < count;
        salt_pair_t p; p.rs = 0; p.pos = 0;
        if (live) p = pairs[it];
        const uint32_t rid = p.rs >> 1;
        const int plen = (live && rid < c.n_reads) ? (int)c.rd_len[rid] : 0;
        const int tlen = plen + 4;                                  // alnse.c:373
        const bool ok = plen > 0 && (uint64_t)p.pos + (uint64_t)tlen <= (uint64_t)c.l;   // editdistance.c:178
        const int toff = ok ? lv_stage_own(c, p, tlen, T, P, TW, PW) : 0;
        __syncwarp();
        int result = -1;
        bool defer = false;
        if (ok) {
            int k = k_fixed >= 0 ? k_fixed : plen / 10;             // alnse.c:1090
            k = imin(k, LV_MAXK - 1);                               // LandauVishkin.c:31
            // first pass of two (df.pairs set): only the first K levels are walked here; a pair that needs more goes to
            // the list of the second pass, so that the lanes of a warp stop within a few levels of each other
            if (df.pairs && k > K) defer = true;
            k = imin(k, K);
            const int L0 = lv_extend0(T, P, plen, tlen, toff);
            if (L0 == plen) result = 0;
            else {
                int Lp[2 * K + 1];
#pragma unroll
                for (int i = 0; i < 2 * K + 1; ++i) Lp[i] = -2;
                Lp[K] = L0;
                for (int e = 1; e <= k; ++e) {
                    int Ln[2 * K + 1];
                    bool hit = false;
                    int pend_di = -1, pend_best = 0;                // one deferred long extension per level
#pragma unroll
                    for (int di = 0; di < 2 * K + 1; ++di) {
                        const int d = di - K;
                        int v = -2;
                        if (d >= -e && d <= e) {
                            const int left = di > 0 ? Lp[di > 0 ? di - 1 : 0] : -2;
                            const int right = di < 2 * K ? Lp[di < 2 * K ? di + 1 : 0] + 1 : -1;
                            bool more;
                            v = lv_extend_first(T, P, imax(imax(Lp[di] + 1, left), right), d, plen, tlen, more, toff);
                            if (more) {
                                if (pend_di < 0) { pend_di = di; pend_best = v; }
                                else v = lv_extend_more(T, P, v, d, plen, tlen, toff);
                            }
                            hit = hit || v == plen;
                        }
                        Ln[di] = v;
                    }
                    int pend_v = 0;
                    if (pend_di >= 0) {
                        pend_v = lv_extend_more(T, P, pend_best, pend_di - K, plen, tlen, toff);
                        hit = hit || pend_v == plen;
                    }
                    if (hit) { result = e; break; }
#pragma unroll
                    for (int i = 0; i < 2 * K + 1; ++i) Lp[i] = i == pend_di ? pend_v : Ln[i];
                }
            }
        }
        defer = defer && result < 0;                                // within the first K levels: final
        if (df.pairs) {
            const unsigned bal = __ballot_sync(0xffffffffu, live && defer);
            uint32_t at = 0;
            if (bal) {
                if (lane == 0) at = atomicAdd(df.count, (uint32_t)__popc(bal));
                at = __shfl_sync(0xffffffffu, at, 0) + (uint32_t)__popc(bal & ((1u << lane) - 1u));
            }
            if (live && defer) { df.pairs[at] = p; df.slots[at] = slots ? slots[it] : (uint32_t)it; }
        }
        if (live && !defer) out[slots ? slots[it] : it] = (int8_t)result;
        __syncwarp();
    }
}

// --------------------------------------------------------------------------------------
// lv_filter: an exact pigeonhole pre-filter in front of Landau-Vishkin.  One thread per pair,
// registers only, no divergence.
//
// computeEditDistance returns e <= k only if the pattern aligns to a prefix of the text with
// e edits, every other column being a match under salt's AND test (LandauVishkin.c:42-53,
// :85-95), on diagonals |d| <= e.  Cut the pattern into its nf = plen/8 full 8-base words.
// An edit spoils at most one word, so at least nf - k words are edit-free, and an edit-free
// word matches the text on ONE diagonal d, |d| <= k.  Hence
//
//      #{ w < nf : exists d in [-k, min(k,(k+4)/2)], all 8 nibbles of P_w & T(8w+d ..) non-zero }  >=  nf - k
//
// is necessary for a result >= 0.  Text outside [0, textLen) is read from the real reference
// instead of the zeros the reference sees; that can only add matches, so the filter stays
// conservative (it never rejects a pair the reference accepts).  A random decoy passes with
// probability ~1e-5 (needs two 8-mer hits at L=100, k=10); it would otherwise walk all k levels
// of Landau-Vishkin, the worst case.  Survivors are compacted into a worklist for the LV kernels.
//
// Work per pair: the 2k+1 diagonals are swept by moving a register copy of the window one
// nibble per step (16 funnel shifts) and testing the 12 words of a 96-base tile (AND, has-zero-
// nibble, running minimum): ~64 integer ops per diagonal.  Reads longer than 96 bases take
// several tiles.  Work items as in lv_kernel.
// --------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t lv_haszero(uint32_t x)      // non-zero iff some nibble of x is zero
{
    return (x - 0x11111111u) & ~x & 0x88888888u;
}

// One word offset A of the filter's diagonal sweep: shifts s_first..s_last (nibbles) of the window words
// TA[j + A], TA[j + A + 1] against the tile's 12 pattern words (one-hot reads: a word matches iff P & ~W == 0).
template <int A>
__device__ __forceinline__ void lv_filter_block(const uint32_t (&TA)[20], const uint32_t (&P)[12], uint32_t (&m)[12],
                                                int s_first, int s_last)
{
    for (int s = s_first; s <= s_last; s += 2) {
        const int sh0 = 4 * s, sh1 = 4 * (s + 1 <= s_last ? s + 1 : s);
#pragma unroll
        for (int j = 0; j < 12; ++j) {
            const uint32_t x0 = P[j] & ~__funnelshift_r(TA[j + A], TA[j + A + 1], sh0);
            const uint32_t x1 = P[j] & ~__funnelshift_r(TA[j + A], TA[j + A + 1], sh1);
            m[j] = salt_min3u(m[j], x0, x1);
        }
    }
}

__global__ void __launch_bounds__(128, 6)           // 80 registers, no spills; 7 and 8 spill (4 CTAs unbounded: 112 registers)
lv_filter_kernel(DevCtx c, const salt_pair_t *__restrict__ pairs, size_t n, int k_fixed,
                 const uint32_t *__restrict__ slots, const uint32_t *__restrict__ wl_count,
                 int8_t *__restrict__ out, salt_pair_t *__restrict__ pass_pairs, uint32_t *__restrict__ pass_slots,
                 uint32_t *__restrict__ pass_count)
{
    constexpr int TW = 12;                              // pattern words per tile (96 bases)
    const size_t count = slots ? (size_t)*wl_count : n;
    const size_t step = (size_t)gridDim.x * blockDim.x;
    const int lane = threadIdx.x & 31;
    // mixref is preceded by zero padding (engine.cu), so a window may start 32 bases before pos = 0
    const uint2 *__restrict__ mix = reinterpret_cast<const uint2 *>(c.mixref);
    for (size_t base = (size_t)blockIdx.x * blockDim.x + (threadIdx.x & ~31); base < count; base += step) {
        const size_t it = base + lane;
        const bool live = it < count;
        salt_pair_t p; p.rs = 0; p.pos = 0;
        if (live) p = pairs[it];
        const uint32_t rid = p.rs >> 1;
        const int plen = (live && rid < c.n_reads) ? (int)c.rd_len[rid] : 0;
        const int tlen = plen + 4;                                  // alnse.c:373
        const bool ok = plen > 0 && (uint64_t)p.pos + (uint64_t)tlen <= (uint64_t)c.l;   // editdistance.c:178
        int k = k_fixed >= 0 ? k_fixed : plen / 10;                 // alnse.c:1090
        k = imin(k, LV_MAXK - 1);                                   // LandauVishkin.c:31
        const int nf = plen >> 3;
        const int need = nf - k;
        // The alignment must end with the pattern consumed inside the text: final diagonal <= tlen - plen
        // = 4.  Being on diagonal d > 4 costs d edits to get there and d - 4 to come back, so positive
        // diagonals beyond (k + 4) / 2 cannot be part of a k-difference alignment.
        const int kp = imin(k, (k + 4) >> 1);
        const int kw = __reduce_max_sync(0xffffffffu, ok ? k : 0);
        const int kpw = __reduce_max_sync(0xffffffffu, ok ? kp : 0);
        const int nfw = __reduce_max_sync(0xffffffffu, (ok && need > 0) ? nf : 0);
        int found = 0;
        const uint32_t *__restrict__ prow = reinterpret_cast<const uint32_t *>(c.rd4 + (size_t)p.rs * c.W64);
        for (int t0 = 0; t0 < nfw; t0 += TW) {                      // warp-uniform tile loop
            // ---- window of the tile: TA[u] holds text nibbles 8u-32 .. 8u-25 relative to pos + 8*t0
            const int64_t g0 = (int64_t)p.pos + 8 * t0 - 32;        // first nibble, may be negative (front padding)
            const int a = (int)(g0 & 15);
            const uint2 *__restrict__ src = mix + (g0 >> 4);
            uint32_t raw[22];
#pragma unroll
            for (int i = 0; i < 11; ++i) {
                const uint2 v = ok ? src[i] : make_uint2(0u, 0u);
                raw[2 * i] = v.x; raw[2 * i + 1] = v.y;
            }
            uint32_t TA[20];
            {
                const int sh = (a & 7) * 4;
                const bool hi = (a & 8) != 0;
#pragma unroll
                for (int u = 0; u < 20; ++u) {
                    const uint32_t lo = hi ? raw[u + 1] : raw[u];
                    const uint32_t up = hi ? raw[u + 2] : raw[u + 1];
                    TA[u] = __funnelshift_r(lo, up, sh);
                }
            }
            uint32_t P[TW], m[TW];
            bool hasN = false;
#pragma unroll
            for (int j = 0; j < TW; ++j) {
                P[j] = (ok && t0 + j < nf) ? prow[t0 + j] : 0u;     // words past the last full one never match
                hasN = hasN || ((P[j] & (P[j] >> 1) & 0x11111111u) != 0u);
            }
            // A read base is one-hot, so "all eight nibbles of P & W non-zero" is "P & ~W == 0": one LOP3 and
            // a running minimum per word and diagonal.  N (15) matches any non-empty mask, which that test
            // does not express: a warp holding such a read takes the general has-zero-nibble form (warp-uniform).
            if (!__any_sync(0xffffffffu, hasN)) {
#pragma unroll
                for (int j = 0; j < TW; ++j) m[j] = (t0 + j < nf && ok) ? (P[j] & ~TA[j + 4]) : 1u;           // diagonal 0
                // Diagonal +d looks at text nibble 8j + d of the tile, i.e. TA[j + 4 + (d >> 3)] shifted by d & 7
                // nibbles; diagonal -d at nibble 8j + t - 32 with t = 32 - d, i.e. TA[j + (t >> 3)] shifted by t & 7.
                // Each block below covers the eight shifts of one word offset (static register indices), two
                // diagonals per step so that one three-input minimum folds both.  The sweep runs to the warp's
                // largest k: a diagonal beyond a lane's own k can only add matches, the filter stays exact.
                lv_filter_block<4>(TA, P, m, 1, imin(7, kpw));
                if (kpw >= 8) lv_filter_block<5>(TA, P, m, 0, imin(7, kpw - 8));
                if (kpw >= 16) lv_filter_block<6>(TA, P, m, 0, imin(7, kpw - 16));
                lv_filter_block<3>(TA, P, m, 8 - imin(8, kw), kw >= 1 ? 7 : -1);
                if (kw >= 9) lv_filter_block<2>(TA, P, m, 16 - imin(16, kw), 7);
                if (kw >= 17) lv_filter_block<1>(TA, P, m, 24 - imin(24, kw), 7);
                if (kw >= 25) lv_filter_block<0>(TA, P, m, 32 - imin(32, kw), 7);
                // an absent word (P = 0) would pass "P & ~W == 0": it was seeded with 1 and min() keeps 0 out only
                // if every test is masked too
#pragma unroll
                for (int j = 0; j < TW; ++j) if (!(t0 + j < nf && ok)) m[j] = 1u;
            } else {
#pragma unroll
            for (int j = 0; j < TW; ++j) {
                m[j] = lv_haszero(P[j] & TA[j + 4]);                // diagonal 0
            }
            // ---- diagonals +1 .. +k: the window moves down one nibble per step
            {
                uint32_t W[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) W[j] = TA[j + 4];
                for (int d = 1; d <= kpw; ++d) {
#pragma unroll
                    for (int j = 0; j < 15; ++j) W[j] = __funnelshift_r(W[j], W[j + 1], 4);
                    W[15] >>= 4;
                    if (d <= kp) {
#pragma unroll
                        for (int j = 0; j < TW; ++j) m[j] = min(m[j], lv_haszero(P[j] & W[j]));
                    }
                }
            }
            // ---- diagonals -1 .. -k: the window moves up one nibble per step
            {
                uint32_t W[16];                                     // W[j] <-> TA[j]: tile words sit at j = 4..15
#pragma unroll
                for (int j = 0; j < 16; ++j) W[j] = TA[j];
                for (int d = 1; d <= kw; ++d) {
#pragma unroll
                    for (int j = 15; j > 0; --j) W[j] = __funnelshift_l(W[j - 1], W[j], 4);
                    W[0] <<= 4;
                    if (d <= k) {
#pragma unroll
                        for (int j = 0; j < TW; ++j) m[j] = min(m[j], lv_haszero(P[j] & W[j + 4]));
                    }
                }
            }
            }
#pragma unroll
            for (int j = 0; j < TW; ++j) found += (m[j] == 0u && t0 + j < nf) ? 1 : 0;
        }
        const bool pass = ok && (need <= 0 || found >= need);
        // ---- survivors go to the LV worklist, everything else is decided here
        const unsigned bal = __ballot_sync(0xffffffffu, pass);
        uint32_t w0 = 0;
        if (lane == 0 && bal) w0 = atomicAdd(pass_count, (uint32_t)__popc(bal));
        w0 = __shfl_sync(0xffffffffu, w0, 0);
        const size_t slot = slots ? slots[live ? it : 0] : it;
        if (pass) {
            const uint32_t at = w0 + (uint32_t)__popc(bal & ((1u << lane) - 1u));
            pass_pairs[at] = p; pass_slots[at] = (uint32_t)slot;
        } else if (live) {
            out[slot] = (int8_t)-1;
        }
    }
}

// --------------------------------------------------------------------------------------
// lv_cigar: one warp per pair, 64 diagonals (2 per lane), furthest-reaching and action
// tables kept in shared memory for the backtrace, which lane 0 performs.
// Work items: (pairs[i], k_each[i]) -> cigars + i*stride, or, when `worklist` is non-null,
// read ids whose primary (rec[rid]) is gapped -> cigars + (list index)*stride (query.c:282-295).
// --------------------------------------------------------------------------------------
// shared-memory words one warp of lv_cigar needs: L (int16) and A (char) tables of `rows` levels x nd
// diagonals, then the staged text and pattern
__host__ __device__ inline size_t lv_cigar_tab_words(int rows, int nd) { return ((size_t)rows * nd * 3 + 3) / 4; }

template <int DPL>
__global__ void __launch_bounds__(128)
lv_cigar_kernel(DevCtx c, const salt_pair_t *__restrict__ pairs, const uint8_t *__restrict__ k_each, size_t n,
                const uint32_t *__restrict__ worklist, const uint32_t *__restrict__ wl_count,
                const salt_verify_out_t *__restrict__ rec, int rows,
                char *__restrict__ cigars, int stride, int8_t *__restrict__ out)
{
    SALT_DYN_SMEM(uint32_t, smem);
    constexpr int G = 32, ND = 32 * DPL, C = ND / 2;
    const int TW = lv_tw((int)c.l_max), PW = lv_pw((int)c.l_max);
    const int wl = threadIdx.x / 32, lane = threadIdx.x % 32;
    const int warps = blockDim.x / 32;
    const size_t per_warp = lv_cigar_tab_words(rows, ND) + TW + PW;
    uint32_t *base = smem + (size_t)wl * per_warp;
    int16_t *tabL = reinterpret_cast<int16_t *>(base);                  // [rows][ND]
    char *tabA = reinterpret_cast<char *>(tabL + (size_t)rows * ND);    // [rows][ND]
    uint32_t *T = base + lv_cigar_tab_words(rows, ND);
    uint32_t *P = T + TW;
    const unsigned gmask = 0xffffffffu;
    const size_t groups = (size_t)gridDim.x * warps;
    const size_t count = worklist ? (size_t)*wl_count : n;

    for (size_t it = (size_t)blockIdx.x * warps + wl; it < count; it += groups) {
        salt_pair_t p; int k; size_t slot;
        if (worklist) {
            const uint32_t rid = worklist[it];
            const salt_verify_out_t r = rec[rid];
            p.rs = (rid << 1) | (r.strand & 1); p.pos = r.pos; k = r.n_diff; slot = it;   // compact: slot = list index
        } else {
            p = pairs[it]; k = k_each[it]; slot = it;
        }
        char *buf = cigars + slot * (size_t)stride;
        const uint32_t rid = p.rs >> 1;
        const int plen = rid < c.n_reads ? (int)c.rd_len[rid] : 0;
        const int tlen = plen + 4;
        int result = -1;
        const bool ok = plen > 0 && k < LV_MAXK && k < rows && k < C && (uint64_t)p.pos + (uint64_t)tlen <= (uint64_t)c.l;
        if (ok) {
            lv_stage<G>(c, p.rs, p.pos, plen, tlen, T, P, TW, PW, lane);
            __syncwarp();
            const int L0 = lv_level0<G>(T, P, plen, tlen, lane, gmask, 0);
            if (L0 == plen) {
                if (lane == 0) { CigarOut o{buf, stride}; result = o.put(plen, 'M') ? 0 : -2; }
            } else {
                int Lp[DPL];
#pragma unroll
                for (int q = 0; q < DPL; ++q) {
                    const int d = lane * DPL + q - C;
                    Lp[q] = d == 0 ? L0 : -2;
                    tabL[d + C] = (int16_t)Lp[q];
                }
                int found_e = -1, found_d = 0;
                for (int e = 1; e <= k; ++e) {
                    int lft = __shfl_up_sync(gmask, Lp[DPL - 1], 1);
                    int rgt = __shfl_down_sync(gmask, Lp[0], 1);
                    if (lane == 0) lft = -2;
                    if (lane == G - 1) rgt = -2;
                    int Ln[DPL];
                    int myrank = 1 << 20;
#pragma unroll
                    for (int q = 0; q < DPL; ++q) {
                        const int d = lane * DPL + q - C;
                        const int left = q == 0 ? lft : Lp[q - 1];
                        const int right = (q == DPL - 1 ? rgt : Lp[q + 1]) + 1;
                        int best = Lp[q] + 1; char a = 'X';                // LandauVishkin.c:249-260
                        if (left > best) { best = left; a = 'D'; }
                        if (right > best) { best = right; a = 'I'; }
                        const bool act = d >= -e && d <= e;
                        best = lv_extend_group<G>(T, P, best, d, act, plen, tlen, lane, gmask, 0);
                        if (act) {
                            if (best == plen) myrank = imin(myrank, lv_cigar_rank(d));
                            Ln[q] = best;
                            tabA[e * ND + d + C] = a;
                        } else {
                            Ln[q] = -2;
                        }
                        tabL[e * ND + d + C] = (int16_t)Ln[q];
                    }
                    int r = myrank;
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) r = imin(r, __shfl_xor_sync(gmask, r, o));
                    if (r < (1 << 20)) {
                        found_e = e;
                        found_d = r == 0 ? 0 : ((r & 1) ? -(r + 1) / 2 : r / 2);
                        break;
                    }
#pragma unroll
                    for (int q = 0; q < DPL; ++q) Lp[q] = Ln[q];
                }
                __syncwarp();
                if (lane == 0) {
                    if (found_e < 0) { if (stride > 0) buf[0] = '\0'; result = -1; }
                    else result = lv_cigar_emit(tabL, tabA, ND, found_e, found_d, buf, stride);
                }
            }
            __syncwarp();
        } else if (lane == 0 && stride > 0 && plen > 0 && k < LV_MAXK) {
            buf[0] = '\0';
        }
        if (lane == 0 && out) out[slot] = (int8_t)result;
    }
}

// CTA-local counting sort of one tile of thread-per-pair work by a small difficulty key (0..31): thread t
// gets back the tile index of the item it should process, so that the 32 items of a warp have nearly the
// same key and the warp's lanes stay busy for the same number of levels (sorting the whole bench list by
// its result e makes lv_tpp 1.9x faster, tools/lvsort.py).  `area` = 64 + blockDim.x/2 words of shared
// memory; every thread of the CTA must call (two barriers inside, the caller owns the one before reuse).
__device__ __forceinline__ int cta_sort_by_key(uint32_t *area, int key)
{
    uint32_t *hist = area, *start = area + 32;
    uint16_t *perm = reinterpret_cast<uint16_t *>(area + 64);
    const int t = (int)threadIdx.x;
    if (t < 32) hist[t] = 0u;
    __syncthreads();
    const uint32_t rk = atomicAdd(&hist[key], 1u);
    __syncthreads();
    if (t < 32) {
        const uint32_t v = hist[t];
        uint32_t inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t up = __shfl_up_sync(0xffffffffu, inc, o);
            if (t >= o) inc += up;
        }
        start[t] = inc - v;
    }
    __syncthreads();
    perm[start[key] + rk] = (uint16_t)t;
    __syncthreads();
    return (int)perm[t];
}
__host__ __device__ constexpr int cta_sort_words(int threads) { return 64 + threads / 2; }

// --------------------------------------------------------------------------------------
// lv_cigar_tpp: computeEditDistanceWithCigar (LandauVishkin.c:200-462) with one THREAD per pair, for the
// case the verify stage produces: many pairs whose k is the small number of differences already found.
// Furthest-reaching values of the current level live in registers as in lv_tpp; the history the
// backtrace needs is a per-thread triangular table in shared memory (level e keeps diagonals -e..e, entry
// e*e + d + e: (K+1)^2 int16 values and as many action bytes).  Every lane emits its own string.
// Same staging, deferred long extensions and bank layout as lv_tpp.
// --------------------------------------------------------------------------------------
__host__ __device__ constexpr int lv_cigar_tpp_tab_words(int K, int lbytes) { return ((K + 1) * (K + 1) * (lbytes + 1) + 3) / 4; }

template <int K, class LT>
__global__ void __launch_bounds__(128)
lv_cigar_tpp_kernel(DevCtx c, const salt_pair_t *__restrict__ pairs, const uint8_t *__restrict__ k_each, size_t n,
                    const uint32_t *__restrict__ worklist, const uint32_t *__restrict__ wl_count,
                    const salt_verify_out_t *__restrict__ rec,
                    char *__restrict__ cigars, int stride, int8_t *__restrict__ out)
{
    SALT_DYN_SMEM(uint32_t, smem);
    const int TW = lv_twr((int)c.l_max), PW = lv_pwt((int)c.l_max);
    constexpr int TABW = lv_cigar_tpp_tab_words(K, (int)sizeof(LT));
    const int rstride = (TW + PW + TABW) | 1;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t *sort_area = smem;
    uint32_t *T = smem + cta_sort_words((int)blockDim.x) + ((size_t)warp * 32 + lane) * rstride;
    uint32_t *P = T + TW;
    LT *tabL = reinterpret_cast<LT *>(P + PW);                         // [(K+1)^2]: uint8 while reads are <= 253 bases, else int16
    char *tabA = reinterpret_cast<char *>(tabL + (K + 1) * (K + 1));   // [(K+1)^2]
    const LvTriIdx at;
    const size_t count = worklist ? (size_t)*wl_count : n;
    const size_t step = (size_t)gridDim.x * blockDim.x;

    for (size_t tile = (size_t)blockIdx.x * blockDim.x; tile < count; tile += step) {      // CTA-uniform
        // the tile's items are dealt to the threads in order of their k, so a warp runs as many levels as its lanes need
        int k0 = 31;
        {
            const size_t i0 = tile + threadIdx.x;
            if (i0 < count) k0 = imin(worklist ? (int)rec[worklist[i0]].n_diff : (int)k_each[i0], 30);
        }
        const size_t it = tile + (size_t)cta_sort_by_key(sort_area, k0);
        const bool live = it < count;
        salt_pair_t p; p.rs = 0; p.pos = 0; int k = 0;
        const size_t slot = it;
        if (live) {
            if (worklist) {
                const uint32_t rid = worklist[it];
                const salt_verify_out_t r = rec[rid];
                p.rs = (rid << 1) | (r.strand & 1); p.pos = r.pos; k = r.n_diff;   // compact: slot = list index
            } else {
                p = pairs[it]; k = k_each[it];
            }
        }
        char *buf = cigars + slot * (size_t)stride;
        const uint32_t rid = p.rs >> 1;
        const int plen = (live && rid < c.n_reads) ? (int)c.rd_len[rid] : 0;
        const int tlen = plen + 4;
        const bool ok = plen > 0 && k <= K && (uint64_t)p.pos + (uint64_t)tlen <= (uint64_t)c.l;
        const int toff = ok ? lv_stage_own(c, p, tlen, T, P, TW, PW) : 0;
        __syncwarp();
        int result = -1;
        if (ok) {
            const int L0 = lv_extend0(T, P, plen, tlen, toff);
            if (L0 == plen) {
                CigarOut o{buf, stride};
                result = o.put(plen, 'M') ? 0 : -2;
            } else {
                int Lp[2 * K + 1];
#pragma unroll
                for (int i = 0; i < 2 * K + 1; ++i) Lp[i] = -2;
                Lp[K] = L0;
                tabL[0] = (LT)L0;
                int found_e = -1, found_d = 0;
                for (int e = 1; e <= k; ++e) {
                    int Ln[2 * K + 1];
                    int myrank = 1 << 20;
                    int pend_di = -1, pend_best = 0;
#pragma unroll
                    for (int di = 0; di < 2 * K + 1; ++di) {
                        const int d = di - K;
                        int v = -2;
                        if (d >= -e && d <= e) {
                            const int left = di > 0 ? Lp[di > 0 ? di - 1 : 0] : -2;
                            const int right = di < 2 * K ? Lp[di < 2 * K ? di + 1 : 0] + 1 : -1;
                            int best = Lp[di] + 1; char a = 'X';                // LandauVishkin.c:249-260
                            if (left > best) { best = left; a = 'D'; }
                            if (right > best) { best = right; a = 'I'; }
                            bool more;
                            v = lv_extend_first(T, P, best, d, plen, tlen, more, toff);
                            if (more) {
                                if (pend_di < 0) { pend_di = di; pend_best = v; }
                                else v = lv_extend_more(T, P, v, d, plen, tlen, toff);
                            }
                            if (v == plen) myrank = imin(myrank, lv_cigar_rank(d));
                            tabA[at(e, d)] = a;
                            tabL[at(e, d)] = (LT)v;
                        }
                        Ln[di] = v;
                    }
                    int pend_v = 0;
                    if (pend_di >= 0) {
                        const int d = pend_di - K;
                        pend_v = lv_extend_more(T, P, pend_best, d, plen, tlen, toff);
                        if (pend_v == plen) myrank = imin(myrank, lv_cigar_rank(d));
                        tabL[at(e, d)] = (LT)pend_v;
                    }
                    if (myrank < (1 << 20)) {
                        found_e = e;
                        found_d = myrank == 0 ? 0 : ((myrank & 1) ? -(myrank + 1) / 2 : myrank / 2);
                        break;
                    }
#pragma unroll
                    for (int i = 0; i < 2 * K + 1; ++i) Lp[i] = i == pend_di ? pend_v : Ln[i];
                }
                if (found_e < 0) { if (stride > 0) buf[0] = '\0'; result = -1; }
                else result = lv_cigar_emit_t(tabL, tabA, at, found_e, found_d, buf, stride);
            }
        } else if (live && stride > 0 && plen > 0 && k < LV_MAXK) {
            buf[0] = '\0';
        }
        if (live && out) out[slot] = (int8_t)result;
        __syncwarp();
    }
}

// --------------------------------------------------------------------------------------
// nogap_fused: the whole ungapped stage of one read in one group of G lanes
// (alnse_check_nogap on strand 0 then strand 1, alnse.c:734-782 + :1079-1083).
//   counting  lane t keeps 64-bit words t, t+G, .. of the packed read (16 bases each) in
//             registers while the read's candidate list is walked.  A candidate costs one
//             coalesced, 8-byte-aligned window load per lane (G*8 contiguous bytes per candidate:
//             one or two L1 wavefronts); the next lane's word arrives by shuffle and two funnel
//             shifts bring the window to the read's nibble phase.  A read base is one-hot, so
//             popc(window & read) counts matches directly; reads containing N (nibble 15) take
//             the nibble-OR path (group-uniform choice).  Four windows are in flight together and
//             their per-lane counts travel through ONE butterfly, packed PK per register.
//             G and WPL are chosen so that 16*G*WPL >= l_max + 16: the last lane never needs a
//             word from beyond the group.
//   accepting candidates are taken G at a time, lane j owning candidate j.  The reference's
//             running threshold (code_kmismatch, alnse.c:348-370) is "accept n_j iff
//             n_j <= min(threshold so far, min of earlier valid n)", i.e. an exclusive prefix
//             minimum over the lanes; the primary is the first candidate reaching the stage's
//             minimum, and the first accepted hit of a stage always replaces the primary
//             (flag_match is local to alnse_check_nogap).  Candidate lists must be sorted
//             ascending (alnse_locate sorts them), so "pos == previous unskipped pos"
//             (alnse.c:762) is a comparison with the neighbouring lane.
//   reads with no ungapped hit append their candidates to the Landau-Vishkin worklist.
// --------------------------------------------------------------------------------------
// State of one read's ungapped stage, carried across strands and candidate chunks.
struct NogapState {
    uint32_t prim_pos; int prim_n, prim_strand, max_diff; bool any;
};

// Window loads of four candidates (lanes jb+J .. jb+J+3 of the group own them).  The owner lane
// has already turned its candidate into (aligned 64-bit word index, nibble phase) -- see
// nogap_strand -- so a lane only adds its own offset.  Loads are unconditional: a missing or
// out-of-range candidate was clamped to the last valid window start; its count is computed and
// then ignored by the acceptance step.
template <int G, int WPL, int J>
__device__ __forceinline__ void nogap_load4(const uint2 *__restrict__ mixl, uint32_t myidx, uint32_t myph, int jb, int wlim,
                                            uint32_t (&ph)[4], uint2 (&q)[4][WPL])
{
    constexpr unsigned FULL = 0xffffffffu;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
        const uint32_t idx = __shfl_sync(FULL, myidx, jb + J + u, G);
        ph[u] = __shfl_sync(FULL, myph, jb + J + u, G);
        const uint2 *__restrict__ wp = mixl + idx;
#pragma unroll
        for (int w = 0; w < WPL; ++w) {
            // a window of L bases starting at most 15 bases into its aligned word ends in word (L + 14) / 16: lanes whose
            // word lies beyond that do not touch memory (their read words are zero anyway).  With the reference in DRAM
            // (GRCh38-sized) the kernel is bound by the sectors it fetches, so the unused tail of the span is time.
            if (w == 0 && WPL == 1 && G <= 8) q[u][w] = wp[0];
            else q[u][w] = (w * G < wlim) ? wp[w * G] : make_uint2(0u, 0u);
        }
    }
}

// Match counts of four candidates, summed over the group; lane jb+J+u returns candidate u's count
// in `mine`.  rw = the lane's read words, rwh = the same read moved up by 8 bases (used when the
// window starts in the upper half of its aligned 64-bit word, so one funnel shift always suffices).
template <int G, int WPL, int J, bool HASN>
__device__ __forceinline__ void nogap_count4(const uint32_t (&ph)[4], const uint2 (&q)[4][WPL], const uint2 (&rw)[WPL],
                                             const uint2 (&rwh)[WPL], int lane, int jb, int &mine)
{
    constexpr int PK = (16 * G * WPL <= 256) ? 4 : 2;   // counts per register while they fit their field
    constexpr int PB = 32 / PK;
    constexpr uint32_t PM = PK == 4 ? 0xffu : 0xffffu;
    constexpr unsigned FULL = 0xffffffffu;
    uint32_t packed[4 / PK];
#pragma unroll
    for (int k = 0; k < 4 / PK; ++k) packed[k] = 0u;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
        const int sh = (int)ph[u];                              // funnel shifts use the low 5 bits: 4*(pos & 7)
        const bool hi = (int)ph[u] < 0;                         // bit 31: window starts in the upper 32-bit half
        uint32_t m = 0;
#pragma unroll
        for (int w = 0; w < WPL; ++w) {
            uint32_t n0 = __shfl_down_sync(FULL, q[u][w].x, 1, G);
            if (w + 1 < WPL) {                                  // the last lane's neighbour is lane 0's next word
                const uint32_t w0 = __shfl_sync(FULL, q[u][w + 1 < WPL ? w + 1 : w].x, 0, G);
                if (lane == G - 1) n0 = w0;
            }
            uint32_t x0 = __funnelshift_r(q[u][w].x, q[u][w].y, sh) & (hi ? rwh[w].x : rw[w].x);
            uint32_t x1 = __funnelshift_r(q[u][w].y, n0, sh) & (hi ? rwh[w].y : rw[w].y);
            if (HASN) {
                x0 |= x0 >> 1; x0 |= x0 >> 2; x0 &= 0x11111111u;
                x1 |= x1 >> 1; x1 |= x1 >> 2; x1 &= 0x11111111u;
            }
            m += (uint32_t)(__popc(x0) + __popc(x1));
        }
        packed[u / PK] += m << (PB * (u % PK));
    }
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1)
#pragma unroll
        for (int k = 0; k < 4 / PK; ++k) packed[k] += __shfl_xor_sync(FULL, packed[k], o, G);
    const int sel = lane - jb - J;                              // lane jb+J+u keeps candidate u's count
    const uint32_t pv = (PK == 4 || sel < 2) ? packed[0] : packed[4 / PK - 1];
    if (sel >= 0 && sel < 4) mine = (int)((pv >> (PB * (sel % PK))) & PM);
}

// One strand of one read per group of G lanes; every loop bound is made warp-uniform (maximum
// over the warp's groups, idle groups predicated off) so that all shuffles run on the full
// warp mask without divergence checks.
template <int G, int WPL, bool HASN>
__device__ __forceinline__ int nogap_strand(const DevCtx &c, const uint2 *__restrict__ mixl, const uint2 (&rw)[WPL],
                                            const uint2 (&rwh)[WPL], const uint32_t *__restrict__ loci,
                                            int8_t *__restrict__ accs, uint32_t lbs, uint32_t les, uint32_t first,
                                            int L, int T0, bool fits, uint32_t lim, int s, int lane, int gshift,
                                            NogapState &st)
{
    constexpr unsigned FULL = 0xffffffffu;
    constexpr unsigned gfull = (unsigned)((1ull << G) - 1ull);
    constexpr int BIG = 255;
    const int nchunks = (int)((les - lbs + (G - 1)) / G);
    const int nchunks_w = __reduce_max_sync(FULL, nchunks);
    const int wlim = (L + 14) / 16 + 1 - lane;              // this lane loads word lane + w*G iff w*G < wlim
    bool matched = false;
    int nhits = 0;
    for (int ci = 0; ci < nchunks_w; ++ci) {
        const uint32_t base = lbs + (uint32_t)ci * G;
        const int cnt = ci < nchunks ? (int)min((uint32_t)G, les - base) : 0;
        const uint32_t truepos = ci == 0 ? first : (lane < cnt ? loci[base + lane] : 0xFFFFFFFFu);
        // the element before this lane's candidate in the (sorted) list: alnse.c:762 skips repeats
        const uint32_t prev = (lane < cnt && base + lane > lbs) ? loci[base + lane - 1] : 0xFFFFFFFFu;
        // a read longer than the reference counts nothing
        const uint32_t mypos = fits ? truepos : 0xFFFFFFFFu;
        const uint32_t cpos = min(mypos, lim);                  // what the lanes load for this candidate
        const uint32_t myidx = cpos >> 4;
        const uint32_t myph = ((cpos & 7u) << 2) | ((cpos & 8u) << 28);
        const int cnt_w = __reduce_max_sync(FULL, cnt);
        int mymatches = 0;
        // ---- counting: eight candidates in flight, all lanes of a group cooperate on each window
        for (int jb = 0; jb < (G == 8 ? 1 : cnt_w); jb += 8) {
            const int jbb = G == 8 ? 0 : jb;
            const bool more = cnt_w > jbb + 4;
            uint32_t pa[4], pb[4];
            uint2 qa[4][WPL], qb[4][WPL];
            nogap_load4<G, WPL, 0>(mixl, myidx, myph, jbb, wlim, pa, qa);
            if (more) nogap_load4<G, WPL, 4>(mixl, myidx, myph, jbb, wlim, pb, qb);
            nogap_count4<G, WPL, 0, HASN>(pa, qa, rw, rwh, lane, jbb, mymatches);
            if (more) nogap_count4<G, WPL, 4, HASN>(pb, qb, rw, rwh, lane, jbb, mymatches);
        }
        // ---- accepting: lane j decides candidate j
        const bool skip = lane >= cnt || truepos == prev || truepos >= c.l;          // alnse.c:762
        const int nmis = L - mymatches;
        const int key = (!skip && mypos <= lim && nmis <= T0) ? nmis : BIG;
        int pm = key;                                                                // inclusive prefix minimum
#pragma unroll
        for (int o = 1; o < G; o <<= 1) pm = imin(pm, __shfl_up_sync(FULL, pm, o, G));   // lanes < o get themselves back
        int ex = __shfl_up_sync(FULL, pm, 1, G);                                     // exclusive
        if (lane == 0) ex = BIG;
        const bool accepted = key != BIG && key <= imin(st.max_diff, ex);
        const unsigned bal = (__ballot_sync(FULL, accepted) >> gshift) & gfull;
        if (lane < cnt) accs[base + lane] = (int8_t)(accepted ? key : -1);
        const int cmin = __shfl_sync(FULL, pm, G - 1, G);                            // minimum over the chunk's valid keys
        // first accepted candidate that reaches the chunk minimum becomes the primary
        const unsigned at = (__ballot_sync(FULL, accepted && key == cmin) >> gshift) & gfull;
        const uint32_t cand_pos = __shfl_sync(FULL, truepos, at ? __ffs((int)at) - 1 : 0, G);
        if (bal) {
            if (cmin < st.max_diff || !matched) { st.prim_pos = cand_pos; st.prim_n = cmin; st.prim_strand = s; }
            st.max_diff = imin(st.max_diff, cmin);
            matched = true;
            nhits += __popc(bal);
        }
    }
    st.any = st.any || matched;
    return nhits;
}

template <int G, int WPL, bool HASN>
__device__ __forceinline__ void nogap_read(const DevCtx &c, const uint2 *__restrict__ mixl, const uint2 (&rw0)[WPL],
                                           const uint2 (&rw1)[WPL], const uint32_t *__restrict__ loci0,
                                           const uint32_t *__restrict__ loci1, int8_t *__restrict__ acc, size_t n0,
                                           uint32_t lb0, uint32_t le0, uint32_t lb1, uint32_t le1, uint32_t first0,
                                           uint32_t first1, int L, int T0, bool fits, uint32_t lim, int lane, int gshift,
                                           NogapState &st, int &hits0, int &hits1)
{
    constexpr unsigned FULL = 0xffffffffu;
    uint2 rwh[WPL];
#pragma unroll
    for (int w = 0; w < WPL; ++w) {                     // the read moved up by one 32-bit word (8 bases)
        uint32_t up = __shfl_up_sync(FULL, rw0[w].y, 1, G);
        if (lane == 0) up = 0u;
        if (w > 0) { const uint32_t wrap = __shfl_sync(FULL, rw0[w > 0 ? w - 1 : 0].y, G - 1, G); if (lane == 0) up = wrap; }
        rwh[w] = make_uint2(up, rw0[w].x);
    }
    hits0 = nogap_strand<G, WPL, HASN>(c, mixl, rw0, rwh, loci0, acc, lb0, le0, first0, L, T0, fits, lim, 0, lane, gshift, st);
#pragma unroll
    for (int w = 0; w < WPL; ++w) {
        uint32_t up = __shfl_up_sync(FULL, rw1[w].y, 1, G);
        if (lane == 0) up = 0u;
        if (w > 0) { const uint32_t wrap = __shfl_sync(FULL, rw1[w > 0 ? w - 1 : 0].y, G - 1, G); if (lane == 0) up = wrap; }
        rwh[w] = make_uint2(up, rw1[w].x);
    }
    hits1 = nogap_strand<G, WPL, HASN>(c, mixl, rw1, rwh, loci1, acc + n0, lb1, le1, first1, L, T0, fits, lim, 1, lane, gshift, st);
}

template <int G, int WPL>
__global__ void __launch_bounds__(64, (G == 8 && WPL == 1) ? 18 : (WPL == 2 && G <= 16) ? 12 : 4)
nogap_fused_kernel(DevCtx c, const uint32_t *__restrict__ offs0, const uint32_t *__restrict__ loci0,
                   const uint32_t *__restrict__ offs1, const uint32_t *__restrict__ loci1, size_t n0,
                   int T0, int8_t *__restrict__ acc, salt_verify_out_t *__restrict__ rec,
                   salt_pair_t *__restrict__ lv_pairs, uint32_t *__restrict__ lv_slots, uint32_t *__restrict__ lv_count,
                   uint32_t *__restrict__ lv_reads)
{
    constexpr unsigned FULL = 0xffffffffu;
    const size_t rr = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) / G;
    const bool live = rr < (size_t)c.n_reads;           // a dead group runs along with empty lists
    const uint32_t r = live ? (uint32_t)rr : 0u;
    const int lane = threadIdx.x % G;
    const int gshift = (threadIdx.x & 31) / G * G;
    uint32_t lb0 = 0, le0 = 0, lb1 = 0, le1 = 0;
    if (live) { lb0 = offs0[r]; le0 = offs0[r + 1]; lb1 = offs1[r]; le1 = offs1[r + 1]; }
    const int L = live ? (int)c.rd_len[r] : 0;
    // both strands' read words and first candidates are requested before any of them is used
    uint2 rw0[WPL], rw1[WPL];
    const uint2 *__restrict__ rrow = reinterpret_cast<const uint2 *>(c.rd4 + (size_t)r * 2 * c.W64);
#pragma unroll
    for (int w = 0; w < WPL; ++w) {
        const bool in = live && (uint32_t)(lane + w * G) < c.W64;
        rw0[w] = in ? rrow[lane + w * G] : make_uint2(0u, 0u);
        rw1[w] = in ? rrow[c.W64 + lane + w * G] : make_uint2(0u, 0u);
    }
    const uint32_t first0 = lb0 + lane < le0 ? loci0[lb0 + lane] : 0xFFFFFFFFu;
    const uint32_t first1 = lb1 + lane < le1 ? loci1[lb1 + lane] : 0xFFFFFFFFu;
    const uint2 *__restrict__ mixl = reinterpret_cast<const uint2 *>(c.mixref) + lane;
    const bool fits = c.l >= (uint32_t)L && L > 0;
    const uint32_t lim = fits ? c.l - (uint32_t)L : 0u; // pos + L <= l  <=>  pos <= lim

    NogapState st;
    st.prim_pos = 0xFFFFFFFFu; st.prim_n = 255; st.prim_strand = 3; st.max_diff = T0; st.any = false;
    int hits0 = 0, hits1 = 0;
    bool hasN = false;                                  // N is its own complement: one test serves both strands
#pragma unroll
    for (int w = 0; w < WPL; ++w)
        hasN = hasN || ((rw0[w].x & (rw0[w].x >> 1) & 0x11111111u) != 0u) || ((rw0[w].y & (rw0[w].y >> 1) & 0x11111111u) != 0u);
    if (__any_sync(FULL, hasN))          // warp-uniform: the general path is exact for every read
        nogap_read<G, WPL, true>(c, mixl, rw0, rw1, loci0, loci1, acc, n0, lb0, le0, lb1, le1, first0, first1, L, T0, fits, lim,
                                 lane, gshift, st, hits0, hits1);
    else
        nogap_read<G, WPL, false>(c, mixl, rw0, rw1, loci0, loci1, acc, n0, lb0, le0, lb1, le1, first0, first1, L, T0, fits, lim,
                                  lane, gshift, st, hits0, hits1);

    const bool need_lv = live && !st.any;                // alnse.c:1022 / :1089: gapped stage for this read
    const uint32_t c0 = le0 - lb0, c1 = le1 - lb1;
    uint32_t w = 0;
    if (need_lv && lane == 0) {
        if (c0 + c1) w = atomicAdd(lv_count, c0 + c1);
        lv_reads[atomicAdd(lv_count + 1, 1u)] = r;       // the reads scan_gap has to visit
    }
    w = __shfl_sync(FULL, w, 0, G);
    if (need_lv) {
        for (uint32_t i = lane; i < c0; i += G) {
            salt_pair_t p; p.rs = r << 1; p.pos = loci0[lb0 + i];
            lv_pairs[w + i] = p; lv_slots[w + i] = lb0 + i;
        }
        for (uint32_t i = lane; i < c1; i += G) {
            salt_pair_t p; p.rs = (r << 1) | 1u; p.pos = loci1[lb1 + i];
            lv_pairs[w + c0 + i] = p; lv_slots[w + c0 + i] = (uint32_t)n0 + lb1 + i;
        }
    }
    if (live && lane == 0) {
        uint4 o;                                         // salt_verify_out_t as one 16-byte store
        o.x = st.prim_pos;
        o.y = (uint32_t)(st.prim_strand & 255) | ((uint32_t)(st.prim_n & 255) << 8) |
              ((st.any ? 0u : 255u) << 16) | ((need_lv ? 1u : 0u) << 24);
        o.z = (uint32_t)hits0; o.w = (uint32_t)hits1;
        *reinterpret_cast<uint4 *>(rec + r) = o;
    }
}

// --------------------------------------------------------------------------------------
// Acceptance scan of the gapped stage.  The LV kernels return e iff e <= T0; the reference calls
// them with a threshold that tightens as candidates are accepted, which is equivalent to accepting
// candidate i iff e_i <= min(T0, min_{j<i} e_j) in list order, strand 0 then strand 1
// (code_kdiff, alnse.c:372-393).  Eight lanes per read that reached the stage walk its lists, eight
// candidates per step.
// --------------------------------------------------------------------------------------

// One strand of one read, eight lanes over eight candidates at a time: the same acceptance as nogap_strand
// (exclusive prefix minimum over the lanes, then the running threshold), fed with the LV results in acc.
__device__ __forceinline__ void scan_strand(const uint32_t *__restrict__ loci, uint32_t b, uint32_t e,
                                            int8_t *__restrict__ accs, uint32_t l_mref, uint32_t guard, int strand,
                                            int lane, int gshift, unsigned gmask, int &max_diff, salt_verify_out_t &q)
{
    constexpr int G = 8, BIG = 255;
    bool matched = false;                                           // per stage call (alnse.c:883)
    for (uint32_t base = b; base < e; base += G) {
        const uint32_t idx = base + (uint32_t)lane;
        const bool in = idx < e;
        const uint32_t pos = in ? loci[idx] : 0xFFFFFFFFu;
        const uint32_t prev = (in && idx > b) ? loci[idx - 1] : 0xFFFFFFFFu;
        // alnse.c:894: repeats of the previous locus and pos + l_seq + 4 >= l (uint32 arithmetic) are skipped
        const bool skip = !in || pos == prev || (uint32_t)(pos + guard) >= l_mref;
        const int nd = in ? (int)accs[idx] : -1;
        const int key = (!skip && nd >= 0) ? nd : BIG;
        int pm = key;                                               // inclusive prefix minimum over the group
#pragma unroll
        for (int o = 1; o < G; o <<= 1) {
            const int up = __shfl_up_sync(gmask, pm, o, G);
            if (lane >= o) pm = imin(pm, up);
        }
        int ex = __shfl_up_sync(gmask, pm, 1, G);
        if (lane == 0) ex = BIG;
        const bool accepted = key != BIG && key <= imin(max_diff, ex);
        const unsigned bal = (__ballot_sync(gmask, accepted) >> gshift) & 0xffu;
        if (in) accs[idx] = (int8_t)(accepted ? key : -1);
        const int cmin = __shfl_sync(gmask, pm, G - 1, G);
        const unsigned at = (__ballot_sync(gmask, accepted && key == cmin) >> gshift) & 0xffu;
        const uint32_t cand_pos = __shfl_sync(gmask, pos, at ? __ffs((int)at) - 1 : 0, G);
        if (bal) {
            if (cmin < max_diff || !matched) {                      // first success, then every strict improvement
                q.is_gap = 1; q.n_diff = (uint8_t)cmin; q.strand = (uint8_t)strand; q.pos = cand_pos;
            }
            max_diff = imin(max_diff, cmin);
            matched = true;
            q.n_hits[strand] += __popc(bal);
        }
    }
}

__global__ void __launch_bounds__(128)
scan_gap_kernel(DevCtx c, const uint32_t *__restrict__ offs0, const uint32_t *__restrict__ loci0,
                const uint32_t *__restrict__ offs1, const uint32_t *__restrict__ loci1, size_t n0,
                int lv_T0, int8_t *__restrict__ acc, salt_verify_out_t *__restrict__ rec,
                const uint32_t *__restrict__ lv_reads, const uint32_t *__restrict__ lv_read_count,
                uint32_t *__restrict__ cig_list, uint32_t *__restrict__ cig_count)
{
    // eight lanes per read that reached the gapped stage (the list nogap_fused wrote)
    constexpr int G = 8;
    const uint32_t count = *lv_read_count;
    const int lane = threadIdx.x % G;
    const int gshift = (threadIdx.x & 31) / G * G;
    const unsigned gmask = 0xffu << gshift;
    const uint32_t groups = gridDim.x * blockDim.x / G;
    for (uint32_t i = (blockIdx.x * blockDim.x + threadIdx.x) / G; i < count; i += groups) {
        const uint32_t r = lv_reads[i];
        salt_verify_out_t q = rec[r];
        const int L = c.rd_len[r];
        int max_diff = lv_T0 >= 0 ? lv_T0 : L / 10;                  // alnse.c:1090 / :1027
        const uint32_t guard = (uint32_t)L + 4u;
        // strand 1 starts from the threshold strand 0 left (alnse.c:1092-1094)
        const uint32_t b0 = offs0[r], e0 = offs0[r + 1], b1 = offs1[r], e1 = offs1[r + 1];   // all four before the first use
        scan_strand(loci0, b0, e0, acc, c.l, guard, 0, lane, gshift, gmask, max_diff, q);
        scan_strand(loci1, b1, e1, acc + n0, c.l, guard, 1, lane, gshift, gmask, max_diff, q);
        if (lane == 0) {
            rec[r] = q;
            if (q.is_gap == 1 && cig_list) cig_list[atomicAdd(cig_count, 1u)] = r;
        }
    }
}

// --------------------------------------------------------------------------------------
// launchers
// --------------------------------------------------------------------------------------
#define SALT_LAUNCH_CHECK() do { cudaError_t e_ = cudaGetLastError(); if (e_ != cudaSuccess) return e_; } while (0)

cudaError_t launch_pack_reads(const uint8_t *codes, const uint32_t *offs, uint32_t n_reads, uint32_t W64,
                              uint64_t *rd4, uint16_t *rd_len, cudaStream_t st)
{
    const size_t total = (size_t)n_reads * 2 * 8;        // eight lanes per read-strand
    if (!total) return cudaSuccess;
    SALT_LAUNCH(pack_reads_kernel, (unsigned)((total + 255) / 256), 256, 0, st, codes, offs, n_reads, W64, rd4, rd_len);
    SALT_LAUNCH_CHECK();
    return cudaSuccess;
}

template <int G, int WPL>
static cudaError_t launch_mismatch_t(const DevCtx &c, const salt_pair_t *pairs, size_t n, int max_err, int8_t *out, cudaStream_t st)
{
    const size_t groups = (n + 3) / 4;
    size_t blocks = (groups * G + 255) / 256;
    const size_t cap = 148 * 64;                         // grid-stride beyond that
    if (blocks > cap) blocks = cap;
    auto kern = mismatch_kernel<G, WPL>;
    SALT_LAUNCH(kern, (unsigned)blocks, 256, 0, st, c, pairs, n, max_err, out);
    SALT_LAUNCH_CHECK();
    return cudaSuccess;
}

cudaError_t launch_mismatch(const DevCtx &c, const salt_pair_t *pairs, size_t n, int max_err, int8_t *out, cudaStream_t st)
{
    if (!n) return cudaSuccess;
    const int lm = (int)c.l_max;                        // l_max + 16 <= 16*G*WPL, as for nogap_fused
    if (lm <= 112) return launch_mismatch_t<8, 1>(c, pairs, n, max_err, out, st);
    if (lm <= 240) return launch_mismatch_t<8, 2>(c, pairs, n, max_err, out, st);
    if (lm <= 496) return launch_mismatch_t<16, 2>(c, pairs, n, max_err, out, st);
    if (lm <= 1008) return launch_mismatch_t<32, 2>(c, pairs, n, max_err, out, st);
    return launch_mismatch_t<32, 3>(c, pairs, n, max_err, out, st);
}

template <int G, int DPL>
static cudaError_t launch_lv_t(const DevCtx &c, const salt_pair_t *pairs, size_t n, int k,
                               const uint32_t *worklist, const uint32_t *wl_count, size_t wl_cap,
                               int8_t *out, int sm_count, cudaStream_t st)
{
    constexpr int threads = 128;
    const int gpc = threads / G;
    const size_t smem = (size_t)gpc * (lv_tw((int)c.l_max) + lv_pw((int)c.l_max)) * 4;
    const size_t items = worklist ? wl_cap : n;
    size_t blocks = (items + gpc - 1) / gpc;
    const size_t cap = (size_t)sm_count * 16;          // persistent upper bound: groups stride over the rest
    if (blocks > cap) blocks = cap;
    if (blocks == 0) return cudaSuccess;
    auto kern = lv_kernel<G, DPL>;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    SALT_LAUNCH(kern, (unsigned)blocks, threads, smem, st, c, pairs, n, k, worklist, wl_count, out);
    SALT_LAUNCH_CHECK();
    return cudaSuccess;
}

template <int K>
static cudaError_t launch_lv_tpp(const DevCtx &c, const salt_pair_t *pairs, size_t n, int k,
                                 const uint32_t *worklist, const uint32_t *wl_count, size_t wl_cap,
                                 int8_t *out, int sm_count, cudaStream_t st, LvDefer df = LvDefer{nullptr, nullptr, nullptr})
{
    const int stride = (lv_twr((int)c.l_max) + lv_pwt((int)c.l_max)) | 1;
    int threads = 128;
    size_t smem = (size_t)threads * stride * 4;
    if (smem > 100 * 1024) { threads = 64; smem = (size_t)threads * stride * 4; }
    const size_t items = worklist ? wl_cap : n;
    size_t blocks = (items + threads - 1) / threads;
    const size_t cap = (size_t)sm_count * 16;          // persistent upper bound: warps stride over the rest
    if (blocks > cap) blocks = cap;
    if (blocks == 0) return cudaSuccess;
    auto kern = lv_tpp_kernel<K>;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    SALT_LAUNCH(kern, (unsigned)blocks, threads, smem, st, c, pairs, n, k, worklist, wl_count, out, df);
    SALT_LAUNCH_CHECK();
    return cudaSuccess;
}

// mapping: 0 = choose (thread-per-pair up to k = 15, warp-per-pair beyond), 1 = force warp-per-pair
static cudaError_t launch_lv_core(const DevCtx &c, const salt_pair_t *pairs, size_t n, int k,
                                  const uint32_t *worklist, const uint32_t *wl_count, size_t wl_cap,
                                  int8_t *out, int sm_count, cudaStream_t st, int mapping)
{
    int kmax = k >= 0 ? k : (int)c.l_max / 10;
    if (kmax > LV_MAXK - 1) kmax = LV_MAXK - 1;
    if (mapping != 1 && c.l_max <= 512) {
        if (kmax <= 3) return launch_lv_tpp<3>(c, pairs, n, k, worklist, wl_count, wl_cap, out, sm_count, st);
        if (kmax <= 8) return launch_lv_tpp<8>(c, pairs, n, k, worklist, wl_count, wl_cap, out, sm_count, st);
        if (kmax <= 10) return launch_lv_tpp<10>(c, pairs, n, k, worklist, wl_count, wl_cap, out, sm_count, st);
        if (kmax <= 15) return launch_lv_tpp<15>(c, pairs, n, k, worklist, wl_count, wl_cap, out, sm_count, st);
    }
    if (kmax <= 3) return launch_lv_t<8, 1>(c, pairs, n, k, worklist, wl_count, wl_cap, out, sm_count, st);
    if (kmax <= 7) return launch_lv_t<16, 1>(c, pairs, n, k, worklist, wl_count, wl_cap, out, sm_count, st);
    if (kmax <= 15) return launch_lv_t<32, 1>(c, pairs, n, k, worklist, wl_count, wl_cap, out, sm_count, st);
    return launch_lv_t<32, 2>(c, pairs, n, k, worklist, wl_count, wl_cap, out, sm_count, st);
}

// Landau-Vishkin over a pair list.  With a filter scratch (room for the whole list) the pigeonhole
// filter runs first and only its survivors reach the LV kernels; results are identical either way.
cudaError_t launch_lv(const DevCtx &c, const salt_pair_t *pairs, size_t n, int k,
                      const uint32_t *worklist, const uint32_t *wl_count, size_t wl_cap,
                      int8_t *out, int sm_count, cudaStream_t st, int mapping, const LvFilterScratch *f)
{
    const size_t items = worklist ? wl_cap : n;
    if (!items) return cudaSuccess;
    if (!f || !f->pairs || !f->slots || !f->count)
        return launch_lv_core(c, pairs, n, k, worklist, wl_count, wl_cap, out, sm_count, st, mapping);
    cudaError_t e = cudaMemsetAsync(f->count, 0, 4, st);
    if (e != cudaSuccess) return e;
    const int threads = 128;
    size_t blocks = (items + threads - 1) / threads;
    const size_t cap = (size_t)sm_count * 32;            // persistent upper bound: warps stride over the rest
    if (blocks > cap) blocks = cap;
    SALT_LAUNCH(lv_filter_kernel, (unsigned)blocks, threads, 0, st, c, pairs, n, k, worklist, wl_count, out,
                f->pairs, f->slots, f->count);
    SALT_LAUNCH_CHECK();
    // Survivors are mostly true hits and their +-1..3 shifted twins, which stop after a few levels:
    // one thread per pair (long extensions deferred to the end of a level) beats one warp per pair on
    // both the verify stage's worklists and flat decoy-heavy lists (profiles/r1r_lv_split.txt).
    int kmax = k >= 0 ? k : (int)c.l_max / 10;
    if (kmax > LV_MAXK - 1) kmax = LV_MAXK - 1;
    // (measured: 0.76 -> 0.67 ms on a flat 7.8 M-pair list at k = 10, but 0.166 -> 0.181 ms on the verify stage's 1.2 M-pair
    // worklist, where the second launch and the re-staging cost more than the idle lanes: large lists only)
    if (mapping != 1 && c.l_max <= 512 && kmax > 3 && kmax <= 15 && !worklist && n >= 3000000 && f->pairs2 && f->slots2 && f->count2) {
        // two passes: three levels for everybody (most survivors end there), the full depth only for the rest -- the
        // lanes of a warp then finish within a few levels of each other instead of waiting for the one deep pair
        if ((e = cudaMemsetAsync(f->count2, 0, 4, st)) != cudaSuccess) return e;
        LvDefer df{f->pairs2, f->slots2, f->count2};
        if ((e = launch_lv_tpp<3>(c, f->pairs, items, k, f->slots, f->count, items, out, sm_count, st, df)) != cudaSuccess) return e;
        return launch_lv_core(c, f->pairs2, items, k, f->slots2, f->count2, items, out, sm_count, st, mapping);
    }
    return launch_lv_core(c, f->pairs, items, k, f->slots, f->count, items, out, sm_count, st, mapping);
}

template <int DPL>
static cudaError_t launch_lv_cigar_t(const DevCtx &c, const salt_pair_t *pairs, const uint8_t *k_each, size_t n,
                                     const uint32_t *worklist, const uint32_t *wl_count, size_t wl_cap,
                                     const salt_verify_out_t *rec, int rows, char *cigars, int stride, int8_t *out,
                                     int sm_count, cudaStream_t st)
{
    constexpr int threads = 128, warps = threads / 32;
    const size_t per_warp = (lv_cigar_tab_words(rows, 32 * DPL) + lv_tw((int)c.l_max) + lv_pw((int)c.l_max)) * 4;
    const size_t smem = per_warp * warps;
    const size_t items = worklist ? wl_cap : n;
    size_t blocks = (items + warps - 1) / warps;
    const size_t cap = (size_t)sm_count * 16;
    if (blocks > cap) blocks = cap;
    if (blocks == 0) return cudaSuccess;
    auto kern = lv_cigar_kernel<DPL>;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    SALT_LAUNCH(kern, (unsigned)blocks, threads, smem, st, c, pairs, k_each, n, worklist, wl_count, rec, rows, cigars, stride, out);
    SALT_LAUNCH_CHECK();
    return cudaSuccess;
}

template <int K, class LT>
static cudaError_t launch_lv_cigar_tpp_t(const DevCtx &c, const salt_pair_t *pairs, const uint8_t *k_each, size_t n,
                                         const uint32_t *worklist, const uint32_t *wl_count, size_t wl_cap,
                                         const salt_verify_out_t *rec, char *cigars, int stride, int8_t *out,
                                         int sm_count, cudaStream_t st)
{
    const int rstride = (lv_twr((int)c.l_max) + lv_pwt((int)c.l_max) + lv_cigar_tpp_tab_words(K, (int)sizeof(LT))) | 1;
    int threads = 128;
    size_t smem = ((size_t)threads * rstride + cta_sort_words(threads)) * 4;
    if (smem > 73 * 1024) { threads = 64; smem = ((size_t)threads * rstride + cta_sort_words(threads)) * 4; }
    const size_t items = worklist ? wl_cap : n;
    size_t blocks = (items + threads - 1) / threads;
    const size_t cap = (size_t)sm_count * 16;
    if (blocks > cap) blocks = cap;
    if (blocks == 0) return cudaSuccess;
    auto kern = lv_cigar_tpp_kernel<K, LT>;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    SALT_LAUNCH(kern, (unsigned)blocks, threads, smem, st, c, pairs, k_each, n, worklist, wl_count, rec, cigars, stride, out);
    SALT_LAUNCH_CHECK();
    return cudaSuccess;
}

// furthest-reaching values fit a byte while reads are at most 253 bases: a third less shared memory per thread, 4 CTAs per
// SM instead of 3 at K = 10 -- the kernel is latency-bound, occupancy is what it lacks
template <int K>
static cudaError_t launch_lv_cigar_tpp(const DevCtx &c, const salt_pair_t *pairs, const uint8_t *k_each, size_t n,
                                       const uint32_t *worklist, const uint32_t *wl_count, size_t wl_cap,
                                       const salt_verify_out_t *rec, char *cigars, int stride, int8_t *out,
                                       int sm_count, cudaStream_t st)
{
    if (c.l_max <= 253)
        return launch_lv_cigar_tpp_t<K, uint8_t>(c, pairs, k_each, n, worklist, wl_count, wl_cap, rec, cigars, stride, out, sm_count, st);
    return launch_lv_cigar_tpp_t<K, int16_t>(c, pairs, k_each, n, worklist, wl_count, wl_cap, rec, cigars, stride, out, sm_count, st);
}

// kmax: upper bound of the k the items carry (levels kept in shared memory = kmax + 1; up to 15
// differences fit 32 diagonals, one per lane)
cudaError_t launch_lv_cigar(const DevCtx &c, const salt_pair_t *pairs, const uint8_t *k_each, size_t n,
                            const uint32_t *worklist, const uint32_t *wl_count, size_t wl_cap,
                            const salt_verify_out_t *rec, int kmax, char *cigars, int stride, int8_t *out,
                            int sm_count, cudaStream_t st, int mapping)
{
    if (kmax < 0) kmax = 0;
    if (kmax > LV_MAXK - 1) kmax = LV_MAXK - 1;
    // One thread per pair issues a quarter of the instructions but is a long serial chain per warp: it pays
    // once there are enough pairs to fill the machine with such warps (2 M-read chunk, 75 k gapped primaries:
    // 88 us against 112 us), while a 100 k-read chunk's ~4 k pairs finish sooner spread one per warp
    // (23 us against 46 us; profiles/r1r_launches_bench_summary.txt).  A verify-stage worklist holds the
    // gapped primaries, a few percent of the chunk's reads.
    const size_t expect = worklist ? wl_cap / 25 : n;
    if (mapping == 0 && expect < 32768) mapping = 1;
    if (mapping != 1 && c.l_max <= 512) {               // thread per pair: see lv_cigar_tpp_kernel
        if (kmax <= 4) return launch_lv_cigar_tpp<4>(c, pairs, k_each, n, worklist, wl_count, wl_cap, rec, cigars, stride, out, sm_count, st);
        if (kmax <= 10) return launch_lv_cigar_tpp<10>(c, pairs, k_each, n, worklist, wl_count, wl_cap, rec, cigars, stride, out, sm_count, st);
        if (kmax <= 15) return launch_lv_cigar_tpp<15>(c, pairs, k_each, n, worklist, wl_count, wl_cap, rec, cigars, stride, out, sm_count, st);
    }
    if (kmax <= 15)
        return launch_lv_cigar_t<1>(c, pairs, k_each, n, worklist, wl_count, wl_cap, rec, kmax + 1, cigars, stride, out, sm_count, st);
    return launch_lv_cigar_t<2>(c, pairs, k_each, n, worklist, wl_count, wl_cap, rec, kmax + 1, cigars, stride, out, sm_count, st);
}

template <int G, int WPL>
static cudaError_t launch_nogap_t(const DevCtx &c, const uint32_t *offs0, const uint32_t *loci0,
                                  const uint32_t *offs1, const uint32_t *loci1, size_t n0, int T0,
                                  int8_t *acc, salt_verify_out_t *rec, salt_pair_t *lv_pairs, uint32_t *lv_slots,
                                  uint32_t *lv_count, uint32_t *lv_reads, cudaStream_t st)
{
    auto kern = nogap_fused_kernel<G, WPL>;
    const size_t threads = (size_t)c.n_reads * G;
    SALT_LAUNCH(kern, (unsigned)((threads + 63) / 64), 64, 0, st, c, offs0, loci0, offs1, loci1, n0, T0, acc, rec,
                lv_pairs, lv_slots, lv_count, lv_reads);
    SALT_LAUNCH_CHECK();
    return cudaSuccess;
}

cudaError_t launch_nogap_fused(const DevCtx &c, const uint32_t *offs0, const uint32_t *loci0,
                               const uint32_t *offs1, const uint32_t *loci1, size_t n0, int T0,
                               int8_t *acc, salt_verify_out_t *rec, salt_pair_t *lv_pairs, uint32_t *lv_slots,
                               uint32_t *lv_count, uint32_t *lv_reads, cudaStream_t st)
{
    if (!c.n_reads) return cudaSuccess;
    // G lanes x WPL 64-bit words per lane cover 16*G*WPL bases; one spare word so the last lane
    // never needs a neighbour: l_max + 16 <= 16*G*WPL
    // Two words per lane on half as many lanes measured 30 % faster than one word per lane for 150 and 250 bp
    // (per-candidate work is shared by fewer lanes); at 100 bp 8 x 1 and 4 x 2 measured the same.
    const int lm = (int)c.l_max;
    if (lm <= 112) return launch_nogap_t<8, 1>(c, offs0, loci0, offs1, loci1, n0, T0, acc, rec, lv_pairs, lv_slots, lv_count, lv_reads, st);
    if (lm <= 240) return launch_nogap_t<8, 2>(c, offs0, loci0, offs1, loci1, n0, T0, acc, rec, lv_pairs, lv_slots, lv_count, lv_reads, st);
    if (lm <= 496) return launch_nogap_t<16, 2>(c, offs0, loci0, offs1, loci1, n0, T0, acc, rec, lv_pairs, lv_slots, lv_count, lv_reads, st);
    if (lm <= 1008) return launch_nogap_t<32, 2>(c, offs0, loci0, offs1, loci1, n0, T0, acc, rec, lv_pairs, lv_slots, lv_count, lv_reads, st);
    return launch_nogap_t<32, 3>(c, offs0, loci0, offs1, loci1, n0, T0, acc, rec, lv_pairs, lv_slots, lv_count, lv_reads, st);
}

cudaError_t launch_scan_gap(const DevCtx &c, const uint32_t *offs0, const uint32_t *loci0,
                            const uint32_t *offs1, const uint32_t *loci1, size_t n0, int lv_T0,
                            int8_t *acc, salt_verify_out_t *rec, const uint32_t *lv_reads, const uint32_t *lv_read_count,
                            uint32_t *cig_list, uint32_t *cig_count, int sm_count, cudaStream_t st)
{
    if (!c.n_reads) return cudaSuccess;
    size_t blocks = ((size_t)c.n_reads * 8 + 127) / 128;      // eight lanes per listed read
    const size_t cap = (size_t)sm_count * 32;            // the list is usually a few percent of the reads
    if (blocks > cap) blocks = cap;
    SALT_LAUNCH(scan_gap_kernel, (unsigned)blocks, 128, 0, st, c, offs0, loci0, offs1, loci1, n0, lv_T0, acc, rec,
                lv_reads, lv_read_count, cig_list, cig_count);
    SALT_LAUNCH_CHECK();
    return cudaSuccess;
}

}  // namespace salt
