// sw_core.cuh -- per-thread strip update of the affine-gap local alignment DP
// (what ssw.c:371-547 sw_sse2_word computes, restated without the striped layout).
//
// Two alignment tasks are packed into the two signed 16-bit halves of every 32-bit
// register ("inter-task SIMD in a word").  A group of G threads owns one task pair; thread
// j holds rows [j*S, (j+1)*S) of the read in registers (H of the previous column and E for
// the current column) and walks the reference window column by column, one column behind
// thread j-1 (systolic wavefront).  Per cell and task pair:
//
//     h  = max(Hdiag + s, E)           VIADDMNMX.S16x2
//     h  = max(h, F, 0)                VIMNMX3.S16x2
//     E' = max(h - gapO, E - gapE)     VIADD + VIADDMNMX.S16x2
//     F' = max(h - gapO, F - gapE)     VIADD + VIADDMNMX.S16x2
// (h - gapO computed once for both saves an instruction per cell pair and measured 10 % SLOWER: it puts a third
// instruction on the F chain that runs down the strip's rows, and at five CTAs per SM that chain is what a warp waits on.)
//
// All of H, E, F carry a constant bias SW_BIAS in both halves ("0" is SW_BIAS).  With every half
// >= SW_BIAS - gapO > gapE, subtracting gapE from both halves is one 32-bit integer subtraction
// that never borrows across the halves: a plain 32-bit VIADD instead of the SIMD VIADD.16x2, which
// measured 6 % faster on B200 (forcing it onto the FMA pipe as IMAD measured slower again).
// E and F are allowed to go below the bias (>= SW_BIAS - gapO); the reference floors them at 0
// with saturating unsigned subtraction (ssw.c:458-466), which yields the same H because h is
// floored itself.  Substitution scores come from an 8-byte row of the score table per
// reference symbol, looked up with PRMT using per-row read-code selectors, so no per-task
// query profile is kept in memory.
//
// Rows are numbered so that the read's LAST padded row (8*ceil(L/8)-1, ssw.c:352) is row
// G*S-1; rows above the read are "top pad" rows with score -128 for every symbol, which
// stay at H = 0 and hand the read's first row exactly the boundary the reference has.
#pragma once
#include "common.cuh"

namespace salt {

// read-code alphabet used by the selectors
constexpr int SW_CODE_N = 4;        // read N
constexpr int SW_CODE_TOP = 5;      // row above the read: score -128 against everything
constexpr int SW_CODE_TAIL = 6;     // reference's zero-score padded rows [L, 8*ceil(L/8))
constexpr int SW_SYM_PADCOL = 16;   // column past the end of a task's window
constexpr int SW_BIAS = 0x2000;     // value of "0" in both int16 halves of H, E, F inside the DP kernels
constexpr uint32_t SW_BIAS2 = (uint32_t)SW_BIAS | ((uint32_t)SW_BIAS << 16);

SALT_HD uint32_t s16x2(int lo, int hi) { return ((uint32_t)(uint16_t)(int16_t)lo) | ((uint32_t)(uint16_t)(int16_t)hi << 16); }
SALT_HD int s16lo(uint32_t x) { return (int)(int16_t)(x & 0xffffu); }
SALT_HD int s16hi(uint32_t x) { return (int)(int16_t)(x >> 16); }

#if defined(__CUDA_ARCH__)
SALT_HD uint32_t vaddmax_relu(uint32_t a, uint32_t b, uint32_t c) { return __viaddmax_s16x2_relu(a, b, c); }
SALT_HD uint32_t vaddmax(uint32_t a, uint32_t b, uint32_t c) { return __viaddmax_s16x2(a, b, c); }
SALT_HD uint32_t vmax2(uint32_t a, uint32_t b) { return __vmaxs2(a, b); }
SALT_HD uint32_t vmax3(uint32_t a, uint32_t b, uint32_t c) { return __vimax3_s16x2(a, b, c); }
SALT_HD uint32_t vadd2(uint32_t a, uint32_t b) { return __vadd2(a, b); }
SALT_HD uint32_t prmt(uint32_t a, uint32_t b, uint32_t s) { return __byte_perm(a, b, s); }
// byte_perm has no sign-replication mode; PRMT does (selector bit 3)
SALT_HD uint32_t prmt_sx(uint32_t a, uint32_t b, uint32_t s)
{
    uint32_t r;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(s));
    return r;
}
#else
SALT_HD int mx(int a, int b) { return a > b ? a : b; }
SALT_HD uint32_t vaddmax_relu(uint32_t a, uint32_t b, uint32_t c)
{
    return s16x2(mx(mx((int16_t)(s16lo(a) + s16lo(b)), s16lo(c)), 0), mx(mx((int16_t)(s16hi(a) + s16hi(b)), s16hi(c)), 0));
}
SALT_HD uint32_t vaddmax(uint32_t a, uint32_t b, uint32_t c)
{
    return s16x2(mx((int16_t)(s16lo(a) + s16lo(b)), s16lo(c)), mx((int16_t)(s16hi(a) + s16hi(b)), s16hi(c)));
}
SALT_HD uint32_t vmax2(uint32_t a, uint32_t b) { return s16x2(mx(s16lo(a), s16lo(b)), mx(s16hi(a), s16hi(b))); }
SALT_HD uint32_t vmax3(uint32_t a, uint32_t b, uint32_t c) { return vmax2(vmax2(a, b), c); }
SALT_HD uint32_t vadd2(uint32_t a, uint32_t b) { return s16x2(s16lo(a) + s16lo(b), s16hi(a) + s16hi(b)); }
SALT_HD uint32_t prmt_sx(uint32_t a, uint32_t b, uint32_t s)
{
    uint64_t src = (uint64_t)a | ((uint64_t)b << 32);
    uint32_t r = 0;
    for (int i = 0; i < 4; ++i) {
        unsigned sel = (s >> (4 * i)) & 15u;
        unsigned byte = (unsigned)(src >> (8 * (sel & 7u))) & 255u;
        if (sel & 8u) byte = (byte & 128u) ? 255u : 0u;
        r |= byte << (8 * i);
    }
    return r;
}
SALT_HD uint32_t prmt(uint32_t a, uint32_t b, uint32_t s) { return prmt_sx(a, b, s & 0x7777u); }
#endif

// Register state of one thread: S rows x 2 tasks.
template <int S>
struct SwStrip {
    static constexpr int NQ = (S + 3) / 4;
    uint32_t H[S];           // H(row, previous column), packed pair
    uint32_t E[S];           // E(row, current column)
    uint32_t sel0[NQ];       // read-code selectors, 4 rows per word (low 16 bits), task 0 / task 1
    uint32_t sel1[NQ];

    SALT_HD void clear()
    {
#pragma unroll
        for (int i = 0; i < S; ++i) { H[i] = SW_BIAS2; E[i] = SW_BIAS2; }
    }

    // One column, everything in the biased domain.  t0/t1: 8-byte score-table rows of the two tasks'
    // reference symbols (lo,hi words).  diag: H(row j*S-1, previous column); F: F entering row j*S.
    // negO: (-gapO, -gapO) as s16x2; negE32: -(gapE | gapE << 16) as a 32-bit integer.
    // Returns the strip maximum; leaves the new bottom H in H[S-1] and F leaving the strip in F.
    SALT_HD uint32_t column(uint32_t t0lo, uint32_t t0hi, uint32_t t1lo, uint32_t t1hi,
                            uint32_t diag, uint32_t &F, uint32_t negO, uint32_t negE32)
    {
        uint32_t sm = SW_BIAS2, hprev = SW_BIAS2;
#pragma unroll
        for (int q = 0; q < NQ; ++q) {
            // 4 int8 scores per task for rows 4q..4q+3
            const uint32_t x0 = prmt(t0lo, t0hi, sel0[q]);
            const uint32_t x1 = prmt(t1lo, t1hi, sel1[q]);
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const int i = 4 * q + r;
                if (i < S) {
                    // (x0.byte r sign-extended) | (x1.byte r sign-extended) << 16
                    const uint32_t s = prmt_sx(x0, x1, (uint32_t)(r | ((r | 8) << 4) | ((4 + r) << 8) | ((4 + r) | 8) << 12));
                    uint32_t h = vaddmax(diag, s, E[i]);
                    h = vmax3(h, F, SW_BIAS2);
                    diag = H[i];
                    H[i] = h;
                    if (i & 1) sm = vmax3(sm, hprev, h);         // the strip maximum takes two rows per instruction
                    else hprev = h;
                    E[i] = vaddmax(h, negO, E[i] + negE32);
                    F = vaddmax(h, negO, F + negE32);
                }
            }
        }
        if (S & 1) sm = vmax2(sm, hprev);
        return sm;
    }
};

// banded_sw direction byte: bit0 = E came from H (code 3 vs 2), bit1 = F came from H (5 vs 4),
// bits 2..4 = code of the H cell (1..5)  (ssw.c:601-623)
SALT_HD uint8_t sw_dir_pack(int de, int df, int dh) { return (uint8_t)((de == 3) | ((df == 5) << 1) | (dh << 2)); }
SALT_HD int sw_dir_get(uint8_t b, int which) { return which == 0 ? ((b & 1) ? 3 : 2) : (which == 1 ? ((b & 2) ? 5 : 4) : (b >> 2)); }

}  // namespace salt
