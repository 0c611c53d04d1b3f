// lv_core.cuh -- per-diagonal pieces of the SNP-aware Landau-Vishkin kernels.
//
// Text (reference window) and pattern (read) are kept as packed 4-bit symbols, 8 per 32-bit
// word, least-significant nibble first -- the mixRef layout (metaref.c:54-56).  Text nibble =
// allele mask, 0 beyond textLen; pattern nibble = one-hot base (N = 15), 0 beyond patternLen.
// With that encoding the reference's byte tests (LandauVishkin.c:79-95) become nibble tests:
//   gate   *p == *t          ->  nibbles equal
//   match  (p & t) != 0      ->  nibble AND non-zero
#pragma once
#include "common.cuh"

namespace salt {

constexpr int LV_MAXK = 31;            // LandauVishkin.c:13
constexpr int LV_ND = 64;              // diagonals -32..31 held by the widest kernel

// Longest common extension on diagonal d starting at pattern offset `best`
// (LandauVishkin.c:77-104 / :262-289).
//   - equality gate on the first symbol (:79); if it fails `best` is returned unchanged
//   - both symbols zero (pattern exhausted over ref N / text exhausted): the reference's
//     8-byte x==0 shortcut walks to pend and clamps -> endl(d)
//   - otherwise extend while the AND is non-zero, clamp to endl(d)
SALT_HD int lv_extend(const uint32_t *T, const uint32_t *P, int best, int d, int plen, int tlen)
{
    const int e = imin(plen, tlen - d);
    uint32_t pc = nib8(P, best), tc = nib8(T, d + best);
    const uint32_t pb = pc & 15u, tb = tc & 15u;
    if (pb != tb) return best;
    if (pb == 0) return e;
    for (;;) {
        const uint32_t z = zero_nibbles(pc & tc);
        if (z) { best += first_set_nibble(z); break; }
        best += 8;
        if (best >= e) break;
        pc = nib8(P, best); tc = nib8(T, d + best);
    }
    return imin(best, e);
}

// lv_extend split in two for thread-per-pair kernels: the gate and the first 8 symbols here,
// the rest in lv_extend_more.  Most diagonals stop inside the first word; the few that run on
// (`more`) are finished after the level's diagonal loop, where the lanes of a warp that have one
// iterate together instead of one after the other.  `toff`: nibble offset of text position 0 inside T
// (a kernel may keep the window as the raw 16-byte-aligned reference words it loaded).
SALT_HD int lv_extend_first(const uint32_t *T, const uint32_t *P, int best, int d, int plen, int tlen, bool &more,
                            int toff = 0)
{
    const int e = imin(plen, tlen - d);
    const uint32_t pc = nib8(P, best), tc = nib8(T, toff + d + best);
    const uint32_t pb = pc & 15u, tb = tc & 15u;
    more = false;
    if (pb != tb) return best;
    if (pb == 0) return e;
    const uint32_t z = zero_nibbles(pc & tc);
    if (z) return imin(best + first_set_nibble(z), e);
    if (best + 8 >= e) return e;
    more = true;
    return best + 8;
}

// continuation of lv_extend_first from `best` < endl(d), all symbols before it matched
SALT_HD int lv_extend_more(const uint32_t *T, const uint32_t *P, int best, int d, int plen, int tlen, int toff = 0)
{
    const int e = imin(plen, tlen - d);
    for (;;) {
        const uint32_t z = zero_nibbles(nib8(P, best) & nib8(T, toff + d + best));
        if (z) { best += first_set_nibble(z); break; }
        best += 8;
        if (best >= e) break;
    }
    return imin(best, e);
}

// Level-0 extension from (0,0), no gate (LandauVishkin.c:41-58).  Single-thread form; the
// kernels use a cooperative version with the same result.
SALT_HD int lv_extend0(const uint32_t *T, const uint32_t *P, int plen, int tlen, int toff = 0)
{
    const int e = imin(plen, tlen);
    int i = 0;
    while (i < e) {
        const uint32_t z = zero_nibbles(nib8(P, i) & nib8(T, toff + i));
        if (z) { i += first_set_nibble(z); break; }
        i += 8;
    }
    return imin(i, e);
}

// rank of diagonal d in the CIGAR variant's search order 0,-1,+1,-2,+2,... (LandauVishkin.c:248)
SALT_HD int lv_cigar_rank(int d) { return d == 0 ? 0 : (d < 0 ? -2 * d - 1 : 2 * d); }

// "%d%c" with writeCigar's COMPACT_CIGAR_STRING bookkeeping (LandauVishkin.c:139-152):
// snprintf truncates to len-1 characters + NUL and the call fails when it had to truncate.
struct CigarOut {
    char *buf; int len;
    SALT_HD bool put(int count, char code)
    {
        if (count <= 0) return true;
        char tmp[12]; int nd = 0;
        int c = count;
        while (c > 0) { tmp[nd++] = (char)('0' + c % 10); c /= 10; }
        const int w = nd + 1;
        const int room = len - 1;                       // characters snprintf may write
        for (int i = 0; i < w && i < room; ++i) buf[i] = (i < nd) ? tmp[nd - 1 - i] : code;
        if (len > 0) buf[w < room ? w : room] = '\0';
        if (w > len - 1) return false;
        buf += w; len -= w;
        return true;
    }
};

// Table layouts of the CIGAR variant: entry (e, d) of the furthest-reaching / action tables.
struct LvRectIdx {                       // row-major [e][nd], diagonal d at column d + nd/2
    int nd;
    SALT_HD int operator()(int e, int d) const { return e * nd + d + nd / 2; }
};
struct LvTriIdx {                        // level e holds diagonals -e..e only: row e starts at e*e
    SALT_HD int operator()(int e, int d) const { return e * e + d + e; }
};

// Backtrace + emission of computeEditDistanceWithCigar, useM = 1 (LandauVishkin.c:380-462).
// Lt/At are the furthest-reaching table and action table addressed through `at`.
// Returns e or -2 (buffer too small).
template <class Idx, class LT>
SALT_HD int lv_cigar_emit_t(const LT *Lt, const char *At, Idx at, int e, int d, char *buf, int buflen)
{
    char act[LV_MAXK + 1]; int run[LV_MAXK + 1];
    int cd = d;
    for (int ce = e; ce >= 1; --ce) {
        const char a = At[at(ce, cd)];
        act[ce] = a;
        const int here = Lt[at(ce, cd)];
        if (a == 'I') { run[ce] = here - Lt[at(ce - 1, cd + 1)] - 1; cd += 1; }
        else if (a == 'D') { run[ce] = here - Lt[at(ce - 1, cd - 1)]; cd -= 1; }
        else { run[ce] = here - Lt[at(ce - 1, cd)] - 1; }
    }
    CigarOut o{buf, buflen};
    int accM = Lt[at(0, 0)];
    int ce = 1;
    while (ce <= e) {
        const char a = act[ce]; int cnt = 1;
        while (ce + 1 <= e && run[ce] == 0 && act[ce + 1] == a) { ++cnt; ++ce; }
        if (a == 'X') accM += cnt;
        else {
            if (accM != 0) { if (!o.put(accM, 'M')) return -2; accM = 0; }
            if (!o.put(cnt, a)) return -2;
        }
        if (run[ce] > 0) accM += run[ce];
        ++ce;
    }
    if (accM != 0) { if (!o.put(accM, 'M')) return -2; }
    if (o.len > 0) *o.buf = '\0';
    return e;
}

SALT_HD int lv_cigar_emit(const int16_t *Lt, const char *At, int nd, int e, int d, char *buf, int buflen)
{
    return lv_cigar_emit_t(Lt, At, LvRectIdx{nd}, e, d, buf, buflen);
}

}  // namespace salt
