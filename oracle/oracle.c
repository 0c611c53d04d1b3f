/*
 * oracle/oracle.c -- TEST INFRASTRUCTURE ONLY.
 *
 * A plain-C, single-threaded CPU restatement of the algorithms on salt's
 * verification/extension hot path.  It exists so that tests/, the smoke test
 * and bench.py's cpu_baseline / --impl reference legs have something to check
 * the CUDA engine against.  Nothing under salt_b200/ may include, link or call
 * this file; the product path has no CPU fallback.
 *
 * Parity pin: every function below is fuzzed against the reference's own
 * unmodified sources compiled into oracle/_ref/libsaltref.so (see
 * oracle/Makefile and tests/test_oracle_vs_ref.py), and against the golden
 * vectors in tests/golden/ that were generated from that library.
 *
 * Each function cites the reference file:line whose behaviour it restates
 * (paths relative to the reference tree, Align_src/ unless noted).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>

#define ORC_EXPORT __attribute__((visibility("default")))

/* ------------------------------------------------------------------------ */
/* mixRef access: 4 bit/base, base p in bits 4*(p%8).. of word p/8          */
/* (metaref.c:54-56, alnse.c:783)                                           */
/* ------------------------------------------------------------------------ */
static inline unsigned orc_nib(const uint32_t *mixref, uint32_t p)
{
    return (mixref[p >> 3] >> (4u * (p & 7u))) & 15u;
}

/* read code 0..3 -> one-hot allele bit, 4 (N) -> 15 (editdistance.c:40) */
static inline unsigned orc_onehot(unsigned code)
{
    return code > 3 ? 15u : (1u << code);
}

/* ------------------------------------------------------------------------ */
/* ed_mismatch (editdistance.c:88-163): SNP-aware Hamming distance with a   */
/* threshold.  Returns n if n <= max_err else -1.                           */
/* Reference precondition (its word loop runs away otherwise): the window   */
/* spans at least two mixRef words, i.e. ref_st%8 + l_comp > 8.             */
/* ------------------------------------------------------------------------ */
ORC_EXPORT int orc_ed_mismatch(const uint32_t *mixref, uint32_t ref_st,
                               const uint8_t *seq, uint32_t l_comp, int max_err)
{
    int n = 0;
    for (uint32_t i = 0; i < l_comp; ++i) {
        if ((orc_nib(mixref, ref_st + i) & orc_onehot(seq[i])) == 0) {
            if (++n > max_err) return -1;
        }
    }
    return n;
}

/* ------------------------------------------------------------------------ */
/* Landau-Vishkin with salt's AND-match (LandauVishkin.c:19-122).           */
/* text/pattern are 1 byte per base; out-of-range positions read as 0, as   */
/* the reference's calloc padding makes them (editdistance.c:183-184).      */
/* ------------------------------------------------------------------------ */
typedef struct {
    const uint8_t *t; int tl;
    const uint8_t *p; int pl;
} lv_in_t;

static inline int lvT(const lv_in_t *in, int i) { return (i >= 0 && i < in->tl) ? in->t[i] : 0; }
static inline int lvP(const lv_in_t *in, int i) { return (i >= 0 && i < in->pl) ? in->p[i] : 0; }
static inline int lv_endl(const lv_in_t *in, int d)
{
    int a = in->pl, b = in->tl - d;
    return a < b ? a : b;
}

/* longest common extension on diagonal d from pattern offset `best`
 * (LandauVishkin.c:77-104 and :262-289).  The byte-equality gate at :79/:264
 * is part of the observable behaviour. */
static int lv_extend(const lv_in_t *in, int best, int d)
{
    int pb = lvP(in, best), tb = lvT(in, d + best);
    if (pb != tb) return best;
    int e = lv_endl(in, d);
    if (pb == 0) return e;               /* identical zero bytes: 8-byte shortcut then clamp */
    int j = best;
    while (lvP(in, j) & lvT(in, d + j)) ++j;
    return j < e ? j : e;
}

#define LV_MAXK 31
#define LV_W (2 * LV_MAXK + 1)

ORC_EXPORT int orc_lv(const uint8_t *text, int textLen, const uint8_t *pattern, int patternLen, int k)
{
    short L[LV_MAXK + 1][LV_W];
    lv_in_t in = { text, textLen, pattern, patternLen };
    for (int i = 0; i <= LV_MAXK; ++i) for (int j = 0; j < LV_W; ++j) L[i][j] = -2;
    if (k > LV_MAXK - 1) k = LV_MAXK - 1;               /* :31 */
    if (!text) return -1;
    int e0 = lv_endl(&in, 0);
    int i = 0;
    while (i < e0 && (lvP(&in, i) & lvT(&in, i))) ++i;  /* :41-58 */
    L[0][LV_MAXK] = (short)i;
    if (i == e0) return patternLen > e0 ? patternLen - e0 : 0;   /* :60-63 */
    for (int e = 1; e <= k; ++e) {
        /* diagonal order 0,+1,-1,+2,-2,... (:67); the result does not depend on it */
        for (int d = 0; d != e + 1; d = (d > 0 ? -d : -d + 1)) {
            int best = L[e - 1][LV_MAXK + d] + 1;
            int left = L[e - 1][LV_MAXK + d - 1];
            int right = L[e - 1][LV_MAXK + d + 1] + 1;
            if (left > best) best = left;
            if (right > best) best = right;
            best = lv_extend(&in, best, d);
            if (best == patternLen) return e;
            L[e][LV_MAXK + d] = (short)best;
        }
    }
    return -1;
}

/* snprintf("%d%c") with the truncation/return-code semantics of
 * writeCigar's COMPACT_CIGAR_STRING branch (LandauVishkin.c:139-152). */
static int cig_put(char **buf, int *len, int count, char code)
{
    if (count <= 0) return 1;
    if (*len == 0) { *(*buf - 1) = '\0'; return 0; }
    int w = snprintf(*buf, (size_t)*len, "%d%c", count, code);
    if (w > *len - 1) return 0;
    *buf += w; *len -= w;
    return 1;
}

/* computeEditDistanceWithCigar, useM=1, COMPACT_CIGAR_STRING
 * (LandauVishkin.c:176-470).  Returns e, -1 (not within k), -2 (buffer),
 * -3 if k >= 31 (the reference asserts, :183). */
ORC_EXPORT int orc_lv_cigar(const uint8_t *text, int textLen, const uint8_t *pattern, int patternLen,
                            int k, char *cigarBuf, int cigarBufLen)
{
    if (k >= LV_MAXK) return -3;
    if (!text) return -1;
    short L[LV_MAXK + 1][LV_W];
    char A[LV_MAXK + 1][LV_W];
    lv_in_t in = { text, textLen, pattern, patternLen };
    for (int i = 0; i <= LV_MAXK; ++i) for (int j = 0; j < LV_W; ++j) { L[i][j] = -2; A[i][j] = 0; }
    int e0 = lv_endl(&in, 0);
    int i = 0;
    while (i < e0 && (lvP(&in, i) & lvT(&in, i))) ++i;
    L[0][LV_MAXK] = (short)i;
    if (i == e0) {                                        /* :224-242 */
        if (!cig_put(&cigarBuf, &cigarBufLen, patternLen, 'M')) return -2;
        return 0;
    }
    for (int e = 1; e <= k; ++e) {
        /* diagonal order 0,-1,+1,-2,+2,... (:248) decides which alignment is reported */
        for (int d = 0; d != -(e + 1); d = (d >= 0 ? -(d + 1) : -d)) {
            int best = L[e - 1][LV_MAXK + d] + 1; char a = 'X';
            int left = L[e - 1][LV_MAXK + d - 1];
            if (left > best) { best = left; a = 'D'; }
            int right = L[e - 1][LV_MAXK + d + 1] + 1;
            if (right > best) { best = right; a = 'I'; }
            A[e][LV_MAXK + d] = a;
            best = lv_extend(&in, best, d);
            L[e][LV_MAXK + d] = (short)best;              /* stored before the test (:291) */
            if (best != patternLen) continue;

            char act[LV_MAXK + 1]; int run[LV_MAXK + 1];
            int cd = d;
            for (int ce = e; ce >= 1; --ce) {             /* :380-399 */
                act[ce] = A[ce][LV_MAXK + cd];
                if (act[ce] == 'I') {
                    run[ce] = L[ce][LV_MAXK + cd] - L[ce - 1][LV_MAXK + cd + 1] - 1; cd += 1;
                } else if (act[ce] == 'D') {
                    run[ce] = L[ce][LV_MAXK + cd] - L[ce - 1][LV_MAXK + cd - 1]; cd -= 1;
                } else {
                    run[ce] = L[ce][LV_MAXK + cd] - L[ce - 1][LV_MAXK + cd] - 1;
                }
            }
            int accM = L[0][LV_MAXK];
            int ce = 1;
            while (ce <= e) {                              /* :413-452 */
                char ac = act[ce]; int cnt = 1;
                while (ce + 1 <= e && run[ce] == 0 && act[ce + 1] == ac) { ++cnt; ++ce; }
                if (ac == 'X') accM += cnt;
                else {
                    if (accM != 0) { if (!cig_put(&cigarBuf, &cigarBufLen, accM, 'M')) return -2; accM = 0; }
                    if (!cig_put(&cigarBuf, &cigarBufLen, cnt, ac)) return -2;
                }
                if (run[ce] > 0) accM += run[ce];
                ++ce;
            }
            if (accM != 0) { if (!cig_put(&cigarBuf, &cigarBufLen, accM, 'M')) return -2; }
            *(cigarBuf - (cigarBufLen == 0 ? 1 : 0)) = '\0';
            return e;
        }
    }
    *(cigarBuf - (cigarBufLen == 0 ? 1 : 0)) = '\0';      /* :468 */
    return -1;
}

/* window unpack shared by ed_diff / ed_diff_withcigar (editdistance.c:183-218) */
static void orc_unpack(const uint32_t *mixref, uint32_t ref_st, uint32_t l_ref,
                       const uint8_t *seq, uint32_t l_seq, uint8_t *t, uint8_t *q)
{
    for (uint32_t i = 0; i < l_ref; ++i) t[i] = (uint8_t)orc_nib(mixref, ref_st + i);
    for (uint32_t i = 0; i < l_seq; ++i) q[i] = (uint8_t)orc_onehot(seq[i]);
}

/* ed_diff (editdistance.c:174-232) */
ORC_EXPORT int orc_ed_diff(const uint32_t *mixref, uint32_t l_mref, uint32_t ref_st, uint32_t l_ref,
                           const uint8_t *seq, uint32_t l_seq, int max_k_diff)
{
    if (ref_st > l_mref || ref_st + l_ref > l_mref) return -1;      /* :178 */
    uint8_t *t = (uint8_t *)calloc(l_ref + 16, 1), *q = (uint8_t *)calloc(l_seq + 16, 1);
    orc_unpack(mixref, ref_st, l_ref, seq, l_seq, t, q);
    int r = orc_lv(t, (int)l_ref, q, (int)l_seq, max_k_diff);
    free(t); free(q);
    return r;
}

/* ed_diff_withcigar (editdistance.c:234-284), useM=1 / COMPACT_CIGAR_STRING */
ORC_EXPORT int orc_ed_diff_withcigar(const uint32_t *mixref, uint32_t ref_st, uint32_t l_ref,
                                     const uint8_t *seq, uint32_t l_seq, int max_k_diff,
                                     char *cigarBuf, int cigarLen)
{
    uint8_t *t = (uint8_t *)calloc(l_ref + 16, 1), *q = (uint8_t *)calloc(l_seq + 16, 1);
    orc_unpack(mixref, ref_st, l_ref, seq, l_seq, t, q);
    int r = orc_lv_cigar(t, (int)l_ref, q, (int)l_seq, max_k_diff, cigarBuf, cigarLen);
    free(t); free(q);
    return r;
}

/* ------------------------------------------------------------------------ */
/* SSW: scalar emulation of the SSE2 int16 striped kernel                   */
/* (ssw.c:347-369 qP_word, :371-547 sw_sse2_word).  Eight lanes, segLen     */
/* vectors per column, saturating ops spelled out.                          */
/* ------------------------------------------------------------------------ */
typedef struct { int16_t v[8]; } v8;

static inline int16_t sat_adds(int a, int b) { int s = a + b; return (int16_t)(s > 32767 ? 32767 : (s < -32768 ? -32768 : s)); }
static inline int16_t sat_subu(uint16_t a, uint16_t b) { return (int16_t)(a > b ? a - b : 0); }

static inline v8 v8_zero(void) { v8 r; memset(&r, 0, sizeof r); return r; }
static inline v8 v8_set1(int16_t x) { v8 r; for (int i = 0; i < 8; ++i) r.v[i] = x; return r; }
static inline v8 v8_shl1(v8 a) { v8 r; r.v[0] = 0; for (int i = 1; i < 8; ++i) r.v[i] = a.v[i - 1]; return r; }
static inline v8 v8_adds(v8 a, v8 b) { v8 r; for (int i = 0; i < 8; ++i) r.v[i] = sat_adds(a.v[i], b.v[i]); return r; }
static inline v8 v8_subu(v8 a, v8 b) { v8 r; for (int i = 0; i < 8; ++i) r.v[i] = sat_subu((uint16_t)a.v[i], (uint16_t)b.v[i]); return r; }
static inline v8 v8_max(v8 a, v8 b) { v8 r; for (int i = 0; i < 8; ++i) r.v[i] = a.v[i] > b.v[i] ? a.v[i] : b.v[i]; return r; }
static inline int v8_any_gt(v8 a, v8 b) { for (int i = 0; i < 8; ++i) if (a.v[i] > b.v[i]) return 1; return 0; }
static inline int v8_eq(v8 a, v8 b) { return memcmp(&a, &b, sizeof a) == 0; }
static inline uint16_t v8_hmax(v8 a) { int16_t m = a.v[0]; for (int i = 1; i < 8; ++i) if (a.v[i] > m) m = a.v[i]; return (uint16_t)m; }

typedef struct { uint16_t score; int32_t ref; int32_t read; } orc_end_t;

static v8 *orc_profile(const int8_t *read, const int8_t *mat, int readLen, int n)
{
    int segLen = (readLen + 7) / 8;
    v8 *prof = (v8 *)malloc((size_t)n * segLen * sizeof(v8));
    for (int nt = 0; nt < n; ++nt)
        for (int i = 0; i < segLen; ++i)
            for (int s = 0; s < 8; ++s) {
                int j = i + s * segLen;
                prof[nt * segLen + i].v[s] = (int16_t)(j >= readLen ? 0 : mat[nt * n + read[j]]);
            }
    return prof;
}

static void orc_sw_word(const int8_t *ref, int dir, int refLen, int readLen, int gapO, int gapE,
                        const v8 *prof, uint16_t terminate, int maskLen, orc_end_t bests[2])
{
    uint16_t max = 0;
    int end_read = readLen - 1, end_ref = 0, segLen = (readLen + 7) / 8;
    uint16_t *maxColumn = (uint16_t *)calloc((size_t)refLen, 2);
    v8 *Hs = (v8 *)calloc((size_t)segLen, sizeof(v8)), *Hl = (v8 *)calloc((size_t)segLen, sizeof(v8));
    v8 *E = (v8 *)calloc((size_t)segLen, sizeof(v8)), *Hmax = (v8 *)calloc((size_t)segLen, sizeof(v8));
    v8 vGapO = v8_set1((int16_t)gapO), vGapE = v8_set1((int16_t)gapE);
    v8 vMaxScore = v8_zero(), vMaxMark = v8_zero();
    int begin = 0, end = refLen, step = 1;
    if (dir == 1) { begin = refLen - 1; end = -1; step = -1; }
    for (int i = begin; i != end; i += step) {
        v8 vF = v8_zero(), vMaxColumn = v8_zero();
        v8 vH = v8_shl1(Hs[segLen - 1]);
        const v8 *vP = prof + (int)ref[i] * segLen;
        v8 *tmp = Hl; Hl = Hs; Hs = tmp;
        for (int j = 0; j < segLen; ++j) {                 /* :450-476 */
            vH = v8_adds(vH, vP[j]);
            v8 e = E[j];
            vH = v8_max(vH, e);
            vH = v8_max(vH, vF);
            vMaxColumn = v8_max(vMaxColumn, vH);
            Hs[j] = vH;
            vH = v8_subu(vH, vGapO);
            e = v8_subu(e, vGapE);
            e = v8_max(e, vH);
            E[j] = e;
            vF = v8_subu(vF, vGapE);
            vF = v8_max(vF, vH);
            vH = Hl[j];
        }
        for (int k = 0; k < 8; ++k) {                      /* lazy F, :479-489 */
            int done = 0;
            vF = v8_shl1(vF);
            for (int j = 0; j < segLen; ++j) {
                vH = Hs[j];
                vH = v8_max(vH, vF);
                Hs[j] = vH;
                vH = v8_subu(vH, vGapO);
                vF = v8_subu(vF, vGapE);
                if (!v8_any_gt(vF, vH)) { done = 1; break; }
            }
            if (done) break;
        }
        vMaxScore = v8_max(vMaxScore, vMaxColumn);          /* :492-507 */
        if (!v8_eq(vMaxMark, vMaxScore)) {
            vMaxMark = vMaxScore;
            uint16_t t = v8_hmax(vMaxScore);
            if (t > max) {
                max = t; end_ref = i;
                memcpy(Hmax, Hs, (size_t)segLen * sizeof(v8));
            }
        }
        maxColumn[i] = v8_hmax(vMaxColumn);
        if (maxColumn[i] == terminate) break;
    }
    for (int i = 0; i < segLen * 8; ++i) {                  /* :512-521 */
        uint16_t h = (uint16_t)Hmax[i / 8].v[i % 8];
        if (h == max) {
            int r = i / 8 + i % 8 * segLen;
            if (r < end_read) end_read = r;
        }
    }
    bests[0].score = max; bests[0].ref = end_ref; bests[0].read = end_read;
    bests[1].score = 0; bests[1].ref = 0; bests[1].read = 0;
    int edge = (end_ref - maskLen) > 0 ? (end_ref - maskLen) : 0;       /* :537-550 */
    for (int i = 0; i < edge; ++i)
        if (maxColumn[i] > bests[1].score) { bests[1].score = maxColumn[i]; bests[1].ref = i; }
    edge = (end_ref + maskLen) > refLen ? refLen : (end_ref + maskLen);
    for (int i = edge; i < refLen; ++i)
        if (maxColumn[i] > bests[1].score) { bests[1].score = maxColumn[i]; bests[1].ref = i; }
    free(maxColumn); free(Hs); free(Hl); free(E); free(Hmax);
}

/* banded_sw (ssw.c:549-727): scalar banded affine DP with a direction matrix,
 * band doubling until the banded maximum reaches `score`, then traceback from
 * the bottom-right corner.  Output: BAM-style (len<<4 | op) cigar. */
static inline int bu(int w, int i, int j) { int x = i - w; if (x < 0) x = 0; return j - x + 1; }
static inline int bd(int w, int i, int j, int p) { int x = i - w; if (x < 0) x = 0; return (j - x) * 3 + p; }

static int orc_banded(const int8_t *ref, const int8_t *read, int refLen, int readLen, int score,
                      int gapO, int gapE, int band, const int8_t *mat, int n,
                      uint32_t **cig_out, int *ncig_out)
{
    int *hb = NULL, *eb = NULL, *hc = NULL;
    int8_t *dirs = NULL;
    int max = 0, width, wd;
    /* the reference grows its buffers with realloc and never clears them;
     * we size generously and zero once -- every cell that is read was written
     * in the same band iteration or is one of the explicitly zeroed edges. */
    do {
        width = band * 2 + 3; wd = band * 2 + 1;
        hb = (int *)realloc(hb, (size_t)(width + 2) * sizeof(int));
        eb = (int *)realloc(eb, (size_t)(width + 2) * sizeof(int));
        hc = (int *)realloc(hc, (size_t)(width + 2) * sizeof(int));
        dirs = (int8_t *)realloc(dirs, (size_t)wd * readLen * 3 + 16);
        for (int j = 1; j < width - 1; ++j) hb[j] = 0;
        for (int i = 0; i < readLen; ++i) {
            int beg = 0, end = refLen - 1, u = 0;
            if (i - band > beg) beg = i - band;
            if (i + band < end) end = i + band;
            int edge = end + 1 < width - 1 ? end + 1 : width - 1;
            int f = 0;
            hb[0] = eb[0] = hb[edge] = eb[edge] = hc[0] = 0;
            int8_t *dl = dirs + (size_t)wd * i * 3;
            for (int j = beg; j <= end; ++j) {
                u = bu(band, i, j);
                int ue = bu(band, i - 1, j), ub = bu(band, i, j - 1), ud = bu(band, i - 1, j - 1);
                int de = bd(band, i, j, 0), df = bd(band, i, j, 1), dh = bd(band, i, j, 2);
                int t1 = i == 0 ? -gapO : hb[ue] - gapO;
                int t2 = i == 0 ? -gapE : eb[ue] - gapE;
                eb[u] = t1 > t2 ? t1 : t2;
                dl[de] = t1 > t2 ? 3 : 2;
                t1 = hc[ub] - gapO;
                t2 = f - gapE;
                f = t1 > t2 ? t1 : t2;
                dl[df] = t1 > t2 ? 5 : 4;
                int e1 = eb[u] > 0 ? eb[u] : 0;
                int f1 = f > 0 ? f : 0;
                t1 = e1 > f1 ? e1 : f1;
                t2 = hb[ud] + mat[ref[j] * n + read[i]];
                hc[u] = t1 > t2 ? t1 : t2;
                if (hc[u] > max) max = hc[u];
                if (t1 <= t2) dl[dh] = 1;
                else dl[dh] = e1 > f1 ? dl[de] : dl[df];
            }
            for (int j = 1; j <= u; ++j) hb[j] = hc[j];
        }
        band *= 2;
    } while (max < score);
    band /= 2;
    wd = band * 2 + 1;

    int cap = 16 + 2 * (readLen + refLen), l = 0;
    uint32_t *c = (uint32_t *)malloc((size_t)cap * sizeof(uint32_t));
    int i = readLen - 1, j = refLen - 1, e = 0, f = 0, prev = 0, state = 2;
    const int8_t *dl = dirs + (size_t)wd * i * 3;
    while (i > 0) {                                        /* :642-690 */
        int code = dl[bd(band, i, j, state)];
        switch (code) {
        case 1: --i; --j; state = 2; dl -= wd * 3; f = 0; break;
        case 2: --i; state = 0; dl -= wd * 3; f = 1; break;
        case 3: --i; state = 2; dl -= wd * 3; f = 1; break;
        case 4: --j; state = 1; f = 2; break;
        case 5: --j; state = 2; f = 2; break;
        default: free(c); free(hb); free(eb); free(hc); free(dirs); return -1;
        }
        if (f == prev) ++e;
        else { c[l++] = (uint32_t)e << 4 | (uint32_t)prev; prev = f; e = 1; }
    }
    if (f == 0) c[l++] = (uint32_t)(e + 1) << 4;           /* :691-709 */
    else { c[l++] = (uint32_t)e << 4 | (uint32_t)f; c[l++] = 16; }
    uint32_t *out = (uint32_t *)malloc((size_t)(l > 0 ? l : 1) * sizeof(uint32_t));
    for (int s = 0; s < l; ++s) out[s] = c[l - 1 - s];
    free(c); free(hb); free(eb); free(hc); free(dirs);
    *cig_out = out; *ncig_out = l;
    return 0;
}

/* result record mirroring s_align (ssw.h:37-47) without the heap pointer */
typedef struct {
    uint16_t score1, score2;
    int32_t ref_begin1, ref_end1, read_begin1, read_end1, ref_end2;
    int32_t cigarLen;
} orc_align_t;

/* ssw_init(score_size=1) + ssw_align (ssw.c:742-763, :771-856).
 * cigar_out must hold cigar_cap entries; cigarLen reports the true length
 * (entries beyond cigar_cap are dropped).  Returns 0, or -1 where the
 * reference would return NULL. */
ORC_EXPORT int orc_ssw_align(const int8_t *read, int readLen, const int8_t *mat, int n,
                             const int8_t *ref, int refLen, int gapO, int gapE, int flag,
                             int filters, int filterd, int maskLen,
                             orc_align_t *r, uint32_t *cigar_out, int cigar_cap)
{
    orc_end_t b[2];
    memset(r, 0, sizeof *r);
    r->ref_begin1 = -1; r->read_begin1 = -1;
    v8 *prof = orc_profile(read, mat, readLen, n);
    orc_sw_word(ref, 0, refLen, readLen, gapO, gapE, prof, (uint16_t)-1, maskLen, b);
    free(prof);
    r->score1 = b[0].score; r->ref_end1 = b[0].ref; r->read_end1 = b[0].read;
    if (maskLen >= 15) { r->score2 = b[1].score; r->ref_end2 = b[1].ref; }
    else { r->score2 = 0; r->ref_end2 = -1; }
    if (flag == 0 || (flag == 2 && r->score1 < filters)) return 0;

    int rl = r->read_end1 + 1;
    int8_t *rev = (int8_t *)malloc((size_t)rl);
    for (int i = 0; i < rl; ++i) rev[i] = read[r->read_end1 - i];
    prof = orc_profile(rev, mat, rl, n);
    orc_sw_word(ref, 1, r->ref_end1 + 1, rl, gapO, gapE, prof, r->score1, maskLen, b);
    free(prof); free(rev);
    r->ref_begin1 = b[0].ref;
    r->read_begin1 = r->read_end1 - b[0].read;
    if ((7 & flag) == 0 || ((2 & flag) != 0 && r->score1 < filters) ||
        ((4 & flag) != 0 && (r->ref_end1 - r->ref_begin1 > filterd || r->read_end1 - r->read_begin1 > filterd)))
        return 0;

    int sub_ref = r->ref_end1 - r->ref_begin1 + 1, sub_read = r->read_end1 - r->read_begin1 + 1;
    int band = abs(sub_ref - sub_read) + 1;
    uint32_t *cig = NULL; int nc = 0;
    if (orc_banded(ref + r->ref_begin1, read + r->read_begin1, sub_ref, sub_read, r->score1,
                   gapO, gapE, band, mat, n, &cig, &nc) != 0) return -1;
    r->cigarLen = nc;
    for (int i = 0; i < nc && i < cigar_cap; ++i) cigar_out[i] = cig[i];
    free(cig);
    return 0;
}

/* salt's scoring matrices, regenerated from their rule rather than copied
 * (alnpe.c:52-73).  score_mat2 is indexed [ref_mask*16 + read_onehot] by
 * ssw but was laid out [read_onehot][ref_mask]; we reproduce the bytes. */
ORC_EXPORT void orc_score_mat2(int8_t out[256])
{
    for (int r = 0; r < 16; ++r)
        for (int c = 0; c < 16; ++c) {
            int onehot = (r == 1 || r == 2 || r == 4 || r == 8);
            out[r * 16 + c] = (int8_t)((onehot && (c & r)) ? 1 : -3);
        }
}
ORC_EXPORT void orc_score_mat(int8_t out[25])
{
    for (int r = 0; r < 5; ++r)
        for (int c = 0; c < 5; ++c)
            out[r * 5 + c] = (int8_t)((r == 4 || c == 4) ? -1 : (r == c ? 1 : -3));
}

/* mate-rescue wrapper: snpaln_sw_snpaware (alnpe.c:261-328) minus the query_t
 * bookkeeping.  Window [start,end] inclusive on mixRef; read codes 0..4 map to
 * 1<<code (N -> 16).  mat must have n*n+1 readable entries when N can meet
 * mask 15 (the reference reads one past score_mat2 there). */
ORC_EXPORT int orc_rescue_mixref(const uint32_t *mixref, uint32_t start, uint32_t end,
                                 const uint8_t *seq, int l_seq, const int8_t *mat,
                                 int gapO, int gapE, int filters, int filterd,
                                 orc_align_t *r, uint32_t *cigar_out, int cigar_cap)
{
    int l_ref = (int)(end - start + 1);
    int8_t *ref = (int8_t *)calloc((size_t)l_ref, 1), *read = (int8_t *)calloc((size_t)l_seq, 1);
    for (int i = 0; i < l_ref; ++i) ref[i] = (int8_t)orc_nib(mixref, start + (uint32_t)i);
    for (int i = 0; i < l_seq; ++i) read[i] = (int8_t)(1 << seq[i]);
    int rc = orc_ssw_align(read, l_seq, mat, 16, ref, l_ref, gapO, gapE, 2, filters, filterd, l_seq / 2,
                           r, cigar_out, cigar_cap);
    free(ref); free(read);
    return rc;
}

/* snpaln_sw (alnpe.c:330-393): same on the 2-bit pac (MSB-first, alnpe.c:47) with the 5x5 matrix */
ORC_EXPORT int orc_rescue_pac(const uint8_t *pac, uint32_t start, uint32_t end,
                              const uint8_t *seq, int l_seq, const int8_t *mat,
                              int gapO, int gapE, int filters, int filterd,
                              orc_align_t *r, uint32_t *cigar_out, int cigar_cap)
{
    int l_ref = (int)(end - start + 1);
    int8_t *ref = (int8_t *)calloc((size_t)l_ref, 1);
    for (int i = 0; i < l_ref; ++i) {
        uint32_t p = start + (uint32_t)i;
        ref[i] = (int8_t)((pac[p >> 2] >> ((~p & 3) << 1)) & 3);
    }
    int rc = orc_ssw_align((const int8_t *)seq, l_seq, mat, 5, ref, l_ref, gapO, gapE, 2, filters, filterd,
                           l_seq / 2, r, cigar_out, cigar_cap);
    free(ref);
    return rc;
}

/* ------------------------------------------------------------------------ */
/* SNP-aware reference construction                                         */
/* (Index_src/mixRef.c:93-197 build_mixRef; Index_src/hapmap.c:92-158)      */
/* ------------------------------------------------------------------------ */
static inline unsigned base_mask(unsigned char c)           /* mixRef.c:36-53 */
{
    switch (c) {
    case 'A': case 'a': return 1; case 'C': case 'c': return 2;
    case 'G': case 'g': return 4; case 'T': case 't': return 8;
    default: return 0;
    }
}
static inline unsigned base_code(unsigned char c)           /* hapmap.c:22-39: else 4 ('-' -> 5) */
{
    switch (c) {
    case 'A': case 'a': return 0; case 'C': case 'c': return 1;
    case 'G': case 'g': return 2; case 'T': case 't': return 3;
    case '-': return 5;
    default: return 4;
    }
}

/* Set bases [off, off+l) of `words` from ASCII; existing nibbles are replaced. */
ORC_EXPORT void orc_mixref_put_seq(uint32_t *words, uint32_t off, const char *bases, uint32_t l)
{
    for (uint32_t i = 0; i < l; ++i) {
        uint32_t p = off + i;
        words[p >> 3] &= ~(15u << (4 * (p & 7)));
        words[p >> 3] |= base_mask((unsigned char)bases[i]) << (4 * (p & 7));
    }
}

/* allele string "A/G[/T]" -> mask, exactly as the parser does: every second
 * character, 1<<code, truncated to 4 bits when OR-ed in (mixRef.c:150). */
ORC_EXPORT unsigned orc_allele_mask(const char *s)
{
    unsigned m = 0;
    size_t n = strlen(s);
    for (size_t j = 0; j < n; j += 2) m |= 1u << base_code((unsigned char)s[j]);
    return m & 15u;
}

ORC_EXPORT void orc_mixref_or_snp(uint32_t *words, uint32_t pos0, unsigned mask)
{
    words[pos0 >> 3] |= (mask & 15u) << (4 * (pos0 & 7));
}

/* ------------------------------------------------------------------------ */
/* Acceptance logic of the verification loop                                */
/* (alnse.c:348-393 code_kmismatch/code_kdiff, :734-782 alnse_check_nogap,  */
/*  :871-901 alnse_check_withgap, :1014-1036 / :1077-1097 orchestration,    */
/*  query.c:297-333 query_set_hits)                                         */
/* ------------------------------------------------------------------------ */
typedef struct { uint32_t pos; uint8_t n_diff; uint8_t is_gap; uint16_t strand; } orc_hit_t;

typedef struct {
    uint32_t pos; int strand; uint8_t n_diff; uint8_t is_gap;   /* primary (query_t fields) */
    int b0, b1; uint32_t mapq;
    int n_hits[2];                                               /* aux->hits per strand (all accepted) */
    int n_alt[2];                                                /* query->hits[strand] after query_set_hits */
} orc_verify_t;

/* ------------------------------------------------------------------------ */
/* SAM tail: MD / NM / XV of one alignment (sam.c:246-328, sam_add_md_nm).   */
/* `seq` is the read as aligned (query->seq or ->rseq), codes 0..4; `cigar`  */
/* is query->cigar->s (M/I/D runs only, soft clips are printed elsewhere).   */
/* Appends exactly what the reference appends to the SAM line; returns the   */
/* text length, or -3 where the reference's assert(ref_pos < l_pac) fires.   */
/* Quirks kept: no "0" between adjacent mismatches, a deletion flushes the   */
/* match run, XV lists read offsets (relative to seq_start) of mismatching   */
/* bases that hit a non-reference SNP allele, at most 64 of them.            */
/* ------------------------------------------------------------------------ */
static inline unsigned orc_pac(const uint8_t *pac, uint32_t p) { return pac[p >> 2] >> ((~p & 3) << 1) & 3; }   /* sam.c:244 */

ORC_EXPORT int orc_md_nm(const uint32_t *mixref, const uint8_t *pac, uint32_t l_pac, const uint8_t *seq, int l_seq,
                         uint32_t pos, uint32_t seq_start, const char *cigar, char *out, int cap)
{
    if (pos == 0xFFFFFFFFu) { if (cap > 0) out[0] = 0; return 0; }        /* sam.c:248 */
    char buf[8192]; int o = 0;
#define MD_PUT(...) do { o += snprintf(buf + o, sizeof buf - (size_t)o, __VA_ARGS__); } while (0)
    int nm = 0, n_match = 0, n_rs = 0, rs[64];
    uint32_t ref_pos = pos;
    int si = (int)seq_start;                                               /* index into seq */
    MD_PUT("\tMD:Z:");
    const char *c = cigar;
    while (*c) {
        char *end; long n = strtol(c, &end, 10); c = end;
        const char op = *c;
        if (op == 'M') {
            for (long i = 0; i < n; ++i) {
                if (ref_pos >= l_pac) return -3;                           /* sam.c:270 assert */
                const unsigned bt = orc_pac(pac, ref_pos);
                const unsigned rd = (si >= 0 && si < l_seq) ? seq[si] : 4u;
                if (bt == rd) n_match += 1;
                else {
                    const unsigned meta = orc_nib(mixref, ref_pos);
                    if ((meta & (1u << rd)) != 0 && n_rs < 64) rs[n_rs++] = si - (int)seq_start;   /* sam.c:281-286 */
                    nm += 1;
                    if (n_match != 0) MD_PUT("%d", n_match);
                    n_match = 0;
                    MD_PUT("%c", "ACGTN"[bt]);
                }
                ref_pos += 1; si += 1;
            }
        } else if (op == 'I') { nm += (int)n; si += (int)n; }
        else if (op == 'D') {
            if (n_match != 0) MD_PUT("%d", n_match);
            n_match = 0; nm += (int)n;
            MD_PUT("^");
            for (long i = 0; i < n; ++i) { MD_PUT("%c", "ACGTN"[orc_pac(pac, ref_pos)]); ref_pos += 1; }
        }
        if (*c) c += 1;
    }
    if (n_match != 0) MD_PUT("%d", n_match);
    MD_PUT("\tNM:i:%u", (unsigned)nm);
    if (n_rs > 0) {
        MD_PUT("\tXV:i:");
        for (int i = 0; i < n_rs; ++i) { if (i) MD_PUT(","); MD_PUT("%d", rs[i]); }
    }
#undef MD_PUT
    if (cap > 0) { int m = o < cap - 1 ? o : cap - 1; memcpy(out, buf, (size_t)m); out[m] = 0; }
    return o;
}

static uint32_t orc_gen_mapq(uint32_t b0, uint32_t b1)       /* query.c:270-281 */
{
    if (b0 == 0) return 0;
    uint32_t mapq = (uint32_t)(255.0 * ((double)abs((int)(b0 - b1)) / (double)b0));
    return mapq < 254 ? mapq : 254;
}

/* One stage (one strand) of nogap / withgap checking.  Returns max_diff after
 * the stage, or -1 (NO_MATCH).  hits[] must have room for n entries. */
static int orc_stage(const uint32_t *mixref, uint32_t l_mref, const uint8_t *seq, uint32_t l_seq,
                     const uint32_t *loci, uint32_t n, int max_diff, int strand, int gapped,
                     orc_verify_t *q, orc_hit_t *hits, int *n_hits)
{
    int matched = 0;
    uint32_t last = (uint32_t)-1;
    for (uint32_t i = 0; i < n; ++i) {
        uint32_t pos = loci[i];
        if (pos == last) continue;
        if (!gapped) { if (pos >= l_mref) continue; }          /* alnse.c:762 */
        else { if (pos + l_seq + 4 >= l_mref) continue; }      /* alnse.c:894 (uint32 arithmetic) */
        int nd = gapped ? orc_ed_diff(mixref, l_mref, pos, l_seq + 4, seq, l_seq, max_diff)
                        : orc_ed_mismatch(mixref, pos, seq, l_seq, max_diff);
        if (nd >= 0) {
            if (nd < max_diff || !matched) {
                max_diff = nd;
                q->is_gap = (uint8_t)gapped; q->n_diff = (uint8_t)nd; q->strand = strand; q->pos = pos;
            }
            matched = 1;
            hits[*n_hits].pos = pos; hits[*n_hits].n_diff = (uint8_t)nd;
            hits[*n_hits].is_gap = (uint8_t)gapped; hits[*n_hits].strand = (uint16_t)strand;
            ++*n_hits;
        }
        last = pos;
    }
    return matched ? max_diff : -1;
}

/* Verification of one read given its sorted candidate loci per strand.
 * nogap_T0 = 3 (alnse.c:1016,1079); lv_T0 = l_seq/10 for SE (alnse.c:1090), 3 for PE (alnse.c:1027).
 * hits0/hits1 receive aux->hits per strand; alt0/alt1 receive query->hits[strand]. */
ORC_EXPORT void orc_verify_read(const uint32_t *mixref, uint32_t l_mref,
                                const uint8_t *seq, const uint8_t *rseq, uint32_t l_seq,
                                const uint32_t *loci0, uint32_t n0, const uint32_t *loci1, uint32_t n1,
                                int nogap_T0, int lv_T0, int max_hits,
                                orc_verify_t *q, orc_hit_t *hits0, orc_hit_t *hits1,
                                orc_hit_t *alt0, orc_hit_t *alt1)
{
    memset(q, 0, sizeof *q);
    q->pos = 0xFFFFFFFFu; q->strand = 3; q->n_diff = 255; q->is_gap = 255; q->b0 = -1; q->b1 = -1;
    int max_diff = nogap_T0;
    int m0 = orc_stage(mixref, l_mref, seq, l_seq, loci0, n0, max_diff, 0, 0, q, hits0, &q->n_hits[0]);
    if (m0 != -1 && m0 < max_diff) max_diff = m0;
    int m1 = orc_stage(mixref, l_mref, rseq, l_seq, loci1, n1, max_diff, 1, 0, q, hits1, &q->n_hits[1]);
    if (m1 != -1 && m1 < max_diff) max_diff = m1;
    if (m0 == -1 && m1 == -1) {
        max_diff = lv_T0;
        int d0 = orc_stage(mixref, l_mref, seq, l_seq, loci0, n0, max_diff, 0, 1, q, hits0, &q->n_hits[0]);
        if (d0 != -1 && d0 < max_diff) max_diff = d0;
        (void)orc_stage(mixref, l_mref, rseq, l_seq, loci1, n1, max_diff, 1, 1, q, hits1, &q->n_hits[1]);
    }
    /* query_set_hits (query.c:297-333), including the a->n_diff quirk (:317-318) */
    uint32_t primary = q->pos;
    int tot = 0;
    q->b0 = q->n_diff; q->b1 = 100000;
    for (int s = 0; s < 2; ++s) {
        const orc_hit_t *a = s == 0 ? hits0 : hits1;
        orc_hit_t *dst = s == 0 ? alt0 : alt1;
        int n = q->n_hits[s];
        uint32_t last_pos = (uint32_t)-1;       /* never updated in the reference */
        for (int j = 0; j < n; ++j) {
            uint32_t pos = a[j].pos;
            if (pos == last_pos || pos == primary) continue;
            if (a[0].n_diff <= q->n_diff) {
                if (a[0].n_diff <= q->b1) q->b1 = a[0].n_diff;
                dst[q->n_alt[s]++] = a[j];
                ++tot;
            }
            if (tot == max_hits) goto done;
        }
    }
done:
    q->mapq = orc_gen_mapq((uint32_t)q->b0, (uint32_t)q->b1);
}

/* ------------------------------------------------------------------------ */
/* Threaded batch drivers for the CPU baseline (bench.py cpu_baseline and    */
/* --impl reference).  The per-pair functions are passed in as pointers so   */
/* the same loop can time either this file's restatement ("port") or the     */
/* reference's own compiled functions from oracle/_ref ("reference").  The   */
/* loop itself is the reference's: one worker per thread, reads dealt        */
/* round-robin (alnse.c:1316-1350 alnse_core1: i % n_threads == tid).        */
/* ------------------------------------------------------------------------ */
#include <pthread.h>
#include <time.h>

typedef int (*orc_mm_fn)(const uint32_t *, uint32_t, const uint8_t *, uint32_t, int);
typedef int (*orc_diff_fn)(const uint32_t *, uint32_t, uint32_t, uint32_t, const uint8_t *, uint32_t, int);
typedef int (*orc_cig_fn)(const uint32_t *, uint32_t, uint32_t, const uint8_t *, uint32_t, int, char *, int, int, int);

typedef struct {
    const uint32_t *mixref; uint32_t l;
    const uint8_t *codes; const uint32_t *roffs; uint32_t n_reads;
    const uint32_t *offs[2]; const uint32_t *loci[2];
    int nogap_T0, lv_T0;
    orc_mm_fn mm; orc_diff_fn df; orc_cig_fn cg;
    orc_verify_t *out; int8_t *acc[2]; char *cigars; int cigar_stride;
    int tid, n_threads;
} vb_job_t;

static int vb_stage(const vb_job_t *J, const uint8_t *seq, uint32_t l_seq, int s, uint32_t r, int max_diff, int gapped, orc_verify_t *q)
{
    int matched = 0;
    uint32_t last = (uint32_t)-1;
    for (uint32_t i = J->offs[s][r]; i < J->offs[s][r + 1]; ++i) {
        uint32_t pos = J->loci[s][i];
        if (pos == last || (gapped ? (uint32_t)(pos + l_seq + 4) >= J->l : pos >= J->l)) { if (J->acc[s]) J->acc[s][i] = -1; continue; }
        int nd = gapped ? J->df(J->mixref, J->l, pos, l_seq + 4, seq, l_seq, max_diff) : J->mm(J->mixref, pos, seq, l_seq, max_diff);
        if (nd >= 0) {
            if (nd < max_diff || !matched) { max_diff = nd; q->is_gap = (uint8_t)gapped; q->n_diff = (uint8_t)nd; q->strand = s; q->pos = pos; }
            matched = 1; q->n_hits[s]++;
        }
        if (J->acc[s]) J->acc[s][i] = (int8_t)nd;
        last = pos;
    }
    return matched ? max_diff : -1;
}

static void *vb_worker(void *arg)
{
    const vb_job_t *J = (const vb_job_t *)arg;
    uint8_t rbuf[2048];
    for (uint32_t r = (uint32_t)J->tid; r < J->n_reads; r += (uint32_t)J->n_threads) {
        const uint8_t *seq = J->codes + J->roffs[r];
        uint32_t L = J->roffs[r + 1] - J->roffs[r];
        for (uint32_t i = 0; i < L; ++i) { uint8_t c = seq[L - 1 - i]; rbuf[i] = c < 4 ? (uint8_t)(3 - c) : c; }
        orc_verify_t *q = J->out + r;
        memset(q, 0, sizeof *q);
        q->pos = 0xFFFFFFFFu; q->strand = 3; q->n_diff = 255; q->is_gap = 255;
        int max_diff = J->nogap_T0;
        int m0 = vb_stage(J, seq, L, 0, r, max_diff, 0, q);
        if (m0 != -1 && m0 < max_diff) max_diff = m0;
        int m1 = vb_stage(J, rbuf, L, 1, r, max_diff, 0, q);
        if (m0 == -1 && m1 == -1) {
            max_diff = J->lv_T0 >= 0 ? J->lv_T0 : (int)L / 10;
            int d0 = vb_stage(J, seq, L, 0, r, max_diff, 1, q);
            if (d0 != -1 && d0 < max_diff) max_diff = d0;
            (void)vb_stage(J, rbuf, L, 1, r, max_diff, 1, q);
        }
        if (q->is_gap == 1 && J->cigars && J->cg)      /* query_gen_cigar, query.c:282-295 */
            J->cg(J->mixref, q->pos, L + 4, q->strand ? rbuf : seq, L, q->n_diff,
                  J->cigars + (size_t)r * J->cigar_stride, J->cigar_stride, 1, 0);
    }
    return NULL;
}

/* Returns wall seconds of the threaded region. */
ORC_EXPORT double orc_verify_batch(const uint32_t *mixref, uint32_t l, const uint8_t *codes, const uint32_t *roffs,
                                   uint32_t n_reads, const uint32_t *offs0, const uint32_t *loci0,
                                   const uint32_t *offs1, const uint32_t *loci1, int nogap_T0, int lv_T0,
                                   void *mm, void *df, void *cg, int n_threads,
                                   orc_verify_t *out, int8_t *acc0, int8_t *acc1, char *cigars, int cigar_stride)
{
    if (n_threads < 1) n_threads = 1;
    vb_job_t *jobs = (vb_job_t *)calloc((size_t)n_threads, sizeof(vb_job_t));
    pthread_t *th = (pthread_t *)calloc((size_t)n_threads, sizeof(pthread_t));
    struct timespec a, b;
    clock_gettime(CLOCK_MONOTONIC, &a);
    for (int t = 0; t < n_threads; ++t) {
        vb_job_t *J = jobs + t;
        J->mixref = mixref; J->l = l; J->codes = codes; J->roffs = roffs; J->n_reads = n_reads;
        J->offs[0] = offs0; J->offs[1] = offs1; J->loci[0] = loci0; J->loci[1] = loci1;
        J->nogap_T0 = nogap_T0; J->lv_T0 = lv_T0;
        J->mm = mm ? (orc_mm_fn)mm : orc_ed_mismatch;
        J->df = df ? (orc_diff_fn)df : orc_ed_diff;
        J->cg = (orc_cig_fn)cg;
        J->out = out; J->acc[0] = acc0; J->acc[1] = acc1; J->cigars = cigars; J->cigar_stride = cigar_stride;
        J->tid = t; J->n_threads = n_threads;
        pthread_create(&th[t], NULL, vb_worker, J);
    }
    for (int t = 0; t < n_threads; ++t) pthread_join(th[t], NULL);
    clock_gettime(CLOCK_MONOTONIC, &b);
    free(jobs); free(th);
    return (double)(b.tv_sec - a.tv_sec) + 1e-9 * (double)(b.tv_nsec - a.tv_nsec);
}

/* SSW batch: ssw_init + ssw_align + destroys per task through function pointers (the
 * reference's own, from oracle/_ref), or this file's orc_ssw_align when they are NULL. */
typedef void *(*orc_sswinit_fn)(const int8_t *, int32_t, const int8_t *, int32_t, int8_t);
typedef void *(*orc_sswalign_fn)(const void *, const int8_t *, int32_t, uint8_t, uint8_t, uint8_t, uint16_t, int32_t, int32_t);
typedef void (*orc_free_fn)(void *);
typedef struct {
    const uint32_t *mixref; const uint8_t *codes; int L; const uint32_t *starts; const uint32_t *ends; uint32_t n;
    const int8_t *mat; int gapO, gapE;
    orc_sswinit_fn init; orc_sswalign_fn align; orc_free_fn idestroy, adestroy;
    uint64_t *checksum; int tid, n_threads;
} sb_job_t;

static void *sb_worker(void *arg)
{
    const sb_job_t *J = (const sb_job_t *)arg;
    int8_t ref[4096], read[2048];
    uint64_t sum = 0;
    for (uint32_t t = (uint32_t)J->tid; t < J->n; t += (uint32_t)J->n_threads) {
        int l_ref = (int)(J->ends[t] - J->starts[t] + 1);
        if (l_ref > 4096) l_ref = 4096;
        for (int i = 0; i < l_ref; ++i) ref[i] = (int8_t)orc_nib(J->mixref, J->starts[t] + (uint32_t)i);
        for (int i = 0; i < J->L; ++i) read[i] = (int8_t)(1 << J->codes[(size_t)t * J->L + i]);
        if (J->init) {
            void *p = J->init(read, J->L, J->mat, 16, 1);
            uint16_t *res = (uint16_t *)J->align(p, ref, l_ref, (uint8_t)J->gapO, (uint8_t)J->gapE, 2, 0, 20, J->L / 2);
            sum += res[0];
            J->adestroy(res); J->idestroy(p);
        } else {
            orc_align_t r; uint32_t cig[256];
            orc_ssw_align(read, J->L, J->mat, 16, ref, l_ref, J->gapO, J->gapE, 2, 0, 20, J->L / 2, &r, cig, 256);
            sum += r.score1;
        }
    }
    J->checksum[J->tid] = sum;
    return NULL;
}

ORC_EXPORT double orc_ssw_batch(const uint32_t *mixref, const uint8_t *codes, int L, const uint32_t *starts,
                                const uint32_t *ends, uint32_t n, const int8_t *mat, int gapO, int gapE,
                                void *init, void *align, void *idestroy, void *adestroy, int n_threads, uint64_t *checksum)
{
    if (n_threads < 1) n_threads = 1;
    sb_job_t *jobs = (sb_job_t *)calloc((size_t)n_threads, sizeof(sb_job_t));
    pthread_t *th = (pthread_t *)calloc((size_t)n_threads, sizeof(pthread_t));
    uint64_t *sums = (uint64_t *)calloc((size_t)n_threads, sizeof(uint64_t));
    struct timespec a, b;
    clock_gettime(CLOCK_MONOTONIC, &a);
    for (int t = 0; t < n_threads; ++t) {
        sb_job_t *J = jobs + t;
        J->mixref = mixref; J->codes = codes; J->L = L; J->starts = starts; J->ends = ends; J->n = n;
        J->mat = mat; J->gapO = gapO; J->gapE = gapE;
        J->init = (orc_sswinit_fn)init; J->align = (orc_sswalign_fn)align;
        J->idestroy = (orc_free_fn)idestroy; J->adestroy = (orc_free_fn)adestroy;
        J->checksum = sums; J->tid = t; J->n_threads = n_threads;
        pthread_create(&th[t], NULL, sb_worker, J);
    }
    for (int t = 0; t < n_threads; ++t) pthread_join(th[t], NULL);
    clock_gettime(CLOCK_MONOTONIC, &b);
    uint64_t tot = 0;
    for (int t = 0; t < n_threads; ++t) tot += sums[t];
    if (checksum) *checksum = tot;
    free(jobs); free(th); free(sums);
    return (double)(b.tv_sec - a.tv_sec) + 1e-9 * (double)(b.tv_nsec - a.tv_nsec);
}

/* orc_ed_diff_withcigar with the reference's 10-argument signature (editdistance.h:22), so the
 * batch driver can call either implementation through one pointer type. */
ORC_EXPORT int orc_ed_diff_withcigar_ref_abi(const uint32_t *mixref, uint32_t ref_st, uint32_t l_ref,
                                             const uint8_t *seq, uint32_t l_seq, int max_k_diff,
                                             char *cigarBuf, int cigarLen, int useM, int fmt)
{
    (void)useM; (void)fmt;
    return orc_ed_diff_withcigar(mixref, ref_st, l_ref, seq, l_seq, max_k_diff, cigarBuf, cigarLen);
}
