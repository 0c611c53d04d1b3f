/*
 * level0_shim.c -- the reference's own per-call symbols as batch-of-one calls into libsalt_b200.so (SURVEY section 8b,
 * "Level 0"): a drop-in PROOF, not the fast path -- every call is at least one PCIe round trip.
 *
 *   ed_mismatch / ed_diff / ed_diff_withcigar            editdistance.h:20-22   (call sites alnse.c:349, :373; query.c:288; sam.c:218)
 *   computeEditDistance / computeEditDistanceWithCigar   LandauVishkin.h:45, :50 (called by ed_diff / ed_diff_withcigar)
 *   ssw_init / ssw_align / init_destroy / align_destroy  ssw.h:71, :111, :76, :124 (call sites alnpe.c:284-325, :344-391)
 *
 * With these nine symbols the reference's unmodified objects link without editdistance.o, LandauVishkin.o and ssw.o
 * (oracle/Makefile builds that program as oracle/_ref/salt_level0; tests/test_dropin.py compares its SAM).
 *
 * Threads: the reference calls these from its -t workers concurrently (alnse.c:1423-1428).  A handle is single-threaded,
 * so every calling thread lazily takes its own: one attached to the process-wide reference (salt_b200_attach shares the
 * resident mixRef) for the ed_* calls, and one small scratch handle for the calls that bring their own reference bytes
 * (computeEditDistance*, ssw_align), reloaded per call.  Both are released when the thread exits.
 *
 * Where does the process-wide reference come from?  Either the program calls salt_level0_attach(handle, l) itself, or
 * it links level0_wrap.c with -Wl,--wrap=mixRef_restore: that wrapper lets the reference load its PREFIX.ref as always
 * (metaref.c:61-93) and hands what it loaded to salt_level0_attach_words.
 */
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "../../include/salt_b200.h"

#define L0_EXPORT __attribute__((visibility("default")))

static salt_b200_t *g_parent;
static int g_parent_owned;
static uint32_t g_l;
static pthread_key_t g_key;
static pthread_once_t g_once = PTHREAD_ONCE_INIT;

typedef struct { salt_b200_t *main, *scratch; } l0_thread_t;

static void l0_thread_free(void *p)
{
    l0_thread_t *t = (l0_thread_t *)p;
    if (!t) return;
    if (t->main) salt_b200_destroy(t->main);
    if (t->scratch) salt_b200_destroy(t->scratch);
    free(t);
}
static void l0_make_key(void) { pthread_key_create(&g_key, l0_thread_free); }

static l0_thread_t *l0_thread(void)
{
    pthread_once(&g_once, l0_make_key);
    l0_thread_t *t = (l0_thread_t *)pthread_getspecific(g_key);
    if (!t) {
        t = (l0_thread_t *)calloc(1, sizeof *t);
        pthread_setspecific(g_key, t);
    }
    return t;
}

static salt_b200_t *l0_main(void)
{
    l0_thread_t *t = l0_thread();
    if (!t || !g_parent) return NULL;
    if (!t->main) t->main = salt_b200_attach(g_parent);
    return t->main;
}

/* the calling thread's scratch handle holding `words` / `pac` as its reference */
static salt_b200_t *l0_scratch(const uint32_t *words, uint32_t l, const uint8_t *pac)
{
    l0_thread_t *t = l0_thread();
    if (!t) return NULL;
    if (!t->scratch) {
        t->scratch = salt_b200_init(words, l, pac, pac ? (int64_t)l : 0, g_parent ? 0 : 0);
        return t->scratch;
    }
    return salt_b200_reload_ref(t->scratch, words, l, pac, pac ? (int64_t)l : 0) == SALT_OK ? t->scratch : NULL;
}

L0_EXPORT void salt_level0_attach(salt_b200_t *h, uint32_t l_mixref) { g_parent = h; g_l = l_mixref; g_parent_owned = 0; }

/* the same from the words themselves: the library keeps the handle (used by level0_wrap.c's mixRef_restore wrapper) */
L0_EXPORT int salt_level0_attach_words(const uint32_t *words, uint32_t l)
{
    if (g_parent) return SALT_OK;
    g_parent = salt_b200_init(words, l, NULL, 0, 0);
    if (!g_parent) { fprintf(stderr, "[salt_level0] %s\n", salt_b200_last_error()); return SALT_ERR_CUDA; }
    g_l = l; g_parent_owned = 1;
    return SALT_OK;
}

static int set_one_read(salt_b200_t *h, const uint8_t *seq, uint32_t l_seq)
{
    uint32_t offs[2] = {0, l_seq};
    salt_reads_t r; r.codes = seq; r.offs = offs; r.n_reads = 1;
    return salt_b200_set_reads(h, &r);
}

/* editdistance.h:20 */
L0_EXPORT int ed_mismatch(const uint32_t *mixRef, uint32_t ref_st, const uint8_t *seq, uint32_t l_comp, int max_err)
{
    (void)mixRef;
    salt_b200_t *h = l0_main();
    salt_pair_t p = {0u, ref_st};
    int8_t out = -1;
    if (!h || set_one_read(h, seq, l_comp) != SALT_OK) return -1;
    if (salt_b200_mismatch(h, &p, 1, max_err > 127 ? 127 : max_err, &out) != SALT_OK) return -1;
    return out;
}

/* editdistance.h:21 -- salt only ever calls it with l_ref = l_seq + 4 (alnse.c:373) */
L0_EXPORT int ed_diff(const uint32_t *mixRef, uint32_t l_mref, uint32_t ref_st, const uint32_t l_ref,
                      const uint8_t *seq, uint32_t l_seq, int max_k_diff)
{
    (void)mixRef; (void)l_mref;
    salt_b200_t *h = l0_main();
    salt_pair_t p = {0u, ref_st};
    int8_t out = -1;
    if (!h || l_ref != l_seq + 4 || max_k_diff < 0) return -1;
    if (set_one_read(h, seq, l_seq) != SALT_OK) return -1;
    if (salt_b200_lv(h, &p, 1, max_k_diff, &out) != SALT_OK) return -1;
    return out;
}

/* editdistance.h:22 -- useM = 1, COMPACT_CIGAR_STRING at every salt call site (query.c:288, sam.c:218) */
L0_EXPORT int ed_diff_withcigar(const uint32_t *mixRef, uint32_t ref_st, uint32_t l_ref, const uint8_t *seq, uint32_t l_seq,
                                int max_k_diff, char *cigarBuf, int cigarLen, int useM, int cigarFormat)
{
    (void)mixRef;
    salt_b200_t *h = l0_main();
    salt_pair_t p = {0u, ref_st};
    int8_t out = -1;
    uint8_t k = (uint8_t)max_k_diff;
    if (!h || l_ref != l_seq + 4 || !useM || cigarFormat != 0 || max_k_diff < 0 || max_k_diff >= 31) return -1;
    if (set_one_read(h, seq, l_seq) != SALT_OK) return -1;
    if (salt_b200_lv_cigar(h, &p, &k, 1, cigarBuf, cigarLen, &out) != SALT_OK) return -1;
    return out;
}

/* ---- LandauVishkin.h:45 / :50: text and pattern arrive as bytes (text: allele masks 0..15, pattern: one-hot 1/2/4/8 or 15
 * for N, editdistance.c:189-218).  Served for the shape salt uses: textLen == patternLen + 4. ---- */
static int lv_bytes_to_engine(const char *text, int textLen, const char *pattern, int patternLen, uint32_t **words, uint8_t **codes)
{
    if (!text || !pattern || patternLen < 1 || patternLen > 1024 || textLen != patternLen + 4) return -1;
    *words = (uint32_t *)calloc(((size_t)textLen + 7) / 8 + 1, 4);
    *codes = (uint8_t *)malloc((size_t)patternLen);
    if (!*words || !*codes) return -1;
    for (int i = 0; i < textLen; ++i) (*words)[i >> 3] |= ((uint32_t)text[i] & 15u) << (4 * (i & 7));
    for (int i = 0; i < patternLen; ++i) {
        switch ((unsigned char)pattern[i]) {
        case 1: (*codes)[i] = 0; break; case 2: (*codes)[i] = 1; break; case 4: (*codes)[i] = 2; break;
        case 8: (*codes)[i] = 3; break; case 15: (*codes)[i] = 4; break;
        default: return -1;                    /* not a read symbol salt produces */
        }
    }
    return 0;
}

L0_EXPORT int computeEditDistance(const char *text, int textLen, const char *pattern, int patternLen, int k)
{
    uint32_t *words = NULL; uint8_t *codes = NULL;
    int8_t out = -1;
    if (lv_bytes_to_engine(text, textLen, pattern, patternLen, &words, &codes) == 0 && k >= 0) {
        salt_b200_t *h = l0_scratch(words, (uint32_t)textLen, NULL);
        salt_pair_t p = {0u, 0u};
        if (!h || set_one_read(h, codes, (uint32_t)patternLen) != SALT_OK || salt_b200_lv(h, &p, 1, k, &out) != SALT_OK) out = -1;
    }
    free(words); free(codes);
    return out;
}

L0_EXPORT int computeEditDistanceWithCigar(const char *text, int textLen, const char *pattern, int patternLen, int k,
                                           char *cigarBuf, int cigarBufLen, int useM, int cigarFormat)
{
    uint32_t *words = NULL; uint8_t *codes = NULL;
    int8_t out = -1;
    if (useM && cigarFormat == 0 && k >= 0 && k < 31 && cigarBuf && cigarBufLen >= 2 &&
        lv_bytes_to_engine(text, textLen, pattern, patternLen, &words, &codes) == 0) {
        salt_b200_t *h = l0_scratch(words, (uint32_t)textLen, NULL);
        salt_pair_t p = {0u, 0u};
        uint8_t kk = (uint8_t)k;
        if (!h || set_one_read(h, codes, (uint32_t)patternLen) != SALT_OK ||
            salt_b200_lv_cigar(h, &p, &kk, 1, cigarBuf, cigarBufLen, &out) != SALT_OK) out = -1;
    }
    free(words); free(codes);
    return out;
}

/* ---- ssw.h: the profile only remembers its (borrowed) arguments (ssw.c:758-759 keeps read / mat pointers too) ---- */
typedef struct { const int8_t *read; int32_t readLen; const int8_t *mat; int32_t n; } l0_profile_t;
/* s_align, ssw.h:37-47 */
typedef struct {
    uint16_t score1, score2;
    int32_t ref_begin1, ref_end1, read_begin1, read_end1, ref_end2;
    uint32_t *cigar;
    int32_t cigarLen;
} l0_align_t;

L0_EXPORT void *ssw_init(const int8_t *read, const int32_t readLen, const int8_t *mat, const int32_t n, const int8_t score_size)
{
    (void)score_size;                          /* every salt caller passes 1: the int16 profile (ssw.c:748,757) */
    l0_profile_t *p = (l0_profile_t *)calloc(1, sizeof *p);
    if (p) { p->read = read; p->readLen = readLen; p->mat = mat; p->n = n; }
    return p;
}

L0_EXPORT void init_destroy(void *p) { free(p); }
L0_EXPORT void align_destroy(void *a) { if (a) { free(((l0_align_t *)a)->cigar); free(a); } }

L0_EXPORT void *ssw_align(const void *prof, const int8_t *ref, int32_t refLen, const uint8_t gapO, const uint8_t gapE,
                          const uint8_t flag, const uint16_t filters, const int32_t filterd, const int32_t maskLen)
{
    const l0_profile_t *p = (const l0_profile_t *)prof;
    if (!p || !ref || refLen < 1 || p->readLen < 1 || p->readLen > 1024 || (p->n != 16 && p->n != 5)) return NULL;
    const int use_pac = p->n == 5;
    uint32_t *words = (uint32_t *)calloc(((size_t)refLen + 7) / 8 + 1, 4);
    uint8_t *pac = (uint8_t *)calloc(((size_t)refLen + 3) / 4 + 1, 1);
    uint8_t *codes = (uint8_t *)malloc((size_t)p->readLen);
    l0_align_t *a = NULL;
    int ok = words && pac && codes;
    for (int i = 0; ok && i < refLen; ++i) {
        const unsigned s = (unsigned char)ref[i];
        if (use_pac) { if (s > 3) ok = 0; else pac[i >> 2] |= (uint8_t)(s << ((~i & 3) << 1)); }      /* alnpe.c:47 */
        else words[i >> 3] |= (s & 15u) << (4 * (i & 7));
    }
    for (int i = 0; ok && i < p->readLen; ++i) {
        const unsigned s = (unsigned char)p->read[i];
        if (use_pac) { if (s > 4) ok = 0; else codes[i] = (uint8_t)s; }
        else switch (s) {                                                                              /* alnpe.c:283: 1 << code */
            case 1: codes[i] = 0; break; case 2: codes[i] = 1; break; case 4: codes[i] = 2; break;
            case 8: codes[i] = 3; break; case 16: codes[i] = 4; break; default: ok = 0;
        }
    }
    if (ok) {
        salt_b200_t *h = l0_scratch(words, (uint32_t)refLen, use_pac ? pac : NULL);
        salt_win_t w = {0u, 0u, (uint32_t)refLen - 1u};
        salt_ssw_out_t o;
        const int stride = 2 * p->readLen + 8;
        uint32_t *cg = (uint32_t *)calloc((size_t)stride, 4);
        if (h && cg && set_one_read(h, codes, (uint32_t)p->readLen) == SALT_OK &&
            salt_b200_ssw(h, &w, 1, use_pac, p->mat, p->n, gapO, gapE, flag, filters, filterd, maskLen, &o, cg, stride) == SALT_OK &&
            o.cigarLen >= 0 && o.cigarLen <= stride) {
            a = (l0_align_t *)calloc(1, sizeof *a);
            if (a) {
                a->score1 = o.score1; a->score2 = o.score2; a->ref_begin1 = o.ref_begin1; a->ref_end1 = o.ref_end1;
                a->read_begin1 = o.read_begin1; a->read_end1 = o.read_end1; a->ref_end2 = o.ref_end2; a->cigarLen = o.cigarLen;
                if (o.cigarLen > 0) {
                    a->cigar = (uint32_t *)malloc((size_t)o.cigarLen * 4);
                    if (a->cigar) memcpy(a->cigar, cg, (size_t)o.cigarLen * 4);
                }
            }
        }
        free(cg);
    }
    free(words); free(pac); free(codes);
    return a;
}
