/*
 * level0_shim.c -- the reference's own per-pair symbols (editdistance.h:20-22) as batch-of-one calls
 * into libsalt_b200.so.  A drop-in PROOF, not the fast path: every call is one PCIe round trip.
 * It lets the unmodified call sites (alnse.c:349, :373; query.c:288; sam.c:218) consume the
 * engine's results.  The mixRef pointer the reference passes is ignored: the engine compares
 * against the copy uploaded by salt_b200_init, which salt_level0_attach registers.
 *
 * Not thread-safe (one handle, one current read); the reference's worker threads would each need
 * their own handle -- which is exactly why the real integration is the chunk pipeline.
 */
#include <stdint.h>
#include <string.h>
#include "../../include/salt_b200.h"

static salt_b200_t *g_h;
static uint32_t g_l;

#define L0_EXPORT __attribute__((visibility("default")))

L0_EXPORT void salt_level0_attach(salt_b200_t *h, uint32_t l_mixref) { g_h = h; g_l = l_mixref; }

static int set_one_read(const uint8_t *seq, uint32_t l_seq)
{
    uint32_t offs[2] = {0, l_seq};
    salt_reads_t r; r.codes = seq; r.offs = offs; r.n_reads = 1;
    return salt_b200_set_reads(g_h, &r);
}

/* editdistance.h:20 */
L0_EXPORT int ed_mismatch(const uint32_t *mixRef, uint32_t ref_st, const uint8_t *seq, uint32_t l_comp, int max_err)
{
    (void)mixRef;
    salt_pair_t p = {0u, ref_st};
    int8_t out = -1;
    if (!g_h || set_one_read(seq, l_comp) != SALT_OK) return -1;
    if (salt_b200_mismatch(g_h, &p, 1, max_err > 127 ? 127 : max_err, &out) != SALT_OK) return -1;
    return out;
}

/* editdistance.h:21 -- salt only ever calls it with l_ref = l_seq + 4 (alnse.c:373) */
L0_EXPORT int ed_diff(const uint32_t *mixRef, uint32_t l_mref, uint32_t ref_st, const uint32_t l_ref,
                      const uint8_t *seq, uint32_t l_seq, int max_k_diff)
{
    (void)mixRef; (void)l_mref;
    salt_pair_t p = {0u, ref_st};
    int8_t out = -1;
    if (!g_h || l_ref != l_seq + 4 || max_k_diff < 0) return -1;
    if (set_one_read(seq, l_seq) != SALT_OK) return -1;
    if (salt_b200_lv(g_h, &p, 1, max_k_diff, &out) != SALT_OK) return -1;
    return out;
}

/* editdistance.h:22 -- useM = 1, COMPACT_CIGAR_STRING at every salt call site (query.c:288, sam.c:218) */
L0_EXPORT int ed_diff_withcigar(const uint32_t *mixRef, uint32_t ref_st, uint32_t l_ref, const uint8_t *seq, uint32_t l_seq,
                                int max_k_diff, char *cigarBuf, int cigarLen, int useM, int cigarFormat)
{
    (void)mixRef;
    salt_pair_t p = {0u, ref_st};
    int8_t out = -1;
    uint8_t k = (uint8_t)max_k_diff;
    if (!g_h || l_ref != l_seq + 4 || !useM || cigarFormat != 0 || max_k_diff < 0 || max_k_diff >= 31) return -1;
    if (set_one_read(seq, l_seq) != SALT_OK) return -1;
    if (salt_b200_lv_cigar(g_h, &p, &k, 1, cigarBuf, cigarLen, &out) != SALT_OK) return -1;
    return out;
}
