/* tools/micro/sam_text_bench.c -- MEASUREMENT HELPER (host only).  One SAM line per call: salt_sam_se (include/salt_host.h)
 * against the reference's aln_samse through oracle/_ref/libsaltref_sam.so (ref_sam_se, oracle/dropin/sam_harness.c, which also
 * builds the reference's structs per call -- a few hundred ns of the figure it gets), same read, tags supplied / recomputed.
 *   gcc -O2 -I include -o /tmp/sam_text_bench tools/micro/sam_text_bench.c -L salt_b200 -l:libsalt_host.so -l:libsalt_b200.so \
 *       -L oracle/_ref -l:libsaltref_sam.so -Wl,-rpath,$PWD/salt_b200 -Wl,-rpath,$PWD/oracle/_ref && /tmp/sam_text_bench */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include "salt_host.h"

typedef struct {
    const char *name; const uint8_t *seq, *rseq; const char *qual; int l_seq;
    uint32_t pos; int strand; uint32_t mapq; const char *cigar; uint32_t seq_start, seq_end;
    int n_alt[2]; const uint32_t *alt[2];
} ref_sam_read_t;
int ref_sam_se(const uint32_t *mixref, uint32_t l, const uint8_t *pac, int n_seqs, const char *const *names, const int64_t *offsets,
               const ref_sam_read_t *r, int print_xa_cigar, int print_nm_md, const char *rg_id, char *out, int cap);

static double now(void) { struct timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return t.tv_sec + 1e-9 * t.tv_nsec; }

int main(void)
{
    enum { L = 100, GL = 100000, N = 400000 };
    static uint32_t mixref[GL / 8 + 64]; static uint8_t pac[GL / 4 + 64];
    uint8_t codes[GL];
    srand(7);
    for (int i = 0; i < GL; ++i) { codes[i] = (uint8_t)(rand() & 3); mixref[i >> 3] |= (1u << codes[i]) << (4 * (i & 7)); pac[i >> 2] |= (uint8_t)(codes[i] << ((~i & 3) << 1)); }
    uint8_t seq[L], rseq[L]; char qual[L + 1];
    const uint32_t pos = 4321;
    for (int i = 0; i < L; ++i) { seq[i] = codes[pos + i]; qual[i] = (char)(40 + i % 30); }
    seq[17] = (uint8_t)((seq[17] + 1) & 3); seq[60] = (uint8_t)((seq[60] + 2) & 3);
    for (int i = 0; i < L; ++i) rseq[i] = (uint8_t)(3 - seq[L - 1 - i]);
    qual[L] = 0;
    const char *names[2] = {"chr1", "chr2"}; const int64_t offsets[2] = {0, 50000};
    salt_sam_refs_t refs = {2, names, offsets, GL};
    salt_hit_t alt0[1] = {{70000, 2, 0, 0}};
    uint16_t xv[1] = {17};
    salt_sam_read_t q; memset(&q, 0, sizeof q);
    q.name = "read_000123"; q.seq = seq; q.qual = qual; q.l_seq = L; q.pos = pos; q.strand = 0; q.mapq = 37; q.cigar = "100M";
    q.seq_end = L - 1; q.n_alt[0] = 1; q.alt[0] = alt0; q.md = "17A42C39"; q.nm = 2; q.xv = xv; q.n_xv = 0;
    char out[1024], out2[1024];
    ref_sam_read_t r; memset(&r, 0, sizeof r);
    r.name = q.name; r.seq = seq; r.rseq = rseq; r.qual = qual; r.l_seq = L; r.pos = pos; r.strand = 0; r.mapq = 37; r.cigar = "100M"; r.seq_end = L - 1;
    uint32_t a0[3] = {70000, 2, 0}; r.n_alt[0] = 1; r.alt[0] = a0;
    /* the tags as the reference computes them for this read (the tail kernels' job in the product) */
    static char md[256];
    ref_sam_se(mixref, GL, pac, 2, names, offsets, &r, 1, 1, "grp", out2, sizeof out2);
    { const char *m = strstr(out2, "MD:Z:"); size_t k = 0; for (m += 5; *m && *m != '\t'; ++m) md[k++] = *m; md[k] = 0;
      q.md = md; q.nm = (uint32_t)atoi(strstr(out2, "NM:i:") + 5); q.n_xv = strstr(out2, "XV:i:") ? 1 : 0;
      if (q.n_xv) xv[0] = (uint16_t)atoi(strstr(out2, "XV:i:") + 5); }
    double t0 = now(); long tot = 0;
    for (int i = 0; i < N; ++i) tot += salt_sam_se(&refs, &q, 1, "grp", out, sizeof out);
    const double mine = now() - t0;
    t0 = now(); long tot2 = 0;
    for (int i = 0; i < N; ++i) tot2 += ref_sam_se(mixref, GL, pac, 2, names, offsets, &r, 1, 1, "grp", out2, sizeof out2);
    const double theirs = now() - t0;
    printf("{\"lines\": %d, \"line_bytes\": %ld, \"identical\": %s, \"salt_sam_se_ns_per_line\": %.0f, \"aln_samse_ns_per_line\": %.0f, "
           "\"note\": \"aln_samse recomputes MD/NM (sam_add_md_nm) inside the call; salt_sam_se prints the tags the tail kernels delivered\"}\n",
           N, tot / N, strcmp(out, out2) == 0 ? "true" : "false", mine / N * 1e9, theirs / N * 1e9);
    return 0;
}
