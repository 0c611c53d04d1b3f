/* oracle/dropin/sam_harness.c -- TEST INFRASTRUCTURE ONLY.
 *
 * Calls the reference's own sam_add_md_nm (Align_src/sam.c:246-328, compiled unmodified where it lies
 * under /root/reference, see oracle/Makefile -> _ref/libsaltref_sam.so) on one alignment described by
 * plain buffers, so that tests can pin oracle.c's restatement and the CUDA kernel against it.
 * The structs are the reference's (index_t indexio.h:26-33, query_t query.h:37-63); only the fields
 * sam_add_md_nm reads are filled in. */
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
#include "aln.h"
#include "sam.h"
#include "kstring.h"
#include "kvec.h"

void sam_add_md_nm(kstring_t *s, index_t *index, query_t *q);

/* returns the length of the text the reference appended ("\tMD:Z:...\tNM:i:..[\tXV:i:..]"), copied to out */
int ref_sam_md_nm(const uint32_t *mixref, uint32_t l, const uint8_t *pac, const uint8_t *seq, const uint8_t *rseq,
                  int l_seq, uint32_t pos, int strand, uint32_t seq_start, const char *cigar, char *out, int out_cap)
{
    bntann1_t ann; memset(&ann, 0, sizeof ann);
    ann.offset = 0; ann.len = (int32_t)l; ann.name = (char *)"ref"; ann.anno = (char *)"";
    bntseq_t bns; memset(&bns, 0, sizeof bns);
    bns.l_pac = l; bns.n_seqs = 1; bns.anns = &ann; bns.n_holes = 0; bns.ambs = NULL;
    mixRef_t mr; mr.seq = (uint32_t *)mixref; mr.l = l;
    index_t index; memset(&index, 0, sizeof index);
    index.mixRef = &mr; index.bntseq = &bns; index.pac = (uint8_t *)pac;
    kstring_t cg; memset(&cg, 0, sizeof cg);
    kputs(cigar, &cg);
    query_t q; memset(&q, 0, sizeof q);
    q.l_seq = l_seq; q.seq = (uint8_t *)seq; q.rseq = (uint8_t *)rseq; q.pos = pos; q.strand = strand;
    q.seq_start = seq_start; q.seq_end = (uint32_t)l_seq - 1; q.cigar = &cg;
    kstring_t s; memset(&s, 0, sizeof s);
    sam_add_md_nm(&s, &index, &q);
    int n = (int)s.l;
    if (out_cap > 0) {
        int m = n < out_cap - 1 ? n : out_cap - 1;
        if (m > 0) memcpy(out, s.s, (size_t)m);
        out[m > 0 ? m : 0] = '\0';
    }
    free(s.s); free(cg.s);
    return n;
}

/* ---- the reference's own line formatters, aln_samse (sam.c:86-180) and alnpe_sam (sam.c:331-455), on reads described by plain
 * buffers: the oracle of salt_sam_se / salt_sam_pe.  Alternates come as (pos, n_diff, is_gap) triples per strand. ---------- */
typedef struct {
    const char *name; const uint8_t *seq, *rseq; const char *qual; int l_seq;
    uint32_t pos; int strand; uint32_t mapq; const char *cigar; uint32_t seq_start, seq_end;
    int n_alt[2]; const uint32_t *alt[2];
} ref_sam_read_t;

static void fill_query(query_t *q, kstring_t *cg, kstring_t *sam, const ref_sam_read_t *r)
{
    int s, i;
    memset(q, 0, sizeof *q); memset(cg, 0, sizeof *cg); memset(sam, 0, sizeof *sam);
    q->name = (char *)r->name; q->l_seq = r->l_seq; q->seq = (uint8_t *)r->seq; q->rseq = (uint8_t *)r->rseq; q->qual = (uint8_t *)r->qual;
    q->pos = r->pos; q->strand = r->strand; q->mapq = (uint8_t)r->mapq; q->seq_start = r->seq_start; q->seq_end = r->seq_end;
    kputs(r->cigar ? r->cigar : "", cg); q->cigar = cg; q->sam = sam;
    for (s = 0; s < 2; ++s)
        for (i = 0; i < r->n_alt[s]; ++i) {
            hit_t h; h.pos = r->alt[s][3 * i]; h.n_diff = (uint8_t)r->alt[s][3 * i + 1]; h.is_gap = (uint8_t)r->alt[s][3 * i + 2]; h.strand = (uint16_t)s;
            kv_push(hit_t, q->hits[s], h);
        }
}

static void fill_index(index_t *index, bntseq_t *bns, bntann1_t *anns, mixRef_t *mr, const uint32_t *mixref, uint32_t l, const uint8_t *pac,
                       int n_seqs, const char *const *names, const int64_t *offsets)
{
    int i;
    memset(index, 0, sizeof *index); memset(bns, 0, sizeof *bns);
    for (i = 0; i < n_seqs; ++i) {
        memset(&anns[i], 0, sizeof anns[i]);
        anns[i].offset = offsets[i]; anns[i].name = (char *)names[i]; anns[i].anno = (char *)"";
        anns[i].len = (int32_t)((i + 1 < n_seqs ? offsets[i + 1] : (int64_t)l) - offsets[i]);
    }
    bns->l_pac = l; bns->n_seqs = n_seqs; bns->anns = anns; bns->n_holes = 0; bns->ambs = NULL;
    mr->seq = (uint32_t *)mixref; mr->l = l;
    index->mixRef = mr; index->bntseq = bns; index->pac = (uint8_t *)pac;
}

static int take(kstring_t *sam, char *out, int cap)
{
    int n = (int)sam->l;
    int m = n < cap - 1 ? n : cap - 1;
    if (m > 0) memcpy(out, sam->s, (size_t)m);
    if (cap > 0) out[m > 0 ? m : 0] = '\0';
    free(sam->s);
    return n;
}

int ref_sam_se(const uint32_t *mixref, uint32_t l, const uint8_t *pac, int n_seqs, const char *const *names, const int64_t *offsets,
               const ref_sam_read_t *r, int print_xa_cigar, int print_nm_md, const char *rg_id, char *out, int cap)
{
    index_t index; bntseq_t bns; mixRef_t mr; bntann1_t anns[64];
    query_t q; kstring_t cg, sam;
    aln_opt_t opt;
    if (n_seqs > 64) return -1;
    fill_index(&index, &bns, anns, &mr, mixref, l, pac, n_seqs, names, offsets);
    fill_query(&q, &cg, &sam, r);
    memset(&opt, 0, sizeof opt);
    opt.print_xa_cigar = print_xa_cigar; opt.print_nm_md = print_nm_md; opt.rg_id = (char *)rg_id;
    aln_samse(&index, &q, &opt);
    free(cg.s); kv_destroy(q.hits[0]); kv_destroy(q.hits[1]);
    return take(&sam, out, cap);
}

int ref_sam_pe(const uint32_t *mixref, uint32_t l, const uint8_t *pac, int n_seqs, const char *const *names, const int64_t *offsets,
               const ref_sam_read_t *r /* two */, uint32_t min_tlen, uint32_t max_tlen, int print_xa_cigar, int print_nm_md, const char *rg_id,
               char *out0, int cap0, char *out1, int cap1, int len[2])
{
    index_t index; bntseq_t bns; mixRef_t mr; bntann1_t anns[64];
    query_t q[2]; kstring_t cg[2], sam[2];
    aln_opt_t opt;
    int i;
    if (n_seqs > 64) return -1;
    fill_index(&index, &bns, anns, &mr, mixref, l, pac, n_seqs, names, offsets);
    for (i = 0; i < 2; ++i) fill_query(&q[i], &cg[i], &sam[i], &r[i]);
    memset(&opt, 0, sizeof opt);
    opt.print_xa_cigar = print_xa_cigar; opt.print_nm_md = print_nm_md; opt.rg_id = (char *)rg_id;
    opt.min_tlen = min_tlen; opt.max_tlen = max_tlen;
    alnpe_sam(&index, q, &opt);
    for (i = 0; i < 2; ++i) { free(cg[i].s); kv_destroy(q[i].hits[0]); kv_destroy(q[i].hits[1]); }
    len[0] = take(&sam[0], out0, cap0);
    len[1] = take(&sam[1], out1, cap1);
    return 0;
}
