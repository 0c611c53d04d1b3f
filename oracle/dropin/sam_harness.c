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
