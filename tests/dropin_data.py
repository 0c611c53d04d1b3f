"""Inputs for the end-to-end SAM comparison (tests/test_dropin.py): a small synthetic genome with a
SNP table in salt-idx's 4-column format, and simulated reads as FASTQ."""
import os

import numpy as np

from salt_b200 import synth


def write_inputs(outdir, glen=150_000, n_reads=6000, L=100, seed=5):
    os.makedirs(outdir, exist_ok=True)
    rng = np.random.default_rng(seed + 991)      # not the genome's own stream
    g = synth.Genome(glen, snp_rate=0.01, n_rate=0.0, seed=seed)
    # two records, so that coordinate translation (bns_coor_pac2real) is exercised
    cut = glen * 3 // 5
    fasta = g.fasta()
    fa = os.path.join(outdir, "ref.fa")
    with open(fa, "w") as f:
        for name, s in (("chrA", fasta[:cut]), ("chrB", fasta[cut:])):
            f.write(">%s\n" % name)
            for i in range(0, len(s), 60):
                f.write(s[i:i + 60] + "\n")
    sn = os.path.join(outdir, "snps.txt")
    with open(sn, "w") as f:
        for chrom, pos1, alleles, ref in g.snp_table():
            p0 = pos1 - 1
            if p0 < cut:
                f.write("chrA\t%d\t%s\t%s\n" % (pos1, alleles, ref))
            else:
                f.write("chrB\t%d\t%s\t%s\n" % (pos1 - cut, alleles, ref))
    reads, pos, strand = synth.sample_reads(g, n_reads, L, seed=seed + 1, sub_rate=0.012, indel_frac=0.15, n_frac=0.001)
    # a few reads that map nowhere and a few with many N
    junk = rng.integers(0, 4, (40, L), dtype=np.uint8)
    reads = np.concatenate([reads, junk])
    reads[5, 10:40] = 4
    fq = os.path.join(outdir, "reads.fq")
    with open(fq, "w") as f:
        for i, r in enumerate(reads):
            f.write("@r%d\n%s\n+\n%s\n" % (i, "".join("ACGTN"[c] for c in r), "I" * L))
    return fa, sn, fq
