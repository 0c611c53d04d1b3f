"""Inputs for the end-to-end SAM comparison (tests/test_dropin.py): a small synthetic genome with a
SNP table in salt-idx's 4-column format, and simulated reads as FASTQ."""
import os

import numpy as np

from salt_b200 import synth


def write_inputs(outdir, glen=150_000, n_reads=6000, L=100, seed=5, two_copies=False):
    os.makedirs(outdir, exist_ok=True)
    rng = np.random.default_rng(seed + 991)      # not the genome's own stream
    g = synth.Genome(glen, snp_rate=0.01, n_rate=0.0, seed=seed)
    if two_copies:                               # the second half repeats the first (SNPs included): every read has an alternate
        half = glen // 2
        g.codes[half:2 * half] = g.codes[:half]; g.masks[half:2 * half] = g.masks[:half]
        g.snp_pos = np.flatnonzero((g.masks & (g.masks - 1)) != 0)
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


def write_pe_inputs(outdir, glen=150_000, n_pairs=3000, L=100, seed=9, two_copies=False):
    """Genome + SNP table as write_inputs; mates 1/2 as two FASTQ files, insert ~ N(500, 40) (the bounds of
    run_pe_test.sh are 350..650).  A tenth of the second mates is too divergent to verify and has to be rescued.
    two_copies: the second half of the genome repeats the first, so every mate has an alternate (XA, pairing over alternates)."""
    fa, sn, _ = write_inputs(outdir, glen=glen, n_reads=10, L=L, seed=seed, two_copies=two_copies)
    rng = np.random.default_rng(seed + 17)
    g = synth.Genome(glen, snp_rate=0.01, n_rate=0.0, seed=seed)
    if two_copies:
        half = glen // 2
        g.codes[half:2 * half] = g.codes[:half]
    codes = (g.codes & 3).astype(np.uint8)

    def noisy(seq, rate, indel):
        out = []
        i = 0
        while len(out) < L and i < len(seq):
            u = rng.random()
            if u < rate:
                out.append((int(seq[i]) + int(rng.integers(1, 4))) & 3); i += 1
            elif u < rate + indel:
                if rng.random() < 0.5:
                    out.append(int(rng.integers(0, 4)))
                else:
                    i += 1
            else:
                out.append(int(seq[i])); i += 1
        while len(out) < L:
            out.append(int(rng.integers(0, 4)))
        return np.array(out[:L], np.uint8)

    f1 = open(os.path.join(outdir, "r1.fq"), "w"); f2 = open(os.path.join(outdir, "r2.fq"), "w")
    for i in range(n_pairs):
        ins = int(np.clip(rng.normal(500, 40), 360, 640))
        p = int(rng.integers(0, glen - ins - 20))
        hard = i % 10 == 3
        a = noisy(codes[p:p + L + 12], 0.01, 0.002)
        b = noisy(codes[p + ins - L:p + ins + 12], 0.09 if hard else 0.01, 0.02 if hard else 0.002)
        b = synth.revcomp(b[None, :])[0]
        if rng.random() < 0.5:                       # fragment from the other strand: swap roles
            a, b = b, a
        if i % 97 == 0:
            b = rng.integers(0, 4, L, dtype=np.uint8)   # a mate that maps nowhere
        for f, r in ((f1, a), (f2, b)):
            f.write("@p%d\n%s\n+\n%s\n" % (i, "".join("ACGTN"[c] for c in r), "I" * L))
    f1.close(); f2.close()
    return fa, sn, os.path.join(outdir, "r1.fq"), os.path.join(outdir, "r2.fq")
