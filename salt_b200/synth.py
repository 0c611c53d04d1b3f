"""Seeded synthetic inputs for the verification hot path.

Everything is numpy and derives from one integer seed:

* genome      -- i.i.d. A/C/G/T codes (optional N runs), as salt's 4-bit allele-mask
                 reference ("mixRef", metaref.h:2-5; layout metaref.c:54-56) and 2-bit pac
                 (bntseq.c:166-250, MSB first, alnpe.c:47)
* SNP set     -- a fraction of positions gets one extra allele OR-ed in (3 % get two),
                 the way Index_src/mixRef.c:147-151 applies the SNP table
* reads       -- uniform start, random strand, donor picks a random allele at SNP sites,
                 substitution errors, optional single indel per read
* candidates  -- per read and strand: the true locus plus decoys (random loci and small
                 shifts of the true locus), sorted and de-duplicated like
                 alnse_locate_alt (alnse.c:633-731) leaves them

No reference code and nothing under oracle/ is used here.
"""
import numpy as np

ONEHOT = np.array([1, 2, 4, 8, 15], np.uint8)


def pack_mixref(masks):
    """uint8 masks (0..15), one per base -> uint32 words, base p in bits 4*(p%8).."""
    n = len(masks)
    pad = (-n) % 8
    m = np.concatenate([masks, np.zeros(pad, np.uint8)]).astype(np.uint32).reshape(-1, 8)
    sh = (4 * np.arange(8, dtype=np.uint32))[None, :]
    return np.bitwise_or.reduce(m << sh, axis=1).astype(np.uint32)


def unpack_mixref(words, start, n):
    idx = np.arange(start, start + n, dtype=np.int64)
    return ((words[idx >> 3] >> (4 * (idx & 7)).astype(np.uint32)) & 15).astype(np.uint8)


def pack_pac(codes):
    """codes 0..3 -> bytes, base p in bits ((~p&3)<<1) of byte p>>2 (MSB first)."""
    n = len(codes)
    pad = (-n) % 4
    c = np.concatenate([codes & 3, np.zeros(pad, np.uint8)]).astype(np.uint8).reshape(-1, 4)
    return (c[:, 0] << 6 | c[:, 1] << 4 | c[:, 2] << 2 | c[:, 3]).astype(np.uint8)


def revcomp(codes):
    """query_seq_reverse(..., is_comp=1) (query.c:46-64): reverse, 3-c for c<4, N stays."""
    r = codes[..., ::-1]
    return np.where(r < 4, 3 - r, r).astype(np.uint8)


class Genome:
    def __init__(self, length, snp_rate=0.01, n_rate=0.0, seed=1, multi_allele=0.03):
        rng = np.random.default_rng(seed)
        self.l = int(length)
        self.codes = rng.integers(0, 4, self.l, dtype=np.uint8)
        masks = (1 << self.codes).astype(np.uint8)
        n_snp = int(self.l * snp_rate)
        self.snp_pos = np.unique(rng.integers(0, self.l, n_snp)) if n_snp else np.zeros(0, np.int64)
        alt = (self.codes[self.snp_pos] + rng.integers(1, 4, len(self.snp_pos), dtype=np.uint8)) & 3
        masks[self.snp_pos] |= (1 << alt).astype(np.uint8)
        two = rng.random(len(self.snp_pos)) < multi_allele
        alt2 = (self.codes[self.snp_pos] + rng.integers(1, 4, len(self.snp_pos), dtype=np.uint8)) & 3
        masks[self.snp_pos[two]] |= (1 << alt2[two]).astype(np.uint8)
        if n_rate > 0:                       # N / IUPAC bases encode as mask 0 (metaref.c:36-53)
            npos = rng.integers(0, self.l, int(self.l * n_rate))
            masks[npos] = 0
        self.masks = masks
        self.mixref = pack_mixref(masks)
        self.pac = pack_pac(self.codes)

    def snp_table(self, chrom="chr1"):
        """The 4-column table Index_src/hapmap.c:92 parses: chrom, 1-based pos, alleles, ref."""
        rows = []
        for p in self.snp_pos:
            m = int(self.masks[p]); ref = "ACGT"[self.codes[p]]
            al = "/".join("ACGT"[b] for b in range(4) if m >> b & 1)
            rows.append((chrom, int(p) + 1, al, ref))
        return rows

    def fasta(self):
        return "".join("ACGT"[c] if m else "N" for c, m in zip(self.codes, self.masks))


class _MaskView:
    """g.masks[idx] for a genome that keeps only the packed words: the allele mask is the mixRef nibble."""

    def __init__(self, words):
        self.words = words

    def __getitem__(self, idx):
        idx = np.asarray(idx, np.int64)
        return ((self.words[idx >> 3] >> (4 * (idx & 7)).astype(np.uint32)) & 15).astype(np.uint8)


class BigGenome:
    """Genome for reference sizes where whole-array temporaries do not fit (BASELINE configs[2]: 3.1 Gbp, snp144-like
    density): generated and packed block by block, same layouts and same SNP rule as Genome; `n_records` equal records
    mimic the 24 FASTA records of GRCh38 (the records are simply concatenated, as bns/mixRef do: l = sum of lengths).
    Keeps codes (1 B/base), mixref (4 bit/base) and pac (2 bit/base); masks are read back from mixref."""

    def __init__(self, length, snp_rate=0.0047, seed=1, multi_allele=0.03, n_records=24, block=1 << 26):
        self.l = int(length)
        self.n_records = n_records
        self.record_len = [self.l // n_records + (1 if i < self.l % n_records else 0) for i in range(n_records)]
        self.codes = np.empty(self.l, np.uint8)
        self.mixref = np.zeros((self.l + 7) // 8, np.uint32)
        self.pac = np.zeros((self.l + 3) // 4, np.uint8)
        n_snp = 0
        for bi, b0 in enumerate(range(0, self.l, block)):          # block is a multiple of 8: whole words and bytes
            b1 = min(self.l, b0 + block)
            rng = np.random.default_rng([seed, bi])
            c = rng.integers(0, 4, b1 - b0, dtype=np.uint8)
            self.codes[b0:b1] = c
            m = np.left_shift(np.uint8(1), c)
            k = int((b1 - b0) * snp_rate)
            if k:
                sp = np.unique(rng.integers(0, b1 - b0, k))
                alt = (c[sp] + rng.integers(1, 4, len(sp), dtype=np.uint8)) & 3
                m[sp] |= np.left_shift(np.uint8(1), alt)
                two = rng.random(len(sp)) < multi_allele
                alt2 = (c[sp] + rng.integers(1, 4, len(sp), dtype=np.uint8)) & 3
                m[sp[two]] |= np.left_shift(np.uint8(1), alt2[two])
                n_snp += len(sp)
            self.mixref[b0 >> 3:(b1 + 7) >> 3] = pack_mixref(m)
            self.pac[b0 >> 2:(b1 + 3) >> 2] = pack_pac(c)
        self.n_snp = n_snp
        self.masks = _MaskView(self.mixref)


def sample_reads(g, n, L, seed=2, sub_rate=0.01, indel_frac=0.0, max_indel=3, n_frac=0.0):
    """Return (reads[n,L] codes in sequencing orientation, true_pos[n], strand[n]).

    true_pos is the leftmost reference coordinate of the forward-strand alignment."""
    rng = np.random.default_rng(seed)
    margin = L + 2 * max_indel + 16
    pos = rng.integers(0, g.l - margin, n)
    strand = rng.integers(0, 2, n, dtype=np.uint8)
    col = np.arange(L)[None, :]
    shift = np.zeros((n, L), np.int64)
    ins_mask = np.zeros((n, L), bool)
    if indel_frac > 0:
        has = rng.random(n) < indel_frac
        at = rng.integers(5, L - 5 - max_indel, n)[:, None]
        ln = rng.integers(1, max_indel + 1, n)[:, None]
        is_del = (rng.random(n) < 0.5)[:, None]
        after = col >= at
        # deletion: columns >= at read from ln bases further on;
        # insertion: columns [at, at+ln) are random bases, the rest shifts back by ln
        ins_mask = has[:, None] & ~is_del & after & (col < at + ln)
        shift = np.where(has[:, None] & ~is_del & (col >= at + ln), -ln, np.where(has[:, None] & is_del & after, ln, 0))
    idx = pos[:, None] + col + shift
    base = g.codes[idx]
    m = g.masks[idx]
    # donor haplotype: at multi-allele sites pick a uniformly random set bit
    multi = (m & (m - 1)) != 0
    if multi.any():
        mm = m[multi].astype(np.int64)
        pick = rng.integers(0, 4, mm.shape)
        # rotate through bit positions until a set allele is found
        out = np.zeros(mm.shape, np.uint8)
        done = np.zeros(mm.shape, bool)
        for k in range(4):
            b = (pick + k) & 3
            ok = ~done & ((mm >> b) & 1).astype(bool)
            out[ok] = b[ok]; done |= ok
        base = base.copy(); base[multi] = out
    reads = base.astype(np.uint8)
    if ins_mask.any():
        reads[ins_mask] = rng.integers(0, 4, int(ins_mask.sum()), dtype=np.uint8)
    if sub_rate > 0:
        e = rng.random((n, L)) < sub_rate
        reads[e] = (reads[e] + rng.integers(1, 4, int(e.sum()), dtype=np.uint8)) & 3
    if n_frac > 0:
        e = rng.random((n, L)) < n_frac
        reads[e] = 4
    rc = revcomp(reads)
    reads = np.where(strand[:, None] == 1, rc, reads).astype(np.uint8)
    return np.ascontiguousarray(reads), pos.astype(np.uint32), strand


def sample_pairs(g, n_pairs, L, seed=6, insert_mean=400, insert_sd=50, sub_rate=0.01, hard_frac=0.05, hard_sub=0.09,
                 junk_frac=0.0):
    """Paired-end reads (insert ~ N(400, 50)): mates interleaved (row 2i = mate 0, 2i+1 = mate 1 of pair i).
    Mate 0 reads the forward strand at p, mate 1 the reverse strand of [p+ins-L, p+ins); half of the pairs come from
    the other strand (roles swapped).  A fraction hard_frac of the second mates carries hard_sub substitutions and an
    indel, so the verification stage cannot place it and the pair needs a Smith-Waterman mate rescue; junk_frac of the
    second mates are random (singleton pairs).  Returns (reads[2n, L], true_pos[2n], strand[2n])."""
    rng = np.random.default_rng(seed)
    ins = np.clip(rng.normal(insert_mean, insert_sd, n_pairs), L + 20, insert_mean + 4 * insert_sd).astype(np.int64)
    p = rng.integers(0, g.l - ins.max() - 16, n_pairs)
    col = np.arange(L)[None, :]
    a = g.codes[p[:, None] + col].astype(np.uint8)
    pos_b = p + ins - L
    b = g.codes[pos_b[:, None] + col].astype(np.uint8)
    for arr in (a, b):
        e = rng.random(arr.shape) < sub_rate
        arr[e] = (arr[e] + rng.integers(1, 4, int(e.sum()), dtype=np.uint8)) & 3
    hard = rng.random(n_pairs) < hard_frac
    if hard.any():
        hb = b[hard]
        e = rng.random(hb.shape) < hard_sub
        hb[e] = (hb[e] + rng.integers(1, 4, int(e.sum()), dtype=np.uint8)) & 3
        cut = rng.integers(L // 3, 2 * L // 3, len(hb))
        for i in range(len(hb)):                       # a 2-base deletion from the read
            hb[i, cut[i]:-2] = hb[i, cut[i] + 2:]
        b[hard] = hb
    junk = rng.random(n_pairs) < junk_frac
    if junk.any():
        b[junk] = rng.integers(0, 4, (int(junk.sum()), L), dtype=np.uint8)
    b = revcomp(b)
    flip = rng.random(n_pairs) < 0.5
    reads = np.empty((2 * n_pairs, L), np.uint8); pos = np.empty(2 * n_pairs, np.uint32); strand = np.empty(2 * n_pairs, np.uint8)
    # fragment from the other strand: mate 0 is the reverse-strand read at the far end, mate 1 the forward read
    reads[0::2] = np.where(flip[:, None], b, a); reads[1::2] = np.where(flip[:, None], a, b)
    pos[0::2] = np.where(flip, pos_b, p); pos[1::2] = np.where(flip, p, pos_b)
    strand[0::2] = flip.astype(np.uint8); strand[1::2] = 1 - flip.astype(np.uint8)
    return np.ascontiguousarray(reads), pos, strand


def edit_rich_reads(g, n, L, n_edits, seed=4):
    """Forward-strand reads carrying about n_edits[i] scattered edits each (substitutions, 1-base
    insertions and deletions at random places) -- the adversarial case for the Landau-Vishkin
    pigeonhole filter, where every edit tries to spoil a different 8-base word.
    Returns (reads[n,L], pos[n]) with strand 0."""
    rng = np.random.default_rng(seed)
    reads = np.zeros((n, L), np.uint8)
    pos = rng.integers(40, g.l - 2 * L - 80, n).astype(np.uint32)
    for i in range(n):
        src = g.codes[int(pos[i]):int(pos[i]) + 2 * L].astype(np.uint8) & 3
        out = []
        j = 0
        where = set(rng.choice(np.arange(2, L - 2), size=min(int(n_edits[i]), L - 4), replace=False).tolist())
        while len(out) < L:
            if len(out) in where:
                where.discard(len(out))
                kind = rng.integers(0, 3)
                if kind == 0:
                    out.append((int(src[j]) + int(rng.integers(1, 4))) & 3); j += 1      # substitution
                elif kind == 1:
                    out.append(int(rng.integers(0, 4)))                                  # insertion in the read
                else:
                    j += 1                                                               # deletion from the read
                    out.append(int(src[j])); j += 1
            else:
                out.append(int(src[j])); j += 1
        reads[i] = out[:L]
    return reads, pos


def make_candidates(g, true_pos, strand, L, per_strand=8, seed=3, shift_frac=0.25, lv_pad=4):
    """CSR candidate lists per strand: (offs0, loci0, offs1, loci1).

    Each read gets `per_strand` loci on each strand: uniformly random decoys, a few
    true±{1..3} shifts, and the true locus on the strand the read came from.  Lists are
    sorted and de-duplicated per read/strand; loci keep pos+L+lv_pad < l."""
    rng = np.random.default_rng(seed)
    n = len(true_pos)
    hi = g.l - L - lv_pad - 1
    out = []
    for s in (0, 1):
        loci = rng.integers(0, hi, (n, per_strand)).astype(np.int64)
        nshift = max(1, int(per_strand * shift_frac))
        sh = rng.integers(1, 4, (n, nshift)) * rng.choice([-1, 1], (n, nshift))
        loci[:, 1:1 + nshift] = np.clip(true_pos[:, None].astype(np.int64) + sh, 0, hi - 1)
        mine = strand == s
        loci[mine, 0] = true_pos[mine]
        loci.sort(axis=1)
        keep = np.ones_like(loci, bool)
        keep[:, 1:] = loci[:, 1:] != loci[:, :-1]
        cnt = keep.sum(axis=1)
        offs = np.zeros(n + 1, np.int64); np.cumsum(cnt, out=offs[1:])
        out += [offs.astype(np.uint32), loci[keep].astype(np.uint32)]
    return tuple(out)


def fuzz_masks(n, seed, snp=0.03, nfrac=0.01):
    """Mask string for fuzzing: one-hot with `snp` two-allele sites and `nfrac` zero (N)."""
    rng = np.random.default_rng(seed)
    c = rng.integers(0, 4, n)
    m = (1 << c).astype(np.uint8)
    s = rng.random(n) < snp
    m[s] |= (1 << ((c[s] + rng.integers(1, 4, int(s.sum()))) & 3)).astype(np.uint8)
    z = rng.random(n) < nfrac
    m[z] = 0
    return m


def fuzz_read_from_masks(masks, L, rng, sub=0.02, indel=0.01, burst=0.0, nfrac=0.005):
    """Walk along `masks`, copying a random allele per site, with substitutions, indels
    (len 1..6), optional random bursts; returns L codes (0..4)."""
    out = []
    i = 0
    n = len(masks)
    while len(out) < L:
        r = rng.random()
        if r < indel / 2:                     # deletion from the read: skip reference bases
            i += int(rng.integers(1, 7)); continue
        if r < indel:                         # insertion into the read
            out += list(rng.integers(0, 4, int(rng.integers(1, 7)))); continue
        if burst and rng.random() < burst:
            k = int(rng.integers(5, 15))
            out += list(rng.integers(0, 4, k)); i += k; continue
        m = int(masks[i]) if i < n else 0
        bits = [b for b in range(4) if m >> b & 1]
        c = int(rng.choice(bits)) if bits else int(rng.integers(0, 4))
        if rng.random() < sub:
            c = (c + int(rng.integers(1, 4))) & 3
        if rng.random() < nfrac:
            c = 4
        out.append(c); i += 1
    return np.array(out[:L], np.uint8)
