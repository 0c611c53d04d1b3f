#!/usr/bin/env python
"""Generate tests/golden/salt_golden_v1.npz: seeded inputs with the outputs OF THE REFERENCE ITSELF
(oracle/_ref/libsaltref.so = weiquan/salt's unmodified editdistance.c, LandauVishkin.c, ssw.c,
and libsaltref_idx.so = its Index_src/mixRef.c + hapmap.c), so that boxes without /root/reference
can check both the oracle restatement (CPU tests) and the CUDA engine (GPU tests) against the
reference's own numbers.

Run in the build container only (needs /root/reference to build oracle/_ref):
    python tests/golden/make_golden.py
The reference has no golden vectors of its own; the two known-answer vectors that
exist (test/test_ssw_snp.c and the LV equality-gate vector) are stored too.
"""
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import orc  # noqa: E402
from salt_b200 import synth  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "salt_golden_v1.npz")


def main():
    assert orc.ref_available(), "oracle/_ref is not built: run `make -C oracle` where /root/reference exists"
    ref = orc.Ref()
    o = orc.Oracle()
    d = {}
    # ---- world: small genome with SNPs, N runs, reads with substitutions/indels/N, decoy candidates
    for tag, L, n_reads in (("a", 100, 240), ("b", 150, 80), ("c", 250, 40), ("d", 37, 40)):
        g = synth.Genome(24000, snp_rate=0.02, n_rate=0.002, seed=50 + L)
        reads, pos, strand = synth.sample_reads(g, n_reads, L, seed=60 + L, sub_rate=0.02, indel_frac=0.35, n_frac=0.003)
        offs0, loci0, offs1, loci1 = synth.make_candidates(g, pos, strand, L, per_strand=5, seed=70 + L)
        # edge loci: duplicates, at/after the end of the reference
        loci0 = loci0.copy(); loci1 = loci1.copy()
        loci0[offs0[3] + 1] = loci0[offs0[3]]
        e = int(offs1[n_reads]); loci1[e - 1] = g.l + 7; loci1[e - 2] = g.l - L + 1
        d[tag + "_mixref"] = g.mixref; d[tag + "_l"] = np.uint32(g.l); d[tag + "_pac"] = g.pac
        d[tag + "_reads"] = reads; d[tag + "_pos"] = pos; d[tag + "_strand"] = strand
        d[tag + "_offs0"] = offs0; d[tag + "_loci0"] = loci0; d[tag + "_offs1"] = offs1; d[tag + "_loci1"] = loci1
        # per-pair reference outputs on the flat pair list (strand 0 lists, then strand 1)
        rid = np.concatenate([np.repeat(np.arange(n_reads), np.diff(offs0.astype(np.int64))),
                              np.repeat(np.arange(n_reads), np.diff(offs1.astype(np.int64)))]).astype(np.uint32)
        st = np.concatenate([np.zeros(len(loci0), np.uint8), np.ones(len(loci1), np.uint8)])
        lo = np.concatenate([loci0, loci1])
        mm3 = np.full(len(lo), -1, np.int8); mm0 = mm3.copy(); lvk = mm3.copy(); lv3 = mm3.copy()
        for i in range(len(lo)):
            seq = np.ascontiguousarray(synth.revcomp(reads[rid[i]]) if st[i] else reads[rid[i]])
            p = int(lo[i])
            if p + L <= g.l:                                     # the reference's callers guarantee this (alnse.c:529)
                mm3[i] = ref.ed_mismatch(g.mixref, p, seq, 3)
                mm0[i] = ref.ed_mismatch(g.mixref, p, seq, 0)
            lvk[i] = ref.ed_diff(g.mixref, g.l, p, seq, L // 10)
            lv3[i] = ref.ed_diff(g.mixref, g.l, p, seq, 3)
        d[tag + "_pair_rid"] = rid; d[tag + "_pair_strand"] = st; d[tag + "_pair_pos"] = lo
        d[tag + "_mm3"] = mm3; d[tag + "_mm0"] = mm0; d[tag + "_lvk"] = lvk; d[tag + "_lv3"] = lv3
        # CIGARs at the true loci
        cg_e = np.zeros(n_reads, np.int8); cg_s = np.zeros((n_reads, 128), np.uint8)
        for r in range(n_reads):
            seq = np.ascontiguousarray(synth.revcomp(reads[r]) if strand[r] else reads[r])
            e_, s_ = ref.ed_diff_withcigar(g.mixref, int(pos[r]), seq, min(30, L // 10 + 2), 128)
            cg_e[r] = e_; b = s_.encode(); cg_s[r, :len(b)] = np.frombuffer(b, np.uint8)
        d[tag + "_cig_e"] = cg_e; d[tag + "_cig_s"] = cg_s
        # the whole verification stage driven over the REFERENCE's functions (SE rule and PE rule)
        codes = np.ascontiguousarray(reads).reshape(-1)
        roffs = (np.arange(n_reads + 1, dtype=np.uint64) * L).astype(np.uint32)
        for rule, lvT in (("se", -1), ("pe", 3)):
            _, recs, a0, a1, cig = o.verify_batch(g.mixref, g.l, codes, roffs, offs0, loci0, offs1, loci1, 3, lvT, 1, ref=ref)
            rec = np.array([(q.pos, q.strand, q.n_diff, q.is_gap, q.n_hits[0], q.n_hits[1]) for q in recs], np.int64)
            d["%s_%s_rec" % (tag, rule)] = rec; d["%s_%s_acc0" % (tag, rule)] = a0; d["%s_%s_acc1" % (tag, rule)] = a1
            d["%s_%s_cig" % (tag, rule)] = cig
        # mate-rescue windows through the reference's ssw (mixRef and pac flavours)
        rng = np.random.default_rng(80 + L)
        W = {100: 401, 150: 401, 250: 301, 37: 120}[L]
        nw = min(n_reads, 60)
        wins = np.zeros((nw, 3), np.uint32); rec_m = np.zeros((nw, 8), np.int64); rec_p = np.zeros((nw, 8), np.int64)
        cg_m = np.zeros((nw, 64), np.uint32); cg_p = np.zeros((nw, 64), np.uint32)
        mat2 = np.concatenate([ref.score_mat2_ref, [-3]]).astype(np.int8)      # index 256 (read N vs mask 15) -> last entry
        for i in range(nw):
            start = max(0, int(pos[i]) - int(rng.integers(0, W - L))) if i % 9 else int(rng.integers(0, g.l - W))
            end = min(g.l - 1, start + W - 1)
            seq = np.ascontiguousarray(synth.revcomp(reads[i]) if strand[i] else reads[i])
            wins[i] = ((i << 1) | int(strand[i]), start, end)
            rc, t, c = ref.rescue_mixref(g.mixref, start, end, seq, mat2)
            rec_m[i] = t; cg_m[i, :len(c)] = c
            seq4 = seq.copy()
            rc, t, c = ref.rescue_pac(g.pac, start, end, seq4, ref.score_mat_ref)
            rec_p[i] = t; cg_p[i, :len(c)] = c
        d[tag + "_wins"] = wins; d[tag + "_ssw_mix"] = rec_m; d[tag + "_ssw_mix_cig"] = cg_m
        d[tag + "_ssw_pac"] = rec_p; d[tag + "_ssw_pac_cig"] = cg_p
    d["score_mat2"] = ref.score_mat2_ref; d["score_mat"] = ref.score_mat_ref
    # ---- known-answer vectors
    t = np.array([1, 1, 1, 1, 1, 3] + [1] * 18, np.uint8); p = np.array([1, 1, 1, 1, 4, 2] + [1] * 14, np.uint8)
    d["gate_text"] = t; d["gate_pattern"] = p
    d["gate_e"] = np.int32(ref.lv(t, p, 5)); e_, s_ = ref.lv_cigar(t, p, 5)
    d["gate_cigar"] = np.frombuffer(s_.encode(), np.uint8); assert (int(d["gate_e"]), e_, s_) == (2, 2, "20M")
    # ---- the reference's own build_mixRef on a two-record FASTA with a SNP table
    rng = np.random.default_rng(99)
    recs = []
    for name, n in (("chrA", 1777), ("chrB", 905)):
        s = "".join(rng.choice(list("ACGT"), n)); s = s[:300] + "NNNNN" + s[305:600] + "ry" + s[602:]
        recs.append((name, s))
    rows = []
    for name, s in recs:
        for p1 in sorted(rng.choice(np.arange(1, len(s) + 1), 40, replace=False).tolist()):
            al = "/".join(rng.choice(list("ACGT"), rng.choice([2, 2, 3]), replace=False))
            rows.append((name, p1, al, s[p1 - 1]))
    with tempfile.TemporaryDirectory() as td:
        fa = os.path.join(td, "g.fa"); sn = os.path.join(td, "s.txt"); outp = os.path.join(td, "o.ref")
        with open(fa, "w") as f:
            for name, s in recs:
                f.write(">%s\n" % name)
                for i in range(0, len(s), 70):
                    f.write(s[i:i + 70] + "\n")
        with open(sn, "w") as f:
            for r in rows:
                f.write("%s\t%d\t%s\t%s\n" % r)
        words, tot, rc = orc.RefIdx().build_mixref(fa, sn, outp)
    d["idx_fasta"] = np.frombuffer("".join(">%s\n%s\n" % r for r in recs).encode(), np.uint8)
    d["idx_snps"] = np.frombuffer("".join("%s\t%d\t%s\t%s\n" % r for r in rows).encode(), np.uint8)
    if tot % 8:          # nibbles past l in the last word are uninitialised heap in the reference (realloc, mixRef.c:135)
        words = words.copy(); words[-1] &= np.uint32((1 << (4 * (tot % 8))) - 1)
    d["idx_words"] = words; d["idx_l"] = np.uint32(tot)
    np.savez_compressed(OUT, **d)
    print("wrote", OUT, os.path.getsize(OUT), "bytes,", len(d), "arrays")


if __name__ == "__main__":
    main()
