"""Parity cases shared by the GPU tests (real CUDA library) and the CPU logic tests (the SIMT
emulator build of the same kernel sources).  Every case compares the engine, called through
the C ABI, with the oracle on the same seeded inputs; all comparisons are bit-exact."""
import numpy as np

from salt_b200 import api, synth


def make_world(seed, glen=60000, L=100, n_reads=64, per_strand=6, snp_rate=0.02, n_rate=0.002,
               sub_rate=0.02, indel_frac=0.3, n_frac=0.003):
    g = synth.Genome(glen, snp_rate=snp_rate, n_rate=n_rate, seed=seed)
    reads, pos, strand = synth.sample_reads(g, n_reads, L, seed=seed + 1, sub_rate=sub_rate,
                                            indel_frac=indel_frac, n_frac=n_frac)
    offs0, loci0, offs1, loci1 = synth.make_candidates(g, pos, strand, L, per_strand=per_strand, seed=seed + 2)
    return g, reads, pos, strand, (offs0, loci0, offs1, loci1)


def flat_pairs(cands, n_reads):
    offs0, loci0, offs1, loci1 = cands
    rid0 = np.repeat(np.arange(n_reads, dtype=np.uint32), np.diff(offs0.astype(np.int64)))
    rid1 = np.repeat(np.arange(n_reads, dtype=np.uint32), np.diff(offs1.astype(np.int64)))
    p0 = api.Engine.make_pairs(rid0, np.zeros(len(rid0), np.uint32), loci0)
    p1 = api.Engine.make_pairs(rid1, np.ones(len(rid1), np.uint32), loci1)
    return np.concatenate([p0, p1])


def read_of(reads, rs):
    r = reads[rs >> 1]
    return synth.revcomp(r) if rs & 1 else r


def check_mismatch(eng, oracle, g, reads, pairs, max_err):
    got = eng.mismatch(pairs, max_err)
    for i, p in enumerate(pairs):
        seq = np.ascontiguousarray(read_of(reads, int(p["rs"])))
        want = oracle.ed_mismatch(g.mixref, int(p["pos"]), seq, max_err) if int(p["pos"]) + len(seq) <= g.l else -1
        assert got[i] == want, ("mismatch", i, int(p["rs"]), int(p["pos"]), got[i], want)
    return got


def check_lv(eng, oracle, g, reads, pairs, k):
    got = eng.lv(pairs, k)
    for i, p in enumerate(pairs):
        seq = np.ascontiguousarray(read_of(reads, int(p["rs"])))
        kk = k if k >= 0 else len(seq) // 10
        want = oracle.ed_diff(g.mixref, g.l, int(p["pos"]), seq, kk)
        assert got[i] == want, ("lv", i, int(p["rs"]), int(p["pos"]), kk, got[i], want)
    return got


def check_lv_filter(eng, oracle, g, L, k, n, seed):
    """Reads with about k scattered edits at their true locus (and +-1..3 shifted loci): the
    pigeonhole filter must never reject what the reference accepts, and filter on/off must agree."""
    rng = np.random.default_rng(seed)
    kk = k if k >= 0 else L // 10
    n_edits = rng.integers(max(0, kk - 3), kk + 3, n)
    reads, pos = synth.edit_rich_reads(g, n, L, n_edits, seed=seed + 1)
    eng.set_reads(reads)
    rid = np.arange(n, dtype=np.uint32)
    shift = rng.integers(-3, 4, n)
    pairs = np.concatenate([api.Engine.make_pairs(rid, np.zeros(n, np.uint32), pos),
                            api.Engine.make_pairs(rid, np.zeros(n, np.uint32), (pos.astype(np.int64) + shift).astype(np.uint32))])
    eng.set_lv_filter(1)
    got = check_lv(eng, oracle, g, reads, pairs, k)
    eng.set_lv_filter(0)
    assert np.array_equal(eng.lv(pairs, k), got)
    eng.set_lv_filter(1)
    return int((got >= 0).sum()), int((got == kk).sum())


def check_lv_cigar(eng, oracle, g, reads, pairs, k_each, stride):
    out, buf = eng.lv_cigar(pairs, k_each, stride, fill=0x7e)
    n_gapped = 0
    for i, p in enumerate(pairs):
        seq = np.ascontiguousarray(read_of(reads, int(p["rs"])))
        want = oracle.ed_diff_withcigar(g.mixref, int(p["pos"]), seq, int(k_each[i]), stride)
        assert out[i] == want[0], ("lv_cigar rc", i, out[i], want)
        if want[0] != -2:
            assert api.cstr(buf[i]) == want[1], ("lv_cigar str", i, api.cstr(buf[i]), want)
            # bytes after the terminator are the caller's (here: fill pattern), like the reference leaves them
            assert (buf[i][len(want[1]) + 1:] == 0x7e).all()
        n_gapped += ("I" in want[1]) or ("D" in want[1])
    return n_gapped


def check_verify(eng, oracle, g, reads, cands, nogap_T0=3, lv_T0=-1, stride=128):
    offs0, loci0, offs1, loci1 = cands
    rec, acc0, acc1, cig = eng.verify(offs0, loci0, offs1, loci1, nogap_T0, lv_T0, stride)
    n = len(reads)
    stats = dict(lv_ran=0, gapped=0, mapped=0)
    for r in range(n):
        seq = np.ascontiguousarray(reads[r]); rseq = np.ascontiguousarray(synth.revcomp(reads[r]))
        l0 = loci0[offs0[r]:offs0[r + 1]]; l1 = loci1[offs1[r]:offs1[r + 1]]
        prim, hits, _ = oracle.verify_read(g.mixref, g.l, seq, rseq, l0, l1, nogap_T0,
                                           (len(seq) // 10 if lv_T0 < 0 else lv_T0))
        got = rec[r]
        assert (int(got["pos"]), int(got["strand"]), int(got["n_diff"]), int(got["is_gap"])) == prim[:4], (r, got, prim)
        for s, (acc, lo, of) in enumerate(((acc0, l0, offs0), (acc1, l1, offs1))):
            a = acc[of[r]:of[r + 1]]
            got_hits = [(int(p), int(v)) for p, v in zip(lo, a) if v >= 0]
            want_hits = [(h[0], h[1]) for h in hits[s]]
            assert got_hits == want_hits, (r, s, got_hits, want_hits)
            assert int(got["n_hits"][s]) == len(want_hits)
        stats["lv_ran"] += int(got["lv_ran"]); stats["mapped"] += prim[0] != 0xFFFFFFFF
        if prim[3] == 1:
            stats["gapped"] += 1
            want = oracle.ed_diff_withcigar(g.mixref, prim[0], rseq if prim[1] else seq, prim[2], stride)
            assert api.cstr(cig[r]) == want[1], (r, api.cstr(cig[r]), want)
    return stats


def check_verify_batch(eng, reads, cands, chunk_reads, nogap_T0=3, lv_T0=-1, stride=128):
    """The asynchronous chunk pipeline must return exactly what the one-shot stage returns
    (which check_verify pins against the oracle)."""
    offs0, loci0, offs1, loci1 = cands
    n, L = reads.shape
    eng.set_reads(reads)
    want = eng.verify(offs0, loci0, offs1, loci1, nogap_T0, lv_T0, stride)
    roffs = (np.arange(n + 1, dtype=np.uint64) * L).astype(np.uint32)
    got = eng.verify_batch(reads, roffs, offs0, loci0, offs1, loci1, chunk_reads, nogap_T0, lv_T0, stride)
    for a, b, name in zip(got, want, ("rec", "acc0", "acc1", "cigars")):
        assert a.tobytes() == b.tobytes(), name
    return want[0]


def check_verify_packed(eng, reads_list, cands, chunk_reads, nogap_T0=3, lv_T0=-1, stride=128, variants=None):
    """The compact transport (salt_packed_chunk_t: 2/4-bit bases, lengths, per-read candidate counts) must give
    exactly what the plain format gives, for every packing flavour, through views that start mid-byte.
    reads_list: list of 1-D code arrays (ragged allowed)."""
    offs0, loci0, offs1, loci1 = cands
    lens = np.array([len(r) for r in reads_list], np.int64)
    roffs = np.concatenate([[0], np.cumsum(lens)]).astype(np.uint32)
    codes = np.concatenate(reads_list).astype(np.uint8) if len(reads_list) else np.zeros(0, np.uint8)
    eng.set_reads(codes, roffs)
    want = eng.verify(offs0, loci0, offs1, loci1, nogap_T0, lv_T0, stride)
    n_variants = 0
    for bits in (2, 4):
        for count_bits in (16, 32):
            for base_start in (0, 3):
                if variants is not None and (bits, count_bits, base_start) not in variants:
                    continue
                pk, keep = eng.packed_chunk(codes, roffs, offs0, loci0, offs1, loci1, bits=bits, count_bits=count_bits,
                                            base_start=base_start)
                got = eng.verify_batch_packed(pk, len(loci0), len(loci1), chunk_reads, nogap_T0, lv_T0, stride)
                for a, b, name in zip(got, want, ("rec", "acc0", "acc1", "cigars")):
                    assert a.tobytes() == b.tobytes(), (name, bits, count_bits, base_start)
                n_variants += 1
    # the synchronous read upload in the compact format feeds the per-pair entry points the same reads
    pk, keep = eng.packed_chunk(codes, roffs, offs0, loci0, offs1, loci1, bits=2, base_start=1)
    eng.set_reads_packed(pk)
    got = eng.verify(offs0, loci0, offs1, loci1, nogap_T0, lv_T0, stride)
    for a, b, name in zip(got, want, ("rec", "acc0", "acc1", "cigars")):
        assert a.tobytes() == b.tobytes(), ("set_reads_packed", name)
    return n_variants


class SparseWorld:
    """A reference of l bases that is zero (N) everywhere except around a few loci: large coordinates without
    gigabytes of random bases.  .mixref / .l like synth.Genome; reads come from the filled windows."""

    def __init__(self, l, centres, L, seed, span=480):
        rng = np.random.default_rng(seed)
        self.l = int(l)
        self.mixref = np.zeros((self.l + 7) // 8 + 1, np.uint32)        # one spare word: the oracle may look at position l
        self.windows = []
        for c in centres:
            a = max(0, min(int(c) - span // 2, self.l - 1))
            b = min(self.l, a + span)
            m = synth.fuzz_masks(b - a, int(rng.integers(1 << 30)), snp=0.03, nfrac=0.002)
            for i, v in enumerate(m):
                p = a + i
                self.mixref[p >> 3] |= np.uint32(int(v) << (4 * (p & 7)))
            self.windows.append((a, b, m))
        self.L = L

    def reads_and_candidates(self, per_window, seed, extra_loci=()):
        """per_window reads from every window (forward strand, a few edits), each with its true locus, small shifts,
        decoys inside other windows and the `extra_loci` (end-of-reference / beyond-the-end cases) on both strands."""
        rng = np.random.default_rng(seed)
        L = self.L
        reads, cand0, cand1 = [], [], []
        for (a, b, m) in self.windows:
            for _ in range(per_window):
                if b - a < L + 12:
                    continue
                o = int(rng.integers(0, b - a - L - 8))
                kind = rng.random()
                rd = synth.fuzz_read_from_masks(m[o:], L, rng, sub=0.015, indel=0.012 if kind < 0.4 else 0.0, nfrac=0.003)
                if rng.random() < 0.5:
                    reads.append(rd); mine = cand0
                    other = cand1
                else:
                    reads.append(synth.revcomp(rd)); mine = cand1
                    other = cand0
                true = a + o
                decoys = [w[0] + int(rng.integers(0, max(1, w[1] - w[0] - L))) for w in self.windows[::3]]
                lst = sorted(set([true, max(0, true - 2), true + 1, true + 3] + decoys + list(extra_loci)))
                mine.append(np.array(lst, np.uint32))
                other.append(np.array(sorted(set(decoys[:4] + list(extra_loci))), np.uint32))
        reads = np.array(reads, np.uint8)

        def csr(lists):
            offs = np.zeros(len(lists) + 1, np.uint32)
            offs[1:] = np.cumsum([len(x) for x in lists])
            return offs, (np.concatenate(lists) if lists else np.zeros(0, np.uint32)).astype(np.uint32)
        o0, l0 = csr(cand0); o1, l1 = csr(cand1)
        return reads, (o0, l0, o1, l1)


def check_host_chunks(eng, hostlib, oracle, g, reads, cands, chunk_reads, nogap_T0=3, lv_T0=-1, max_hits=5, with_tail=False):
    """The host-side C layer (include/salt_host.h): chunk queues through the pipeline slots, then
    query_set_hits / gen_mapq / query_gen_cigar per read -- against the oracle's verify_read."""
    from salt_b200 import host_api
    offs0, loci0, offs1, loci1 = cands
    n, L = reads.shape
    n_slots = 4
    chunks = [host_api.Chunk(hostlib, chunk_reads, chunk_reads * L, chunk_reads * 64 + 4096) for _ in range(n_slots)]
    owner = [None] * n_slots
    n_gapped = 0

    def drain(si):
        nonlocal n_gapped
        if owner[si] is None:
            return
        b = owner[si]; ch = chunks[si]
        ch.wait(eng, si)
        if with_tail:
            ch.tail(eng, si)
        for i in range(ch.H.salt_chunk_n_reads(ch.c)):
            r = b + i
            seq = np.ascontiguousarray(reads[r]); rseq = np.ascontiguousarray(synth.revcomp(reads[r]))
            l0 = loci0[offs0[r]:offs0[r + 1]]; l1 = loci1[offs1[r]:offs1[r + 1]]
            prim, hits, alts = oracle.verify_read(g.mixref, g.l, seq, rseq, l0, l1, nogap_T0,
                                                  (L // 10 if lv_T0 < 0 else lv_T0), max_hits)
            gp, galts, gcig = ch.result(i, max_hits)
            assert gp == prim, (r, gp, prim)
            assert galts == alts, (r, galts, alts)
            for s in (0, 1):
                assert ch.hits(i, s) == hits[s], (r, s)
            if prim[0] == 0xFFFFFFFF:
                assert gcig == ""
            elif prim[3] == 0:
                assert gcig == "%dM" % L
            else:
                want = oracle.ed_diff_withcigar(g.mixref, prim[0], rseq if prim[1] else seq, prim[2], 128)
                assert gcig == want[1], (r, gcig, want)
                n_gapped += 1
            if with_tail:                      # salt_chunk_tail: the MD/NM/XV tags of the primary (sam.c:246-328)
                md, nm, xv = ch.md(i)
                if prim[0] == 0xFFFFFFFF:
                    assert (md, nm, xv) == ("", 0, [])
                else:
                    text = "\tMD:Z:%s\tNM:i:%u" % (md, nm) + ("\tXV:i:" + ",".join(map(str, xv)) if xv else "")
                    assert text == oracle.md_nm(g.mixref, g.pac, g.l, rseq if prim[1] else seq, prim[0], 0, gcig), (r, text)
        owner[si] = None

    k = 0
    for b in range(0, n, chunk_reads):
        si = k % n_slots; k += 1
        drain(si)
        ch = chunks[si]; ch.reset()
        for r in range(b, min(n, b + chunk_reads)):
            ch.add_read(reads[r], loci0[offs0[r]:offs0[r + 1]], loci1[offs1[r]:offs1[r + 1]])
        ch.submit(eng, si, nogap_T0, lv_T0)
        owner[si] = b
    for si in range(n_slots):
        drain(si)
    for ch in chunks:
        ch.close()
    return n_gapped


def make_windows(g, reads, pos, strand, L, rng, width=401):
    """Rescue-like windows: the read's true locus somewhere inside a `width`-base window,
    plus a few decoy windows and windows clipped at the ends of the reference."""
    n = len(reads)
    wins = np.zeros(n, api.WIN_DT)
    for i in range(n):
        kind = rng.random()
        if kind < 0.8:
            start = max(0, int(pos[i]) - int(rng.integers(0, max(1, width - L))))
        elif kind < 0.9:
            start = int(rng.integers(0, g.l - width))
        else:
            start = int(rng.choice([0, g.l - width, g.l - int(rng.integers(L // 2, width))]))
        end = min(g.l - 1, start + width - 1)
        wins[i] = ((i << 1) | int(strand[i]), start, end)
    return wins


def check_ssw(eng, oracle, g, reads, wins, use_pac, mat, n_sym, gapO=3, gapE=1, flag=2, filters=0, filterd=20,
              mask_len=-1, cigar_stride=64):
    out, cig = eng.ssw(wins, mat, n_sym, use_pac, gapO, gapE, flag, filters, filterd, mask_len, cigar_stride)
    gapped = 0
    pad_mat = np.concatenate([np.asarray(mat, np.int8), np.asarray(mat, np.int8)[-1:]])   # index n*n -> last entry
    for i, w in enumerate(wins):
        seq = np.ascontiguousarray(read_of(reads, int(w["rs"])))
        ml = len(seq) // 2 if mask_len < 0 else mask_len
        if use_pac:
            idx = np.arange(int(w["start"]), int(w["end"]) + 1)
            ref = ((g.pac[idx >> 2] >> ((~idx & 3) << 1).astype(np.uint8)) & 3).astype(np.int8)
            rc, rec, wc = oracle.ssw_align(seq.astype(np.int8), mat, n_sym, ref, gapO, gapE, flag, filters, filterd, ml)
        else:
            ref = synth.unpack_mixref(g.mixref, int(w["start"]), int(w["end"]) - int(w["start"]) + 1).astype(np.int8)
            rd = (1 << seq.astype(np.int32)).astype(np.int8)
            rc, rec, wc = oracle.ssw_align(rd, pad_mat, n_sym, ref, gapO, gapE, flag, filters, filterd, ml)
        got = tuple(int(out[i][f]) for f in ("score1", "score2", "ref_begin1", "ref_end1", "read_begin1",
                                              "read_end1", "ref_end2", "cigarLen"))
        assert got == rec, ("ssw rec", i, got, rec, tuple(w))
        assert np.array_equal(cig[i][:rec[7]], wc), ("ssw cigar", i, cig[i][:rec[7]], wc)
        gapped += any((int(c) & 15) != 0 for c in wc)
    return gapped


def check_ssw_wide_bands(eng, oracle, seed, n_reads=150, L=100, glen=20003):
    """Every read lacks `gap` >= 16 reference bases in its middle, so banded_sw's first band (|refLen - readLen| + 1,
    ssw.c:845) is already wider than the main pass serves: all of them go through the overflow pass, more of them
    than it has threads.  Plus windows that end AT l, which alnpe.c:213-252 clamps to (position l reads as symbol 0)."""
    rng = np.random.default_rng(seed)
    g = synth.Genome(glen, snp_rate=0.01, n_rate=0.0, seed=seed)           # glen % 8 != 0: position l is in the last word
    codes = np.log2(np.maximum(synth.unpack_mixref(g.mixref, 0, g.l) & -synth.unpack_mixref(g.mixref, 0, g.l).astype(np.int8), 1)).astype(np.uint8)
    reads = np.zeros((n_reads, L), np.uint8); wins = np.zeros(n_reads, api.WIN_DT)
    for i in range(n_reads):
        gap = int(rng.integers(16, 40)); a = int(rng.integers(40, 61))
        p = int(rng.integers(200, g.l - 600))
        reads[i] = np.concatenate([codes[p:p + a], codes[p + a + gap:p + gap + L]])
        start = p - int(rng.integers(0, 150)); end = start + 400
        if i % 10 == 0:
            end = g.l; start = g.l - 400
            q = g.l - L - gap - int(rng.integers(0, 100))
            reads[i] = np.concatenate([codes[q:q + a], codes[q + a + gap:q + gap + L]])
        wins[i] = (i << 1, start, end)
    eng.set_reads(reads)
    gapped = check_ssw(eng, oracle, g, reads, wins, False, api.salt_score_mat2(), 16)
    return gapped


def check_ssw_narrow_bands(eng, oracle, seed, n_reads=120, L=100, glen=20011):
    """banded_sw's first band is |refLen - readLen| + 1 (ssw.c:845) and doubles until the band holds score1.  Reads built to
    walk every branch of the narrow-band pass: no gap (band 1), a 1- or 2-base gap (bands 2, 3), a deletion and an
    insertion of two bases each far apart (equal lengths: band 1 fails, band 2 succeeds), of three to six bases each (bands 1,
    2 and then 4 fail: handed on to the 4..7 pass and to the warp-per-task kernel), a 3+-base gap (first band 4+: general kernel), a 3..14-base gap (first band up to 15, the widest the warp-per-task pass
    serves) and reads whose alignment is only a few bases long (soft-clipped to less than the band rows)."""
    rng = np.random.default_rng(seed)
    g = synth.Genome(glen, snp_rate=0.01, n_rate=0.0, seed=seed)
    m = synth.unpack_mixref(g.mixref, 0, g.l)
    codes = np.log2(np.maximum(m & -m.astype(np.int8), 1)).astype(np.uint8)
    reads = np.zeros((n_reads, L), np.uint8); wins = np.zeros(n_reads, api.WIN_DT)
    for i in range(n_reads):
        p = int(rng.integers(300, g.l - 700))
        kind = i % 8
        if kind == 0:
            rd = codes[p:p + L].copy()
        elif kind in (1, 2):                                      # one gap of 1..2 bases, deletion or insertion
            k = kind; a = int(rng.integers(30, 70))
            if rng.random() < 0.5: rd = np.concatenate([codes[p:p + a], codes[p + a + k:p + k + L]])
            else: rd = np.concatenate([codes[p:p + a], rng.integers(0, 4, k).astype(np.uint8), codes[p + a:p + L - k]])
        elif kind in (3, 4, 5):                                   # compensating deletion + insertion of k bases
            k = 2 + (i // 8 + kind) % 5; a = int(rng.integers(20, 30)); b = int(rng.integers(60, 70))      # 2..6: bands 1, 2 (and 4) fail
            rd = np.concatenate([codes[p:p + a], codes[p + a + k:p + b + k], (3 - codes[p + b + k:p + b + 2 * k]), codes[p + b + k:p + L]])[:L]
        elif kind == 6:                                           # one gap of 3..14 bases: first bands 4..15, one warp per task
            k = int(rng.integers(3, 15)); a = int(rng.integers(30, 70))
            rd = np.concatenate([codes[p:p + a], codes[p + a + k:p + k + L]])
        else:                                                     # a few matching bases inside noise
            k = int(rng.integers(6, 14)); a = int(rng.integers(0, L - k))
            rd = rng.integers(0, 4, L).astype(np.uint8); rd[a:a + k] = codes[p + a:p + a + k]
        assert len(rd) == L
        e = rng.random(L) < 0.01
        rd = rd.copy(); rd[e] = (rd[e] + 1) & 3
        reads[i] = rd
        start = p - int(rng.integers(0, 150))
        wins[i] = (i << 1, start, start + 400)
    eng.set_reads(reads)
    gapped = check_ssw(eng, oracle, g, reads, wins, False, api.salt_score_mat2(), 16, cigar_stride=96)
    if g.pac is not None:
        check_ssw(eng, oracle, g, reads, wins, True, api.salt_score_mat2(), 16, cigar_stride=96)
    return gapped


def check_long_cigars(eng, oracle, seed, n_reads=40, L=250):
    """gapped primaries whose CIGAR strings run to 32+ characters (several indels per read): the rows that travel back in the
    slimmed 32-byte form must fall back to the full row, in the one-shot stage and through the chunk pipeline"""
    rng = np.random.default_rng(seed)
    glen = 60000
    masks = synth.fuzz_masks(glen, seed + 1, snp=0.01, nfrac=0.0)
    class G:            # the two fields the checks read
        pass
    g = G(); g.l = glen; g.mixref = synth.pack_mixref(masks); g.pac = None
    reads, l0, l1 = [], [], []
    for i in range(n_reads):
        p = int(rng.integers(100, glen - 2 * L))
        rd = synth.fuzz_read_from_masks(masks[p:], L, rng, sub=0.01, indel=0.03, nfrac=0.0)
        reads.append(rd)
        l0.append(np.array(sorted({p, max(0, p - 3), int(rng.integers(0, glen - 2 * L))}), np.uint32)); l1.append(np.zeros(0, np.uint32))
    reads = np.array(reads, np.uint8)
    offs0 = np.concatenate([[0], np.cumsum([len(x) for x in l0])]).astype(np.uint32); offs1 = np.zeros(n_reads + 1, np.uint32)
    cands = (offs0, np.concatenate(l0), offs1, np.zeros(0, np.uint32))
    return g, reads, cands


def check_tail_primaries(eng, oracle, g, reads, cands, nogap_T0=3, lv_T0=-1):
    """salt_b200_tail_primaries (tags of a verified chunk's primaries, nothing uploaded, MD strings packed) against
    salt_b200_md_nm on the same alignments (which check_md_nm pins against the oracle and the golden vectors)."""
    offs0, loci0, offs1, loci1 = cands
    eng.set_reads(reads)
    rec, acc0, acc1, cig = eng.verify(offs0, loci0, offs1, loci1, nogap_T0, lv_T0, 128)
    out, offs, packed, xv = eng.tail_primaries(0)
    n, L = reads.shape
    cigs = [api.cstr(cig[i]) if rec["is_gap"][i] == 1 else "%dM" % L for i in range(n)]
    rs = (np.arange(n, dtype=np.uint32) << 1) | (rec["strand"].astype(np.uint32) & 1)
    want, md, xv2 = eng.md_nm(rs, rec["pos"], np.zeros(n, np.uint32), cigs, md_stride=320)
    n_gapped = 0
    for i in range(n):
        a = bytes(packed[offs[i]:offs[i + 1]])
        assert a.endswith(b"\0")
        assert a[:-1].decode() == api.cstr(md[i]), (i, a, api.cstr(md[i]))
        assert tuple(out[i]) == tuple(want[i]), (i, out[i], want[i])
        assert np.array_equal(xv[i][:out["n_xv"][i]], xv2[i][:want["n_xv"][i]])
        n_gapped += rec["is_gap"][i] == 1
    assert int(offs[-1]) == len(packed)
    # the asynchronous flavour: tail queued right behind a verify that is still in flight, both completed by _tail_wait
    import ctypes as C
    roffs = (np.arange(n + 1) * L).astype(np.uint32)
    codes = np.ascontiguousarray(reads).reshape(-1)
    r = api.ReadsT(codes.ctypes.data, roffs.ctypes.data, n)
    cd = api.CandsT(); cd.offs[0], cd.offs[1] = offs0.ctypes.data, offs1.ctypes.data
    cd.loci[0], cd.loci[1] = loci0.ctypes.data, loci1.ctypes.data
    rec2 = np.zeros(n, api.VERIFY_DT); cig2 = np.zeros((n, 128), np.uint8)
    out2 = np.zeros(n, api.MDNM_OUT_DT); offs2 = np.zeros(n + 1, np.uint32); packed2 = np.zeros(len(packed) + 64, np.uint8)
    xvb = np.zeros((n, 64), np.uint16)
    Lb = eng.L
    for cap in (len(packed2), len(packed)):             # roomy, and exactly enough (the eager copy must not overrun it)
        eng._ck(Lb.salt_b200_verify_submit(eng.h, 2, C.byref(r), C.byref(cd), nogap_T0, lv_T0, rec2.ctypes.data, None, None, cig2.ctypes.data, 128))
        eng._ck(Lb.salt_b200_tail_submit(eng.h, 2, out2.ctypes.data, offs2.ctypes.data, packed2.ctypes.data, cap, xvb.ctypes.data, 64))
        nb = C.c_size_t(0)
        eng._ck(Lb.salt_b200_tail_wait(eng.h, 2, C.byref(nb)))
        assert nb.value == len(packed) and rec2.tobytes() == rec.tobytes()
        assert out2.tobytes() == out.tobytes() and np.array_equal(offs2, offs) and packed2[:nb.value].tobytes() == packed.tobytes()
    return n_gapped


def check_chunk_pair(eng, hostlib, g, n_pairs, L, seed, min_tlen=250, max_tlen=550):
    """salt_chunk_pair (the paired-end stage of a whole chunk: plans, one Smith-Waterman batch per flavour, apply, CIGARs
    of promoted alternates, MD/NM) against the same stage composed pair by pair from the pieces that the reference's own
    pairing pins (tests/test_pair_plan.py, the PE drop-in): salt_chunk_result, salt_pair_plan, salt_b200_ssw,
    salt_pair_apply, salt_b200_md_nm."""
    import ctypes as C
    from salt_b200 import host_api
    reads, pos, strand = synth.sample_pairs(g, n_pairs, L, seed=seed, hard_frac=0.15, junk_frac=0.04)
    offs0, loci0, offs1, loci1 = synth.make_candidates(g, pos, strand, L, per_strand=4, seed=seed + 1)
    n = 2 * n_pairs
    roffs = (np.arange(n + 1) * L).astype(np.uint32)
    ch = host_api.Chunk(hostlib, n + 8, (n + 8) * L, len(loci0) + len(loci1) + 64)
    assert ch.add_reads(reads, roffs, offs0, loci0, offs1, loci1) == 0
    ch.submit(eng, 0, 3, 3); ch.wait(eng, 0)
    finals, tail_out, tail_md, st = ch.pair(eng, 0, n_pairs, min_tlen, max_tlen, g.l)
    assert st.pairs == n_pairs
    m16 = api.salt_score_mat2(); m5 = api.salt_score_mat()
    n_rescued = 0
    for p in range(n_pairs):
        rr = []
        for m in (0, 1):
            r = host_api.ReadResultT()
            assert hostlib.salt_chunk_result(ch.c, 2 * p + m, 5, C.byref(r)) == 0
            rr.append(r)
        plan = host_api.pair_plan(hostlib, rr[0], L, rr[1], L, min_tlen, max_tlen, g.l, raw=True)
        ssw, cigs = [], []
        for w in plan.win[:plan.n_win]:
            wins = np.zeros(1, api.WIN_DT); wins[0] = (((2 * p + w.mate) << 1) | w.strand, w.start, w.end)
            o, cg = eng.ssw(wins, m5 if w.flavour == 5 else m16, w.flavour, w.flavour == 5)
            ssw.append(tuple(int(o[0][f]) for f in ("score1", "score2", "ref_begin1", "ref_end1", "read_begin1", "read_end1")))
            cigs.append([(int(c) >> 4, int(c) & 15) for c in cg[0][:int(o[0]["cigarLen"])]])
        rc, want = host_api.pair_apply(hostlib, plan, rr[0], L, rr[1], L, ssw, cigs, stride=64)
        assert rc == finals[p].paired, (p, rc, finals[p].paired)
        for m in (0, 1):
            f = finals[p].mate[m]
            got = (f.pos, f.strand, f.n_diff, f.is_gap, f.seq_start, f.seq_end, f.b0, f.b1, f.mapq, f.cigar_kind)
            assert got == want[m][:10], (p, m, got, want[m])
            if f.cigar_kind != 4:
                assert f.cigar.decode() == want[m][10], (p, m, f.cigar, want[m][10])
            n_rescued += f.cigar_kind == 3
            # the tags of this mate
            if f.pos != 0xFFFFFFFF:
                o, md, xv = eng.md_nm(np.array([((2 * p + m) << 1) | (f.strand & 1)], np.uint32), np.array([f.pos], np.uint32),
                                      np.array([f.seq_start], np.uint32), [f.cigar.decode()], md_stride=128, xv_stride=0)
                assert int(tail_out["nm"][2 * p + m]) == int(o["nm"][0]) and api.cstr(tail_md[2 * p + m]) == api.cstr(md[0]), (p, m)
            else:
                assert int(tail_out["md_len"][2 * p + m]) == 0
    assert st.rescued == n_rescued
    ch.close()
    return st


def random_cigar(rng, qlen, max_ops=5):
    """A random M/I/D run string consuming exactly qlen read bases (what query->cigar->s may hold)."""
    ops = []
    left = qlen
    n_ops = int(rng.integers(0, max_ops))
    for _ in range(n_ops):
        if left < 12:
            break
        m = int(rng.integers(1, left - 8))
        ops.append("%dM" % m); left -= m
        if rng.random() < 0.5:
            ops.append("%dD" % int(rng.integers(1, 5)))
        else:
            i = int(rng.integers(1, min(5, left - 2)))
            ops.append("%dI" % i); left -= i
    ops.append("%dM" % left)
    return "".join(ops)


def mdnm_cases(g, reads, pos, strand, seed, clip_frac=0.3):
    """(seq as aligned, rseq partner, pos, strand, seq_start, cigar) per read, with random soft clips and gaps."""
    rng = np.random.default_rng(seed)
    out = []
    for r in range(len(reads)):
        L = reads.shape[1]
        seq = np.ascontiguousarray(reads[r]); rseq = np.ascontiguousarray(synth.revcomp(reads[r]))
        s0 = int(rng.integers(1, 20)) if rng.random() < clip_frac else 0
        s1 = int(rng.integers(1, 20)) if rng.random() < clip_frac else 0
        if L - s0 - s1 < 16:                       # short reads: keep an aligned part worth the name
            s0 = s1 = 0
        cig = random_cigar(rng, L - s0 - s1)
        out.append((seq, rseq, int(pos[r]) + s0, int(strand[r]), s0, cig))
    return out


def check_md_nm(eng, oracle, g, reads, pos, strand, seed, md_stride=128):
    """salt_b200_md_nm against the oracle's sam_add_md_nm text, alignment by alignment (random M/I/D strings,
    soft-clip starts, both strands), plus the unmapped / too-small-buffer / past-the-end codes."""
    cases = mdnm_cases(g, reads, pos, strand, seed)
    eng.set_reads(reads)
    rid = np.arange(len(reads), dtype=np.uint32)
    rs = (rid << 1) | np.array([c[3] for c in cases], np.uint32)
    p = np.array([c[2] for c in cases], np.uint32); s0 = np.array([c[4] for c in cases], np.uint32)
    cigs = [c[5] for c in cases]
    out, md, xv = eng.md_nm(rs, p, s0, cigs, md_stride=md_stride)
    n_xv = 0
    for i, (seq, rseq, pp, st, ss, cig) in enumerate(cases):
        want = oracle.md_nm(g.mixref, g.pac, g.l, rseq if st else seq, pp, ss, cig)
        assert out["md_len"][i] >= 0, (i, out[i])
        got = api.Engine.md_nm_text(out, md, xv, i)
        assert got == want, (i, pp, st, ss, cig, got, want)
        n_xv += int(out["n_xv"][i] > 0)
    # unmapped: no tags; an M run past the end of the reference: -3 (the reference asserts); tiny MD buffer: -2
    out2, md2, _ = eng.md_nm(rs[:3], np.array([0xFFFFFFFF, g.l - 50, p[2]], np.uint32), np.zeros(3, np.uint32),
                             ["100M", "100M", cigs[2]], md_stride=128)
    assert out2["md_len"][0] == 0 and out2["nm"][0] == 0 and md2[0, 0] == 0
    assert out2["md_len"][1] == -3
    assert oracle.md_nm(g.mixref, g.pac, g.l, cases[1][1] if cases[1][3] else cases[1][0], g.l - 50, 0, "100M") == -3
    out3, md3, _ = eng.md_nm(rs, p, s0, cigs, md_stride=4)
    for i in range(len(cases)):
        full = api.cstr(md[i])
        if len(full) > 3:
            assert out3["md_len"][i] == -2 and api.cstr(md3[i]) == full[:3], (i, full, api.cstr(md3[i]))
        else:
            assert out3["md_len"][i] == len(full) and api.cstr(md3[i]) == full
        assert out3["nm"][i] == out["nm"][i]
    return n_xv
