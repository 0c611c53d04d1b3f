"""Pin the oracle: our plain-C restatement (oracle/oracle.c) against the reference's own
unmodified sources compiled into oracle/_ref/libsaltref.so, on fuzzed inputs that cover
ref-N, read-N, multi-allele sites, indels, bursts, small buffers and band doubling."""
import os
import tempfile

import numpy as np
import pytest

from oracle import orc
from salt_b200 import synth

N_FAST = int(os.environ.get("SALT_FUZZ_N", "400"))


def _window_case(rng, L, with_indel=True):
    masks = synth.fuzz_masks(L + 80, int(rng.integers(1 << 30)))
    read = synth.fuzz_read_from_masks(masks[8:], L, rng, sub=rng.choice([0.0, 0.02, 0.06]),
                                      indel=(rng.choice([0.0, 0.01, 0.03]) if with_indel else 0.0))
    return masks, read


def test_score_matrices_match_reference(oracle, ref):
    assert np.array_equal(oracle.score_mat2()[:256], ref.score_mat2_ref)
    assert np.array_equal(oracle.score_mat(), ref.score_mat_ref)


def test_ed_mismatch(oracle, ref):
    rng = np.random.default_rng(11)
    for it in range(N_FAST * 5):
        L = int(rng.choice([37, 100, 150, 250]))
        masks, read = _window_case(rng, L, with_indel=False)
        mix = synth.pack_mixref(masks)
        pos = int(rng.integers(0, 24))
        for T in (0, 3, int(rng.integers(0, 12))):
            assert oracle.ed_mismatch(mix, pos, read, T) == ref.ed_mismatch(mix, pos, read, T)


def test_lv_bytes(oracle, ref):
    """computeEditDistance on raw byte strings, including the equality-gate vector of SURVEY §4."""
    t = np.array([1, 1, 1, 1, 1, 3] + [1] * 18, np.uint8)
    p = np.array([1, 1, 1, 1, 4, 2] + [1] * 14, np.uint8)
    assert ref.lv(t, p, 5) == 2 and oracle.lv(t, p, 5) == 2
    assert ref.lv_cigar(t, p, 5) == (2, "20M") and oracle.lv_cigar(t, p, 5) == (2, "20M")


def test_ed_diff(oracle, ref):
    rng = np.random.default_rng(12)
    for it in range(N_FAST * 4):
        L = int(rng.choice([37, 100, 150, 250]))
        masks, read = _window_case(rng, L)
        mix = synth.pack_mixref(masks)
        l = len(masks)
        pos = int(rng.integers(0, 16))
        k = int(rng.choice([2, 3, 5, 8, 10, 15, 25, 30, 40]))
        a = oracle.ed_diff(mix, l, pos, read, k)
        b = ref.ed_diff(mix, l, pos, read, k)
        assert a == b, (it, L, pos, k)
    # window running off the end of the reference (editdistance.c:178)
    assert oracle.ed_diff(mix, l, l - 50, read, 5) == ref.ed_diff(mix, l, l - 50, read, 5) == -1


def test_ed_diff_withcigar(oracle, ref):
    rng = np.random.default_rng(13)
    n_gapped = 0
    for it in range(N_FAST * 4):
        L = int(rng.choice([37, 100, 150, 250]))
        masks, read = _window_case(rng, L)
        mix = synth.pack_mixref(masks)
        pos = int(rng.integers(4, 12))
        k = int(rng.choice([1, 3, 5, 10, 15, 25, 30]))
        buflen = int(rng.choice([128, 256, 128, 8, 4, 12]))
        a = oracle.ed_diff_withcigar(mix, pos, read, k, buflen)
        b = ref.ed_diff_withcigar(mix, pos, read, k, buflen)
        if b[0] == -2:
            assert a[0] == -2
        else:
            assert a == b, (it, L, pos, k, buflen)
        n_gapped += ("I" in b[1]) or ("D" in b[1])
    assert n_gapped > N_FAST // 4


def _sw_case(rng, L, W):
    masks = synth.fuzz_masks(W, int(rng.integers(1 << 30)), snp=0.03, nfrac=0.005)
    st = int(rng.integers(0, max(1, W - L - 10)))
    read = synth.fuzz_read_from_masks(masks[st:], L, rng, sub=0.04, indel=0.04, burst=0.01, nfrac=0.003)
    return masks, read


def test_ssw_known_answer(oracle, ref):
    """Align_src/test/test_ssw_snp.c:81-87 -- the only fixed vector in the reference tree."""
    mat = np.full(256, -3, np.int8)
    for m in range(16):
        for b in range(4):
            if m >> b & 1:
                mat[m * 16 + b] = 1
    Ref = np.array([1, 3, 5, 7, 2, 4, 8, 9, 10, 11, 12, 13, 14, 15, 1, 2, 4, 6, 1], np.int8)
    Seq = np.array([0, 0, 0, 0, 1, 2, 2, 3, 3, 3, 3, 3, 3, 3, 3, 0, 1, 2, 1, 0], np.int8)
    want = (15, 0, 18, 1, 19)
    for impl in (oracle, ref):
        rc, t, cig = impl.ssw_align(Seq, mat, 16, Ref, gapO=5, gapE=2, flag=2, filters=0, filterd=100, maskLen=10)
        assert rc == 0 and (t[0], t[2], t[3], t[4], t[5]) == want
        assert [(int(c) >> 4, int(c) & 15) for c in cig] == [(19, 0)]


def test_ssw_align_mixref(oracle, ref):
    rng = np.random.default_rng(14)
    mat = oracle.score_mat2()
    gapped = 0
    for it in range(N_FAST):
        L = int(rng.choice([100, 150, 250, 64, 33]))
        W = int(rng.choice([401, 301, 551, L + 1]))
        masks, read = _sw_case(rng, L, max(W, L + 1))
        mix = synth.pack_mixref(masks)
        gapO, gapE = (3, 1) if it % 4 else (int(rng.integers(2, 8)), 1)
        a = oracle.rescue_mixref(mix, 0, W - 1, read, mat, gapO, gapE)
        b = ref.rescue_mixref(mix, 0, W - 1, read, mat, gapO, gapE)
        assert a[0] == b[0] == 0
        assert a[1] == b[1], (it, L, W, a[1], b[1])
        assert np.array_equal(a[2], b[2]), (it, L, W)
        gapped += any((int(c) & 15) != 0 for c in b[2])
    assert gapped > N_FAST // 5


def test_ssw_align_pac(oracle, ref):
    rng = np.random.default_rng(15)
    mat = oracle.score_mat()
    for it in range(N_FAST // 2):
        L = int(rng.choice([100, 150, 250]))
        W = int(rng.choice([401, 301]))
        g = rng.integers(0, 4, W + 8).astype(np.uint8)
        st = int(rng.integers(0, W - L - 8)) if W - L - 8 > 0 else 0
        read = synth.fuzz_read_from_masks((1 << g[st:]).astype(np.uint8), L, rng, sub=0.05, indel=0.03, nfrac=0.01)
        pac = synth.pack_pac(g)
        a = oracle.rescue_pac(pac, 0, W - 1, read, mat)
        b = ref.rescue_pac(pac, 0, W - 1, read, mat)
        assert a[1] == b[1] and np.array_equal(a[2], b[2]), (it, L, W)


def test_ssw_general_params(oracle, ref):
    """Random matrices / gap penalties (gapO > gapE and the degenerate gapO <= gapE), flags, maskLen."""
    rng = np.random.default_rng(16)
    for it in range(N_FAST):
        n = 5
        mat = rng.integers(-6, 0, (n, n)).astype(np.int8)
        mat[np.arange(n), np.arange(n)] = rng.integers(1, 6, n)
        L = int(rng.integers(20, 140)); W = int(rng.integers(L, 300))
        g = rng.integers(0, 4, W).astype(np.int8)
        st = int(rng.integers(0, max(1, W - L)))
        read = g[st:st + L].copy()
        if len(read) < L:
            read = np.concatenate([read, rng.integers(0, 4, L - len(read)).astype(np.int8)])
        e = rng.random(L) < 0.08
        read[e] = rng.integers(0, 5, int(e.sum()))
        gapE = int(rng.integers(1, 4)); gapO = gapE + int(rng.integers(-1, 6))
        gapO = max(gapO, 1)
        flag = int(rng.choice([0, 1, 2, 2, 2]))
        maskLen = int(rng.choice([L // 2, 15, 7]))
        a = oracle.ssw_align(read, mat.ravel(), n, g, gapO, gapE, flag, 0, 20, maskLen)
        b = ref.ssw_align(read, mat.ravel(), n, g, gapO, gapE, flag, 0, 20, maskLen)
        assert a[1] == b[1] and np.array_equal(a[2], b[2]), (it, L, W, gapO, gapE, flag)


def test_build_mixref(oracle):
    if not os.path.exists(os.path.join(os.path.dirname(orc.__file__), "_ref", "libsaltref_idx.so")):
        pytest.skip("oracle/_ref not built")
    refidx = orc.RefIdx()
    rng = np.random.default_rng(17)
    recs, rows = [], []
    for ci in range(3):
        n = int(rng.integers(50, 400))
        s = "".join(rng.choice(list("ACGTNacgtRY"), n, p=[.22, .22, .22, .22, .02, .02, .02, .02, .02, .01, .01]))
        recs.append(("chr%d" % ci, s))
        for p in sorted(set(rng.integers(1, n + 1, 12).tolist())):
            al = "/".join(rng.choice(list("ACGT"), int(rng.integers(2, 4)), replace=False))
            rows.append(("chr%d" % ci, p, al, s[p - 1]))
    with tempfile.TemporaryDirectory() as td:
        fa, sn, out = (os.path.join(td, x) for x in ("g.fa", "snp.txt", "g.ref"))
        with open(fa, "w") as f:
            for name, s in recs:
                f.write(">%s\n" % name)
                for i in range(0, len(s), 60):
                    f.write(s[i:i + 60] + "\n")
        with open(sn, "w") as f:
            for r in rows:
                f.write("%s\t%d\t%s\t%s\n" % r)
        words_ref, l_ref, rc = refidx.build_mixref(fa, sn, out)
    words, l = oracle.build_mixref(recs, rows)
    # the reference realloc()s its word array (Index_src/mixRef.c:135): nibbles past l in the last word are
    # whatever the heap held, so only the l valid nibbles are comparable
    if l % 8:
        words_ref = words_ref.copy(); words_ref[-1] &= np.uint32((1 << (4 * (l % 8))) - 1)
    assert l == l_ref and np.array_equal(words, words_ref)


def test_md_nm(oracle, refsam):
    """sam_add_md_nm (sam.c:246-328): MD/NM/XV text for random M/I/D strings, soft-clip starts, both strands,
    read-N and SNP sites matched by non-reference alleles -- oracle.c's restatement vs the reference's own code."""
    import parity_cases as pc
    g = synth.Genome(60000, snp_rate=0.03, seed=5)
    reads, pos, strand = synth.sample_reads(g, N_FAST * 3, 100, seed=6, sub_rate=0.03, indel_frac=0.3, n_frac=0.01)
    n_xv = 0
    for seq, rseq, p, st, s0, cig in pc.mdnm_cases(g, reads, pos, strand, 7):
        want = refsam.md_nm(g.mixref, g.l, g.pac, seq, rseq, p, st, s0, cig)
        got = oracle.md_nm(g.mixref, g.pac, g.l, rseq if st else seq, p, s0, cig)
        assert got == want, (p, st, s0, cig, got, want)
        n_xv += "XV:i:" in want
    assert n_xv > 20
    # unmapped reads get nothing (sam.c:248)
    assert refsam.md_nm(g.mixref, g.l, g.pac, reads[0], reads[0], 0xFFFFFFFF, 0, 0, "100M") == ""
    assert oracle.md_nm(g.mixref, g.pac, g.l, reads[0], 0xFFFFFFFF, 0, "100M") == ""
