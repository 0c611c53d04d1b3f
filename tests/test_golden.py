"""Golden vectors produced by the reference's own compiled code (tests/golden/make_golden.py):
the oracle restatement must reproduce them on CPU, the CUDA engine on the GPU.  Bit-exact."""
import os

import numpy as np
import pytest

from salt_b200 import api, synth

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "salt_golden_v1.npz")
GOLD_MDNM = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "salt_golden_mdnm_v1.npz")
TAGS = (("a", 100), ("b", 150), ("c", 250), ("d", 37))


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLD)


def _seq(G, tag, r, s):
    rd = G[tag + "_reads"][r]
    return np.ascontiguousarray(synth.revcomp(rd) if s else rd)


def _world(G, tag):
    return (G[tag + "_mixref"], int(G[tag + "_l"]), G[tag + "_pac"], G[tag + "_reads"], G[tag + "_pos"], G[tag + "_strand"],
            (G[tag + "_offs0"], G[tag + "_loci0"], G[tag + "_offs1"], G[tag + "_loci1"]))


def _parse_idx(G):
    fa = bytes(G["idx_fasta"]).decode().strip().split("\n")
    recs = [(fa[i][1:], fa[i + 1]) for i in range(0, len(fa), 2)]
    rows = []
    for ln in bytes(G["idx_snps"]).decode().strip().split("\n"):
        c, p, al, rf = ln.split("\t")
        rows.append((c, int(p), al, rf))
    return recs, rows


@pytest.fixture(scope="module")
def gold_mdnm():
    return np.load(GOLD_MDNM)


def _txt(row):
    return bytes(row).split(b"\0")[0].decode()


# ------------------------------------------------------------------ CPU: oracle vs golden
@pytest.mark.parametrize("tag,L", TAGS)
def test_oracle_md_nm(oracle, gold, gold_mdnm, tag, L):
    """MD/NM/XV text of the reference's own sam_add_md_nm (tests/golden/make_golden_mdnm.py)"""
    M = gold_mdnm
    n_xv = 0
    for i in range(len(M[tag + "_pos"])):
        s = int(M[tag + "_strand"][i])
        got = oracle.md_nm(gold[tag + "_mixref"], gold[tag + "_pac"], int(gold[tag + "_l"]), _seq(gold, tag, i, s),
                           int(M[tag + "_pos"][i]), int(M[tag + "_seq_start"][i]), _txt(M[tag + "_cigar"][i]))
        assert got == _txt(M[tag + "_text"][i]), (tag, i)
        n_xv += "XV:i:" in got
    assert n_xv >= 3


@pytest.mark.parametrize("tag,L", TAGS)
def test_oracle_pairs(oracle, gold, tag, L):
    mix, l, pac, reads, pos, strand, cands = _world(gold, tag)
    rid, st, lo = gold[tag + "_pair_rid"], gold[tag + "_pair_strand"], gold[tag + "_pair_pos"]
    for i in range(len(lo)):
        seq = _seq(gold, tag, rid[i], st[i]); p = int(lo[i])
        if p + L <= l:
            assert oracle.ed_mismatch(mix, p, seq, 3) == gold[tag + "_mm3"][i]
            assert oracle.ed_mismatch(mix, p, seq, 0) == gold[tag + "_mm0"][i]
        assert oracle.ed_diff(mix, l, p, seq, L // 10) == gold[tag + "_lvk"][i]
        assert oracle.ed_diff(mix, l, p, seq, 3) == gold[tag + "_lv3"][i]
    for r in range(len(reads)):
        e, s = oracle.ed_diff_withcigar(mix, int(pos[r]), _seq(gold, tag, r, strand[r]), min(30, L // 10 + 2), 128)
        assert e == gold[tag + "_cig_e"][r] and s == api.cstr(gold[tag + "_cig_s"][r])


@pytest.mark.parametrize("tag,L", TAGS)
def test_oracle_verify_stage(oracle, gold, tag, L):
    mix, l, pac, reads, pos, strand, (o0, l0, o1, l1) = _world(gold, tag)
    for rule, lvT in (("se", L // 10), ("pe", 3)):
        rec = gold["%s_%s_rec" % (tag, rule)]; a0 = gold["%s_%s_acc0" % (tag, rule)]; a1 = gold["%s_%s_acc1" % (tag, rule)]
        for r in range(len(reads)):
            prim, hits, _ = oracle.verify_read(mix, l, _seq(gold, tag, r, 0), _seq(gold, tag, r, 1),
                                               l0[o0[r]:o0[r + 1]], l1[o1[r]:o1[r + 1]], 3, lvT)
            assert prim[:4] == tuple(int(x) for x in rec[r][:4]), (rule, r)
            for s, (acc, lo, of) in enumerate(((a0, l0, o0), (a1, l1, o1))):
                want = [(int(p), int(v)) for p, v in zip(lo[of[r]:of[r + 1]], acc[of[r]:of[r + 1]]) if v >= 0]
                assert [(h[0], h[1]) for h in hits[s]] == want, (rule, r, s)


@pytest.mark.parametrize("tag,L", TAGS)
def test_oracle_ssw(oracle, gold, tag, L):
    mix, l, pac, reads, pos, strand, _ = _world(gold, tag)
    wins = gold[tag + "_wins"]
    for i, (rs, start, end) in enumerate(wins):
        seq = _seq(gold, tag, int(rs) >> 1, int(rs) & 1)
        rc, rec, cig = oracle.rescue_mixref(mix, int(start), int(end), seq, oracle.score_mat2())
        assert rec == tuple(int(x) for x in gold[tag + "_ssw_mix"][i]), (i, rec)
        assert np.array_equal(cig, gold[tag + "_ssw_mix_cig"][i][:rec[7]])
        rc, rec, cig = oracle.rescue_pac(pac, int(start), int(end), seq, oracle.score_mat())
        assert rec == tuple(int(x) for x in gold[tag + "_ssw_pac"][i]), (i, rec)
        assert np.array_equal(cig, gold[tag + "_ssw_pac_cig"][i][:rec[7]])


def test_oracle_known_answers_and_builder(oracle, gold):
    assert np.array_equal(oracle.score_mat2()[:256], gold["score_mat2"]) and np.array_equal(oracle.score_mat(), gold["score_mat"])
    assert oracle.lv(gold["gate_text"], gold["gate_pattern"], 5) == int(gold["gate_e"]) == 2
    assert oracle.lv_cigar(gold["gate_text"], gold["gate_pattern"], 5) == (2, bytes(gold["gate_cigar"]).decode())
    recs, rows = _parse_idx(gold)
    words, tot = oracle.build_mixref(recs, rows)
    assert tot == int(gold["idx_l"]) and np.array_equal(words, gold["idx_words"])


# ------------------------------------------------------------------ GPU: engine vs golden
def _engine(G, tag):
    mix, l, pac = G[tag + "_mixref"], int(G[tag + "_l"]), G[tag + "_pac"]
    eng = api.Engine(mix, l, pac, l, device=0)
    eng.set_reads(G[tag + "_reads"])
    return eng


@pytest.mark.gpu
@pytest.mark.parametrize("tag,L", TAGS)
def test_gpu_pairs(gold, tag, L):
    eng = _engine(gold, tag)
    pairs = api.Engine.make_pairs(gold[tag + "_pair_rid"], gold[tag + "_pair_strand"], gold[tag + "_pair_pos"])
    assert np.array_equal(eng.mismatch(pairs, 3), gold[tag + "_mm3"])
    assert np.array_equal(eng.mismatch(pairs, 0), gold[tag + "_mm0"])
    for filt in (1, 0):
        eng.set_lv_filter(filt)
        for mapping in (0, 1, 2):
            eng.set_lv_mapping(mapping)
            assert np.array_equal(eng.lv(pairs, -1), gold[tag + "_lvk"]), (filt, mapping)
            assert np.array_equal(eng.lv(pairs, 3), gold[tag + "_lv3"]), (filt, mapping)
    n = len(gold[tag + "_reads"])
    tp = api.Engine.make_pairs(np.arange(n, dtype=np.uint32), gold[tag + "_strand"], gold[tag + "_pos"])
    out, buf = eng.lv_cigar(tp, np.full(n, min(30, L // 10 + 2), np.uint8), 128)
    assert np.array_equal(out, gold[tag + "_cig_e"])
    for r in range(n):
        assert api.cstr(buf[r]) == api.cstr(gold[tag + "_cig_s"][r])


@pytest.mark.gpu
@pytest.mark.parametrize("tag,L", TAGS)
def test_gpu_md_nm(gold, gold_mdnm, tag, L):
    M = gold_mdnm
    eng = _engine(gold, tag)
    n = len(M[tag + "_pos"])
    rs = (np.arange(n, dtype=np.uint32) << 1) | M[tag + "_strand"].astype(np.uint32)
    out, md, xv = eng.md_nm(rs, M[tag + "_pos"], M[tag + "_seq_start"], [_txt(c) for c in M[tag + "_cigar"]], md_stride=4 * L + 128)
    for i in range(n):
        assert api.Engine.md_nm_text(out, md, xv, i) == _txt(M[tag + "_text"][i]), (tag, i)


@pytest.mark.gpu
@pytest.mark.parametrize("tag,L", TAGS)
def test_gpu_verify_stage(gold, tag, L):
    eng = _engine(gold, tag)
    o0, l0, o1, l1 = (gold[tag + x] for x in ("_offs0", "_loci0", "_offs1", "_loci1"))
    for rule, lvT in (("se", -1), ("pe", 3)):
        rec, a0, a1, cig = eng.verify(o0, l0, o1, l1, 3, lvT)
        want = gold["%s_%s_rec" % (tag, rule)]
        got = np.stack([rec["pos"], rec["strand"], rec["n_diff"], rec["is_gap"], rec["n_hits"][:, 0], rec["n_hits"][:, 1]], 1).astype(np.int64)
        assert np.array_equal(got, want), rule
        assert np.array_equal(a0, gold["%s_%s_acc0" % (tag, rule)]) and np.array_equal(a1, gold["%s_%s_acc1" % (tag, rule)])
        wc = gold["%s_%s_cig" % (tag, rule)]
        for r in range(len(rec)):
            if rec["is_gap"][r] == 1:
                assert api.cstr(cig[r]) == api.cstr(wc[r]), (rule, r)


@pytest.mark.gpu
@pytest.mark.parametrize("tag,L", TAGS)
def test_gpu_ssw(gold, tag, L):
    eng = _engine(gold, tag)
    w = gold[tag + "_wins"]
    wins = np.zeros(len(w), api.WIN_DT); wins["rs"], wins["start"], wins["end"] = w[:, 0], w[:, 1], w[:, 2]
    fields = ("score1", "score2", "ref_begin1", "ref_end1", "read_begin1", "read_end1", "ref_end2", "cigarLen")
    for use_pac, mat, n_sym, key in ((False, gold["score_mat2"], 16, "_ssw_mix"), (True, gold["score_mat"], 5, "_ssw_pac")):
        out, cig = eng.ssw(wins, mat, n_sym, use_pac, cigar_stride=64)
        got = np.stack([out[f].astype(np.int64) for f in fields], 1)
        assert np.array_equal(got, gold[tag + key]), key
        for i in range(len(w)):
            n = int(gold[tag + key][i][7])
            assert np.array_equal(cig[i][:n], gold[tag + key + "_cig"][i][:n])


@pytest.mark.gpu
def test_gpu_build_mixref(gold, oracle):
    recs, rows = _parse_idx(gold)
    bases = "".join(s for _, s in recs)
    offs = np.cumsum([0] + [len(s) for _, s in recs])
    # global 0-based positions, consuming one same-chrom block per record like the reference does
    name2off = {n: int(o) for (n, _), o in zip(recs, offs)}
    pos = np.array([name2off[r[0]] + r[1] - 1 for r in rows], np.uint32)
    mask = np.array([oracle.lib.orc_allele_mask(r[2].encode()) for r in rows], np.uint8)
    eng = api.Engine.from_bases(bases, pos, mask, device=0)
    assert np.array_equal(eng.get_mixref(), gold["idx_words"])
