"""Full-size runs of the CUDA path (BASELINE.json configs[1] shape, scaled to what a test may take):
size-independent properties the reference's semantics imply, plus sampled bit-exact checks against
the oracle.  The small exhaustive parity cases live in test_gpu_parity.py / test_golden.py."""
import types

import numpy as np
import pytest

import parity_cases as pc
from salt_b200 import api, synth

pytestmark = pytest.mark.gpu

N_READS = 400_000


@pytest.fixture(scope="module")
def big():
    import bench
    args = types.SimpleNamespace(reads=N_READS, genome=50_000_000, read_len=100, cands=8, snp_rate=0.01)
    wl = bench.make_workload(args, seed=23)
    g = wl["g"]
    eng = api.Engine(g.mixref, g.l, g.pac, g.l, device=0)
    eng.set_reads(wl["reads"])
    out = eng.verify(wl["offs0"], wl["loci0"], wl["offs1"], wl["loci1"], 3, -1)
    return wl, eng, out


def test_chunk_split_invariance(big):
    wl, eng, (rec, a0, a1, cig) = big
    n, L = wl["reads"].shape
    roffs = (np.arange(n + 1, dtype=np.uint64) * L).astype(np.uint32)
    for chunk in (100_000, 33_333):
        got = eng.verify_batch(wl["reads"], roffs, wl["offs0"], wl["loci0"], wl["offs1"], wl["loci1"], chunk, 3, -1)
        for a, b, name in zip(got, (rec, a0, a1, cig), ("rec", "acc0", "acc1", "cigars")):
            assert a.tobytes() == b.tobytes(), (chunk, name)
    eng.set_reads(wl["reads"])


def test_acceptance_invariants(big):
    """code_kmismatch / code_kdiff (alnse.c:348-393): accepted values never increase along a read's
    lists (strand 0 then strand 1), the primary carries their minimum, hit counts add up, and the
    gapped stage ran exactly for reads without an ungapped hit."""
    wl, eng, (rec, a0, a1, cig) = big
    n = len(rec)
    for acc, offs, s in ((a0, wl["offs0"], 0), (a1, wl["offs1"], 1)):
        cnt = np.add.reduceat((acc >= 0).astype(np.int64), offs[:-1].astype(np.int64))
        cnt[np.diff(offs.astype(np.int64)) == 0] = 0
        assert np.array_equal(cnt, rec["n_hits"][:, s])
    # running minimum over the concatenated (strand 0, strand 1) accepted values per read
    rid0 = np.repeat(np.arange(n), np.diff(wl["offs0"].astype(np.int64)))
    rid1 = np.repeat(np.arange(n), np.diff(wl["offs1"].astype(np.int64)))
    rid = np.concatenate([rid0, rid1]); order = np.argsort(rid, kind="stable")
    vals = np.concatenate([a0, a1])[order].astype(np.int64); rids = rid[order]
    keep = vals >= 0
    v, r = vals[keep], rids[keep]
    same = r[1:] == r[:-1]
    assert (v[1:][same] <= v[:-1][same]).all()
    best = np.full(n, 255, np.int64); np.minimum.at(best, r, v)
    mapped = rec["pos"] != 0xFFFFFFFF
    assert np.array_equal(best[mapped], rec["n_diff"][mapped].astype(np.int64))
    assert (best[~mapped] == 255).all() and (rec["lv_ran"][~mapped] == 1).all()
    assert np.array_equal(rec["lv_ran"] == 1, (rec["is_gap"] != 0))
    nogap_limit = rec["n_diff"][rec["is_gap"] == 0]
    assert (nogap_limit <= 3).all()
    assert (rec["n_diff"][rec["is_gap"] == 1] <= 10).all()
    assert mapped.mean() > 0.97


def test_mismatch_kernel_agrees_with_stage(big):
    wl, eng, (rec, a0, a1, cig) = big
    m = 300_000
    rid0 = np.repeat(np.arange(len(rec), dtype=np.uint32), np.diff(wl["offs0"].astype(np.int64)))[:m]
    pairs = api.Engine.make_pairs(rid0, np.zeros(m, np.uint32), wl["loci0"][:m])
    mm = eng.mismatch(pairs, 3)
    acc = a0[:m]
    ung = rec["lv_ran"][rid0] == 0
    # every accepted ungapped hit is the pair's mismatch count; a rejected pair either exceeds 3 or lost to the running threshold
    assert np.array_equal(mm[ung & (acc >= 0)], acc[ung & (acc >= 0)])
    assert (mm[ung & (acc < 0) & (mm >= 0)] >= 0).all()
    assert ((acc >= 0) <= ((mm >= 0) | ~ung)).all()


def test_lv_monotone_and_filter(big):
    wl, eng, _ = big
    m = 600_000
    rid0 = np.repeat(np.arange(N_READS, dtype=np.uint32), np.diff(wl["offs0"].astype(np.int64)))[:m]
    pairs = api.Engine.make_pairs(rid0, np.zeros(m, np.uint32), wl["loci0"][:m])
    r10 = eng.lv(pairs, 10); r5 = eng.lv(pairs, 5); r2 = eng.lv(pairs, 2)
    assert np.array_equal(r5[r5 >= 0], r10[r5 >= 0]) and (r10[(r5 < 0) & (r10 >= 0)] > 5).all()
    assert np.array_equal(r2[r2 >= 0], r5[r2 >= 0]) and (r5[(r2 < 0) & (r5 >= 0)] > 2).all()
    eng.set_lv_filter(0)
    assert np.array_equal(eng.lv(pairs, 10), r10)
    eng.set_lv_mapping(1)
    assert np.array_equal(eng.lv(pairs[:200_000], 10), r10[:200_000])
    eng.set_lv_mapping(0); eng.set_lv_filter(1)
    assert (r10 >= 0).sum() > 10_000


def test_sampled_reads_against_oracle(big, oracle):
    wl, eng, (rec, a0, a1, cig) = big
    g = wl["g"]
    rng = np.random.default_rng(3)
    gapped = np.flatnonzero(rec["is_gap"] == 1)
    sample = np.concatenate([rng.choice(len(rec), 1500, replace=False), rng.choice(gapped, min(500, len(gapped)), replace=False)])
    for r in sample:
        seq = np.ascontiguousarray(wl["reads"][r]); rseq = np.ascontiguousarray(synth.revcomp(wl["reads"][r]))
        l0 = wl["loci0"][wl["offs0"][r]:wl["offs0"][r + 1]]; l1 = wl["loci1"][wl["offs1"][r]:wl["offs1"][r + 1]]
        prim, hits, _ = oracle.verify_read(g.mixref, g.l, seq, rseq, l0, l1, 3, 10)
        got = rec[r]
        assert (int(got["pos"]), int(got["strand"]), int(got["n_diff"]), int(got["is_gap"])) == prim[:4], (r, got, prim)
        if prim[3] == 1:
            want = oracle.ed_diff_withcigar(g.mixref, prim[0], rseq if prim[1] else seq, prim[2], 128)
            assert api.cstr(cig[r]) == want[1]


def test_ssw_full_windows(big, oracle):
    """200k rescue windows: an error-free read scores its length, and a sample is bit-exact."""
    wl, eng, _ = big
    g = wl["g"]; L = 100; W = 401; nt = 200_000
    rng = np.random.default_rng(5)
    start = np.maximum(0, wl["pos"][:nt].astype(np.int64) - rng.integers(0, W - L, nt))
    wins = np.zeros(nt, api.WIN_DT)
    wins["rs"] = (np.arange(nt, dtype=np.uint32) << 1) | wl["strand"][:nt]
    wins["start"] = start; wins["end"] = np.minimum(g.l - 1, start + W - 1)
    out, cg = eng.ssw(wins, api.salt_score_mat2(), 16, False, cigar_stride=32)
    assert (out["score1"] <= L).all() and (out["score1"] >= 20).mean() > 0.99
    assert (out["ref_end1"] >= out["ref_begin1"]).all() and (out["read_end1"] - out["read_begin1"] < L).all()
    span = sum(((cg[:, i] >> 4) * ((cg[:, i] & 15) != 2)).astype(np.int64) for i in range(32))   # M + I consume the read
    ok = out["cigarLen"] <= 32
    assert np.array_equal(span[ok], (out["read_end1"] - out["read_begin1"] + 1)[ok].astype(np.int64))
    pc.check_ssw(eng, oracle, g, wl["reads"], wins[rng.choice(nt, 400, replace=False)], False, api.salt_score_mat2(), 16, cigar_stride=64)


def test_sam_tail_consistency(big, oracle):
    """MD/NM/XV of every mapped primary of the full-size run (sam.c:246-328): the MD string must account for exactly the
    reference bases the CIGAR spans (match counts + letters = M + D bases), NM = mismatches + inserted + deleted bases,
    XV entries are read offsets of mismatches; a random sample is compared with the oracle text."""
    import re
    wl, eng, (rec, a0, a1, cig) = big
    n, L = wl["reads"].shape
    eng.set_reads(wl["reads"])
    mapped = np.nonzero(rec["pos"] != 0xFFFFFFFF)[0]
    cigs = [api.cstr(cig[r]) if rec["is_gap"][r] == 1 else "%dM" % L for r in mapped]
    rs = (mapped.astype(np.uint32) << 1) | rec["strand"][mapped].astype(np.uint32)
    out, md, xv = eng.md_nm(rs, rec["pos"][mapped], np.zeros(len(mapped), np.uint32), cigs, md_stride=256)
    assert (out["md_len"] >= 0).all()
    op = re.compile(r"(\d+)([MID])"); tok = re.compile(r"(\d+)|\^([ACGT]+)|([ACGT])")
    rng = np.random.default_rng(5)
    check = set(rng.choice(len(mapped), 3000, replace=False).tolist()) | set(np.nonzero(rec["is_gap"][mapped] == 1)[0][:3000].tolist())
    for i in check:
        m_len = sum(int(a) for a, b in op.findall(cigs[i]) if b == "M")
        i_len = sum(int(a) for a, b in op.findall(cigs[i]) if b == "I")
        d_len = sum(int(a) for a, b in op.findall(cigs[i]) if b == "D")
        s = api.cstr(md[i])
        assert len(s) == out["md_len"][i]
        matches = mism = dels = 0
        for num, de, mm in tok.findall(s):
            if num:
                matches += int(num)
            elif de:
                dels += len(de)
            else:
                mism += 1
        # the reference prints no "0" between a deletion and a mismatch right after it, so "^CCA" may be two deleted
        # bases and a mismatch: only the total number of letters is well defined
        letters = mism + dels
        assert matches + letters == m_len + d_len and dels >= d_len, (cigs[i], s)
        assert out["nm"][i] == letters + i_len
        assert out["n_xv"][i] <= letters - d_len and (xv[i, :out["n_xv"][i]] < L).all()
    for i in list(check)[:400]:
        r = int(mapped[i]); s = int(rec["strand"][r])
        seq = np.ascontiguousarray(synth.revcomp(wl["reads"][r]) if s else wl["reads"][r])
        g = wl["g"]
        assert api.Engine.md_nm_text(out, md, xv, i) == oracle.md_nm(g.mixref, g.pac, g.l, seq, int(rec["pos"][r]), 0, cigs[i])
