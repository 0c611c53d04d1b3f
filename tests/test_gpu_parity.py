"""GPU parity tests: the real CUDA library through the C ABI against the oracle, bit-exact.
Sizes are what the oracle finishes in seconds; the full-size properties are in test_gpu_full.py."""
import numpy as np
import pytest

import parity_cases as pc
from salt_b200 import api, synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def world100():
    return pc.make_world(31, glen=400000, L=100, n_reads=3000, per_strand=8, indel_frac=0.2, sub_rate=0.02)


def _engine(g, pac=True):
    return api.Engine(g.mixref, g.l, g.pac if pac else None, g.l if pac else 0, device=0)


@pytest.mark.parametrize("L", [100, 150, 250, 37, 64])
def test_mismatch(oracle, L):
    g, reads, pos, strand, cands = pc.make_world(100 + L, glen=300000, L=L, n_reads=600, per_strand=6, indel_frac=0.0, sub_rate=0.006)
    eng = _engine(g)
    eng.set_reads(reads)
    pairs = pc.flat_pairs(cands, len(reads))
    extra = api.Engine.make_pairs([0, 1, 2], [0, 1, 0], [g.l - L, g.l - L + 1, g.l - 1])
    pairs = np.concatenate([pairs, extra])
    got = pc.check_mismatch(eng, oracle, g, reads, pairs, 3)
    assert (got >= 0).sum() >= len(reads) // 2
    pc.check_mismatch(eng, oracle, g, reads, pairs[:500], 0)
    pc.check_mismatch(eng, oracle, g, reads, pairs[:500], 40)


@pytest.mark.parametrize("L,k", [(100, -1), (100, 2), (100, 3), (100, 5), (100, 8), (150, -1), (250, -1), (100, 30), (64, 6)])
def test_lv(oracle, L, k):
    g, reads, pos, strand, cands = pc.make_world(200 + L + k, glen=300000, L=L, n_reads=300, per_strand=5, indel_frac=0.6, sub_rate=0.03)
    eng = _engine(g)
    eng.set_reads(reads)
    pairs = pc.flat_pairs(cands, len(reads))
    extra = api.Engine.make_pairs([0, 1], [0, 1], [g.l - L - 4, g.l - L - 3])
    pairs = np.concatenate([pairs, extra])
    got = pc.check_lv(eng, oracle, g, reads, pairs, k)
    eng.set_lv_mapping(1)          # the warp-per-pair kernel must agree with the thread-per-pair one
    assert np.array_equal(eng.lv(pairs, k), got)
    assert (got > 0).sum() >= 10


@pytest.mark.parametrize("L,k", [(100, -1), (100, 3), (100, 8), (64, 6), (150, -1), (250, -1), (40, 4), (100, 30), (300, 25)])
def test_lv_filter_adversarial(oracle, L, k):
    """about k scattered edits per read: the pigeonhole filter may not reject what the reference accepts"""
    g = synth.Genome(300000, snp_rate=0.02, seed=7)
    eng = _engine(g)
    found, at_k = pc.check_lv_filter(eng, oracle, g, L, k, 1500, seed=900 + L + k)
    assert found >= 300


@pytest.mark.parametrize("L", [100, 150, 250])
def test_lv_cigar(oracle, L):
    g, reads, pos, strand, cands = pc.make_world(300 + L, glen=300000, L=L, n_reads=400, per_strand=3, indel_frac=0.8, sub_rate=0.02)
    eng = _engine(g)
    eng.set_reads(reads)
    rng = np.random.default_rng(L)
    pairs = api.Engine.make_pairs(np.arange(len(reads), dtype=np.uint32), strand, pos)
    k_each = rng.choice([2, 5, 10, 25, 30], len(pairs)).astype(np.uint8)
    assert pc.check_lv_cigar(eng, oracle, g, reads, pairs, k_each, 128) >= 100
    pc.check_lv_cigar(eng, oracle, g, reads, pairs, k_each, 6)
    # thresholds the thread-per-pair kernels take (k <= 4, <= 10, <= 15), then the same through the warp kernel
    for mapping in (0, 1):
        eng.set_lv_mapping(mapping)
        for ks in ([1, 2, 3, 4], [2, 5, 10], [12, 15]):
            kk = rng.choice(ks, len(pairs)).astype(np.uint8)
            pc.check_lv_cigar(eng, oracle, g, reads, pairs, kk, 128)
            pc.check_lv_cigar(eng, oracle, g, reads, pairs, kk, 5)
    eng.set_lv_mapping(0)
    pc.check_lv_cigar(eng, oracle, g, reads, pc.flat_pairs(cands, len(reads))[:300], np.full(300, 10, np.uint8), 256)


@pytest.mark.parametrize("L,lv_T0", [(100, -1), (100, 3), (150, -1), (250, -1)])
def test_verify_stage(oracle, L, lv_T0):
    g, reads, pos, strand, cands = pc.make_world(400 + L, glen=300000, L=L, n_reads=1500, per_strand=6, indel_frac=0.3, sub_rate=0.025)
    offs0, loci0, offs1, loci1 = cands
    loci0 = loci0.copy(); loci0[offs0[3] + 1] = loci0[offs0[3]]
    eng = _engine(g)
    eng.set_reads(reads)
    st = pc.check_verify(eng, oracle, g, reads, (offs0, loci0, offs1, loci1), 3, lv_T0)
    assert st["lv_ran"] >= 100 and st["gapped"] >= 30 and st["mapped"] >= 800


def test_verify_long_lists(oracle):
    """lists longer than one lane group, duplicates straddling chunk boundaries, loci past the end"""
    g, reads, pos, strand, cands = pc.make_world(91, glen=100000, L=100, n_reads=300, per_strand=70, indel_frac=0.3)
    offs0, loci0, offs1, loci1 = cands
    loci0 = loci0.copy(); loci1 = loci1.copy()
    for r in range(300):
        b = int(offs0[r])
        if offs0[r + 1] - b > 40:
            loci0[b + 15] = loci0[b + 16] = loci0[b + 17]
            loci0[b + 31] = loci0[b + 32]
        e = int(offs1[r + 1])
        if e - offs1[r] > 4:
            loci1[e - 1] = g.l + 5; loci1[e - 2] = g.l - 50; loci1[e - 3] = g.l - 50
    eng = _engine(g)
    eng.set_reads(reads)
    pc.check_verify(eng, oracle, g, reads, (offs0, loci0, offs1, loci1), 3, -1)
    pc.check_verify(eng, oracle, g, reads, (offs0, loci0, offs1, loci1), 3, 3)


def test_verify_batch_pipeline(oracle):
    """the asynchronous chunk pipeline (4 slots) against the one-shot stage, several chunk sizes"""
    g, reads, pos, strand, cands = pc.make_world(321, glen=400000, L=100, n_reads=5000, per_strand=6, indel_frac=0.3, sub_rate=0.02)
    eng = _engine(g)
    rec = pc.check_verify_batch(eng, reads, cands, 333)
    assert (rec["is_gap"] == 1).sum() >= 100
    pc.check_verify_batch(eng, reads, cands, 1024)
    pc.check_verify_batch(eng, reads, cands, 100000)


@pytest.mark.parametrize("L", [100, 250])
def test_md_nm(oracle, L):
    """SAM tail kernel (MD/NM/XV, sam.c:246-328) against the oracle, 4000 alignments per length"""
    g = synth.Genome(400000, snp_rate=0.03, seed=15 + L)
    reads, pos, strand = synth.sample_reads(g, 4000, L, seed=16, sub_rate=0.03, indel_frac=0.3, n_frac=0.01)
    eng = _engine(g)
    assert pc.check_md_nm(eng, oracle, g, reads, pos, strand, 17, md_stride=2 * L + 64) >= 200


def test_host_layer_chunks(oracle):
    """include/salt_host.h over the CUDA engine: pinned chunk queues through the slots, then
    query_set_hits / gen_mapq / query_gen_cigar per read, against the oracle"""
    from salt_b200 import host_api
    hostlib = host_api.load()
    g, reads, pos, strand, cands = pc.make_world(654, glen=300000, L=100, n_reads=2000, per_strand=6, indel_frac=0.3, sub_rate=0.02)
    eng = _engine(g)
    assert pc.check_host_chunks(eng, hostlib, oracle, g, reads, cands, 300, with_tail=True) >= 50
    assert pc.check_host_chunks(eng, hostlib, oracle, g, reads[:600], tuple(
        c[:601] if i % 2 == 0 else c for i, c in enumerate(cands)), 128, 3, 3, max_hits=3) >= 5


def test_level0_symbol_shim(oracle):
    """the reference's own per-pair symbols (editdistance.h:20-22) served by the engine, batch of one"""
    import ctypes as C
    import os
    g, reads, pos, strand, cands = pc.make_world(88, glen=100000, L=100, n_reads=40, per_strand=3, indel_frac=0.5)
    eng = _engine(g)
    S = C.CDLL(os.path.join(os.path.dirname(api.LIB_PATH), "libsalt_level0.so"))
    S.salt_level0_attach.argtypes = [C.c_void_p, C.c_uint32]
    S.salt_level0_attach(eng.h, g.l)
    vp = C.c_void_p
    S.ed_mismatch.argtypes = [vp, C.c_uint32, vp, C.c_uint32, C.c_int]
    S.ed_diff.argtypes = [vp, C.c_uint32, C.c_uint32, C.c_uint32, vp, C.c_uint32, C.c_int]
    S.ed_diff_withcigar.argtypes = [vp, C.c_uint32, C.c_uint32, vp, C.c_uint32, C.c_int, C.c_char_p, C.c_int, C.c_int, C.c_int]
    n_gap = 0
    for r in range(len(reads)):
        seq = np.ascontiguousarray(synth.revcomp(reads[r]) if strand[r] else reads[r]); p = int(pos[r])
        assert S.ed_mismatch(None, p, seq.ctypes.data, 100, 3) == oracle.ed_mismatch(g.mixref, p, seq, 3)
        assert S.ed_diff(None, g.l, p, 104, seq.ctypes.data, 100, 10) == oracle.ed_diff(g.mixref, g.l, p, seq, 10)
        buf = C.create_string_buffer(128)
        e = S.ed_diff_withcigar(None, p, 104, seq.ctypes.data, 100, 10, buf, 128, 1, 0)
        want = oracle.ed_diff_withcigar(g.mixref, p, seq, 10, 128)
        assert (e, buf.value.decode()) == want
        n_gap += "I" in want[1] or "D" in want[1]
    assert n_gap >= 5
    # LandauVishkin.h:45 / :50 on caller-supplied bytes (text = allele masks, pattern = one-hot / 15), as ed_diff calls them
    S.computeEditDistance.argtypes = [vp, C.c_int, vp, C.c_int, C.c_int]
    S.computeEditDistanceWithCigar.argtypes = [vp, C.c_int, vp, C.c_int, C.c_int, C.c_char_p, C.c_int, C.c_int, C.c_int]
    onehot = np.array([1, 2, 4, 8, 15], np.uint8)
    for r in range(0, len(reads), 2):
        seq = np.ascontiguousarray(synth.revcomp(reads[r]) if strand[r] else reads[r]); p = int(pos[r])
        text = np.ascontiguousarray(synth.unpack_mixref(g.mixref, p, 104)); pat = np.ascontiguousarray(onehot[seq])
        assert S.computeEditDistance(text.ctypes.data, 104, pat.ctypes.data, 100, 10) == oracle.ed_diff(g.mixref, g.l, p, seq, 10)
        buf = C.create_string_buffer(128)
        e = S.computeEditDistanceWithCigar(text.ctypes.data, 104, pat.ctypes.data, 100, 10, buf, 128, 1, 0)
        assert (e, buf.value.decode()) == oracle.ed_diff_withcigar(g.mixref, p, seq, 10, 128)
    # ssw.h:71 / :111 / :76 / :124 with the reference's ownership rules: borrowed read / mat, calloc'd s_align, malloc'd cigar
    class SAlign(C.Structure):
        _fields_ = [("score1", C.c_uint16), ("score2", C.c_uint16), ("ref_begin1", C.c_int32), ("ref_end1", C.c_int32),
                    ("read_begin1", C.c_int32), ("read_end1", C.c_int32), ("ref_end2", C.c_int32), ("cigar", C.POINTER(C.c_uint32)),
                    ("cigarLen", C.c_int32)]
    S.ssw_init.restype = vp; S.ssw_init.argtypes = [vp, C.c_int32, vp, C.c_int32, C.c_int8]
    S.ssw_align.restype = C.POINTER(SAlign)
    S.ssw_align.argtypes = [vp, vp, C.c_int32, C.c_uint8, C.c_uint8, C.c_uint8, C.c_uint16, C.c_int32, C.c_int32]
    S.init_destroy.argtypes = [vp]; S.align_destroy.argtypes = [C.POINTER(SAlign)]
    m16 = api.salt_score_mat2(); m5 = api.salt_score_mat()
    pad16 = np.concatenate([m16, m16[-1:]])
    n_sw_gap = 0
    for r in range(0, len(reads), 3):
        seq = np.ascontiguousarray(synth.revcomp(reads[r]) if strand[r] else reads[r]); p = max(0, int(pos[r]) - 150)
        for n_sym in (16, 5):
            if n_sym == 16:
                ref = synth.unpack_mixref(g.mixref, p, 401).astype(np.int8); rd = (1 << seq.astype(np.int32)).astype(np.int8); mat = m16
            else:
                idx = np.arange(p, p + 401)
                ref = ((g.pac[idx >> 2] >> ((~idx & 3) << 1).astype(np.uint8)) & 3).astype(np.int8); rd = seq.astype(np.int8); mat = m5
            ref = np.ascontiguousarray(ref); rd = np.ascontiguousarray(rd)
            prof = S.ssw_init(rd.ctypes.data, 100, mat.ctypes.data, n_sym, 1)
            a = S.ssw_align(prof, ref.ctypes.data, 401, 3, 1, 2, 0, 20, 50)
            assert a
            rc, want, wc = oracle.ssw_align(rd, pad16 if n_sym == 16 else m5, n_sym, ref, 3, 1, 2, 0, 20, 50)
            got = (a.contents.score1, a.contents.score2, a.contents.ref_begin1, a.contents.ref_end1, a.contents.read_begin1,
                   a.contents.read_end1, a.contents.ref_end2, a.contents.cigarLen)
            assert got == want, (r, n_sym, got, want)
            assert [a.contents.cigar[i] for i in range(want[7])] == [int(c) for c in wc]
            n_sw_gap += any((int(c) & 15) != 0 for c in wc)
            S.align_destroy(a); S.init_destroy(prof)
    assert n_sw_gap >= 2


def test_verify_ragged_and_empty(oracle):
    """ragged read lengths in one chunk, reads without candidates, an empty chunk of lists"""
    rng = np.random.default_rng(5)
    g = synth.Genome(200000, snp_rate=0.02, seed=12)
    lens = rng.choice([36, 50, 75, 100, 101, 125, 150], 400)
    reads, poss, strands = [], [], []
    for i, L in enumerate(lens):
        r, p, s = synth.sample_reads(g, 1, int(L), seed=1000 + i, sub_rate=0.02, indel_frac=0.3)
        reads.append(r[0]); poss.append(p[0]); strands.append(s[0])
    codes = np.concatenate(reads); roffs = np.concatenate([[0], np.cumsum(lens)]).astype(np.uint32)
    offs = [[0], [0]]; loci = [[], []]
    for i, L in enumerate(lens):
        for s in (0, 1):
            c = set(rng.integers(0, g.l - 160, 4).tolist()) if i % 7 else set()
            if strands[i] == s and i % 7:
                c.add(int(poss[i]))
            loci[s] += sorted(c); offs[s].append(len(loci[s]))
    eng = _engine(g)
    eng.set_reads(codes, roffs)
    o0, o1 = np.array(offs[0], np.uint32), np.array(offs[1], np.uint32)
    l0, l1 = np.array(loci[0], np.uint32), np.array(loci[1], np.uint32)
    rec, acc0, acc1, cig = eng.verify(o0, l0, o1, l1, 3, -1)
    for r in range(len(lens)):
        seq = np.ascontiguousarray(reads[r]); rseq = np.ascontiguousarray(synth.revcomp(reads[r]))
        prim, hits, _ = oracle.verify_read(g.mixref, g.l, seq, rseq, l0[o0[r]:o0[r + 1]], l1[o1[r]:o1[r + 1]], 3, len(seq) // 10)
        got = rec[r]
        assert (int(got["pos"]), int(got["strand"]), int(got["n_diff"]), int(got["is_gap"])) == prim[:4], (r, got, prim)


@pytest.mark.parametrize("L,width", [(100, 401), (150, 401), (250, 301), (250, 551), (64, 200), (300, 700)])
def test_ssw_mixref(oracle, L, width):
    g, reads, pos, strand, _ = pc.make_world(500 + L, glen=200000, L=L, n_reads=301, per_strand=2, indel_frac=0.7, sub_rate=0.04)
    eng = _engine(g)
    eng.set_reads(reads)
    rng = np.random.default_rng(L)
    wins = pc.make_windows(g, reads, pos, strand, L, rng, width)
    assert pc.check_ssw(eng, oracle, g, reads, wins, False, api.salt_score_mat2(), 16, cigar_stride=96) >= 60


def test_ssw_pac_and_params(oracle):
    L = 100
    g, reads, pos, strand, _ = pc.make_world(600, glen=200000, L=L, n_reads=200, per_strand=2, indel_frac=0.7, sub_rate=0.04,
                                              n_rate=0.0, n_frac=0.01)
    eng = _engine(g)
    eng.set_reads(reads)
    rng = np.random.default_rng(6)
    wins = pc.make_windows(g, reads, pos, strand, L, rng, 401)
    pc.check_ssw(eng, oracle, g, reads, wins, True, api.salt_score_mat(), 5)
    pc.check_ssw(eng, oracle, g, reads, wins, True, api.salt_score_mat(), 5, gapO=5, gapE=2, mask_len=15)
    pc.check_ssw(eng, oracle, g, reads, wins, True, api.salt_score_mat(), 5, flag=0)
    pc.check_ssw(eng, oracle, g, reads, wins, True, api.salt_score_mat(), 5, flag=1, mask_len=7)
    with pytest.raises(api.SaltError):
        eng.ssw(wins, api.salt_score_mat(), 5, True, gapO=1, gapE=1)


def test_ssw_known_answer():
    """Align_src/test/test_ssw_snp.c:81-87 through the engine: needs its own matrix and window."""
    mat = np.full(256, -3, np.int8)
    for m in range(16):
        for b in range(4):
            if m >> b & 1:
                mat[m * 16 + (1 << b)] = 1       # our read symbols are one-hot, the test file's are 0..3
    Ref = np.array([1, 3, 5, 7, 2, 4, 8, 9, 10, 11, 12, 13, 14, 15, 1, 2, 4, 6, 1], np.uint8)
    Seq = np.array([0, 0, 0, 0, 1, 2, 2, 3, 3, 3, 3, 3, 3, 3, 3, 0, 1, 2, 1, 0], np.uint8)
    masks = np.concatenate([Ref, np.zeros(45, np.uint8)])
    eng = api.Engine(synth.pack_mixref(masks), len(masks))
    eng.set_reads(Seq[None, :])
    wins = np.zeros(1, api.WIN_DT); wins[0] = (0, 0, 18)
    out, cig = eng.ssw(wins, mat, 16, False, gapO=5, gapE=2, flag=2, filters=0, filterd=100, mask_len=10)
    o = out[0]
    assert (int(o["score1"]), int(o["ref_begin1"]), int(o["ref_end1"]), int(o["read_begin1"]), int(o["read_end1"])) == (15, 0, 18, 1, 19)
    assert int(o["cigarLen"]) == 1 and int(cig[0][0]) == (19 << 4)


def test_build_mixref_on_device(oracle):
    g = synth.Genome(300000, snp_rate=0.03, n_rate=0.01, seed=9)
    rows = g.snp_table()
    fasta = g.fasta()
    want, l = oracle.build_mixref([("chr1", fasta)], rows)
    pos = np.array([r[1] - 1 for r in rows], np.uint32)
    mask = np.array([oracle.lib.orc_allele_mask(r[2].encode()) for r in rows], np.uint8)
    eng = api.Engine.from_bases(fasta, pos, mask)
    assert np.array_equal(eng.get_mixref(), want)


def test_error_paths():
    """misuse is reported through return codes and salt_b200_last_error(), never by exiting"""
    import ctypes as C
    g, reads, pos, strand, cands = pc.make_world(5, glen=50000, L=100, n_reads=20, per_strand=2)
    eng = _engine(g, pac=False)
    L_ = eng.L
    pairs = api.Engine.make_pairs([0], [0], [10])
    with pytest.raises(api.SaltError) as e:
        eng.mismatch(pairs, 3)                                   # no reads set yet
    assert e.value.code == -101
    eng.set_reads(reads)
    with pytest.raises(api.SaltError):
        eng.mismatch(pairs, 500)                                 # max_err out of range
    with pytest.raises(api.SaltError):
        eng.lv_cigar(pairs, np.array([31], np.uint8))            # LandauVishkin.c:183 asserts k < 31
    wins = np.zeros(1, api.WIN_DT); wins["end"] = 300
    with pytest.raises(api.SaltError) as e:
        eng.ssw(wins, api.salt_score_mat2(), 16, False, gapO=1, gapE=1)      # gapO <= gapE is not reproduced
    assert e.value.code == -104
    with pytest.raises(api.SaltError):
        eng.ssw(wins, api.salt_score_mat(), 5, True)             # no pac uploaded
    # SAM tail: needs the 2-bit pac, sane strides, an existing slot
    with pytest.raises(api.SaltError) as e:
        eng.md_nm(np.array([0], np.uint32), np.array([10], np.uint32), np.array([0], np.uint32), ["100M"])      # engine built without pac
    assert e.value.code == -101 and "pac" in str(e.value)
    eng2 = _engine(g)
    eng2.set_reads(reads)
    for kw in (dict(md_stride=1), dict(xv_stride=65), dict(slot=7)):
        with pytest.raises(api.SaltError):
            eng2.md_nm(np.array([0], np.uint32), np.array([10], np.uint32), np.array([0], np.uint32), ["100M"], **kw)
    out, md, xv = eng2.md_nm(np.array([0, 1 << 20], np.uint32), np.array([10, 10], np.uint32), np.zeros(2, np.uint32), ["100M", "100M"])
    assert out["md_len"][0] > 0 and out["md_len"][1] == 0 and md[1, 0] == 0            # a read id outside the chunk: no tags
    eng2.close()
    # a slot cannot take a second chunk before it was waited for; bad slot numbers are refused
    offs0, loci0, offs1, loci1 = cands
    n = len(reads)
    codes = np.ascontiguousarray(reads).reshape(-1); roffs = (np.arange(n + 1) * 100).astype(np.uint32)
    r = api.ReadsT(codes.ctypes.data, roffs.ctypes.data, n)
    c = api.CandsT(); c.offs[0], c.offs[1] = offs0.ctypes.data, offs1.ctypes.data; c.loci[0], c.loci[1] = loci0.ctypes.data, loci1.ctypes.data
    rec = np.zeros(n, api.VERIFY_DT)
    assert L_.salt_b200_verify_submit(eng.h, 9, C.byref(r), C.byref(c), 3, -1, rec.ctypes.data, None, None, None, 0) == -101
    assert L_.salt_b200_verify_submit(eng.h, 1, C.byref(r), C.byref(c), 3, -1, rec.ctypes.data, None, None, None, 0) == 0
    assert L_.salt_b200_verify_submit(eng.h, 1, C.byref(r), C.byref(c), 3, -1, rec.ctypes.data, None, None, None, 0) == -101
    assert L_.salt_b200_verify_wait(eng.h, 1) == 0
    assert (rec["pos"] != 0xFFFFFFFF).sum() >= 10
    # a window that leaves the reference comes back flagged, not crashed
    wins = np.zeros(2, api.WIN_DT); wins["start"] = [10, g.l - 50]; wins["end"] = [310, g.l + 20]
    o, _ = eng.ssw(wins, api.salt_score_mat2(), 16, False)
    assert int(o["cigarLen"][0]) >= 0 and int(o["cigarLen"][1]) == -1      # declined per item, the batch goes through


@pytest.mark.gpu
def test_ssw_narrow_bands_every_branch(oracle):
    """bands 1..3 in registers, the doubling 1 -> 2, the hand-over to the general kernel; mixRef and pac scoring"""
    g = synth.Genome(20011, snp_rate=0.01, n_rate=0.0, seed=9)
    eng = _engine(g)
    assert pc.check_ssw_narrow_bands(eng, oracle, 9, n_reads=1600) >= 800
    eng.close()


@pytest.mark.gpu
def test_ssw_wide_bands_and_end_at_l(oracle):
    """bands wider than the main pass, more of them than the overflow pass has threads; windows clamped to end == l"""
    g = synth.Genome(20003, snp_rate=0.01, n_rate=0.0, seed=77)
    eng = _engine(g, pac=False)
    assert pc.check_ssw_wide_bands(eng, oracle, 77, n_reads=700) >= 600
    eng.close()


@pytest.mark.gpu
def test_verify_packed_transport(oracle):
    """compact transport (2/4-bit bases, lengths, per-read counts) == plain format, all flavours, ragged reads, N bases"""
    g, reads, pos, strand, cands = pc.make_world(321, L=150, n_reads=5000, per_strand=5, indel_frac=0.3, glen=400000, n_frac=0.01)
    eng = _engine(g)
    assert pc.check_verify_packed(eng, [r for r in reads], cands, 700) == 8
    rng = np.random.default_rng(5)
    ragged = [r[:int(rng.integers(37, 151))] for r in reads]
    pc.check_verify_packed(eng, ragged, cands, 1100, lv_T0=3)
    pc.check_verify_packed(eng, ragged, cands, 100000)
    eng.close()


@pytest.mark.gpu
def test_verify_max_list_lengths(oracle):
    """the longest lists the reference allows: max_locate = 1000 candidates per read in SE (aln.c:47, alnse.c:678) and
    MAX_LOC_POS = 262144 per strand in PE (alnse.c:42,533), with duplicates everywhere a lane group / chunk boundary
    could fall; one of the huge-list reads only matches with a gap, so the gapped stage sees all 524 288 pairs"""
    L = 100
    g, reads, pos, strand, cands = pc.make_world(4242, glen=3_000_000, L=L, n_reads=24, per_strand=4, indel_frac=0.0, sub_rate=0.01)
    rng = np.random.default_rng(99)
    # read 1 gets a 2-base deletion so that no ungapped candidate passes
    src = synth.unpack_mixref(g.mixref, int(pos[1]), L + 8)
    code = np.array([[b for b in range(4) if (int(m) >> b) & 1][0] if m else 0 for m in src], np.uint8)
    rd = np.concatenate([code[:50], code[52:L + 2]])
    reads[1] = synth.revcomp(rd) if strand[1] else rd
    lists = [[], []]
    hi = g.l - L - 5
    for r in range(len(reads)):
        want = 262144 if r < 2 else (500 if r < 12 else 6)
        for s_ in (0, 1):
            loci = np.unique(rng.integers(0, hi, want + want // 8))[:want].astype(np.int64)
            if strand[r] == s_:
                loci[rng.integers(0, len(loci))] = int(pos[r])
                loci[rng.integers(0, len(loci))] = int(pos[r]) + 1
            loci.sort()
            step = 4093 if want > 1000 else 37
            loci[step::step] = loci[step - 1:-1:step]            # duplicates (alnse.c:762 skips pos1 == pos0)
            if want >= 500:
                loci[-1] = g.l + 7; loci[-2] = g.l - 20           # past the end / window leaving the reference
            lists[s_].append(loci.astype(np.uint32))

    def csr(ls):
        offs = np.zeros(len(ls) + 1, np.uint32); offs[1:] = np.cumsum([len(x) for x in ls])
        return offs, np.concatenate(ls)
    o0, l0 = csr(lists[0]); o1, l1 = csr(lists[1])
    eng = _engine(g)
    eng.set_reads(reads)
    st = pc.check_verify(eng, oracle, g, reads, (o0, l0, o1, l1), 3, -1)
    assert st["lv_ran"] >= 1 and st["gapped"] >= 1 and st["mapped"] >= 20
    pc.check_verify(eng, oracle, g, reads, (o0, l0, o1, l1), 3, 3)
    # the same lists through the chunk pipeline and the compact transport (32-bit counts: 262144 > 65535)
    pc.check_verify_batch(eng, reads, (o0, l0, o1, l1), 5)
    pc.check_verify_packed(eng, [r for r in reads], (o0, l0, o1, l1), 5, variants=[(2, 32, 3)])
    eng.close()


@pytest.mark.gpu
def test_large_coordinates(oracle):
    """a reference longer than 2^31 bases: loci just below / across 2^31, in the last KiB before l, pos + L + 4 >= l,
    pos >= l and pos near 2^32 -- the 32-bit position arithmetic of every kernel on the path"""
    L = 100
    l = (1 << 31) + 12_345_677
    centres = [300, (1 << 31) - 200, (1 << 31) - 40, (1 << 31) + 250, (1 << 30) + 77, 3_000_000_001 % l, l - 900, l - 245, l - 60,
               (1 << 31) + 8_000_000, 123_456_789, l - 5000]
    w = pc.SparseWorld(l, centres, L, seed=5)
    extra = [l - L - 4, l - L - 3, l - L, l - L + 1, l - 1, l, l + 5, 0xFFFFFFF0, 0xFFFFFFFF - L]
    reads, cands = w.reads_and_candidates(12, seed=6, extra_loci=extra)
    eng = api.Engine(w.mixref, w.l, None, 0, device=0)
    eng.set_reads(reads)
    st = pc.check_verify(eng, oracle, w, reads, cands, 3, -1)
    assert st["mapped"] >= 60 and st["gapped"] >= 5
    pc.check_verify(eng, oracle, w, reads, cands, 3, 3)
    # per-pair entry points on the same loci
    pairs = pc.flat_pairs(cands, len(reads))
    pc.check_mismatch(eng, oracle, w, reads, pairs, 3)
    pc.check_lv(eng, oracle, w, reads, pairs[::3], 10)
    # mate-rescue windows at large coordinates, one of them clamped to end == l
    wins = np.zeros(len(reads), api.WIN_DT)
    offs0, loci0, offs1, loci1 = cands
    for i in range(len(reads)):
        a, b, _ = w.windows[i // 12] if i // 12 < len(w.windows) else w.windows[-1]
        wins[i] = ((i << 1) | (0 if offs0[i + 1] - offs0[i] > offs1[i + 1] - offs1[i] else 1), a, min(b + 30, l))
    gapped = pc.check_ssw(eng, oracle, w, reads, wins, False, api.salt_score_mat2(), 16)
    assert gapped >= 3
    eng.close()


@pytest.mark.gpu
@pytest.mark.parametrize("L", [100, 150])
def test_chunk_pair_stage(L):
    """the paired-end stage of a whole chunk (salt_chunk_pair) == the stage composed pair by pair"""
    from salt_b200 import host_api
    hostlib = host_api.load()
    g = synth.Genome(300000, snp_rate=0.01, seed=70 + L)
    eng = _engine(g)
    st = pc.check_chunk_pair(eng, hostlib, g, 400, L, seed=9 + L)
    assert st.windows16 + st.windows5 >= 40 and st.rescued >= 20 and st.proper >= 250
    eng.close()


@pytest.mark.gpu
def test_chunk_pair_stage_two_driver_threads():
    """two host threads, each with its own attached handle (salt_b200_attach) and its own chunk, run the paired-end stage of
    different chunks at the same time on one device: each must equal its pair-by-pair composition"""
    import threading
    from salt_b200 import host_api
    hostlib = host_api.load()
    g = synth.Genome(300000, snp_rate=0.01, seed=171)
    eng = _engine(g)
    engs = [eng, eng.attach(), eng.attach()]
    out = [None] * len(engs); errs = []

    def drive(i):
        try:
            for rep in range(3):
                out[i] = pc.check_chunk_pair(engs[i], hostlib, g, 300, 100, seed=40 + 7 * i + rep)
        except BaseException as ex:                      # noqa: BLE001
            errs.append((i, repr(ex)))
    ths = [threading.Thread(target=drive, args=(i,)) for i in range(len(engs))]
    for t in ths: t.start()
    for t in ths: t.join()
    assert not errs, errs
    assert all(o is not None and o.rescued >= 10 for o in out)
    for e in engs[1:]:
        e.close()
    eng.close()


@pytest.mark.gpu
def test_multi_gpu_in_process_matches_single():
    """salt_multi_*: one process, one handle per device, contiguous shares on their own host threads -- the output must
    equal the single-device output byte for byte (two handles on one device when the box has a single GPU)"""
    import ctypes as C
    from salt_b200 import host_api
    hostlib = host_api.load()
    g, reads, pos, strand, cands = pc.make_world(555, glen=400000, L=100, n_reads=9000, per_strand=6, indel_frac=0.3, n_frac=0.005)
    offs0, loci0, offs1, loci1 = cands
    n, L = reads.shape
    roffs = (np.arange(n + 1) * L).astype(np.uint32)
    eng = _engine(g)
    pk, keep = eng.packed_chunk(reads, roffs, offs0, loci0, offs1, loci1, bits=2)
    want = eng.verify_batch_packed(pk, len(loci0), len(loci1), chunk_reads=700)
    ndev = int(eng.L.salt_b200_device_count())
    for devs in ([0, 1 % ndev], [0, 1 % ndev, 2 % ndev]):
        darr = (C.c_int * len(devs))(*devs)
        m = hostlib.salt_multi_init(g.mixref.ctypes.data, g.l, g.pac.ctypes.data, g.l, darr, len(devs))
        assert m and hostlib.salt_multi_n(m) == len(devs)
        rec = np.zeros(n, api.VERIFY_DT); acc0 = np.empty(len(loci0), np.int8); acc1 = np.empty(len(loci1), np.int8)
        cig = np.zeros((n, 128), np.uint8)
        rc = hostlib.salt_multi_verify_batch_packed(m, C.byref(pk), 700, 3, -1, rec.ctypes.data, acc0.ctypes.data, acc1.ctypes.data,
                                                    cig.ctypes.data, 128)
        assert rc == 0
        for a, b, name in zip((rec, acc0, acc1, cig), want, ("rec", "acc0", "acc1", "cigars")):
            assert a.tobytes() == b.tobytes(), (name, devs)
        hostlib.salt_multi_destroy(m)
    eng.close()


@pytest.mark.gpu
@pytest.mark.parametrize("L", [100, 250])
def test_tail_primaries(oracle, L):
    """tags of a verified chunk's primaries from device-resident records and CIGARs, MD strings packed"""
    g, reads, pos, strand, cands = pc.make_world(808 + L, L=L, n_reads=3000, per_strand=4, indel_frac=0.4, glen=300000, sub_rate=0.03)
    eng = _engine(g)
    assert pc.check_tail_primaries(eng, oracle, g, reads, cands) >= 200
    eng.close()


@pytest.mark.gpu
def test_long_cigars_through_the_slim_download(oracle):
    """CIGAR strings of 32+ characters: the slimmed eager download falls back to the full row"""
    g, reads, cands = pc.check_long_cigars(None, oracle, 31, n_reads=400)
    eng = api.Engine(g.mixref, g.l, None, 0, device=0)
    eng.set_reads(reads)
    st = pc.check_verify(eng, oracle, g, reads, cands, 3, -1)
    rec, _, _, cig = eng.verify(*cands)
    assert st["gapped"] >= 120 and sum(len(api.cstr(c)) >= 32 for c in cig) >= 30
    pc.check_verify_batch(eng, reads, cands, 64)
    eng.close()
