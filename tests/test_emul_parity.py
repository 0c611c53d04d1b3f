"""CPU logic tests: the engine's kernel sources compiled on the SIMT emulator (tests/emul) and
driven through the same C ABI, checked bit-exactly against the oracle.  The emulator is test
infrastructure; the GPU tests in test_gpu_parity.py check the real CUDA library."""
import ctypes as C
import os
import sys

import numpy as np
import pytest

import parity_cases as pc
from salt_b200 import api, synth

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "emul"))


@pytest.fixture(scope="module")
def emul_lib():
    import build_emul
    return api._declare(C.CDLL(build_emul.build()))


def _engine(emul_lib, g, with_pac=False):
    return api.Engine(g.mixref, g.l, g.pac if with_pac else None, g.l if with_pac else 0, lib=emul_lib)


@pytest.mark.parametrize("L", [100, 150, 250, 37])
def test_mismatch(emul_lib, oracle, L):
    g, reads, pos, strand, cands = pc.make_world(100 + L, L=L, n_reads=24, per_strand=4, indel_frac=0.0, sub_rate=0.004)
    eng = _engine(emul_lib, g)
    eng.set_reads(reads)
    pairs = pc.flat_pairs(cands, len(reads))
    extra = api.Engine.make_pairs([0, 1, 2], [0, 1, 0], [g.l - L, g.l - L + 1, g.l - 1])   # last fits, others run off the end
    pairs = np.concatenate([pairs, extra])
    got = pc.check_mismatch(eng, oracle, g, reads, pairs, 3)
    assert (got >= 0).sum() >= len(reads) // 2
    pc.check_mismatch(eng, oracle, g, reads, pairs[:40], 0)
    pc.check_mismatch(eng, oracle, g, reads, pairs[:40], 12)


@pytest.mark.parametrize("L,k", [(100, -1), (100, 3), (150, -1), (250, -1), (100, 7), (64, 2), (100, 30)])
def test_lv(emul_lib, oracle, L, k):
    g, reads, pos, strand, cands = pc.make_world(200 + L + k, L=L, n_reads=16, per_strand=4, indel_frac=0.6, sub_rate=0.03)
    eng = _engine(emul_lib, g)
    eng.set_reads(reads)
    pairs = pc.flat_pairs(cands, len(reads))
    extra = api.Engine.make_pairs([0, 1], [0, 1], [g.l - L - 4, g.l - L - 3])
    pairs = np.concatenate([pairs, extra])
    got = pc.check_lv(eng, oracle, g, reads, pairs, k)
    eng.set_lv_mapping(1)          # the warp-per-pair kernel must agree with the thread-per-pair one
    assert np.array_equal(eng.lv(pairs, k), got)
    assert (got > 0).sum() >= 3


@pytest.mark.parametrize("L,k", [(100, -1), (100, 3), (64, 6), (150, -1), (250, -1), (40, 4)])
def test_lv_filter_adversarial(emul_lib, oracle, L, k):
    g = synth.Genome(30000, snp_rate=0.02, seed=7)
    eng = _engine(emul_lib, g)
    found, at_k = pc.check_lv_filter(eng, oracle, g, L, k, 40 if L <= 150 else 16, seed=900 + L + k)
    assert found >= 8


@pytest.mark.parametrize("L", [100, 250])
def test_lv_cigar(emul_lib, oracle, L):
    g, reads, pos, strand, cands = pc.make_world(300 + L, L=L, n_reads=24, per_strand=3, indel_frac=0.8, sub_rate=0.02)
    eng = _engine(emul_lib, g)
    eng.set_reads(reads)
    rng = np.random.default_rng(L)
    # true loci (mostly gapped alignments) and a few decoys, with assorted k and tiny buffers
    rid = np.arange(len(reads), dtype=np.uint32)
    pairs = api.Engine.make_pairs(rid, strand, pos)
    k_each = rng.choice([2, 5, 10, 25, 30], len(pairs)).astype(np.uint8)
    gapped = pc.check_lv_cigar(eng, oracle, g, reads, pairs, k_each, 128)
    assert gapped >= 5
    pc.check_lv_cigar(eng, oracle, g, reads, pairs, k_each, 6)
    # thresholds the thread-per-pair kernels take (k <= 4, <= 10, <= 15), then the same through the warp kernel
    for mapping in (0, 1):
        eng.set_lv_mapping(mapping)
        for ks in ([1, 2, 3, 4], [2, 5, 10], [12, 15]):
            kk = rng.choice(ks, len(pairs)).astype(np.uint8)
            pc.check_lv_cigar(eng, oracle, g, reads, pairs, kk, 128)
            pc.check_lv_cigar(eng, oracle, g, reads, pairs, kk, 5)
    eng.set_lv_mapping(0)
    pc.check_lv_cigar(eng, oracle, g, reads, pc.flat_pairs(cands, len(reads))[:24], np.full(24, 10, np.uint8), 256)


@pytest.mark.parametrize("L,lv_T0", [(100, -1), (100, 3), (150, -1), (128, 3), (256, 3), (300, 3), (700, 3)])
def test_verify_stage(emul_lib, oracle, L, lv_T0):
    g, reads, pos, strand, cands = pc.make_world(400 + L, L=L, n_reads=48 if L <= 150 else 12, per_strand=5, indel_frac=0.35,
                                                 sub_rate=0.025 if L <= 150 else 0.004, glen=30000)
    # a read with no candidates at all and duplicated loci must behave like the reference's loops
    offs0, loci0, offs1, loci1 = cands
    loci0 = loci0.copy(); loci0[offs0[3] + 1] = loci0[offs0[3]]
    eng = _engine(emul_lib, g)
    eng.set_reads(reads)
    st = pc.check_verify(eng, oracle, g, reads, (offs0, loci0, offs1, loci1), 3, lv_T0)
    if L in (100, 150):
        assert st["lv_ran"] >= 5 and st["gapped"] >= 2 and st["mapped"] >= 30


def test_verify_long_lists(emul_lib, oracle):
    """lists longer than one lane group (several chunks), duplicates that straddle a chunk boundary,
    loci at/after the end of the reference"""
    g, reads, pos, strand, cands = pc.make_world(91, L=100, n_reads=10, per_strand=45, indel_frac=0.3, glen=30000)
    offs0, loci0, offs1, loci1 = cands
    loci0 = loci0.copy(); loci1 = loci1.copy()
    for r in range(10):
        b = int(offs0[r])
        loci0[b + 15] = loci0[b + 16] = loci0[b + 17]          # run of equal loci across lanes 15|16
        loci0[b + 31] = loci0[b + 32]
        e = int(offs1[r + 1])
        loci1[e - 1] = g.l + 5; loci1[e - 2] = g.l - 50; loci1[e - 3] = g.l - 50   # past the end / too close to it
    eng = _engine(emul_lib, g)
    eng.set_reads(reads)
    pc.check_verify(eng, oracle, g, reads, (offs0, loci0, offs1, loci1), 3, -1)
    pc.check_verify(eng, oracle, g, reads, (offs0, loci0, offs1, loci1), 3, 3)


def test_verify_batch_pipeline(emul_lib, oracle):
    """chunks of 7 reads through the 4 slots: slot reuse, last partial chunk, compact CIGAR return"""
    g, reads, pos, strand, cands = pc.make_world(123, L=100, n_reads=60, per_strand=4, indel_frac=0.4, glen=30000)
    eng = _engine(emul_lib, g)
    rec = pc.check_verify_batch(eng, reads, cands, 7)
    assert (rec["is_gap"] == 1).sum() >= 3
    pc.check_verify_batch(eng, reads, cands, 1000)


def test_verify_packed_transport(emul_lib, oracle):
    """compact transport: every packing flavour, uniform and ragged read lengths, reads with N, views starting mid-byte"""
    g, reads, pos, strand, cands = pc.make_world(321, L=100, n_reads=30, per_strand=3, indel_frac=0.4, glen=30000, n_frac=0.02)
    eng = _engine(emul_lib, g)
    assert pc.check_verify_packed(eng, [r for r in reads], cands, 7) == 8
    # ragged: reads cut to different lengths (their candidate lists stay valid loci)
    rng = np.random.default_rng(5)
    ragged = [r[:int(rng.integers(37, 101))] for r in reads]
    pc.check_verify_packed(eng, ragged, cands, 11, lv_T0=3, variants=[(2, 16, 3), (4, 32, 0)])
    # an empty batch
    z = np.zeros(1, np.uint32); e = np.zeros(0, np.uint32)
    pk, keep = eng.packed_chunk(np.zeros(0, np.uint8), z, z, e, z, e)
    rec, a0, a1, cig = eng.verify_batch_packed(pk, 0, 0, 7)
    assert len(rec) == 0


def test_md_nm(emul_lib, oracle):
    """SAM tail kernel (MD/NM/XV, sam.c:246-328) against the oracle"""
    g = synth.Genome(40000, snp_rate=0.03, seed=15)
    reads, pos, strand = synth.sample_reads(g, 60, 100, seed=16, sub_rate=0.03, indel_frac=0.3, n_frac=0.01)
    eng = _engine(emul_lib, g, with_pac=True)
    assert pc.check_md_nm(eng, oracle, g, reads, pos, strand, 17, md_stride=264) >= 3


def test_host_layer_chunks(emul_lib, oracle):
    """include/salt_host.h over the (emulated) engine: pinned chunk queues, slots, query_set_hits/mapq/cigar"""
    import build_emul
    from salt_b200 import host_api
    hostlib = host_api.load(build_emul.build_host())
    g, reads, pos, strand, cands = pc.make_world(222, L=100, n_reads=50, per_strand=5, indel_frac=0.4, glen=30000)
    eng = _engine(emul_lib, g, with_pac=True)
    assert pc.check_host_chunks(eng, hostlib, oracle, g, reads, cands, 9, with_tail=True) >= 3
    assert pc.check_host_chunks(eng, hostlib, oracle, g, reads, cands, 9, 3, 3, max_hits=2) >= 1


def test_verify_empty_lists(emul_lib, oracle):
    g, reads, pos, strand, cands = pc.make_world(77, L=100, n_reads=6, per_strand=3, glen=20000)
    z = np.zeros(7, np.uint32); e = np.zeros(0, np.uint32)
    eng = _engine(emul_lib, g)
    eng.set_reads(reads)
    rec, acc0, acc1, cig = eng.verify(z, e, z, e)
    assert (rec["pos"] == 0xFFFFFFFF).all() and (rec["lv_ran"] == 1).all()


@pytest.mark.parametrize("L,width", [(100, 401), (150, 401), (250, 301), (64, 200)])
def test_ssw_mixref(emul_lib, oracle, L, width):
    g, reads, pos, strand, _ = pc.make_world(500 + L, L=L, n_reads=18, per_strand=2, indel_frac=0.7,
                                              sub_rate=0.04, glen=20000)
    eng = _engine(emul_lib, g)
    eng.set_reads(reads)
    rng = np.random.default_rng(L)
    wins = pc.make_windows(g, reads, pos, strand, L, rng, width)
    gapped = pc.check_ssw(eng, oracle, g, reads, wins, False, api.salt_score_mat2(), 16)
    assert gapped >= 3


def test_ssw_wide_bands_and_end_at_l(emul_lib, oracle):
    g = synth.Genome(20003, snp_rate=0.01, n_rate=0.0, seed=77)
    eng = _engine(emul_lib, g)
    assert pc.check_ssw_wide_bands(eng, oracle, 77) >= 120


def test_ssw_narrow_bands_every_branch(emul_lib, oracle):
    """bands 1..3 in registers, the doubling 1 -> 2, the hand-over to the general kernel; mixRef and pac scoring"""
    g = synth.Genome(20011, snp_rate=0.01, n_rate=0.0, seed=9)
    eng = _engine(emul_lib, g, with_pac=True)
    assert pc.check_ssw_narrow_bands(eng, oracle, 9) >= 60


def test_ssw_pac_and_params(emul_lib, oracle):
    L = 100
    g, reads, pos, strand, _ = pc.make_world(600, L=L, n_reads=11, per_strand=2, indel_frac=0.7, sub_rate=0.04,
                                              glen=20000, n_rate=0.0, n_frac=0.01)
    eng = _engine(emul_lib, g, with_pac=True)
    eng.set_reads(reads)
    rng = np.random.default_rng(6)
    wins = pc.make_windows(g, reads, pos, strand, L, rng, 401)
    pc.check_ssw(eng, oracle, g, reads, wins, True, api.salt_score_mat(), 5)
    pc.check_ssw(eng, oracle, g, reads, wins, True, api.salt_score_mat(), 5, gapO=5, gapE=2, mask_len=15)
    pc.check_ssw(eng, oracle, g, reads, wins, True, api.salt_score_mat(), 5, flag=0)
    pc.check_ssw(eng, oracle, g, reads, wins, True, api.salt_score_mat(), 5, flag=1, mask_len=7)
    with pytest.raises(api.SaltError):
        eng.ssw(wins, api.salt_score_mat(), 5, True, gapO=1, gapE=1)


def test_build_mixref_on_device(emul_lib, oracle):
    g = synth.Genome(5000, snp_rate=0.05, n_rate=0.01, seed=9)
    rows = g.snp_table()
    fasta = g.fasta()
    want, l = oracle.build_mixref([("chr1", fasta)], rows)
    pos = np.array([r[1] - 1 for r in rows], np.uint32)
    mask = np.array([oracle.lib.orc_allele_mask(r[2].encode()) for r in rows], np.uint8)
    eng = api.Engine.from_bases(fasta, pos, mask, lib=emul_lib)
    assert np.array_equal(eng.get_mixref(), want)
    assert np.array_equal(want, g.mixref)


def test_seed_locate_lists(emul_lib, tmp_path):
    """row f1 on the emulator: the device's candidate lists equal the reference's own alnse_seed_overlap +
    alnse_locate_alt (libsaltref_seed.so) on an index written by the reference's salt-idx -- repeat-rich genome,
    ragged reads, N in reads and reference, six option sets"""
    import seed_cases as sc
    from oracle import orc
    from salt_b200 import index_io
    if not sc.have_ref():
        pytest.skip("oracle/_ref not built (reference tree absent at build time)")
    rng = np.random.default_rng(3)
    g, is_n = sc.repeat_genome(rng, n_units=24, unit_len=900, n_rate=0.001)
    prefix = sc.write_index(str(tmp_path), g, is_n, rng)
    fm = index_io.FmIndex(prefix)
    codes, roffs = sc.sample_reads(g, rng, 14)
    ref = orc.SeedRef(prefix)
    eng = api.Engine(fm.mixref, fm.l, None, 0, lib=emul_lib)
    eng.set_index(fm)
    assert sc.check_lists(eng, ref, fm, codes, roffs) > 500
    # the paired-end program's flavour of locate (alnse_locate)
    tot, flagged = sc.check_lists_pe(eng, ref, fm, codes, roffs, option_sets=sc.PE_OPTION_SETS[:2])
    assert tot > 200
    # and the verification stage on the lists the seeding left on the device == on the same lists uploaded
    opt = api.Engine.seed_opt(fm.l_seed, 0, 50, 500)
    offs0, loci0, offs1, loci1 = eng.seed_locate(opt)
    got = eng.verify_seeded(len(loci0), len(loci1))
    want = eng.verify(offs0, loci0, offs1, loci1)
    for a, b, name in zip(got, want, ("rec", "acc0", "acc1", "cigars")):
        assert a.tobytes() == b.tobytes(), name
    ref.close(); eng.close()


def test_seed_error_paths(emul_lib, tmp_path):
    """the seeding entry points refuse what they cannot reproduce exactly instead of guessing (same codes on the GPU build:
    the checks are host code in engine.cu)"""
    import ctypes as C
    import seed_cases as sc
    from salt_b200 import index_io
    if not sc.have_ref():
        pytest.skip("oracle/_ref not built (reference tree absent at build time)")
    rng = np.random.default_rng(5)
    g, is_n = sc.repeat_genome(rng, n_units=8, unit_len=700, n_rate=0.0)
    prefix = sc.write_index(str(tmp_path), g, is_n, rng)
    fm = index_io.FmIndex(prefix)
    codes, roffs = sc.sample_reads(g, rng, 6)
    n_reads = len(roffs) - 1
    eng = api.Engine(fm.mixref, fm.l, None, 0, lib=emul_lib)
    eng.set_reads(codes, roffs)
    good = api.Engine.seed_opt(fm.l_seed, 0, 50, 500)

    def code_of(opt, **kw):
        try:
            eng.seed_locate(opt, **kw)
        except api.SaltError as ex:
            return ex.code
        return 0
    assert code_of(good) == -101                                   # no index uploaded yet
    eng.set_index(fm)
    assert code_of(good) == 0
    assert code_of(api.Engine.seed_opt(8, 8, 50, 500)) == -101      # seed shorter than the lookup table's k-mer
    bad = api.Engine.seed_opt(fm.l_seed, 0, 50, 500); bad.l_overlap = 0
    assert code_of(bad) == -104                                    # the reference's non-overlap seeding is another function
    assert code_of(api.Engine.seed_opt(fm.l_seed, 0, -1, 500)) == -101
    assert code_of(api.Engine.seed_opt(fm.l_seed, 0, 50, 0)) == -104
    assert code_of(api.Engine.seed_opt(fm.l_seed, 0, 50, 20000)) == -104
    assert code_of(api.Engine.seed_opt(fm.l_seed, 0, 50, 500, locate_mode=2)) == -101
    assert code_of(api.Engine.seed_opt(fm.l_seed, 0, 50, 500, locate_mode=1, list_cap=8)) == -101
    assert code_of(api.Engine.seed_opt(fm.l_seed, 0, 50, 500, locate_mode=1, list_cap=4096)) == 0
    # room for the download smaller than a strand's list: SALT_ERR_NOMEM, and the totals are still reported
    n0 = C.c_size_t(0); n1 = C.c_size_t(0)
    offs0 = np.zeros(n_reads + 1, np.uint32); offs1 = np.zeros(n_reads + 1, np.uint32)
    tiny = np.zeros(1, np.uint32)
    rc = eng.L.salt_b200_seed_locate(eng.h, 0, C.byref(good), api._ptr(offs0), api._ptr(offs1), api._ptr(tiny), 1, api._ptr(tiny), 1,
                                     C.byref(n0), C.byref(n1))
    assert rc == -103 and n0.value > 1
    # verification of seeded lists needs a seeded slot
    eng.set_reads(codes, roffs)
    with pytest.raises(api.SaltError):
        eng.verify_seeded(10, 10)
    eng.close()


def test_transport_and_handle_error_paths(emul_lib):
    """malformed compact chunks, slots out of range, tags of a slot that holds no verified chunk, reloading the reference of a
    handle that only borrows it: refused with SALT_ERR_ARG, nothing guessed"""
    g, reads, pos, strand, cands = pc.make_world(5, L=100, n_reads=12, per_strand=3, indel_frac=0.3, glen=20000)
    eng = api.Engine(g.mixref, g.l, g.pac, g.l, lib=emul_lib)
    offs0, loci0, offs1, loci1 = cands
    roffs = (np.arange(len(reads) + 1) * 100).astype(np.uint32)
    res = eng.packed_chunk(reads.reshape(-1), roffs, offs0, loci0, offs1, loci1)
    pk = res[0] if isinstance(res, tuple) else res

    def code_of(fn):
        try:
            fn()
        except api.SaltError as ex:
            return ex.code
        return 0
    run = lambda: eng.verify_batch_packed(pk, len(loci0), len(loci1), 5)
    assert code_of(run) == 0
    for field, val in (("base_bits", 3), ("base_bits", 0), ("count_bits", 8), ("bases", None), ("l_seq", 0)):
        old = getattr(pk, field)
        setattr(pk, field, val)
        assert code_of(run) == -101, field
        setattr(pk, field, old)
    assert code_of(run) == 0
    eng.set_reads(reads)
    assert code_of(lambda: eng.tail_primaries()) == -101            # nothing verified in the slot since its reads changed
    assert code_of(lambda: eng._ck(eng.L.salt_b200_use_slot(eng.h, 99))) == -101
    assert code_of(lambda: eng._ck(eng.L.salt_b200_use_slot(eng.h, -1))) == -101
    other = eng.attach()
    assert code_of(lambda: other._ck(eng.L.salt_b200_reload_ref(other.h, api._ptr(g.mixref), g.l, None, 0))) == -101
    assert not eng.L.salt_b200_attach(None)
    other.close(); eng.close()


def test_chunk_queue_misuse_is_refused(emul_lib):
    """the host layer's chunk queues keep their own state: results before (or without) a submit, waiting on the wrong slot or
    for nothing, appending to a chunk that is in flight or holds results, a second submit: SALT_ERR_ARG, never stale data"""
    import build_emul
    from salt_b200 import host_api
    H = host_api.load(build_emul.build_host())
    g, reads, pos, strand, cands = pc.make_world(5, L=100, n_reads=12, per_strand=3, indel_frac=0.3, glen=20000)
    eng = api.Engine(g.mixref, g.l, g.pac, g.l, lib=emul_lib)
    offs0, loci0, offs1, loci1 = cands
    n = len(reads); roffs = (np.arange(n + 1) * 100).astype(np.uint32)
    ch = host_api.Chunk(H, n + 2, (n + 2) * 100, len(loci0) + len(loci1) + 8)

    def code_of(fn):
        try:
            fn()
        except api.SaltError as ex:
            return ex.code
        return 0
    add = lambda: ch.add_reads(reads.reshape(-1), roffs, offs0, loci0, offs1, loci1)
    assert code_of(lambda: ch.result(0)) == -101
    assert code_of(lambda: ch.wait(eng, 0)) == -101                 # nothing was submitted
    assert code_of(add) == 0
    assert code_of(lambda: ch.result(0)) == -101                    # queued, not verified
    assert code_of(lambda: ch.pair(eng, 0, n // 2, 250, 550, g.l)) == -101
    assert code_of(lambda: ch.submit(eng, 9, 3, 3)) == -101
    assert code_of(lambda: ch.submit(eng, 0, 3, 3)) == 0
    assert code_of(lambda: ch.submit(eng, 0, 3, 3)) == -101         # already in flight
    assert code_of(add) == -101                                     # ... and its queues are not the caller's to write
    assert code_of(lambda: ch.wait(eng, 1)) == -101                 # it went to slot 0
    assert code_of(lambda: ch.wait(eng, 0)) == 0
    assert code_of(lambda: ch.wait(eng, 0)) == 0                    # waiting twice is harmless
    assert code_of(lambda: ch.result(n)) == -101
    assert code_of(lambda: ch.result(0, 0)) == -101 and code_of(lambda: ch.result(0, 17)) == -101
    assert code_of(lambda: ch.result(0)) == 0
    assert code_of(add) == -101                                     # holds results: reset first
    ch.reset()
    assert code_of(lambda: ch.result(0)) == -101
    assert code_of(add) == 0
    assert code_of(lambda: ch.seed_verify(eng, api.Engine.seed_opt(19, 0, 50, 500))) == -101      # no index uploaded
    ch.close(); eng.close()


def test_fastq_text_to_verified_records(emul_lib):
    """the input side end to end: FASTQ text -> salt_fastq_pack -> the arrays it filled ARE a salt_packed_chunk_t (2-bit bases,
    N positions, lengths) -> salt_b200_verify_batch_packed, against the same reads uploaded one byte per base; the text is also
    cut with salt_fastq_split and every part sent as a chunk of its own"""
    import ctypes as C
    import build_emul
    from salt_b200 import host_api
    from test_fastq_pack import FastqT
    H = host_api.load(build_emul.build_host())
    H.salt_fastq_pack.argtypes = [C.c_char_p, C.c_size_t, C.c_int, C.c_uint32, C.POINTER(FastqT), C.POINTER(C.c_size_t)]
    H.salt_fastq_split.argtypes = [C.c_char_p, C.c_size_t, C.c_int, C.POINTER(C.c_size_t)]
    g, reads, pos, strand, cands = pc.make_world(909, L=100, n_reads=40, per_strand=3, indel_frac=0.3, glen=30000, n_frac=0.03)
    rng = np.random.default_rng(4)
    reads_list = [r[:int(rng.integers(60, 101))] for r in reads]            # ragged
    offs0, loci0, offs1, loci1 = cands
    text = b"".join(b"@r%d/1 c\n%s\n+\n%s\n" % (i, bytes(b"ACGTN"[c] for c in r), b"I" * len(r)) for i, r in enumerate(reads_list))
    eng = api.Engine(g.mixref, g.l, g.pac, g.l, lib=emul_lib)
    lens = np.array([len(r) for r in reads_list]); roffs = np.concatenate([[0], np.cumsum(lens)]).astype(np.uint32)
    eng.set_reads(np.concatenate(reads_list).astype(np.uint8), roffs)
    want = eng.verify(offs0, loci0, offs1, loci1, 3, -1)

    def run_part(t, first_read):
        """parse one text, send what the parser filled as one compact chunk; returns (n_reads, results)"""
        cap = len(t) + 8
        bases = np.zeros(cap // 4 + 2, np.uint8); n_pos = np.zeros(cap, np.uint32)
        m = 64
        arrs = {k: np.zeros(m, dt) for k, dt in (("lens", np.uint16), ("n_ambiguous", np.uint16), ("name_off", np.uint32),
                                                  ("name_len", np.uint16), ("comment_off", np.uint32), ("comment_len", np.uint16),
                                                  ("qual_off", np.uint32))}
        fq = FastqT(bases.ctypes.data, cap, n_pos.ctypes.data, cap, *(arrs[k].ctypes.data for k in
                    ("lens", "n_ambiguous", "name_off", "name_len", "comment_off", "comment_len", "qual_off")), 0, 0, 0)
        used = C.c_size_t(0)
        n = H.salt_fastq_pack(t, len(t), 1, m, C.byref(fq), C.byref(used))
        assert n > 0
        a, b = first_read, first_read + n
        c0 = np.diff(offs0[a:b + 1].astype(np.int64)).astype(np.uint16); c1 = np.diff(offs1[a:b + 1].astype(np.int64)).astype(np.uint16)
        l0 = np.ascontiguousarray(loci0[offs0[a]:offs0[b]]); l1 = np.ascontiguousarray(loci1[offs1[a]:offs1[b]])
        pk = api.PackedChunkT()
        pk.n_reads = n; pk.base_bits = 2; pk.bases = bases.ctypes.data; pk.base_start = 0
        pk.lens = arrs["lens"].ctypes.data; pk.l_seq = 0
        pk.n_pos = n_pos.ctypes.data if fq.n_n else None; pk.n_n = fq.n_n; pk.count_bits = 16
        pk.n_cand[0], pk.n_cand[1] = c0.ctypes.data, c1.ctypes.data
        pk.loci[0], pk.loci[1] = (l0.ctypes.data if len(l0) else None), (l1.ctypes.data if len(l1) else None)
        return n, eng.verify_batch_packed(pk, len(l0), len(l1), 16, 3, -1)

    n, got = run_part(text, 0)
    assert n == len(reads_list)
    for x, y, name in zip(got, want, ("rec", "acc0", "acc1", "cigars")):
        assert x.tobytes() == y.tobytes(), name
    cuts = (C.c_size_t * 5)()
    made = H.salt_fastq_split(text, len(text), 4, cuts)
    assert made >= 2
    first = 0
    for k in range(made):
        n, got = run_part(text[cuts[k]:cuts[k + 1]], first)
        assert got[0].tobytes() == want[0][first:first + n].tobytes(), ("rec of part", k)
        assert got[1].tobytes() == want[1][offs0[first]:offs0[first + n]].tobytes() and got[2].tobytes() == want[2][offs1[first]:offs1[first + n]].tobytes()
        assert got[3].tobytes() == want[3][first:first + n].tobytes()
        first += n
    assert first == len(reads_list)
    eng.close()


def test_chunk_pair_stage(emul_lib, oracle):
    """the paired-end stage of a chunk in one call == the stage composed pair by pair"""
    import build_emul
    from salt_b200 import host_api
    hostlib = host_api.load(build_emul.build_host())
    g = synth.Genome(40000, snp_rate=0.01, seed=77)
    eng = _engine(emul_lib, g, with_pac=True)
    st = pc.check_chunk_pair(eng, hostlib, g, 24, 100, seed=9)
    assert st.windows16 + st.windows5 >= 3 and st.rescued >= 1 and st.proper >= 10
    eng.close()


def test_chunk_pair_stage_on_host_threads(emul_lib, oracle):
    """the host layer's per-read / per-pair loops on its persistent thread pool: same results with 1, 3 and 5 host threads
    (the pool is rebuilt when the count changes), and from two driver threads with their own handles and pools at once"""
    import threading
    import build_emul
    from salt_b200 import host_api
    hostlib = host_api.load(build_emul.build_host())
    g = synth.Genome(40000, snp_rate=0.01, seed=78)
    eng = _engine(emul_lib, g, with_pac=True)
    hostlib.salt_host_set_grain(2)                       # 24 pairs are enough to spread over the threads
    try:
        for nthr in (3, 1, 5):
            hostlib.salt_host_set_threads(nthr)
            st = pc.check_chunk_pair(eng, hostlib, g, 24, 100, seed=9)
            assert st.windows16 + st.windows5 >= 3 and st.rescued >= 1
        hostlib.salt_host_set_threads(4)
        engs = [eng, eng.attach()]
        errs = []

        def drive(i):
            try:
                for rep in range(2):
                    pc.check_chunk_pair(engs[i], hostlib, g, 20, 100, seed=30 + 5 * i + rep)
            except BaseException as ex:                  # noqa: BLE001
                errs.append((i, repr(ex)))
        ths = [threading.Thread(target=drive, args=(i,)) for i in range(2)]
        for t in ths: t.start()
        for t in ths: t.join()
        assert not errs, errs
        engs[1].close()
    finally:
        hostlib.salt_host_set_threads(1); hostlib.salt_host_set_grain(256)
    eng.close()


def test_tail_primaries(emul_lib, oracle):
    g, reads, pos, strand, cands = pc.make_world(808, L=100, n_reads=40, per_strand=3, indel_frac=0.4, glen=30000, sub_rate=0.03)
    eng = _engine(emul_lib, g, with_pac=True)
    assert pc.check_tail_primaries(eng, oracle, g, reads, cands) >= 3
    eng.close()


def test_long_cigars_through_the_slim_download(emul_lib, oracle):
    g, reads, cands = pc.check_long_cigars(None, oracle, 31, n_reads=24)
    eng = api.Engine(g.mixref, g.l, None, 0, lib=emul_lib)
    eng.set_reads(reads)
    st = pc.check_verify(eng, oracle, g, reads, cands, 3, -1)
    rec, _, _, cig = eng.verify(*cands)
    assert st["gapped"] >= 10 and max(len(api.cstr(c)) for c in cig) >= 32
    pc.check_verify_batch(eng, reads, cands, 5)
    eng.close()


def test_multi_handles_in_one_process(emul_lib, oracle):
    """salt_multi_*: shares on persistent worker threads, results in input order (two emulated handles)"""
    import ctypes as C
    import build_emul
    from salt_b200 import host_api
    hostlib = host_api.load(build_emul.build_host())
    g, reads, pos, strand, cands = pc.make_world(556, glen=30000, L=100, n_reads=50, per_strand=3, indel_frac=0.3, n_frac=0.01)
    offs0, loci0, offs1, loci1 = cands
    n, L = reads.shape
    roffs = (np.arange(n + 1) * L).astype(np.uint32)
    eng = _engine(emul_lib, g)
    pk, keep = eng.packed_chunk(reads, roffs, offs0, loci0, offs1, loci1, bits=2)
    want = eng.verify_batch_packed(pk, len(loci0), len(loci1), chunk_reads=7)
    devs = (C.c_int * 3)(0, 0, 0)
    m = hostlib.salt_multi_init(g.mixref.ctypes.data, g.l, None, 0, devs, 3)
    assert m and hostlib.salt_multi_n(m) == 3
    for rep in range(2):                                   # the workers are reused across calls
        rec = np.zeros(n, api.VERIFY_DT); acc0 = np.empty(len(loci0), np.int8); acc1 = np.empty(len(loci1), np.int8)
        cig = np.zeros((n, 128), np.uint8)
        assert hostlib.salt_multi_verify_batch_packed(m, C.byref(pk), 7, 3, -1, rec.ctypes.data, acc0.ctypes.data, acc1.ctypes.data,
                                                      cig.ctypes.data, 128) == 0
        for a, b, name in zip((rec, acc0, acc1, cig), want, ("rec", "acc0", "acc1", "cigars")):
            assert a.tobytes() == b.tobytes(), (name, rep)
    hostlib.salt_multi_destroy(m)
    eng.close()
