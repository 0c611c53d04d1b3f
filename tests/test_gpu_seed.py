"""Row f1 on the GPU: seeding + locate (alnse_seed_overlap + alnse_locate_alt) against the reference's own functions
(oracle/_ref/libsaltref_seed.so) on indexes written by the reference's own salt-idx."""
import os

import numpy as np
import pytest

import seed_cases as sc
from salt_b200 import api, index_io, synth

pytestmark = pytest.mark.gpu


def _ref_or_skip():
    if not sc.have_ref():
        pytest.skip("oracle/_ref not built (reference tree absent at build time)")
    from oracle import orc
    return orc


def test_seed_lists_repeat_genome(tmp_path):
    """repeat-rich genome with N, ragged reads with N, every option set: identical sorted lists"""
    orc = _ref_or_skip()
    rng = np.random.default_rng(3)
    g, is_n = sc.repeat_genome(rng, n_units=60, unit_len=1500, n_rate=0.001)
    prefix = sc.write_index(str(tmp_path), g, is_n, rng, records=2)
    fm = index_io.FmIndex(prefix)
    codes, roffs = sc.sample_reads(g, rng, 4000)
    ref = orc.SeedRef(prefix)
    eng = api.Engine(fm.mixref, fm.l, None, 0, device=0)
    eng.set_index(fm)
    assert sc.check_lists(eng, ref, fm, codes, roffs) > 500000
    # the paired-end program's flavour (alnse_locate): per-interval row caps, MAX_LOC_POS in all; strands with an interval the
    # reference subsamples with rand() are flagged instead of compared
    tot, flagged = sc.check_lists_pe(eng, ref, fm, codes, roffs)
    assert tot > 300000 and flagged < 0.2 * 8 * len(roffs)
    # a small list_cap: lists that fill it are flagged, everything else still matches
    tot2, flagged2 = sc.check_lists_pe(eng, ref, fm, codes, roffs, option_sets=sc.PE_OPTION_SETS[:1], list_cap=64)
    assert flagged2 > 0 and tot2 > 10000
    # misuse is refused, not guessed at (the full list of refusals: tests/test_emul_parity.py::test_seed_error_paths, same host code)
    opt = api.Engine.seed_opt(fm.l_seed, 0, 50, 500)
    eng.set_reads(codes[:roffs[50]], roffs[:51])
    n0, n1 = eng.seed_locate(opt, download=False)
    eng.verify_seeded(n0, n1)
    eng.set_reads(codes[:roffs[50]], roffs[:51])                 # new reads in the slot: the lists seeded from the old ones are gone
    with pytest.raises(api.SaltError):
        eng.verify_seeded(n0, n1)
    with pytest.raises(api.SaltError):
        eng.seed_locate(api.Engine.seed_opt(fm.l_seed, 0, 50, 20000), download=False)
    ref.close(); eng.close()


def test_seed_lists_100k_reads(tmp_path):
    """>= 100 k reads on an index built by oracle/_ref/salt-idx, identical aux->loci per strand; then
    the whole single-end stage from reads alone (salt_b200_align_batch_packed) == verification of the reference's lists"""
    orc = _ref_or_skip()
    rng = np.random.default_rng(11)
    g = synth.Genome(2_000_000, snp_rate=0.01, seed=21)
    # a tenth of the genome is a repeat family, so that lists are not all singletons
    unit = g.codes[:3000].copy()
    codes_g = g.codes.copy()
    for i in range(60):
        p = int(rng.integers(10_000, g.l - 10_000)); u = unit.copy(); m = rng.random(3000) < 0.02
        u[m] = rng.integers(0, 4, int(m.sum())); codes_g[p:p + 3000] = u
    prefix = sc.write_index(str(tmp_path), codes_g, np.zeros(g.l, bool), rng, snp_rate=0.01, records=3)
    fm = index_io.FmIndex(prefix)
    n, L = 120_000, 100
    codes, roffs = sc.sample_reads(codes_g, rng, n, ragged=False, sub=0.01, n_every=97)
    ref = orc.SeedRef(prefix)
    want = ref.run(codes, roffs, fm.l_seed, 0, 50, 1000)                       # the program's defaults (aln.c:46-47)
    eng = api.Engine(fm.mixref, fm.l, None, 0, device=0)
    eng.set_index(fm)
    eng.set_reads(codes, roffs)
    opt = api.Engine.seed_opt(fm.l_seed, 0, 50, 1000)
    got = eng.seed_locate(opt)
    for a, b, name in zip(got, want, ("offs0", "loci0", "offs1", "loci1")):
        assert np.array_equal(a, b), (name, len(a), len(b))
    assert int(want[0][-1]) + int(want[2][-1]) > 150_000
    # verification on the device-resident lists == verification of the uploaded reference lists
    rec_s, acc0_s, acc1_s, cig_s = eng.verify_seeded(len(got[1]), len(got[3]))
    rec_u, acc0_u, acc1_u, cig_u = eng.verify(*want)
    assert rec_s.tobytes() == rec_u.tobytes() and acc0_s.tobytes() == acc0_u.tobytes() and acc1_s.tobytes() == acc1_u.tobytes()
    assert cig_s.tobytes() == cig_u.tobytes()
    assert (rec_u["pos"] != 0xFFFFFFFF).sum() > 100_000
    # reads in, records out: the chunk pipeline with seeding on the device
    pk, keep = eng.packed_chunk(codes, roffs, np.zeros(n + 1, np.uint32), np.zeros(0, np.uint32), np.zeros(n + 1, np.uint32),
                                np.zeros(0, np.uint32), bits=2)
    rec_a, cig_a = eng.align_batch_packed(pk, opt, chunk_reads=25_000)
    assert rec_a.tobytes() == rec_u.tobytes() and cig_a.tobytes() == cig_u.tobytes()
    ref.close(); eng.close()


def test_seed_lists_config0(tmp_path):
    """the bundled two-copy lambda genome (BASELINE configs[0]): every seed has two suffix-array rows"""
    orc = _ref_or_skip()
    c0 = os.path.join(sc.REFDIR, "config0")
    if not os.path.exists(os.path.join(c0, "hapmap.txt")):
        pytest.skip("oracle/_ref/config0 not staged")
    import subprocess
    d = str(tmp_path)
    with open(os.path.join(d, "idx.log"), "w") as log:
        subprocess.check_call([os.path.join(sc.REFDIR, "salt-idx"), "-k", "19", os.path.join(c0, "Genome.fa"),
                               os.path.join(c0, "hapmap.txt"), "idx"], cwd=d, stdout=log, stderr=subprocess.STDOUT)
    fm = index_io.FmIndex(os.path.join(d, "idx"))
    reads = []
    for i, ln in enumerate(open(os.path.join(c0, "Read1.fq"))):
        if i % 4 == 1:
            reads.append(np.array(["ACGTN".index(c) for c in ln.strip()], np.uint8))
    codes = np.concatenate(reads); roffs = np.concatenate([[0], np.cumsum([len(r) for r in reads])]).astype(np.uint32)
    ref = orc.SeedRef(os.path.join(d, "idx"))
    eng = api.Engine(fm.mixref, fm.l, None, 0, device=0)
    eng.set_index(fm)
    assert sc.check_lists(eng, ref, fm, codes, roffs, option_sets=((0, 0, 50, 500, 0), (0, 0, 50, 1000, 0))) > 100_000
    ref.close(); eng.close()
