"""salt_pair_plan (include/salt_host.h): pairing2 / pairing_singleton (alnpe.c:94-257, :395-473) re-staged as a plan.
Hand-made cases here (no GPU: the function is host C); tests/test_dropin.py checks it on the GPU box against the
reference's own pairing for every pair of the paired-end run."""
import os
import sys

import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "emul"))
from salt_b200 import host_api

L_PAC = 1_000_000
A, B = 350, 650           # -a / -b of Test/Run_test/run_pe_test.sh:14 -> insert between mates 150..450 for 100 bp reads


@pytest.fixture(scope="module")
def H():
    import build_emul                                   # the host layer linked against the CPU-run engine: loads without a GPU
    return host_api.load(build_emul.build_host())


def plan(H, r0, r1, l0=100, l1=100, a=A, b=B, l_pac=L_PAC):
    return host_api.pair_plan(H, r0, l0, r1, l1, a, b, l_pac)


def test_primaries_pair_as_they_are(H):
    r0 = host_api.make_result(5000, 0, 1, 0); r1 = host_api.make_result(5400, 1, 2, 1)      # gap 300 between end of 0 and start of 1
    assert plan(H, r0, r1) == (0, [(5000, 0, 1, 0), (5400, 1, 2, 1)], [])
    assert plan(H, r1, r0) == (0, [(5400, 1, 2, 1), (5000, 0, 1, 0)], [])                   # mate 1 forward, mate 0 backward
    # both unmapped: nothing to do
    assert plan(H, host_api.make_result(), host_api.make_result()) == (0, None, [])


def test_alternates_pair_with_fewest_errors(H):
    # primaries too far apart, but two combinations of alternates lie in range: the one with fewer differences wins,
    # ties keep the first found (strict <, alnpe.c:146)
    r0 = host_api.make_result(5000, 0, 0, 0, alt0=[(90000, 2, 0), (70000, 1, 1)])
    r1 = host_api.make_result(200000, 1, 0, 0, alt1=[(70350, 1, 0), (90400, 3, 0)])
    rc, hits, wins = plan(H, r0, r1)
    assert (rc, wins) == (0, []) and hits == [(70000, 0, 1, 1), (70350, 1, 1, 0)]
    # the reverse orientation list is scanned too (mate 1 forward x mate 0 backward)
    r0 = host_api.make_result(5000, 0, 0, 0, alt1=[(30400, 2, 0)])
    r1 = host_api.make_result(200000, 1, 0, 0, alt0=[(30000, 0, 0)])
    assert plan(H, r0, r1)[1] == [(30400, 1, 2, 0), (30000, 0, 0, 0)]


def test_rescue_windows_both_mapped(H):
    # same strand: no proper pair anywhere -> mate 1 around mate 0, then mate 0 around mate 1 (SNP-aware flavour)
    r0 = host_api.make_result(5000, 0, 1, 0); r1 = host_api.make_result(300000, 0, 0, 0)
    rc, hits, wins = plan(H, r0, r1)
    assert rc == 0 and hits is None
    assert wins == [(1, 1, 16, 5000 + 150 + 100, 5000 + 450 + 200), (0, 1, 16, 300000 + 150 + 100, 300000 + 450 + 200)]
    # backward anchors look upstream; near the start of the reference the window is clipped at 0
    r0 = host_api.make_result(5000, 1, 1, 0); r1 = host_api.make_result(300, 1, 0, 0)
    assert plan(H, r0, r1)[2] == [(1, 0, 16, 5000 - 450 - 100, 5000 - 150), (0, 0, 16, 0, 150)]
    # forward anchor at the end of the reference: end clamps to l_pac (not l_pac - 1, alnpe.c:212)
    r0 = host_api.make_result(L_PAC - 400, 0, 1, 0); r1 = host_api.make_result(300, 0, 0, 0)
    assert plan(H, r0, r1)[2][0] == (1, 1, 16, L_PAC - 400 + 250, L_PAC)


def test_singleton_window(H):
    r0 = host_api.make_result(5000, 0, 1, 0); un = host_api.make_result()
    assert plan(H, r0, un) == (0, None, [(1, 1, 5, 5250, 5650)])                           # plain flavour (snpaln_sw)
    assert plan(H, un, r0) == (0, None, [(0, 1, 5, 5250, 5650)])
    end = host_api.make_result(L_PAC - 120, 0, 0, 0)                                       # both ends clamp to l_pac - 1
    assert plan(H, end, un)[2] == [(1, 1, 5, L_PAC - 1, L_PAC - 1)]


def test_apply(H):
    """salt_pair_apply: what both mates look like after pairing (the query_t fields alnpe_sam prints)"""
    # proper pair through alternates: mate 0 takes a gapped alternate (its CIGAR is left to one LV+CIGAR item, kind 4),
    # mate 1 an ungapped one ("100M"); b0/b1/mapq stay what the verification stage left
    r0 = host_api.make_result(5000, 0, 0, 0, alt0=[(70000, 1, 1)]); r0.b0, r0.b1, r0.mapq = 0, 1, 17
    r1 = host_api.make_result(200000, 1, 0, 0, alt1=[(70350, 1, 0)]); r1.b0, r1.b1, r1.mapq = 0, 100000, 37
    pl = host_api.pair_plan(H, r0, 100, r1, 100, A, B, L_PAC, raw=True)
    rc, fin = host_api.pair_apply(H, pl, r0, 100, r1, 100, [], [])
    assert rc == 1
    assert fin[0] == (70000, 0, 1, 1, 0, 99, 0, 1, 17, 4, "") and fin[1] == (70350, 1, 1, 0, 0, 99, 0, 100000, 37, 1, "100M")
    # rescue: the first window finds nothing long enough, the second rescues mate 0 with soft clips on both sides
    r0 = host_api.make_result(5000, 0, 1, 1); r0.cigar = b"40M1I59M"; r0.b0, r0.b1, r0.mapq = 1, 100000, 30
    r1 = host_api.make_result(300000, 0, 2, 0); r1.b0, r1.b1, r1.mapq = 2, 100000, 25
    pl = host_api.pair_plan(H, r0, 100, r1, 100, A, B, L_PAC, raw=True)
    assert pl.n_win == 2
    ssw = [(12, 0, 3, 14, 0, 11), (80, 30, 120, 206, 5, 92)]
    rc, fin = host_api.pair_apply(H, pl, r0, 100, r1, 100, ssw, [[(12, 0)], [(50, 0), (1, 2), (38, 0)]])
    assert rc == 1
    w1 = pl.win[1]
    assert (w1.mate, w1.strand) == (0, 1)
    assert fin[0][:6] == (w1.start + 120, 1, 1, 1, 5, 92) and fin[0][6:8] == (80, 30) and fin[0][9:] == (3, "50M1D38M")
    assert fin[1] == (300000, 0, 2, 0, 0, 99, 2, 100000, 25, 1, "100M")           # the anchor keeps its primary
    # both windows fail: both mates keep their primaries, the gapped one its verification-stage CIGAR
    rc, fin = host_api.pair_apply(H, pl, r0, 100, r1, 100, [(12, 0, 3, 14, 0, 11), (15, 0, 0, 10, 0, 9)], [[(12, 0)], [(10, 0)]])
    assert rc == 0 and fin[0] == (5000, 0, 1, 1, 0, 99, 1, 100000, 30, 2, "40M1I59M") and fin[1][9:] == (1, "100M")


def test_against_reference_pairing(H):
    """salt_pair_plan / salt_pair_apply against the reference's own pairing2 / pairing_singleton (oracle/_ref/
    libsaltref_pair.so: alnpe.c unmodified, its rescue functions replaced by recorders that find nothing) on random
    primaries and alternate lists -- proper pairs through alternates included, which the drop-in's data never produce."""
    import numpy as np
    from oracle import orc
    orc.build()
    if not orc.ref_pair_available():
        pytest.skip("oracle/_ref/libsaltref_pair.so not built (reference tree absent)")
    ref = orc.RefPair()
    rng = np.random.default_rng(77)
    l_pac = 60000
    n_paired = n_alt = n_win = 0
    for it in range(4000):
        l = [int(rng.choice([100, 100, 150, 37])), int(rng.choice([100, 100, 250]))]
        a, b = int(rng.choice([350, 0, 200])), int(rng.choice([650, 300, 1000]))
        centre = int(rng.integers(2000, l_pac - 3000))

        def near():
            return int(np.clip(centre + rng.integers(-1500, 1500), 0, l_pac - 300))
        prim = []
        for m in range(2):
            if rng.random() < 0.12:
                prim.append((0xFFFFFFFF, 3, 255, 255))
            else:
                prim.append((near() if rng.random() < 0.8 else int(rng.integers(0, l_pac - 300)), int(rng.integers(0, 2)), int(rng.integers(0, 6)), 0))
        hits = [[sorted({(near(), int(rng.integers(0, 6)), 0) for _ in range(int(rng.integers(0, 6)))}) if prim[m][0] != 0xFFFFFFFF else []
                 for s in range(2)] for m in range(2)]
        hits = [[[h for h in hs if h[0] != prim[m][0]][:5] for hs in hits[m]] for m in range(2)]
        paired, fq, wins, cigs = ref.pairing(l_pac, a, b, prim, l, hits)
        r = [host_api.make_result(*prim[m], alt0=hits[m][0], alt1=hits[m][1]) for m in range(2)]
        rc, phits, pwins = host_api.pair_plan(H, r[0], l[0], r[1], l[1], a, b, l_pac)
        assert pwins == [tuple(int(x) for x in w) for w in wins], (it, prim, hits, pwins, wins)
        assert (phits is not None) == (paired == 1 and not wins), (it, prim, hits, phits, paired)
        pl = host_api.pair_plan(H, r[0], l[0], r[1], l[1], a, b, l_pac, raw=True)
        nw = len(pwins)
        rc2, fin = host_api.pair_apply(H, pl, r[0], l[0], r[1], l[1], [(0, 0, 0, 0, 0, 0)] * nw, [[]] * nw)      # every rescue fails
        assert rc2 == (1 if phits is not None else 0)
        for m in range(2):
            if fq[m][0] == 0xFFFFFFFF:
                assert fin[m][0] == 0xFFFFFFFF
                continue
            assert fin[m][:6] == fq[m], (it, m, fin[m], fq[m])
            assert fin[m][10] == cigs[m], (it, m, fin[m], cigs[m])
        n_paired += phits is not None; n_win += nw
        n_alt += phits is not None and [h[:2] for h in phits] != [p[:2] for p in prim]
    assert n_paired > 500 and n_alt > 100 and n_win > 1000, (n_paired, n_alt, n_win)
