#!/usr/bin/env python
"""Randomised GPU-vs-oracle stress run (beyond the seeded cases of tests/): random read lengths, candidate densities,
error mixes, thresholds and kernel mappings through the verify stage, the per-pair LV / mismatch / CIGAR entry points,
the SAM tail and the rescue Smith-Waterman.  Stops at the first mismatch (assert).
    python tools/fuzz_gpu.py [seconds] [seed] [ssw]        (third argument "ssw": only the rescue Smith-Waterman, many windows
                                                            per world, random gap penalties and window widths)"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import parity_cases as pc  # noqa: E402
from oracle import orc  # noqa: E402
from salt_b200 import api  # noqa: E402

budget = float(sys.argv[1]) if len(sys.argv) > 1 else 120.0
seed0 = int(sys.argv[2]) if len(sys.argv) > 2 else 1
o = orc.Oracle()
try:
    from salt_b200 import host_api
    hostlib = host_api.load()
except Exception:                                    # noqa: BLE001
    hostlib = None
rng = np.random.default_rng(seed0)
t0 = time.time(); it = 0; tot_reads = 0
ssw_only = len(sys.argv) > 3 and sys.argv[3] == "ssw"
while ssw_only and time.time() - t0 < budget:
    L = int(rng.choice([37, 64, 100, 101, 150, 250]))
    n = int(rng.integers(100, 260))
    g, reads, pos, strand, cands = pc.make_world(int(rng.integers(1 << 30)), glen=int(rng.integers(30000, 120000)), L=L, n_reads=n,
                                                 per_strand=1, snp_rate=float(rng.choice([0.0, 0.01, 0.05])),
                                                 n_rate=float(rng.choice([0.0, 0.002])), sub_rate=float(rng.choice([0.0, 0.01, 0.04])),
                                                 indel_frac=float(rng.choice([0.3, 0.8, 1.0])), n_frac=float(rng.choice([0.0, 0.003])))
    eng = api.Engine(g.mixref, g.l, g.pac, g.l, device=0)
    eng.set_reads(reads)
    wins = pc.make_windows(g, reads, pos, strand, L, rng, int(rng.choice([L + 20, 301, 401, 2 * L + 60])))
    gapO, gapE = [(3, 1), (5, 2), (2, 1), (6, 1), (4, 3)][int(rng.integers(0, 5))]
    pac = bool(it % 2)
    pc.check_ssw(eng, o, g, reads, wins, pac, api.salt_score_mat() if pac else api.salt_score_mat2(), 5 if pac else 16,
                 gapO=gapO, gapE=gapE, flag=int(rng.choice([2, 1, 6])), filters=int(rng.choice([0, 20, 60])),
                 filterd=int(rng.choice([20, 5, 300])), mask_len=int(rng.choice([-1, 15, 40])), cigar_stride=int(rng.choice([96, 64])))
    eng.close()
    if it % 4 == 0:                                  # every branch of the traceback stage on a fresh seed (gaps of 1..14, compensating gaps)
        from salt_b200 import synth
        sd = int(rng.integers(1 << 20))
        g2 = synth.Genome(20011, snp_rate=0.01, n_rate=0.0, seed=sd)
        eng = api.Engine(g2.mixref, g2.l, g2.pac, g2.l, device=0)
        pc.check_ssw_narrow_bands(eng, o, sd, n_reads=160, L=int(rng.choice([100, 150])))
        eng.close()
    it += 1; tot_reads += n
while not ssw_only and time.time() - t0 < budget:
    L = int(rng.choice([37, 50, 64, 75, 100, 101, 125, 150, 151, 200, 250, 300]))
    n = int(rng.integers(40, 220 if L <= 150 else 90))
    per = int(rng.integers(1, 14))
    g, reads, pos, strand, cands = pc.make_world(int(rng.integers(1 << 30)), glen=int(rng.integers(30000, 120000)), L=L, n_reads=n,
                                                 per_strand=per, snp_rate=float(rng.choice([0.0, 0.01, 0.05])),
                                                 n_rate=float(rng.choice([0.0, 0.002])), sub_rate=float(rng.choice([0.0, 0.01, 0.04])),
                                                 indel_frac=float(rng.choice([0.0, 0.3, 0.8])), n_frac=float(rng.choice([0.0, 0.003])))
    eng = api.Engine(g.mixref, g.l, g.pac, g.l, device=0)
    eng.set_reads(reads)
    eng.set_lv_filter(int(rng.integers(0, 2))); eng.set_lv_mapping(int(rng.integers(0, 3)))
    nog = int(rng.choice([0, 1, 3, 5])); lvT = int(rng.choice([-1, -1, 3, 0, 7, int(rng.integers(0, 31))]))
    pc.check_verify(eng, o, g, reads, cands, nog, lvT)
    if it % 4 == 1:                                  # the asynchronous chunk pipeline and the host layer on the same world
        pc.check_verify_batch(eng, reads, cands, int(rng.integers(1, n + 5)), nog, lvT)
        eng.set_reads(reads)
    if it % 8 == 3 and hostlib is not None:
        pc.check_host_chunks(eng, hostlib, o, g, reads, cands, int(rng.integers(5, n + 5)), nog, lvT, max_hits=int(rng.integers(1, 9)), with_tail=True)
        eng.set_reads(reads)
    pairs = pc.flat_pairs(cands, n)
    sel = rng.choice(len(pairs), min(len(pairs), 150), replace=False)
    pc.check_mismatch(eng, o, g, reads, pairs[sel], int(rng.integers(0, 8)))
    pc.check_lv(eng, o, g, reads, pairs[sel], int(rng.choice([-1, 2, 5, 10, 15, 16, 30])))
    tp = api.Engine.make_pairs(np.arange(n, dtype=np.uint32), strand, pos)
    pc.check_lv_cigar(eng, o, g, reads, tp, rng.choice([1, 3, 4, 5, 10, 11, 15, 16, 30], n).astype(np.uint8), int(rng.choice([128, 6, 16])))
    pc.check_md_nm(eng, o, g, reads, pos, strand, int(rng.integers(1 << 30)), md_stride=2 * L + 64)
    if L <= 250 and it % 3 == 0:
        m = min(n, 24)
        wins = pc.make_windows(g, reads[:m], pos[:m], strand[:m], L, rng, int(rng.choice([L + 20, 301, 401])))
        pc.check_ssw(eng, o, g, reads, wins, bool(it % 2), api.salt_score_mat() if it % 2 else api.salt_score_mat2(), 5 if it % 2 else 16,
                     cigar_stride=96)
    eng.close()
    it += 1; tot_reads += n
print("fuzz ok: %d worlds, %d reads, %.0f s, seed %d" % (it, tot_reads, time.time() - t0, seed0))
