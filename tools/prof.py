#!/usr/bin/env python
"""Small driver for ncu captures: one warm-up and one measured pass of the verify stage and of the
mate-rescue SSW pipeline on the bench workload shape (fewer reads so the capture stays short).
    ncu --set full -k regex:<kernel> ... python tools/prof.py"""
import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from salt_b200 import api  # noqa: E402


def main():
    n = int(os.environ.get("PROF_READS", "500000"))
    args = types.SimpleNamespace(reads=n, genome=50_000_000, read_len=int(os.environ.get("PROF_L", "100")), cands=8, snp_rate=0.01)
    wl = bench.make_workload(args, seed=11)
    g = wl["g"]
    eng = api.Engine(g.mixref, g.l, g.pac, g.l)
    eng.set_reads(wl["reads"])
    for it in range(2):
        rec, a0, a1, cig = eng.verify(wl["offs0"], wl["loci0"], wl["offs1"], wl["loci1"], 3, -1)
    print("verify: mapped %d lv_ran %d gapped %d" % ((rec["pos"] != 0xFFFFFFFF).sum(), rec["lv_ran"].sum(), (rec["is_gap"] == 1).sum()))
    nt = min(n, int(os.environ.get("PROF_SW_TASKS", "100000"))); W = 401; L = args.read_len
    rng = np.random.default_rng(5)
    start = np.maximum(0, wl["pos"][:nt].astype(np.int64) - rng.integers(0, W - L, nt))
    wins = np.zeros(nt, api.WIN_DT)
    wins["rs"] = (np.arange(nt, dtype=np.uint32) << 1) | wl["strand"][:nt]
    wins["start"] = start; wins["end"] = np.minimum(g.l - 1, start + W - 1)
    for it in range(2):
        out, cg = eng.ssw(wins, api.salt_score_mat2(), 16, False, cigar_stride=32)
    print("ssw: mean score %.1f" % out["score1"].mean())
    pairs = api.Engine.make_pairs(np.repeat(np.arange(50000, dtype=np.uint32), 8), np.zeros(400000, np.uint32),
                                  wl["loci0"][:400000])
    for k in (3, 10):
        e = eng.lv(pairs, k)
    print("lv: found %d" % (e >= 0).sum())


if __name__ == "__main__":
    main()
