#!/usr/bin/env python
"""Generate tests/golden/salt_golden_mdnm_v1.npz: the text the reference's own sam_add_md_nm (sam.c:246-328,
compiled unmodified into oracle/_ref/libsaltref_sam.so behind oracle/dropin/sam_harness.c) appends for seeded
alignments of the worlds already stored in salt_golden_v1.npz -- random M/I/D strings, soft-clip starts, both
strands.  Boxes without /root/reference check the oracle (CPU) and the CUDA kernel (GPU) against it.

Run in the build container only:   python tests/golden/make_golden_mdnm.py"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import parity_cases as pc  # noqa: E402
from oracle import orc  # noqa: E402

OUT = os.path.join(HERE, "salt_golden_mdnm_v1.npz")


def main():
    assert orc.ref_sam_available(), "oracle/_ref/libsaltref_sam.so is not built: run `make -C oracle` where /root/reference exists"
    ref = orc.RefSam()
    G = np.load(os.path.join(HERE, "salt_golden_v1.npz"))
    d = {}
    for tag, L in (("a", 100), ("b", 150), ("c", 250), ("d", 37)):
        reads, pos, strand = G[tag + "_reads"], G[tag + "_pos"], G[tag + "_strand"]
        cases = pc.mdnm_cases(None, reads, pos, strand, 900 + L, clip_frac=0.3 if L > 60 else 0.0)
        n = len(cases)
        cig = np.zeros((n, 64), np.uint8); txt = np.zeros((n, 4 * L + 128), np.uint8)
        for i, (seq, rseq, p, st, s0, c) in enumerate(cases):
            b = c.encode(); cig[i, :len(b)] = np.frombuffer(b, np.uint8)
            t = ref.md_nm(G[tag + "_mixref"], int(G[tag + "_l"]), G[tag + "_pac"], seq, rseq, p, st, s0, c).encode()
            assert len(t) < txt.shape[1]
            txt[i, :len(t)] = np.frombuffer(t, np.uint8)
        d[tag + "_pos"] = np.array([c[2] for c in cases], np.uint32)
        d[tag + "_strand"] = np.array([c[3] for c in cases], np.uint8)
        d[tag + "_seq_start"] = np.array([c[4] for c in cases], np.uint32)
        d[tag + "_cigar"] = cig; d[tag + "_text"] = txt
    np.savez_compressed(OUT, **d)
    print("wrote", OUT, os.path.getsize(OUT), "bytes,", len(d), "arrays")


if __name__ == "__main__":
    main()
