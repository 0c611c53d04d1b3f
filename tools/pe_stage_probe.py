#!/usr/bin/env python
"""Where the paired-end seeding + locate + verification stage of salt_aln spends its time (143 ms per batch of
100 k mates for lists of 8 loci, against 12 ms for a single-end batch at -r 1).  Times, per call and after a warm-up call, on one
batch of mates: salt_b200_set_reads, salt_b200_seed_locate in the paired-end flavour for several list rooms (the scratch is
n_mates x 2 x room x 4 bytes) with and without the download of the lists (the binding locates twice for that: once for the
totals, once with buffers), salt_b200_seed_status, and the verification of the
lists through the chunk queue (3 / 3).  Prints one JSON object.  Not run yet: written when the round's GPU minutes were spent.

    GENOME=5000000 PAIRS=50000 python tools/pe_stage_probe.py > gpurun_out/pe_stage_probe.json
"""
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import dropin_data  # noqa: E402
from salt_b200 import api, host_api, index_io, synth  # noqa: E402

REFDIR = os.path.join(ROOT, "oracle", "_ref")


def timed(fn, reps=3):
    fn()                                    # warm-up: allocations, module loading
    t = []
    for _ in range(reps):
        t0 = time.perf_counter(); fn(); t.append((time.perf_counter() - t0) * 1e3)
    return round(min(t), 3)


def main(lib=None, hostlib=None):
    glen = int(os.environ.get("GENOME", "5000000")); npairs = int(os.environ.get("PAIRS", "50000"))
    res = {"genome_bp": glen, "mates": 2 * npairs}
    with tempfile.TemporaryDirectory() as d:
        dropin_data.write_inputs(d, glen=glen, n_reads=10)
        g = synth.Genome(glen, snp_rate=0.01, n_rate=0.0, seed=5)
        reads = synth.sample_pairs(g, npairs, 100, seed=9, insert_mean=500, insert_sd=40, hard_frac=0.1, junk_frac=0.01)[0]
        subprocess.run([os.path.join(REFDIR, "salt-idx"), "-k", "19", "ref.fa", "snps.txt", "idx"], cwd=d, stdout=open(os.path.join(d, "idx.log"), "w"),
                       stderr=subprocess.PIPE, check=True)
        fm = index_io.FmIndex(os.path.join(d, "idx"))
        pac = np.ascontiguousarray(np.fromfile(os.path.join(d, "idx.C.pac"), np.uint8)[:(fm.l + 3) // 4])
        eng = api.Engine(fm.mixref, fm.l, pac, fm.l, **({"lib": lib} if lib is not None else {"device": 0}))
        eng.set_index(fm)
        H = hostlib if hostlib is not None else host_api.load()
        n, L = reads.shape
        roffs = (np.arange(n + 1) * L).astype(np.uint32)
        codes = np.ascontiguousarray(reads.reshape(-1))
        res["set_reads_ms"] = timed(lambda: eng.set_reads(codes, roffs))
        res["seed_locate"] = []
        lists = None
        for mode, over, cap in ((0, 1, 0), (0, 5, 0), (1, 5, 256), (1, 5, 1024), (1, 5, 4096), (1, 5, 16384)):
            opt = api.Engine.seed_opt(fm.l_seed, over, 50, 1000 if mode else 500, 0, locate_mode=mode, list_cap=cap)
            row = {"locate_mode": mode, "l_overlap": over, "list_cap": cap,
                   "scratch_MB": round(n * 2 * (cap if mode else opt.max_locate) * 4 / 1e6),
                   "totals_only_ms": timed(lambda: eng.seed_locate(opt, download=False)),
                   "with_download_ms": timed(lambda: eng.seed_locate(opt))}
            got = eng.seed_locate(opt)
            row["loci_per_list"] = round((len(got[1]) + len(got[3])) / (2 * n), 2)
            if mode == 1 and cap == 1024:
                lists = got
                row["seed_status_ms"] = timed(lambda: eng.seed_status())
            res["seed_locate"].append(row)
        offs0, loci0, offs1, loci1 = lists
        ch = host_api.Chunk(H, n + 8, len(codes) + 1024, max(len(loci0), len(loci1)) + 64)

        def verify():
            ch.reset(); ch.add_reads(codes, roffs, offs0, loci0, offs1, loci1); ch.submit(eng, 0, 3, 3); ch.wait(eng, 0)
        res["chunk_add_submit_wait_ms"] = timed(verify)
        res["chunk_pair_ms"] = timed(lambda: ch.pair(eng, 0, npairs, 350, 650, fm.l, with_tail=False))
        ch.close(); eng.close()
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
