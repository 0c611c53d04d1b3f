#!/usr/bin/env python
"""Kernel-level sweep of BASELINE.json configs[3] / [4] on one GPU: (read, candidate) pair lists of
1e6..3e7 pairs, L = 100/150/250, LV k = 2..8 and the reference's own L/10, the ungapped stage and
the mate-rescue SSW, with the 250 bp / 5 % error / indel-rich reads of configs[4] as the last block.
Prints one JSON object; `python tools/sweep.py > profiles/<name>.json` on the GPU box.
Times are CUDA events on the engine's stream, best of 3 after 2 warm-ups, inputs resident in HBM."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from salt_b200 import api, synth  # noqa: E402


def timed(fn, stream, dev, reps=3, warm=2):
    for _ in range(warm):
        fn()
    best = 1e30
    for _ in range(reps):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(dev); e0.record(stream); fn(); e1.record(stream); torch.cuda.synchronize(dev)
        best = min(best, e0.elapsed_time(e1))
    return best


def main():
    dev = torch.device("cuda", 0); torch.cuda.set_device(0)
    stream = torch.cuda.Stream(dev); torch.cuda.set_stream(stream)
    glen = int(os.environ.get("SWEEP_GENOME", "50000000"))
    g = synth.Genome(glen, snp_rate=0.01, seed=3)
    out = {"genome_bp": glen, "snp_rate": 0.01, "blocks": []}
    cases = [(100, 0.01, 0.02, 1_000_000, "configs[3] 100 bp"), (150, 0.01, 0.02, 700_000, "configs[3] 150 bp"),
             (250, 0.01, 0.02, 400_000, "configs[3] 250 bp"), (250, 0.04, 0.6, 400_000, "configs[4] 250 bp, 5% error, indel-rich")]
    for L, sub, indel, n_reads, tag in cases:
        t0 = time.time()
        reads, pos, strand = synth.sample_reads(g, n_reads, L, seed=10 + L, sub_rate=sub, indel_frac=indel, max_indel=6 if indel > 0.5 else 3)
        offs0, loci0, offs1, loci1 = synth.make_candidates(g, pos, strand, L, per_strand=8, seed=20 + L)
        eng = api.Engine(g.mixref, g.l, g.pac, g.l, device=0)
        eng.set_stream(stream.cuda_stream)
        eng.set_reads(reads)
        lib, h = eng.L, eng.h
        n0, n1 = len(loci0), len(loci1)
        rid = np.concatenate([np.repeat(np.arange(n_reads, dtype=np.uint32), np.diff(offs0.astype(np.int64))),
                              np.repeat(np.arange(n_reads, dtype=np.uint32), np.diff(offs1.astype(np.int64)))])
        st = np.concatenate([np.zeros(n0, np.uint32), np.ones(n1, np.uint32)])
        pairs = api.Engine.make_pairs(rid, st, np.concatenate([loci0, loci1]))
        d_pairs = torch.from_numpy(pairs.view(np.uint8)).to(dev)
        d_out = torch.empty(len(pairs), dtype=torch.int8, device=dev)
        blk = {"case": tag, "read_len": L, "reads": n_reads, "pairs": len(pairs), "gen_s": round(time.time() - t0, 1), "lv": {}, "mismatch": {}}
        ms = timed(lambda: lib.salt_b200_mismatch_dev(h, d_pairs.data_ptr(), len(pairs), 3, d_out.data_ptr()), stream, dev)
        blk["mismatch"] = {"ms": ms, "pairs_per_s": len(pairs) / ms * 1e3, "GBps_algorithmic": len(pairs) * ((L + 1) // 2 + 10) / ms / 1e6}
        for k in (2, 3, 4, 5, 6, 7, 8, -1):
            for filt in (1, 0):
                lib.salt_b200_set_lv_filter(h, filt)
                ms = timed(lambda: lib.salt_b200_lv_dev(h, d_pairs.data_ptr(), len(pairs), k, d_out.data_ptr()), stream, dev)
                key = ("k%d" % k if k >= 0 else "kL/10") + ("" if filt else "_nofilter")
                blk["lv"][key] = {"ms": ms, "pairs_per_s": len(pairs) / ms * 1e3, "tcups_equiv": len(pairs) * L * (L + 4) / ms / 1e9}
            lib.salt_b200_set_lv_filter(h, 1)
        found = int((d_out >= 0).sum().item())
        blk["lv"]["found_at_L/10"] = found
        # whole ungapped + gapped stage on the same lists
        d = [torch.from_numpy(np.ascontiguousarray(a)).to(dev) for a in (offs0, loci0, offs1, loci1)]
        d_rec = torch.empty(n_reads * 16, dtype=torch.uint8, device=dev); d_acc = torch.empty(n0 + n1 + 16, dtype=torch.int8, device=dev)
        d_cig = torch.empty(n_reads * 128, dtype=torch.uint8, device=dev)
        d_cr = torch.empty(n_reads + 2, dtype=torch.int32, device=dev); d_cc = torch.zeros(4, dtype=torch.int32, device=dev)
        for rule, lvT in (("se", -1), ("pe", 3)):
            ms = timed(lambda: lib.salt_b200_verify_dev(h, d[0].data_ptr(), d[1].data_ptr(), n0, d[2].data_ptr(), d[3].data_ptr(), n1, 3, lvT,
                                                       d_rec.data_ptr(), d_acc.data_ptr(), d_acc.data_ptr() + n0, d_cig.data_ptr(), 128,
                                                       d_cr.data_ptr(), d_cc.data_ptr()), stream, dev)
            rec = np.frombuffer(d_rec.cpu().numpy().tobytes(), api.VERIFY_DT)
            blk["verify_" + rule] = {"ms": ms, "reads_per_s": n_reads / ms * 1e3, "pairs_per_s": (n0 + n1) / ms * 1e3,
                                     "mapped_frac": float((rec["pos"] != 0xFFFFFFFF).mean()), "gapped_stage_frac": float(rec["lv_ran"].mean())}
        # mate rescue: the window the default insert bounds give (alnpe.c:213-252), and for 250 bp also the 551-wide window of
        # wide insert bounds (-a 250 -b 1050)
        for key, W in [("ssw", {100: 401, 150: 401, 250: 301}[L])] + ([("ssw_551", 551)] if L == 250 else []):
            nt = min(n_reads, 200_000)
            rng = np.random.default_rng(5)
            start = np.maximum(0, pos[:nt].astype(np.int64) - rng.integers(0, W - L, nt))
            wins = np.zeros(nt, api.WIN_DT); wins["rs"] = (np.arange(nt, dtype=np.uint32) << 1) | strand[:nt]
            wins["start"] = start; wins["end"] = np.minimum(g.l - 1, start + W - 1)
            d_w = torch.from_numpy(wins.view(np.uint8)).to(dev); d_so = torch.empty(nt * 28, dtype=torch.uint8, device=dev)
            d_sc = torch.empty(nt * 64, dtype=torch.int32, device=dev)
            mat = api.salt_score_mat2()
            lib.salt_b200_set_max_window(h, (W + 7) // 8 * 8)
            ms = timed(lambda: lib.salt_b200_ssw_dev(h, d_w.data_ptr(), nt, 0, mat.ctypes.data, 16, 3, 1, 2, 0, 20, -1, d_so.data_ptr(),
                                                     d_sc.data_ptr(), 64), stream, dev)
            eng.profile(True)
            lib.salt_b200_ssw_dev(h, d_w.data_ptr(), nt, 0, mat.ctypes.data, 16, 3, 1, 2, 0, 20, -1, d_so.data_ptr(), d_sc.data_ptr(), 64)
            stg = {k_: v for k_, v in eng.profile_read().items() if k_.startswith("ssw")}
            eng.profile(False)
            cells = nt * (8 * ((L + 7) // 8)) * W
            blk[key] = {"tasks": nt, "window": W, "ms": ms, "tasks_per_s": nt / ms * 1e3, "tcups_fwd_cells_pipeline": cells / ms / 1e9,
                        "tcups_fwd_kernel": cells / stg["ssw_dp_fwd"] / 1e9, "stages_ms": stg}
        out["blocks"].append(blk)
        eng.close()
        print("done", tag, file=sys.stderr)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
