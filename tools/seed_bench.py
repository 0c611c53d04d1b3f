"""Row f1 measurement (opt-in; not part of bench.py's default line): single-end seeding + locate on the device against
the reference's own alnse_seed_overlap + alnse_locate_alt on all host threads, and the whole single-end stage from
reads alone (salt_b200_align_batch_packed: upload reads -> seed -> locate -> verify -> records) against
seeding + verification on the host.  The index is written by the reference's own salt-idx (oracle/_ref), which is
input preparation, not part of either timed side.

    python tools/seed_bench.py --genome 20000000 --reads 1000000 > gpurun_out/seed_bench.json
"""
import argparse
import ctypes as C
import json
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--genome", type=int, default=20_000_000)
    ap.add_argument("--reads", type=int, default=1_000_000)
    ap.add_argument("--read-len", type=int, default=100)
    ap.add_argument("--repeat-frac", type=float, default=0.10)
    ap.add_argument("--cpu-sample", type=int, default=200_000)
    ap.add_argument("--chunk", type=int, default=100_000)
    ap.add_argument("--max-seed", type=int, default=50)
    ap.add_argument("--max-locate", type=int, default=1000)
    args = ap.parse_args()
    print(json.dumps(run(args)))


def run(args, device=0):
    """args: genome, reads, read_len, repeat_frac, cpu_sample, chunk, max_seed, max_locate"""
    import shutil
    import torch
    import seed_cases as sc
    from oracle import orc
    from salt_b200 import api, index_io, synth
    rng = np.random.default_rng(5)
    t0 = time.time()
    g = synth.Genome(args.genome, snp_rate=0.01, seed=31)
    codes_g = g.codes.copy()
    unit_len = 3000
    n_copies = int(args.genome * args.repeat_frac / unit_len)
    units = [rng.integers(0, 4, unit_len).astype(np.uint8) for _ in range(max(1, n_copies // 50))]
    for i in range(n_copies):                                   # repeat families of ~50 copies, 2 % divergence per copy
        u = units[i % len(units)].copy(); m = rng.random(unit_len) < 0.02; u[m] = rng.integers(0, 4, int(m.sum()))
        p = int(rng.integers(0, args.genome - unit_len)); codes_g[p:p + unit_len] = u
    d = tempfile.mkdtemp(prefix="salt_seedbench_")
    prefix = sc.write_index(d, codes_g, np.zeros(args.genome, bool), rng, snp_rate=0.01, records=4)
    t_index = time.time() - t0
    fm = index_io.FmIndex(prefix)
    n, L = args.reads, args.read_len
    pos = rng.integers(0, args.genome - L, n)
    reads = codes_g[pos[:, None] + np.arange(L)[None, :]]
    e = rng.random((n, L)) < 0.01
    reads[e] = (reads[e] + rng.integers(1, 4, int(e.sum()))) & 3
    rc_mask = rng.random(n) < 0.5
    reads[rc_mask] = synth.revcomp(reads[rc_mask])
    reads = np.ascontiguousarray(reads, np.uint8)
    roffs = (np.arange(n + 1, dtype=np.uint64) * L).astype(np.uint32)
    eng = api.Engine(fm.mixref, fm.l, None, 0, device=device)
    eng.set_index(fm)
    opt = api.Engine.seed_opt(fm.l_seed, 0, args.max_seed, args.max_locate)
    out = {"genome": args.genome, "reads": n, "read_len": L, "l_seed": fm.l_seed, "max_seed": args.max_seed,
           "max_locate": args.max_locate, "index_build_s": t_index, "repeat_frac": args.repeat_frac}

    # ---- seeding + locate alone, device-resident reads, one chunk at a time through slot 0
    per_chunk = []
    tot = 0
    for b in range(0, n, args.chunk):
        m = min(args.chunk, n - b)
        eng.set_reads(reads[b:b + m])
        eng.seed_locate(opt, download=False)
        torch.cuda.synchronize()
        t = time.perf_counter()
        n0, n1 = eng.seed_locate(opt, download=False)
        per_chunk.append(time.perf_counter() - t)
        tot += n0 + n1
    out["gpu_seed_locate"] = {"reads_per_s": n / sum(per_chunk), "ms_per_chunk": 1e3 * float(np.mean(per_chunk)),
                              "chunk_reads": args.chunk, "candidates": int(tot), "candidates_per_read": tot / n,
                              "how": "wall time of salt_b200_seed_locate per chunk (reads resident, lists left on the device, totals read back)"}

    # ---- whole single-end stage from host buffers: reads in (2 bit/base), records out
    pk_bases, pk_npos = api.pack_bases(reads.reshape(-1), 2)
    h_bases = torch.from_numpy(pk_bases).pin_memory()
    h_rec = torch.empty(n * 16, dtype=torch.uint8).pin_memory(); h_cig = torch.zeros(n * 128, dtype=torch.uint8).pin_memory()
    pkc = api.PackedChunkT()
    pkc.n_reads = n; pkc.base_bits = 2; pkc.bases = h_bases.data_ptr(); pkc.base_start = 0; pkc.lens = None; pkc.l_seq = L
    pkc.n_pos = None; pkc.n_n = 0; pkc.count_bits = 16

    def align():
        rc = eng.L.salt_b200_align_batch_packed(eng.h, C.byref(pkc), C.byref(opt), args.chunk, 3, -1, h_rec.data_ptr(), h_cig.data_ptr(), 128)
        assert rc == 0, eng.L.salt_b200_last_error()
    align(); align()
    torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(3):
        align()
    torch.cuda.synchronize()
    sec = (time.perf_counter() - t) / 3
    rec = np.frombuffer(h_rec.numpy().tobytes(), api.VERIFY_DT)
    out["gpu_align_e2e"] = {"reads_per_s": n / sec, "ms": sec * 1e3, "h2d_bytes": int(h_bases.numel()), "d2h_bytes": int(n * 16),
                            "mapped_frac": float((rec["pos"] != 0xFFFFFFFF).mean()),
                            "how": "salt_b200_align_batch_packed from pinned host buffers: upload reads, seed, locate, verify, download records"}

    # ---- the reference on the host cores (bounded sample)
    ref = orc.SeedRef(prefix)
    cores = os.cpu_count() or 1
    ns = min(args.cpu_sample, n)
    o0, l0, o1, l1, sec_seed = ref.run_mt(reads[:ns], roffs[:ns + 1], fm.l_seed, 0, args.max_seed, args.max_locate, cores)
    out["cpu_seed_locate"] = {"reads_per_s": ns / sec_seed, "cores": cores, "kind": "reference", "sample": "%d reads" % ns}
    o = orc.Oracle(); r_ = orc.Ref() if orc.ref_available() else None
    sec_v, recs, a0, a1, cig = o.verify_batch(fm.mixref, fm.l, reads[:ns].reshape(-1), roffs[:ns + 1], o0, l0, o1, l1, 3, -1, n_threads=cores, ref=r_)
    out["cpu_seed_verify"] = {"reads_per_s": ns / (sec_seed + sec_v), "seed_s": sec_seed, "verify_s": sec_v, "cores": cores}
    # parity on the sample while we are here
    eng.set_reads(reads[:ns])
    got = eng.seed_locate(opt)
    out["lists_identical_on_sample"] = bool(all(np.array_equal(a, b) for a, b in zip(got, (o0, l0, o1, l1))))
    out["speedup_seed_locate"] = out["gpu_seed_locate"]["reads_per_s"] / out["cpu_seed_locate"]["reads_per_s"]
    out["speedup_align_e2e"] = out["gpu_align_e2e"]["reads_per_s"] / out["cpu_seed_verify"]["reads_per_s"]
    ref.close(); eng.close()
    shutil.rmtree(d, ignore_errors=True)
    return out


if __name__ == "__main__":
    main()
