#!/usr/bin/env python
"""Row f4 (input side), host only: salt_fastq_pack -- FASTQ text to the compact transport (2-bit bases, N list, lengths, offsets
of name / comment / quality) -- against the reference's own reader (query_open / query_read_seq + query_destroy, query.c:66-239,
through oracle/_ref/libsaltref_seed.so), one thread each, on the same file held in the page cache.  Prints one JSON object.
    python tools/fastq_bench.py [reads] [read_len]"""
import ctypes as C
import json
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from salt_b200 import host_api          # noqa: E402
from test_fastq_pack import FastqT       # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
    L = int(sys.argv[2]) if len(sys.argv) > 2 else 100
    rng = np.random.default_rng(1)
    d = tempfile.mkdtemp(prefix="salt_fastq_")
    path = os.path.join(d, "reads.fq")
    bases = np.frombuffer(b"ACGT", np.uint8)[rng.integers(0, 4, (n, L))]
    bases[rng.random((n, L)) < 0.001] = ord("N")
    qual = (33 + rng.integers(2, 41, (n, L))).astype(np.uint8)
    with open(path, "wb") as f:
        for i in range(n):
            f.write(b"@read%d/1\n" % i); f.write(bases[i].tobytes()); f.write(b"\n+\n"); f.write(qual[i].tobytes()); f.write(b"\n")
    text = open(path, "rb").read()
    try:
        H = host_api.load()
    except Exception:                                        # no CUDA build here: the parser is plain C, the emulator build has it too
        sys.path.insert(0, os.path.join(ROOT, "tests", "emul"))
        import build_emul
        H = host_api.load(build_emul.build_host())
    H.salt_fastq_pack.argtypes = [C.c_char_p, C.c_size_t, C.c_int, C.c_uint32, C.POINTER(FastqT), C.POINTER(C.c_size_t)]
    cap = n * L + 64
    bases_o = np.zeros(cap // 4 + 2, np.uint8); n_pos = np.zeros(max(1024, n * L // 100), np.uint32)
    arrs = {k: np.zeros(n, dt) for k, dt in (("lens", np.uint16), ("n_ambiguous", np.uint16), ("name_off", np.uint32), ("name_len", np.uint16),
                                              ("comment_off", np.uint32), ("comment_len", np.uint16), ("qual_off", np.uint32))}
    best = None
    for _ in range(3):
        fq = FastqT(bases_o.ctypes.data, cap, n_pos.ctypes.data, len(n_pos), *(arrs[k].ctypes.data for k in
                    ("lens", "n_ambiguous", "name_off", "name_len", "comment_off", "comment_len", "qual_off")), 0, 0, 0)
        used = C.c_size_t(0)
        t0 = time.perf_counter()
        got = H.salt_fastq_pack(text, len(text), 1, n, C.byref(fq), C.byref(used))
        dt = time.perf_counter() - t0
        assert got == n, got
        best = dt if best is None else min(best, dt)
    out = {"reads": n, "read_len": L, "fastq_bytes": len(text),
           "salt_fastq_pack": {"reads_per_s": n / best, "mb_per_s": len(text) / best / 1e6, "threads": 1, "n_positions": int(fq.n_n),
                               "out_bytes_per_read": (L + 3) // 4 + 2 + 2 + 4 + 2 + 4 + 2 + 4}}
    # several host threads: salt_fastq_split cuts the text at record headers, every part is parsed on its own thread into its own
    # arrays (each part travels as chunks of its own)
    import threading
    H.salt_fastq_split.argtypes = [C.c_char_p, C.c_size_t, C.c_int, C.POINTER(C.c_size_t)]
    H.salt_fastq_split.restype = C.c_int
    out["threads"] = []
    for T in [t for t in (2, 4, 8, 16, 32) if t <= (os.cpu_count() or 1)]:
        cuts = (C.c_size_t * (T + 1))()
        t0 = time.perf_counter()
        made = H.salt_fastq_split(text, len(text), T, cuts)
        assert made >= 1, made
        parts = []
        for k in range(made):
            ln = cuts[k + 1] - cuts[k]
            mr = n // made + n // (4 * made) + 1024
            b = np.zeros(ln // 8 + 16, np.uint8); npos = np.zeros(max(1024, ln // 100), np.uint32)
            ar = {key: np.zeros(mr, dt) for key, dt in (("lens", np.uint16), ("n_ambiguous", np.uint16), ("name_off", np.uint32),
                                                        ("name_len", np.uint16), ("comment_off", np.uint32), ("comment_len", np.uint16),
                                                        ("qual_off", np.uint32))}
            parts.append((b, npos, ar, mr))
        t_alloc = time.perf_counter() - t0
        counts = [0] * made
        base_addr = C.cast(C.c_char_p(text), C.c_void_p).value

        def work(k):
            b, npos, ar, mr = parts[k]
            fq = FastqT(b.ctypes.data, len(b) * 4 - 8, npos.ctypes.data, len(npos), *(ar[key].ctypes.data for key in
                        ("lens", "n_ambiguous", "name_off", "name_len", "comment_off", "comment_len", "qual_off")), 0, 0, 0)
            used = C.c_size_t(0)
            counts[k] = H.salt_fastq_pack(C.cast(base_addr + cuts[k], C.c_char_p), cuts[k + 1] - cuts[k], 1, mr, C.byref(fq), C.byref(used))
        bestT = None
        for _ in range(6):
            ths = [threading.Thread(target=work, args=(k,)) for k in range(made)]
            t0 = time.perf_counter()
            for th in ths: th.start()
            for th in ths: th.join()
            dt = time.perf_counter() - t0
            bestT = dt if bestT is None else min(bestT, dt)
        assert sum(counts) == n, (counts, n)
        out["threads"].append({"threads": made, "reads_per_s": n / bestT, "mb_per_s": len(text) / bestT / 1e6})
    ref = os.path.join(ROOT, "oracle", "_ref", "libsaltref_seed.so")
    if os.path.exists(ref):
        R = C.CDLL(ref)
        R.seedref_read_fastq.restype = C.c_int
        R.seedref_read_fastq.argtypes = [C.c_char_p, C.c_uint32, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                         C.c_void_p, C.c_int]
        codes = np.zeros(n * L + 64, np.uint8); roffs = np.zeros(n + 1, np.uint32); amb = np.zeros(n, np.uint16)
        stride = 128 if L < 120 else 320
        names = np.zeros(n * stride, np.uint8); coms = np.zeros(n * stride, np.uint8); quals = np.zeros(n * stride, np.uint8)
        bestr = None
        for _ in range(2):
            t0 = time.perf_counter()
            got = R.seedref_read_fastq(path.encode(), n, codes.ctypes.data, len(codes), roffs.ctypes.data, amb.ctypes.data,
                                       names.ctypes.data, coms.ctypes.data, quals.ctypes.data, stride)
            dt = time.perf_counter() - t0
            assert got == n, got
            bestr = dt if bestr is None else min(bestr, dt)
        out["reference_reader"] = {"reads_per_s": n / bestr, "mb_per_s": len(text) / bestr / 1e6, "threads": 1,
                                   "what": "query_open + query_read_seq + query_destroy per record, plus the harness's copies of seq / name / qual"}
        # same bases?
        nb = n * L
        mine = ((bases_o[np.arange(nb) >> 2] >> (2 * (np.arange(nb) & 3)).astype(np.uint8)) & 3).astype(np.uint8)
        mine[n_pos[:fq.n_n]] = 4
        out["codes_identical"] = bool(np.array_equal(mine, codes[:nb]))
        out["speedup"] = out["salt_fastq_pack"]["reads_per_s"] / out["reference_reader"]["reads_per_s"]
    print(json.dumps(out, indent=1))
    os.remove(path); os.rmdir(d)


if __name__ == "__main__":
    main()
