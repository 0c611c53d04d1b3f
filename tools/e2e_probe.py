"""Where does the end-to-end chunk pipeline's time go?  (measurement tool, not part of the product)
Runs salt_b200_verify_batch_packed on the bench workload with different chunk sizes and with outputs switched off,
and the bare copy pattern (same buffers, same sizes, no kernels) as the floor of this transfer pattern."""
import ctypes as C
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    import torch
    from salt_b200 import api
    sys.argv = [sys.argv[0]] + sys.argv[1:]
    args = bench.parse()
    wl = bench.make_workload(args, seed=11)
    g = wl["g"]; n = args.reads; L = args.read_len
    dev = torch.device("cuda", 0)
    eng = api.Engine(g.mixref, g.l, None, 0, device=0)
    lib, h = eng.L, eng.h

    def pin(a):
        return torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
    pk_bases, pk_npos = api.pack_bases(wl["reads"].reshape(-1), 2)
    h_bases = pin(pk_bases)
    h_c0 = pin(np.diff(wl["offs0"].astype(np.int64)).astype(np.uint16)); h_c1 = pin(np.diff(wl["offs1"].astype(np.int64)).astype(np.uint16))
    h_l0, h_l1 = pin(wl["loci0"]), pin(wl["loci1"])
    n0, n1 = len(wl["loci0"]), len(wl["loci1"])
    h_rec = torch.empty(n * 16, dtype=torch.uint8).pin_memory()
    h_acc0 = torch.empty(n0, dtype=torch.int8).pin_memory(); h_acc1 = torch.empty(n1, dtype=torch.int8).pin_memory()
    h_cig = torch.zeros(n * 128, dtype=torch.uint8).pin_memory()
    pkc = api.PackedChunkT()
    pkc.n_reads = n; pkc.base_bits = 2; pkc.bases = h_bases.data_ptr(); pkc.base_start = 0; pkc.lens = None; pkc.l_seq = L
    pkc.n_pos = None; pkc.n_n = 0; pkc.count_bits = 16
    pkc.n_cand[0], pkc.n_cand[1] = h_c0.data_ptr(), h_c1.data_ptr()
    pkc.loci[0], pkc.loci[1] = h_l0.data_ptr(), h_l1.data_ptr()
    res = {}

    def run(name, chunk, acc=True, cig=True, reps=5):
        def f():
            rc = lib.salt_b200_verify_batch_packed(h, C.byref(pkc), chunk, 3, -1, h_rec.data_ptr(),
                                                   h_acc0.data_ptr() if acc else None, h_acc1.data_ptr() if acc else None,
                                                   h_cig.data_ptr() if cig else None, 128)
            assert rc == 0, lib.salt_b200_last_error()
        f(); f()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            f()
        torch.cuda.synchronize()
        res[name] = (time.perf_counter() - t0) / reps * 1e3
        print(name, "%.3f ms" % res[name], flush=True)

    for chunk in (25_000, 50_000, 100_000, 200_000, 400_000, 1_000_000, 2_000_000):
        run("chunk_%d" % chunk, chunk)
    run("chunk_100000_noacc", 100_000, acc=False)
    run("chunk_100000_noacc_nocig", 100_000, acc=False, cig=False)
    run("chunk_200000_noacc_nocig", 200_000, acc=False, cig=False)

    # bare copy pattern: per chunk the same five H2D copies and three D2H copies on 4 streams, no kernels
    streams = [torch.cuda.Stream(dev) for _ in range(4)]
    d_bufs = [dict(b=torch.empty(h_bases.numel() // 8 + 64, dtype=torch.uint8, device=dev),
                   c0=torch.empty(400_000, dtype=torch.int16, device=dev), c1=torch.empty(400_000, dtype=torch.int16, device=dev),
                   l0=torch.empty(n0 // 4 + 64, dtype=torch.int32, device=dev), l1=torch.empty(n1 // 4 + 64, dtype=torch.int32, device=dev),
                   rec=torch.empty(400_000 * 16, dtype=torch.uint8, device=dev), a=torch.empty(n0 // 2 + 64, dtype=torch.int8, device=dev))
              for _ in range(4)]
    o0 = wl["offs0"].astype(np.int64); o1 = wl["offs1"].astype(np.int64)
    l0v = h_l0.view(torch.int32); l1v = h_l1.view(torch.int32); c0v = h_c0.view(torch.int16); c1v = h_c1.view(torch.int16)

    def copies(chunk, d2h=True, reps=5):
        def f():
            k = 0
            for b in range(0, n, chunk):
                m = min(chunk, n - b); s = streams[k % 4]; d = d_bufs[k % 4]; k += 1
                s.synchronize()
                with torch.cuda.stream(s):
                    d["b"][:m * L // 4].copy_(h_bases[b * L // 4:(b + m) * L // 4], non_blocking=True)
                    d["c0"][:m].copy_(c0v[b:b + m], non_blocking=True); d["c1"][:m].copy_(c1v[b:b + m], non_blocking=True)
                    a0, e0 = int(o0[b]), int(o0[b + m]); a1, e1 = int(o1[b]), int(o1[b + m])
                    d["l0"][:e0 - a0].copy_(l0v[a0:e0], non_blocking=True); d["l1"][:e1 - a1].copy_(l1v[a1:e1], non_blocking=True)
                    if d2h:
                        h_rec[b * 16:(b + m) * 16].copy_(d["rec"][:m * 16], non_blocking=True)
                        h_acc0[a0:e0].copy_(d["a"][:e0 - a0], non_blocking=True); h_acc1[a1:e1].copy_(d["a"][:e1 - a1], non_blocking=True)
            for s in streams:
                s.synchronize()
        f()
        t0 = time.perf_counter()
        for _ in range(reps):
            f()
        return (time.perf_counter() - t0) / reps * 1e3
    for chunk in (100_000, 400_000):
        res["copies_only_chunk_%d" % chunk] = copies(chunk)
        res["copies_only_h2d_chunk_%d" % chunk] = copies(chunk, d2h=False)
        print(chunk, res["copies_only_chunk_%d" % chunk], res["copies_only_h2d_chunk_%d" % chunk], flush=True)
    print(json.dumps(res))


if __name__ == "__main__":
    main()
