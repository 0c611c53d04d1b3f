"""Several GPUs in ONE process (salt is one process, alnse.c:1414-1440): salt_multi_verify_batch_packed from pinned host
buffers over 1, 2, 4, ... visible devices, weak scaling (every device gets the bench's 2 M-read batch).  The ordered merge is
inside the timed region by construction: every share writes its results at its reads' positions.
    python tools/multi_bench.py > gpurun_out/multi_bench.json        (on a box with several GPUs: gpurun --gpus N)"""
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
    from salt_b200 import api, host_api
    sys.argv = [sys.argv[0]] + sys.argv[1:]
    args = bench.parse()
    H = host_api.load()
    wl = bench.make_workload(args, seed=11)
    g = wl["g"]; n1 = args.reads; L = args.read_len
    ndev = torch.cuda.device_count()
    out = {"devices_visible": ndev, "reads_per_device": n1, "chunk_reads": args.chunk, "rows": []}
    base_bases, _ = api.pack_bases(wl["reads"].reshape(-1), 2)
    c0 = np.diff(wl["offs0"].astype(np.int64)).astype(np.uint16); c1 = np.diff(wl["offs1"].astype(np.int64)).astype(np.uint16)
    want = None
    for G in [x for x in (1, 2, 4, 8) if x <= ndev]:
        n = n1 * G
        def pin(a):
            return torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
        nb = n1 * L // 4
        h_bases = pin(np.concatenate([base_bases[:nb]] * G + [np.zeros(8, np.uint8)]))
        h_c0 = pin(np.tile(c0, G)); h_c1 = pin(np.tile(c1, G))
        h_l0 = pin(np.tile(wl["loci0"], G)); h_l1 = pin(np.tile(wl["loci1"], G))
        n0, n1c = len(wl["loci0"]) * G, len(wl["loci1"]) * G
        h_rec = torch.empty(n * 16, dtype=torch.uint8).pin_memory()
        h_a0 = torch.empty(n0, dtype=torch.int8).pin_memory(); h_a1 = torch.empty(n1c, dtype=torch.int8).pin_memory()
        h_cig = torch.zeros(n * 128, dtype=torch.uint8).pin_memory()
        pk = api.PackedChunkT()
        pk.n_reads = n; pk.base_bits = 2; pk.bases = h_bases.data_ptr(); pk.base_start = 0; pk.lens = None; pk.l_seq = L
        pk.n_pos = None; pk.n_n = 0; pk.count_bits = 16
        pk.n_cand[0], pk.n_cand[1] = h_c0.data_ptr(), h_c1.data_ptr(); pk.loci[0], pk.loci[1] = h_l0.data_ptr(), h_l1.data_ptr()
        devs = (C.c_int * G)(*range(G))
        m = H.salt_multi_init(g.mixref.ctypes.data, g.l, g.pac.ctypes.data, g.l, devs, G)
        assert m

        for chunk in (args.chunk, 4 * args.chunk):
            def run():
                rc = H.salt_multi_verify_batch_packed(m, C.byref(pk), chunk, 3, -1, h_rec.data_ptr(), h_a0.data_ptr(), h_a1.data_ptr(),
                                                      h_cig.data_ptr(), 128)
                assert rc == 0
            run(); run()
            t0 = time.perf_counter()
            reps = 5
            for _ in range(reps):
                run()
            sec = (time.perf_counter() - t0) / reps
            rec = h_rec.numpy().reshape(G, -1)
            if want is None:
                want = rec[0].tobytes()
            same = all(rec[k].tobytes() == want for k in range(G))      # every device's share equals the single-device result
            out["rows"].append({"devices": G, "chunk_reads": chunk, "reads": n, "ms": sec * 1e3, "reads_per_s": n / sec,
                                "identical_to_single_device": bool(same)})
        H.salt_multi_destroy(m)
    for r in out["rows"]:
        r1 = next(x["reads_per_s"] for x in out["rows"] if x["devices"] == 1 and x["chunk_reads"] == r["chunk_reads"])
        r["efficiency_vs_1"] = r["reads_per_s"] / (r1 * r["devices"])
    print(json.dumps(out))


if __name__ == "__main__":
    main()
