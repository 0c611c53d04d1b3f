"""LV sweep split: filter on/off x mapping (1 warp per pair, 2 thread per pair) x k on the bench's flat pair list."""
import os, sys, types
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
import bench
from salt_b200 import api
n = int(os.environ.get("N_READS", "500000"))
args = types.SimpleNamespace(reads=n, genome=50_000_000, read_len=100, cands=8, snp_rate=0.01)
wl = bench.make_workload(args, seed=11)
g = wl["g"]; eng = api.Engine(g.mixref, g.l, g.pac, g.l); eng.set_reads(wl["reads"])
lib, h = eng.L, eng.h
dev = torch.device("cuda:0")
n0 = int(wl["offs0"][n]); n1 = int(wl["offs1"][n])
rid0 = np.repeat(np.arange(n, dtype=np.uint32), np.diff(wl["offs0"][:n + 1].astype(np.int64)))
rid1 = np.repeat(np.arange(n, dtype=np.uint32), np.diff(wl["offs1"][:n + 1].astype(np.int64)))
pairs = np.concatenate([api.Engine.make_pairs(rid0, np.zeros(n0, np.uint32), wl["loci0"][:n0]),
                        api.Engine.make_pairs(rid1, np.ones(n1, np.uint32), wl["loci1"][:n1])])
d_pairs = torch.from_numpy(pairs.view(np.uint8)).to(dev)
d_out = torch.empty(len(pairs), dtype=torch.int8, device=dev)
print("pairs", len(pairs))
for k in (2, 3, 5, 8, 10):
    row = {}
    for filt in (1, 0):
        for mapping in (1, 2):
            eng.set_lv_filter(filt); eng.set_lv_mapping(mapping)
            for _ in range(2):
                lib.salt_b200_lv_dev(h, d_pairs.data_ptr(), len(pairs), k, d_out.data_ptr())
            torch.cuda.synchronize()
            import time
            t0 = time.perf_counter()
            for _ in range(5):
                lib.salt_b200_lv_dev(h, d_pairs.data_ptr(), len(pairs), k, d_out.data_ptr())
            lib.salt_b200_sync(h) if hasattr(lib, "salt_b200_sync") else None
            torch.cuda.synchronize()
            row["f%d_m%d" % (filt, mapping)] = round((time.perf_counter() - t0) / 5 * 1e3, 3)
    row["found"] = int((d_out >= 0).sum())
    print("k", k, row)
