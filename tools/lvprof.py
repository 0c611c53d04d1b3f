"""ncu target: the flat LV list of bench.py (k = 10 and k = 3), filter + automatic mapping, two passes each."""
import os, sys, types
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import bench
from salt_b200 import api
n = int(os.environ.get("N_READS", "500000"))
args = types.SimpleNamespace(reads=n, genome=50_000_000, read_len=100, cands=8, snp_rate=0.01)
wl = bench.make_workload(args, seed=11)
g = wl["g"]; eng = api.Engine(g.mixref, g.l, g.pac, g.l); eng.set_reads(wl["reads"])
n0 = int(wl["offs0"][n]); n1 = int(wl["offs1"][n])
rid0 = np.repeat(np.arange(n, dtype=np.uint32), np.diff(wl["offs0"][:n + 1].astype(np.int64)))
rid1 = np.repeat(np.arange(n, dtype=np.uint32), np.diff(wl["offs1"][:n + 1].astype(np.int64)))
pairs = np.concatenate([api.Engine.make_pairs(rid0, np.zeros(n0, np.uint32), wl["loci0"][:n0]),
                        api.Engine.make_pairs(rid1, np.ones(n1, np.uint32), wl["loci1"][:n1])])
for k in (10, 3):
    for _ in range(2):
        e = eng.lv(pairs, k)
    print("k", k, "found", int((e >= 0).sum()))
