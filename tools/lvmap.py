"""verify-stage LV time with the two LV mappings (thread per pair / warp per pair) on the bench workload shape"""
import sys, types, os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import bench
from salt_b200 import api
n = int(os.environ.get("N_READS", "1000000"))
args = types.SimpleNamespace(reads=n, genome=50_000_000, read_len=100, cands=8, snp_rate=0.01)
wl = bench.make_workload(args, seed=11)
g = wl["g"]; eng = api.Engine(g.mixref, g.l, g.pac, g.l); eng.set_reads(wl["reads"])
c = (wl["offs0"], wl["loci0"], wl["offs1"], wl["loci1"])
for mapping in (0, 1, 0, 1):
    eng.set_lv_mapping(mapping)
    eng.verify(*c, 3, -1)
    eng.profile(True); rec = eng.verify(*c, 3, -1)[0]; t = eng.profile_read(); eng.profile(False)
    print("mapping", mapping, {k: round(v, 4) for k, v in t.items()}, "checksum", int(rec["pos"].astype(np.uint64).sum() % 1000003))
