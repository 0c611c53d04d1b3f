import sys, types, numpy as np
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import bench, parity_cases as pc
from salt_b200 import api
from oracle import orc
args = types.SimpleNamespace(reads=200000, genome=50_000_000, read_len=100, cands=8, snp_rate=0.01)
wl = bench.make_workload(args, seed=11)
g = wl["g"]; eng = api.Engine(g.mixref, g.l, g.pac, g.l); eng.set_reads(wl["reads"])
nt = 200000; L = 100; W = 401
rng = np.random.default_rng(5)
start = np.maximum(0, wl["pos"][:nt].astype(np.int64) - rng.integers(0, W - L, nt))
wins = np.zeros(nt, api.WIN_DT); wins["rs"] = (np.arange(nt, dtype=np.uint32) << 1) | wl["strand"][:nt]
wins["start"] = start; wins["end"] = np.minimum(g.l - 1, start + W - 1)
eng.profile(True)
out, cg = eng.ssw(wins, api.salt_score_mat2(), 16, False, cigar_stride=32)
print(eng.profile_read())
print("mean score", out["score1"].mean(), "cigarLen>0", (out["cigarLen"] > 0).mean(), "mean ref_begin", out["ref_begin1"].mean(), "read_begin>=0", (out["read_begin1"] >= 0).mean())
o = orc.Oracle()
pc.check_ssw(eng, o, g, wl["reads"], wins[rng.choice(nt, 1500, replace=False)], False, api.salt_score_mat2(), 16, cigar_stride=64)
print("sample ok")
