"""Upper bound of sorting LV worklists by difficulty: thread-per-pair LV (k = 10, filter off) on the pairs that are
within k, in list order against sorted by their result e; same for the CIGAR kernel with k = e."""
import os, sys, types, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
import bench
from salt_b200 import api
n = 500000
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
e = eng.lv(pairs, 10)
keep = e >= 1                                  # what the verify stage sends on: not exact, within k
surv = pairs[keep]; es = e[keep]
print("survivors", len(surv), "mean e %.2f" % es.mean(), "hist", np.bincount(es, minlength=11).tolist())
eng.set_lv_filter(0); eng.set_lv_mapping(2)
def run(pp, label):
    d_pairs = torch.from_numpy(pp.view(np.uint8)).to(dev); d_out = torch.empty(len(pp), dtype=torch.int8, device=dev)
    for _ in range(3):
        lib.salt_b200_lv_dev(h, d_pairs.data_ptr(), len(pp), 10, d_out.data_ptr())
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(10):
        lib.salt_b200_lv_dev(h, d_pairs.data_ptr(), len(pp), 10, d_out.data_ptr())
    torch.cuda.synchronize()
    print("%-28s %.3f ms" % (label, (time.perf_counter() - t0) * 100))
run(surv, "lv_tpp list order")
run(surv[np.argsort(es, kind="stable")], "lv_tpp sorted by e")
coarse = np.where(es <= 3, 0, 1)
run(surv[np.argsort(coarse, kind="stable")], "lv_tpp two buckets (e<=3)")
coarse3 = np.digitize(es, [3, 6])
run(surv[np.argsort(coarse3, kind="stable")], "lv_tpp three buckets")
# a predictor available before LV runs: do the read's first 8 bases match the window on diagonal 0?  (a +-1..3 shifted twin
# does not, although all its words match on the shifted diagonal)
rid = (surv["rs"] >> 1).astype(np.int64); st = (surv["rs"] & 1).astype(bool); pos = surv["pos"].astype(np.int64)
rd = wl["reads"][rid]
rd = np.where(st[:, None], (np.where(rd[:, ::-1] < 4, 3 - rd[:, ::-1], rd[:, ::-1])), rd)[:, :8]
m = g.masks[pos[:, None] + np.arange(8)[None, :]]
first_ok = (((m >> np.minimum(rd, 3)) & 1).astype(bool) | (rd == 4)).all(axis=1)
print("first word matches on diagonal 0: %.1f%% of survivors; mean e there %.2f, elsewhere %.2f" % (100 * first_ok.mean(), es[first_ok].mean(), es[~first_ok].mean()))
run(surv[np.argsort(~first_ok, kind="stable")], "lv_tpp two buckets (first word)")
