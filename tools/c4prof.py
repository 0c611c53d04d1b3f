import sys, numpy as np
sys.path.insert(0, "/root/repo")
from salt_b200 import api, synth
g = synth.Genome(50_000_000, snp_rate=0.01, seed=3)
for L, sub, indel, n in ((250, 0.04, 0.6, 400_000), (250, 0.01, 0.02, 400_000), (150, 0.01, 0.02, 700_000)):
    reads, pos, strand = synth.sample_reads(g, n, L, seed=10 + L, sub_rate=sub, indel_frac=indel, max_indel=6 if indel > 0.5 else 3)
    c = synth.make_candidates(g, pos, strand, L, per_strand=8, seed=20 + L)
    eng = api.Engine(g.mixref, g.l, g.pac, g.l); eng.set_reads(reads)
    eng.verify(*c, 3, -1)
    eng.profile(True)
    rec = eng.verify(*c, 3, -1)[0]
    print(L, sub, indel, {k: round(v, 3) for k, v in eng.profile_read().items()}, "lv_ran", rec["lv_ran"].mean(), "gapped", (rec["is_gap"] == 1).mean())
    eng.close()
