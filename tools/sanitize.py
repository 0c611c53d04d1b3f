"""compute-sanitizer target: one small pass through every round-2 kernel (transport scans / unpack, seeding + locate in
both flavours, verify on seeded lists, SAM tail of primaries, the paired-end chunk stage, slim CIGAR download).
    compute-sanitizer --tool memcheck python tools/sanitize.py"""
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import parity_cases as pc
    import seed_cases as sc
    from oracle import orc
    from salt_b200 import api, host_api, index_io, synth
    o = orc.Oracle()
    g, reads, pos, strand, cands = pc.make_world(5, glen=120000, L=100, n_reads=600, per_strand=5, indel_frac=0.3, n_frac=0.005)
    eng = api.Engine(g.mixref, g.l, g.pac, g.l, device=0)
    print("packed variants", pc.check_verify_packed(eng, [r for r in reads], cands, 97, variants=[(2, 16, 3), (4, 32, 0)]))
    print("tail", pc.check_tail_primaries(eng, o, g, reads, cands))
    print("pairs", pc.check_chunk_pair(eng, host_api.load(), g, 150, 100, seed=3).rescued)
    g2, r2, c2 = pc.check_long_cigars(None, o, 31, n_reads=100)
    e2 = api.Engine(g2.mixref, g2.l, None, 0, device=0); e2.set_reads(r2); pc.check_verify_batch(e2, r2, c2, 16); e2.close()
    eng.close()
    if sc.have_ref():
        rng = np.random.default_rng(3)
        gg, is_n = sc.repeat_genome(rng, n_units=20, unit_len=800, n_rate=0.001)
        d = tempfile.mkdtemp(prefix="salt_sanitize_")
        prefix = sc.write_index(d, gg, is_n, rng)
        fm = index_io.FmIndex(prefix)
        codes, roffs = sc.sample_reads(gg, rng, 300)
        ref = orc.SeedRef(prefix)
        e3 = api.Engine(fm.mixref, fm.l, None, 0, device=0); e3.set_index(fm)
        print("seed lists", sc.check_lists(e3, ref, fm, codes, roffs, option_sets=sc.OPTION_SETS[:3]))
        print("seed lists pe", sc.check_lists_pe(e3, ref, fm, codes, roffs, option_sets=sc.PE_OPTION_SETS[:2]))
        e3.close(); ref.close()
    print("sanitize target done")


if __name__ == "__main__":
    main()
