"""TEST INFRASTRUCTURE: the SIMT-emulated engine built with AddressSanitizer and zero allocation slack, driven
through verify / mismatch / LV / SSW with candidates at both ends of the reference.  compute-sanitizer is not
available on the GPU pool, so out-of-bounds accesses in kernel logic are hunted here.
    g++ -std=c++17 -O1 -g -fPIC -shared -pthread -fsanitize=address -DSALT_EMUL_SLACK=0 -o /tmp/libsalt_b200_emul_asan.so tests/emul/emul_lib.cpp
    LD_PRELOAD=$(gcc -print-file-name=libasan.so) ASAN_OPTIONS=detect_leaks=0 python tests/emul/asan_check.py"""
import sys, ctypes as C, numpy as np
import os
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import parity_cases as pc
from salt_b200 import api, synth
from oracle import orc
lib = api._declare(C.CDLL("/tmp/libsalt_b200_emul_asan.so"))
o = orc.Oracle()
for L, glen in ((100, 30000), (150, 30000), (250, 30000), (37, 20000), (700, 40000)):
    n = 40 if L <= 150 else 10
    g, reads, pos, strand, cands = pc.make_world(400 + L, L=L, n_reads=n, per_strand=5, indel_frac=0.35, glen=glen,
                                                 sub_rate=0.02 if L <= 150 else 0.004)
    # candidates right at both ends of the reference
    offs0, loci0, offs1, loci1 = cands
    loci0 = loci0.copy(); loci0[0] = 0; loci0[offs0[1] - 1] = max(loci0[offs0[1] - 2], g.l - L)
    e = int(offs1[n]); loci1 = loci1.copy(); loci1[e - 1] = g.l - L - 4; loci1[e - 2] = min(loci1[e-2], g.l - L - 5)
    eng = api.Engine(g.mixref, g.l, g.pac, g.l, lib=lib)
    eng.set_reads(reads)
    pc.check_verify(eng, o, g, reads, (offs0, loci0, offs1, loci1), 3, -1)
    pc.check_verify(eng, o, g, reads, (offs0, loci0, offs1, loci1), 3, 3)
    pairs = pc.flat_pairs((offs0, loci0, offs1, loci1), n)
    pc.check_mismatch(eng, o, g, reads, pairs[:60], 3)
    pc.check_lv(eng, o, g, reads, pairs[:60], -1)
    for mapping in (1, 2):                       # warp-per-pair and thread-per-pair LV, CIGAR kernels of both kinds
        eng.set_lv_mapping(mapping)
        pc.check_lv(eng, o, g, reads, pairs[:40], 3)
        tp = api.Engine.make_pairs(np.arange(n, dtype=np.uint32), strand, pos)
        pc.check_lv_cigar(eng, o, g, reads, tp, np.full(n, min(15, L // 10 + 2), np.uint8), 128)
    eng.set_lv_mapping(0)
    cases_seed = 1000 + L
    pc.check_md_nm(eng, o, g, reads, pos, strand, cases_seed, md_stride=2 * L + 64)
    rng = np.random.default_rng(L)
    if L <= 250:
        wins = pc.make_windows(g, reads[:12], pos[:12], strand[:12], L, rng, 301)
        pc.check_ssw(eng, o, g, reads, wins, False, api.salt_score_mat2(), 16, cigar_stride=96)
    eng.close()
    print("asan ok", L, flush=True)

# ---- round 2: compact transport, SAM tail of primaries, wide-band rescues, seeding + locate (both flavours)
import tempfile
import seed_cases as sc
from salt_b200 import index_io
g, reads, pos, strand, cands = pc.make_world(77, L=100, n_reads=30, per_strand=3, indel_frac=0.4, glen=30000, n_frac=0.02)
eng = api.Engine(g.mixref, g.l, g.pac, g.l, lib=lib)
pc.check_verify_packed(eng, [r for r in reads], cands, 7, variants=[(2, 16, 3), (4, 32, 0)])
rng = np.random.default_rng(5)
pc.check_verify_packed(eng, [r[:int(rng.integers(37, 101))] for r in reads], cands, 11, variants=[(2, 32, 1)])
pc.check_tail_primaries(eng, o, g, reads, cands)
eng.close()
g3 = synth.Genome(20003, snp_rate=0.01, n_rate=0.0, seed=77)
eng = api.Engine(g3.mixref, g3.l, None, 0, lib=lib)
pc.check_ssw_wide_bands(eng, o, 77, n_reads=140)
eng.close()
g4 = synth.Genome(20011, snp_rate=0.01, n_rate=0.0, seed=9)
eng = api.Engine(g4.mixref, g4.l, g4.pac, g4.l, lib=lib)
pc.check_ssw_narrow_bands(eng, o, 9, n_reads=64)          # bands 1..3 in registers, warp-per-task bands, both scoring flavours
eng.close()
print("asan ok transport / tail / wide bands", flush=True)
if sc.have_ref():
    rng = np.random.default_rng(3)
    gg, is_n = sc.repeat_genome(rng, n_units=16, unit_len=700, n_rate=0.001)
    prefix = sc.write_index(tempfile.mkdtemp(prefix="salt_asan_"), gg, is_n, rng)
    fm = index_io.FmIndex(prefix)
    codes, roffs = sc.sample_reads(gg, rng, 16)
    ref = orc.SeedRef(prefix)
    eng = api.Engine(fm.mixref, fm.l, None, 0, lib=lib)
    eng.set_index(fm)
    sc.check_lists(eng, ref, fm, codes, roffs, option_sets=sc.OPTION_SETS[:4])
    sc.check_lists_pe(eng, ref, fm, codes, roffs, option_sets=sc.PE_OPTION_SETS[:2])
    print("asan ok seeding", flush=True)
