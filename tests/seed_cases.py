"""Inputs and checks for row f1 (seeding + locate): indexes built on the spot by the reference's own salt-idx
(oracle/_ref), candidate lists from the reference's own alnse_seed_overlap + alnse_locate_alt (libsaltref_seed.so)."""
import os
import subprocess

import numpy as np

from salt_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REFDIR = os.path.join(ROOT, "oracle", "_ref")


def have_ref():
    return all(os.path.exists(os.path.join(REFDIR, f)) for f in ("salt-idx", "libsaltref_seed.so"))


def repeat_genome(rng, n_units=40, unit_len=1200, div=0.03, n_rate=0.0):
    """copies of one unit with a few per-copy differences between random spacers: wide suffix-array intervals, seed
    extension, ties between intervals of equal width, lists that hit max_locate"""
    unit = rng.integers(0, 4, unit_len)
    parts = []
    for _ in range(n_units):
        u = unit.copy(); m = rng.random(unit_len) < div; u[m] = rng.integers(0, 4, int(m.sum()))
        parts += [rng.integers(0, 4, int(rng.integers(200, 900))), u]
    g = np.concatenate(parts).astype(np.uint8)
    is_n = rng.random(len(g)) < n_rate
    return g, is_n


def write_index(d, g, is_n, rng, snp_rate=0.015, k=19, records=1):
    os.makedirs(d, exist_ok=True)
    cut = [len(g) * i // records for i in range(records + 1)]
    s = "".join("ACGTN"[4 if n else c] for c, n in zip(g, is_n))
    with open(os.path.join(d, "ref.fa"), "w") as f:
        for r in range(records):
            f.write(">chr%d\n" % r)
            part = s[cut[r]:cut[r + 1]]
            for i in range(0, len(part), 60):
                f.write(part[i:i + 60] + "\n")
    sp = np.unique(rng.integers(0, len(g), int(len(g) * snp_rate)))
    with open(os.path.join(d, "snps.txt"), "w") as f:
        for p in sp:
            if is_n[p]:
                continue
            r = max(i for i in range(records) if cut[i] <= p)
            alt = (int(g[p]) + int(rng.integers(1, 4))) & 3
            a, b = sorted(["ACGT"[g[p]], "ACGT"[alt]])
            f.write("chr%d\t%d\t%s/%s\t%s\n" % (r, p - cut[r] + 1, a, b, "ACGT"[g[p]]))
    with open(os.path.join(d, "idx.log"), "w") as log:
        subprocess.check_call([os.path.join(REFDIR, "salt-idx"), "-k", str(k), "ref.fa", "snps.txt", "idx"], cwd=d,
                              stdout=log, stderr=subprocess.STDOUT)
    return os.path.join(d, "idx")


def sample_reads(g, rng, n, ragged=True, sub=0.02, n_every=7):
    reads = []
    for i in range(n):
        L = 100 if (i % 3 or not ragged) else int(rng.integers(60, 251))
        p = int(rng.integers(0, len(g) - L)); r = g[p:p + L].copy()
        e = rng.random(L) < sub; r[e] = rng.integers(0, 4, int(e.sum()))
        if n_every and i % n_every == 0:
            r[rng.integers(0, L)] = 4
        if i % 2:
            r = synth.revcomp(r)
        reads.append(r.astype(np.uint8))
    roffs = np.concatenate([[0], np.cumsum([len(r) for r in reads])]).astype(np.uint32)
    return np.concatenate(reads), roffs


OPTION_SETS = ((0, 1, 50, 500, 0),        # run_se_test.sh:12: -r 1 (a seed at every position: up to l_seq - l_seed + 1 intervals per index), -m 500
               (0, 0, 50, 500, 0),        # the same with the default seed spacing (l_overlap = l_seed)
               (0, 0, 2, 7, 0),           # tight caps: extension runs long, every list is cut
               (0, 5, 0, 1000, 0),        # overlapping seeds, extension until unique
               (0, 0, 1000, 33, 0),       # no extension, wide intervals against a small max_locate
               (0, 7, 3, 1, 1),           # seed_only_ref, one locus per list
               (25, 10, 50, 200, 0))      # a seed longer than the indexed one


def check_lists(eng, ref, fm, codes, roffs, option_sets=OPTION_SETS):
    from salt_b200 import api
    eng.set_reads(codes, roffs)
    total = 0
    for (ls, lo, ms, ml, ro) in option_sets:
        ls = ls or fm.l_seed
        want = ref.run(codes, roffs, ls, lo, ms, ml, ro)
        got = eng.seed_locate(api.Engine.seed_opt(ls, lo, ms, ml, ro))
        for a, b, name in zip(got, want, ("offs0", "loci0", "offs1", "loci1")):
            assert np.array_equal(a, b), (name, (ls, lo, ms, ml, ro), len(a), len(b))
        total += int(want[0][-1]) + int(want[2][-1])
    return total


PE_OPTION_SETS = ((0, 5, 50, 200), (0, 0, 50, 1000), (0, 5, 3, 6), (0, 7, 1000, 40))   # (l_seed, l_overlap, max_seed, max_locate)


def check_lists_pe(eng, ref, fm, codes, roffs, option_sets=PE_OPTION_SETS, list_cap=4096):
    """the paired-end program's flavour (alnse_seed_overlap + alnse_locate, alnse.c:501-631).  Strands the device flags
    (an SNP-context interval wider than max_locate: the reference draws rand() there; or a list that filled list_cap) have
    no single right answer and are only counted."""
    from salt_b200 import api
    eng.set_reads(codes, roffs)
    total = flagged = 0
    for (ls, lo, ms, ml) in option_sets:
        ls = ls or fm.l_seed
        want = ref.run(codes, roffs, ls, lo, ms, ml, 0, locate_mode=1, cap_per_read=20000)
        got = eng.seed_locate(api.Engine.seed_opt(ls, lo, ms, ml, 0, locate_mode=1, list_cap=list_cap))
        st = eng.seed_status()
        for s_ in (0, 1):
            go, gl, wo, wl = got[2 * s_], got[2 * s_ + 1], want[2 * s_], want[2 * s_ + 1]
            for r in range(len(roffs) - 1):
                if st[s_][r]:
                    flagged += 1
                    continue
                a = gl[go[r]:go[r + 1]]; b = wl[wo[r]:wo[r + 1]]
                assert np.array_equal(a, b), ("pe lists", (ls, lo, ms, ml), s_, r, len(a), len(b))
                total += len(b)
    return total, flagged
