#!/usr/bin/env python
"""Fuzz of the aligner program against the reference binary (oracle/_ref/salt): random genomes (optionally with a duplicated half,
so that every read has alternates), single-end and paired-end inputs with ragged read lengths, lower case, runs of N, all-N and
random reads, name suffixes and comments, several option sets, small batches.  Every SAM must equal the reference's except the @PG
line; inputs the reference itself crashes on are reported and skipped.  Runs salt_aln over the SIMT emulator by default
(tests/emul/salt_aln_emul, minutes per seed); --gpu takes salt_b200/salt_aln on a B200.

    python tools/fuzz_aln.py 1 2 3            # seeds
"""
import argparse
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "emul")]
import dropin_data  # noqa: E402
from salt_b200 import synth  # noqa: E402

REFDIR = os.path.join(ROOT, "oracle", "_ref")


def body(path):
    return [ln for ln in open(path, "rb").read().split(b"\n") if not ln.startswith(b"@PG")]


def one_seed(exe, seed, n, batch, threads, lens_sets):
    rng = np.random.default_rng(seed)
    bad_total = 0
    for copies in (False, True):
        d = tempfile.mkdtemp(prefix="fuzz_aln_")
        glen = 20000
        run = lambda cmd, out: subprocess.run(cmd, cwd=d, stdout=open(os.path.join(d, out), "w"), stderr=subprocess.PIPE).returncode
        dropin_data.write_inputs(d, glen=glen, n_reads=10, seed=seed, two_copies=copies)
        g = synth.Genome(glen, snp_rate=0.01, n_rate=0.0, seed=seed)
        if copies:
            g.codes[glen // 2:] = g.codes[:glen // 2]
        codes = (g.codes & 3).astype(np.uint8)
        assert run([os.path.join(REFDIR, "salt-idx"), "-k", "19", "ref.fa", "snps.txt", "idx"], "idx.log") == 0

        def reads(m, lens):
            out = []
            for _ in range(m):
                L = int(rng.choice(lens)); p = int(rng.integers(0, glen - L - 10))
                r = codes[p:p + L].copy()
                e = rng.random(L) < 0.02
                r[e] = (r[e] + rng.integers(1, 4, int(e.sum()))) & 3
                if rng.random() < 0.2 and L > 30:
                    c = int(rng.integers(10, L - 10)); r = np.concatenate([r[:c], r[c + 1:], codes[p + L:p + L + 1]])
                if rng.random() < 0.5:
                    r = synth.revcomp(r[None, :])[0]
                s = "".join("ACGT"[c] for c in r)
                u = rng.random()
                if u < 0.1:
                    s = s.lower()
                elif u < 0.2:
                    k = int(rng.integers(1, 12)); a = int(rng.integers(0, max(1, L - k))); s = s[:a] + "N" * k + s[a + k:]
                elif u < 0.25:
                    s = "N" * L
                elif u < 0.3:
                    s = "".join("ACGT"[c] for c in rng.integers(0, 4, L))
                out.append(s)
            return out

        def write(name, seqs, suffix):
            with open(os.path.join(d, name), "w") as f:
                for i, s in enumerate(seqs):
                    q = "".join(chr(33 + int(x)) for x in rng.integers(0, 41, len(s)))
                    f.write("@q%d%s\n%s\n+\n%s\n" % (i, suffix if i % 2 else " a comment", s, q))

        def compare(kind, flags, files):
            nonlocal bad_total
            if run([os.path.join(REFDIR, "salt")] + flags + ["-t", "2", "idx"] + files, "ref.sam"):
                print(kind, seed, copies, flags, "REFERENCE CRASHED"); return
            p = subprocess.run([exe] + flags + ["-t", str(threads), "-B", str(batch), "idx"] + files, cwd=d,
                               stdout=open(os.path.join(d, "mine.sam"), "w"), stderr=subprocess.PIPE, text=True)
            if p.returncode:
                print(kind, seed, copies, flags, "FAILED", p.stderr[-300:]); bad_total += 1; return
            a, b = body(os.path.join(d, "mine.sam")), body(os.path.join(d, "ref.sam"))
            bad = [i for i in range(max(len(a), len(b))) if i >= len(a) or i >= len(b) or a[i] != b[i]]
            bad_total += len(bad)
            print(kind, seed, copies, " ".join(flags), len(a), len(b), "mismatch", len(bad), flush=True)
            for i in bad[:2]:
                print("  mine", a[i][:300] if i < len(a) else None); print("  want", b[i][:300] if i < len(b) else None)
        for lens in lens_sets:
            write("reads.fq", reads(n, lens), "/1")
            for flags in (["-d", "-r", "1", "-c", "-m", "500"], ["-c"], ["-d", "-v", "-s", "5", "-m", "7", "-g", "x y"], ["-r", "3", "-m", "2"]):
                compare("SE", flags, ["reads.fq"])
            write("r1.fq", reads(n // 2, lens), "/1"); write("r2.fq", reads(n // 2, lens), "/2")
            for flags in (["-p", "-d", "-c", "-a", "100", "-b", "5000", "-r", "5"], ["-p", "-a", "50", "-b", "20000", "-m", "3"]):
                compare("PE", flags, ["r1.fq", "r2.fq"])
        dropin_data.write_pe_inputs(d, glen=glen, n_pairs=n, seed=seed, two_copies=copies)      # fragments with inserts: proper pairs, rescues
        assert run([os.path.join(REFDIR, "salt-idx"), "-k", "19", "ref.fa", "snps.txt", "idx"], "idx.log") == 0
        for flags in (["-d", "-p", "-e", "-l", "100", "-c", "-a", "350", "-b", "650", "-r", "5"], ["-p", "-l", "100", "-a", "350", "-b", "650"]):
            compare("PE", flags, ["r1.fq", "r2.fq"])
    return bad_total


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("seeds", type=int, nargs="+")
    ap.add_argument("--gpu", action="store_true", help="salt_b200/salt_aln on the device instead of the emulated build")
    ap.add_argument("--reads", type=int, default=60); ap.add_argument("--batch", type=int, default=25); ap.add_argument("--threads", type=int, default=3)
    a = ap.parse_args()
    from salt_b200 import build as b
    if a.gpu:
        exe = b.build_aln()
    else:
        import build_emul
        exe = b.build_aln(engine=build_emul.build(), hostlib=build_emul.build_host())
    lens_sets = [[100], [19, 36, 50, 75, 100, 125, 150, 250], [100, 101], [300, 400, 100]]
    bad = sum(one_seed(exe, s, a.reads, a.batch, a.threads, lens_sets) for s in a.seeds)
    print("fuzz_aln: %d mismatching lines or failed runs over %d seeds" % (bad, len(a.seeds)))
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
