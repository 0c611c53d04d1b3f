"""tools/salt_se.py -- the single-end program built from this repository's libraries only (FASTQ parser, seeding + locate +
verification, hit selection, tags, XA CIGARs, SAM lines; no reference code in the loop) -- against the reference program itself
(oracle/_ref/salt) on an index written by the reference's salt-idx.  Runs on the SIMT emulator, so the input is small; the GPU
version of the same comparison at scale is the drop-in's (tests/test_dropin.py)."""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools")); sys.path.insert(0, os.path.join(ROOT, "tests", "emul"))
REFDIR = os.path.join(ROOT, "oracle", "_ref")

import ctypes as C  # noqa: E402

import dropin_data  # noqa: E402
from salt_b200 import api, host_api  # noqa: E402


@pytest.fixture(scope="module")
def emul_lib():
    import build_emul
    return api._declare(C.CDLL(build_emul.build()))


@pytest.mark.parametrize("flags,kw", [
    (["-d", "-r", "1", "-l", "100", "-n", "20", "-c", "-m", "500"], dict(l_overlap=1, max_locate=500, print_xa_cigar=True, print_nm_md=True)),
    (["-l", "100", "-g", "grp7"], dict(rg_id=b"grp7")),
])
def test_single_end_program_from_own_parts(tmp_path, emul_lib, flags, kw):
    if not all(os.path.exists(os.path.join(REFDIR, f)) for f in ("salt", "salt-idx")):
        pytest.skip("oracle/_ref programs not built (reference tree absent at build time)")
    import build_emul
    import salt_se
    d = str(tmp_path)
    dropin_data.write_inputs(d, glen=12000, n_reads=36, seed=11, two_copies=True)     # alternates (XA) for every read
    run = lambda cmd, out: subprocess.run(cmd, cwd=d, stdout=open(os.path.join(d, out), "w"), stderr=subprocess.PIPE, check=True)
    run([os.path.join(REFDIR, "salt-idx"), "-k", "19", "ref.fa", "snps.txt", "idx"], "idx.log")
    run([os.path.join(REFDIR, "salt")] + flags + ["-t", "1", "idx", "reads.fq"], "ref.sam")
    want = [ln for ln in open(os.path.join(d, "ref.sam"), "rb").read().split(b"\n") if not ln.startswith(b"@")]
    while want and want[-1] == b"":
        want.pop()
    H = host_api.load(build_emul.build_host())
    body, names, lens = salt_se.align(emul_lib, H, os.path.join(d, "idx"), os.path.join(d, "reads.fq"), chunk_reads=40, **kw)
    assert len(body) == len(want) == 76
    for i, (a, b) in enumerate(zip(body, want)):
        assert a == b, (i, a, b)
    mapped = [ln for ln in body if ln.split(b"\t")[2] != b"*"]
    assert len(mapped) >= 30
    assert sum(b"\tXA:Z:" in ln for ln in mapped) >= 25
    if kw.get("print_nm_md"):
        assert all(b"\tMD:Z:" in ln and b"\tNM:i:" in ln for ln in mapped)
        assert sum(1 for ln in mapped if b"I" in ln.split(b"\t")[5] or b"D" in ln.split(b"\t")[5]) >= 2        # gapped primaries (and XA CIGARs)
    # the header lines this tool writes are the reference's too (except @PG)
    head = [ln for ln in open(os.path.join(d, "ref.sam"), "rb").read().split(b"\n") if ln.startswith(b"@") and not ln.startswith(b"@PG")]
    mine = [b"@HD\tVN:ec1fec2\tSO:unsorted"] + [b"@SQ\tSN:%s\tLN:%d" % (n, l) for n, l in zip(names, lens)] + \
           [b"@RG\tID:%s" % (kw["rg_id"] if kw.get("rg_id") else b"(null)")]
    assert head == mine


def _pe_reference(d, flags, **data_kw):
    dropin_data.write_pe_inputs(d, **data_kw)
    run = lambda cmd, out: subprocess.run(cmd, cwd=d, stdout=open(os.path.join(d, out), "w"), stderr=subprocess.PIPE, check=True)
    run([os.path.join(REFDIR, "salt-idx"), "-k", "19", "ref.fa", "snps.txt", "idx"], "idx.log")
    run([os.path.join(REFDIR, "salt")] + flags + ["idx", "r1.fq", "r2.fq"], "ref.sam")
    text = open(os.path.join(d, "ref.sam"), "rb").read()
    return [ln + b"\n" for ln in text.split(b"\n") if ln and not ln.startswith(b"@")]        # alnpe_sam's lines, blank separators dropped


@pytest.mark.parametrize("flags,kw,copies", [
    (["-d", "-p", "-e", "-l", "100", "-c", "-a", "350", "-b", "650", "-r", "5"],                                # run_pe_test.sh:14
     dict(min_tlen=350, max_tlen=650, l_overlap=5, print_xa_cigar=True, print_nm_md=True), True),
    (["-p", "-l", "100", "-a", "350", "-b", "650", "-g", "grp7"], dict(min_tlen=350, max_tlen=650, rg_id=b"grp7"), False),
])
def test_paired_end_program_from_own_parts(tmp_path, emul_lib, flags, kw, copies):
    """tools/salt_pe.py -- FASTQ parser, seeding + locate (paired-end flavour), verification, pairing plans, mate rescue,
    tags, XA CIGARs, SAM lines, all from this repository's libraries -- against the reference program's own output"""
    if not all(os.path.exists(os.path.join(REFDIR, f)) for f in ("salt", "salt-idx")):
        pytest.skip("oracle/_ref programs not built (reference tree absent at build time)")
    import build_emul
    import salt_pe
    d = str(tmp_path)
    want = _pe_reference(d, flags + ["-t", "1"], glen=12000, n_pairs=40, seed=21, two_copies=copies)
    H = host_api.load(build_emul.build_host())
    st = {}
    body, names, lens = salt_pe.align(emul_lib, H, os.path.join(d, "idx"), os.path.join(d, "r1.fq"), os.path.join(d, "r2.fq"),
                                      chunk_pairs=25, stats=st, **kw)
    assert len(body) == len(want) == 80
    for i, (a, b) in enumerate(zip(body, want)):
        assert a == b, (i, a, b)
    assert st["flagged_mates"] == 0 and st["declined"] == 0
    assert st["pairs"] == 40 and st["windows16"] + st["windows5"] >= 3 and st["rescued"] >= 2, st
    f = [ln.split(b"\t") for ln in body]
    assert sum(1 for x in f if int(x[1]) & 2) >= 50                       # properly paired records
    assert sum(1 for x in f if b"S" in x[5]) >= 1                         # soft-clipped = rescued by Smith-Waterman
    if copies:
        assert sum(b"\tXA:Z:" in ln for ln in body) >= 40
    if kw.get("print_nm_md"):
        assert all(b"\tMD:Z:" in ln for ln, x in zip(body, f) if x[5] != b"*")


@pytest.mark.gpu
def test_single_end_program_from_own_parts_on_the_gpu(tmp_path):
    """the same comparison on the device, 3 040 reads, two reference records, shipped flags"""
    if not all(os.path.exists(os.path.join(REFDIR, f)) for f in ("salt", "salt-idx")):
        pytest.skip("oracle/_ref programs not built (reference tree absent at build time)")
    import salt_se
    d = str(tmp_path)
    dropin_data.write_inputs(d, glen=150000, n_reads=3000, seed=17, two_copies=True)
    run = lambda cmd, out: subprocess.run(cmd, cwd=d, stdout=open(os.path.join(d, out), "w"), stderr=subprocess.PIPE, check=True)
    run([os.path.join(REFDIR, "salt-idx"), "-k", "19", "ref.fa", "snps.txt", "idx"], "idx.log")
    run([os.path.join(REFDIR, "salt"), "-d", "-r", "1", "-l", "100", "-n", "20", "-c", "-m", "500", "-t", "4", "idx", "reads.fq"], "ref.sam")
    want = [ln for ln in open(os.path.join(d, "ref.sam"), "rb").read().split(b"\n") if not ln.startswith(b"@")]
    while want and want[-1] == b"":
        want.pop()
    body, names, lens = salt_se.align(None, host_api.load(), os.path.join(d, "idx"), os.path.join(d, "reads.fq"), l_overlap=1, max_locate=500,
                                      print_xa_cigar=True, print_nm_md=True, chunk_reads=1000)
    assert len(body) == len(want) == 3040
    for i, (a, b) in enumerate(zip(body, want)):
        assert a == b, (i, a, b)
    assert sum(b"\tXA:Z:" in ln for ln in body) >= 2500


@pytest.mark.gpu
def test_paired_end_program_from_own_parts_on_the_gpu(tmp_path):
    """the same comparison on the device: 3 000 pairs, two reference records, the flags of run_pe_test.sh:14, four reference threads"""
    if not all(os.path.exists(os.path.join(REFDIR, f)) for f in ("salt", "salt-idx")):
        pytest.skip("oracle/_ref programs not built (reference tree absent at build time)")
    import salt_pe
    d = str(tmp_path)
    want = _pe_reference(d, ["-d", "-p", "-e", "-l", "100", "-c", "-a", "350", "-b", "650", "-r", "5", "-t", "4"])
    st = {}
    body, names, lens = salt_pe.align(None, host_api.load(), os.path.join(d, "idx"), os.path.join(d, "r1.fq"), os.path.join(d, "r2.fq"),
                                      min_tlen=350, max_tlen=650, l_overlap=5, print_xa_cigar=True, print_nm_md=True, chunk_pairs=1000, stats=st)
    assert len(body) == len(want) == 6000
    for i, (a, b) in enumerate(zip(body, want)):
        assert a == b, (i, a, b)
    assert st["flagged_mates"] == 0 and st["declined"] == 0 and st["rescued"] >= 100 and st["proper"] >= 2000, st


# ---- salt_aln: the same two programs as one C executable (salt_b200/host/salt_aln.c) --------------------------------------------

def _sam_lines(path):
    return [ln for ln in open(path, "rb").read().split(b"\n") if not ln.startswith(b"@PG")]      # @PG carries date + command line


def _many_n(path, every, n_bases):
    """a stretch of N in every `every`-th record of a FASTQ file (mates with more than five are not aligned, alnpe.c:491)"""
    lines = open(path).read().split("\n")
    for r in range(0, len(lines) // 4, every):
        s = lines[4 * r + 1]
        lines[4 * r + 1] = s[:20] + "N" * n_bases + s[20 + n_bases:]
    open(path, "w").write("\n".join(lines))


def _aln_case(d, exe, paired, flags, threads, batch, env=None, extra=(), files=None):
    run = lambda cmd, out: subprocess.run(cmd, cwd=d, stdout=open(os.path.join(d, out), "w"), stderr=subprocess.PIPE, check=True,
                                          env=dict(os.environ, **env) if env else None)
    files = files or (["r1.fq", "r2.fq"] if paired else ["reads.fq"])
    run([os.path.join(REFDIR, "salt")] + flags + ["-t", "2", "idx"] + files, "ref.sam")
    p = subprocess.run([exe] + flags + list(extra) + ["-t", str(threads)] + (["-B", str(batch)] if batch else []) + ["idx"] + files, cwd=d,
                       stdout=open(os.path.join(d, "mine.sam"), "w"), stderr=subprocess.PIPE, text=True)
    assert p.returncode == 0, p.stderr[-1500:]
    want, got = _sam_lines(os.path.join(d, "ref.sam")), _sam_lines(os.path.join(d, "mine.sam"))
    assert len(want) == len(got)
    for i, (a, b) in enumerate(zip(got, want)):
        assert a == b, (i, a, b)
    return want, p.stderr


def test_salt_aln_program_on_the_emulator(tmp_path):
    """salt_aln over the emulated engine against the reference program: header lines and every record identical, single-end and
    paired-end, several batches, more host threads than the reference run, mates with too many N"""
    if not all(os.path.exists(os.path.join(REFDIR, f)) for f in ("salt", "salt-idx")):
        pytest.skip("oracle/_ref programs not built (reference tree absent at build time)")
    import build_emul
    from salt_b200 import build as b
    exe = b.build_aln(engine=build_emul.build(), hostlib=build_emul.build_host())
    d = str(tmp_path)
    dropin_data.write_inputs(d, glen=12000, n_reads=60, seed=31, two_copies=True)
    subprocess.run([os.path.join(REFDIR, "salt-idx"), "-k", "19", "ref.fa", "snps.txt", "idx"], cwd=d, stdout=open(os.path.join(d, "idx.log"), "w"),
                   stderr=subprocess.PIPE, check=True)
    want, err = _aln_case(d, exe, False, ["-d", "-r", "1", "-l", "100", "-n", "20", "-c", "-m", "500"], 3, 48)
    assert sum(b"\tXA:Z:" in ln for ln in want) >= 40 and sum(b"\tMD:Z:" in ln for ln in want) >= 50
    assert "100 reads" in err
    import gzip
    open(os.path.join(d, "reads.fq.gz"), "wb").write(gzip.compress(open(os.path.join(d, "reads.fq"), "rb").read()))
    _aln_case(d, exe, False, ["-l", "100", "-g", "grp7"], 1, 0, files=["reads.fq.gz"])          # gzip input, as the reference's reader takes it
    dropin_data.write_pe_inputs(d, glen=12000, n_pairs=50, seed=31, two_copies=True)
    _many_n(os.path.join(d, "r2.fq"), 7, 8)
    want, err = _aln_case(d, exe, True, ["-d", "-p", "-e", "-l", "100", "-c", "-a", "350", "-b", "650", "-r", "5"], 3, 40)
    f = [ln.split(b"\t") for ln in want if ln and not ln.startswith(b"@")]
    assert len(f) == 100 and sum(1 for x in f if int(x[1]) & 2) >= 60 and sum(1 for x in f if b"S" in x[5]) >= 3
    assert "pairs 50:" in err and "windows declined 0" in err
    # candidate lists longer than the room they are first given on the device (-L 64): the batch is located again with more
    want, err = _aln_case(d, exe, True, ["-p", "-l", "100", "-a", "350", "-b", "650", "-g", "grp7", "-r", "2"], 2, 0, extra=["-L", "64"])
    assert "located again with more list room 1," in err and "cut at 16384 loci 0" in err


@pytest.mark.gpu
def test_salt_aln_program_on_the_gpu(tmp_path):
    """salt_b200/salt_aln on the device against the reference program at -t 4: 6 040 single-end reads with the shipped flags, 3 000 pairs
    with the flags of run_pe_test.sh:14 (every seventh second mate with eight N: not aligned, rescued through its mate)"""
    if not all(os.path.exists(os.path.join(REFDIR, f)) for f in ("salt", "salt-idx")):
        pytest.skip("oracle/_ref programs not built (reference tree absent at build time)")
    from salt_b200 import build as b
    exe = b.build_aln()
    d = str(tmp_path)
    dropin_data.write_inputs(d, seed=9)              # the genome write_pe_inputs uses: one index for both halves of the test
    subprocess.run([os.path.join(REFDIR, "salt-idx"), "-k", "19", "ref.fa", "snps.txt", "idx"], cwd=d, stdout=open(os.path.join(d, "idx.log"), "w"),
                   stderr=subprocess.PIPE, check=True)
    want, err = _aln_case(d, exe, False, ["-d", "-r", "1", "-l", "100", "-n", "20", "-c", "-m", "500"], 4, 2500)
    assert len(want) > 6000 and "6040 reads" in err
    _aln_case(d, exe, False, ["-l", "100"], 16, 0)
    dropin_data.write_pe_inputs(d)
    _many_n(os.path.join(d, "r2.fq"), 7, 8)
    want, err = _aln_case(d, exe, True, ["-d", "-p", "-e", "-l", "100", "-c", "-a", "350", "-b", "650", "-r", "5"], 4, 2500)
    assert "pairs 3000:" in err and "windows declined 0" in err
    f = [ln.split(b"\t") for ln in want if ln and not ln.startswith(b"@")]
    assert sum(1 for x in f if int(x[1]) & 2) >= 4000 and sum(1 for x in f if b"S" in x[5]) >= 20


def test_salt_aln_ragged_inputs_on_the_emulator(tmp_path):
    """reads of 19 .. 150 bases in one file (the seed length included), lower case, runs of N, all-N and random reads, names
    with a /1 suffix and comments, mates of different lengths; seed options -v -s -m -r, a read group with a blank"""
    if not all(os.path.exists(os.path.join(REFDIR, f)) for f in ("salt", "salt-idx")):
        pytest.skip("oracle/_ref programs not built (reference tree absent at build time)")
    import build_emul
    from salt_b200 import build as b, synth
    exe = b.build_aln(engine=build_emul.build(), hostlib=build_emul.build_host())
    d = str(tmp_path)
    glen, seed = 20000, 4
    rng = np.random.default_rng(seed)
    dropin_data.write_inputs(d, glen=glen, n_reads=10, seed=seed, two_copies=True)
    g = synth.Genome(glen, snp_rate=0.01, n_rate=0.0, seed=seed)
    g.codes[glen // 2:] = g.codes[:glen // 2]
    codes = (g.codes & 3).astype(np.uint8)
    subprocess.run([os.path.join(REFDIR, "salt-idx"), "-k", "19", "ref.fa", "snps.txt", "idx"], cwd=d, stdout=open(os.path.join(d, "idx.log"), "w"),
                   stderr=subprocess.PIPE, check=True)

    def reads(n, lens):
        out = []
        for _ in range(n):
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
    lens = [19, 36, 50, 75, 100, 125, 150]       # (250-base reads too: checked off-line, the emulator needs minutes for them)
    write("reads.fq", reads(32, lens), "/1")
    _aln_case(d, exe, False, ["-d", "-r", "4", "-c", "-m", "500", "-v", "-s", "5", "-g", "x y"], 3, 20)
    write("r1.fq", reads(16, lens), "/1"); write("r2.fq", reads(16, lens), "/2")
    want, err = _aln_case(d, exe, True, ["-p", "-d", "-c", "-a", "100", "-b", "5000", "-r", "9"], 3, 0)
    assert "pairs 16:" in err


def test_salt_aln_error_paths_on_the_emulator(tmp_path):
    """what the program does with input it cannot use: a message and a non-zero exit code, never a partial SAM passed off as whole"""
    if not os.path.exists(os.path.join(REFDIR, "salt-idx")):
        pytest.skip("oracle/_ref programs not built (reference tree absent at build time)")
    import build_emul
    from salt_b200 import build as b
    exe = b.build_aln(engine=build_emul.build(), hostlib=build_emul.build_host())
    d = str(tmp_path)
    run = lambda args: subprocess.run([exe] + args, cwd=d, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    p = run([])
    assert p.returncode == 1 and "Usage: salt_aln" in p.stderr
    p = run(["-p", "idx", "only_one.fq"])
    assert p.returncode == 1 and "Usage: salt_aln" in p.stderr
    p = run(["-X", "1", "idx", "reads.fq"])                             # row H: the reference aborts on that path
    assert p.returncode == 1 and "not served" in p.stderr
    open(os.path.join(d, "reads.fq"), "w").write("@r0\nACGTACGTACGTACGTACGTACGTACGT\n+\nIIIIIIIIIIIIIIIIIIIIIIIIIIII\n")
    p = run(["nothing_here", "reads.fq"])
    assert p.returncode == 1 and "cannot open nothing_here.C.bwt" in p.stderr and p.stdout == ""
    dropin_data.write_pe_inputs(d, glen=12000, n_pairs=6, seed=3)
    subprocess.run([os.path.join(REFDIR, "salt-idx"), "-k", "19", "ref.fa", "snps.txt", "idx"], cwd=d, stdout=open(os.path.join(d, "idx.log"), "w"),
                   stderr=subprocess.PIPE, check=True)
    p = run(["idx", "no_such_reads.fq"])
    assert p.returncode == 1 and "cannot open no_such_reads.fq" in p.stderr
    open(os.path.join(d, "broken.fq"), "w").write("@r0\nACGTACGTACGTACGTACGTACGTACGT\n+\nIIII")            # quality shorter than the sequence
    p = run(["idx", "broken.fq"])
    assert p.returncode == 1 and "malformed FASTQ" in p.stderr
    open(os.path.join(d, "empty.fq"), "w").write("")
    p = run(["idx", "empty.fq"])
    assert p.returncode == 0 and "0 reads" in p.stderr
    assert [ln[:3] for ln in p.stdout.split("\n") if ln] == ["@HD", "@SQ", "@SQ", "@RG", "@PG"]
    # the second file two records short: the complete pairs are aligned, the surplus is reported, not silently dropped
    lines = open(os.path.join(d, "r2.fq")).read().split("\n")
    open(os.path.join(d, "r2_short.fq"), "w").write("\n".join(lines[:16]) + "\n")
    p = run(["-p", "-a", "350", "-b", "650", "idx", "r1.fq", "r2_short.fq"])
    assert p.returncode == 0 and "different numbers of records" in p.stderr and "pairs 4:" in p.stderr
    assert sum(1 for ln in p.stdout.split("\n") if ln and not ln.startswith("@")) == 8
    p = run(["-p", "-b", "0", "idx", "r1.fq", "r2.fq"])                  # alnpe.c:583
    assert p.returncode == 1 and "infer isize" in p.stderr


def test_salt_aln_skips_reads_with_too_many_n_on_the_emulator(tmp_path):
    """single-end reads with more than 200 ambiguous bases are not aligned and leave an empty line (alnse.c:1281, :1296); the others
    of the batch are compacted around them"""
    if not all(os.path.exists(os.path.join(REFDIR, f)) for f in ("salt", "salt-idx")):
        pytest.skip("oracle/_ref programs not built (reference tree absent at build time)")
    import build_emul
    from salt_b200 import build as b, synth
    exe = b.build_aln(engine=build_emul.build(), hostlib=build_emul.build_host())
    d = str(tmp_path)
    glen = 20000
    dropin_data.write_inputs(d, glen=glen, n_reads=10, seed=8)
    codes = (synth.Genome(glen, snp_rate=0.01, n_rate=0.0, seed=8).codes & 3).astype(np.uint8)
    rng = np.random.default_rng(1)
    with open(os.path.join(d, "reads.fq"), "w") as f:
        for i in range(8):
            L = 250; p = int(rng.integers(0, glen - L - 5)); s = "".join("ACGT"[c] for c in codes[p:p + L])
            if i in (2, 7):
                s = s[:20] + "N" * (201 + i) + s[221 + i:]
            if i == 5:
                s = s[:20] + "N" * 200 + s[220:]                  # exactly the limit: still aligned
            f.write("@r%d\n%s\n+\n%s\n" % (i, s, "I" * L))
    subprocess.run([os.path.join(REFDIR, "salt-idx"), "-k", "19", "ref.fa", "snps.txt", "idx"], cwd=d, stdout=open(os.path.join(d, "idx.log"), "w"),
                   stderr=subprocess.PIPE, check=True)
    want, err = _aln_case(d, exe, False, ["-c", "-l", "250"], 2, 5)
    assert sum(1 for ln in want[4:-1] if ln == b"") == 2 and "8 reads" in err
