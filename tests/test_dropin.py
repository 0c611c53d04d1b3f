"""End-to-end SAM parity (BASELINE.json configs[0] shape): the reference's own `salt` binary and
`salt_dropin` -- the same reference sources with ONE function (alnse_core) replaced by
oracle/dropin/alnse_core_gpu.c, which sends the verification stage to libsalt_b200.so -- must print
identical SAM for the same index and reads.  Both binaries and the reference indexer are built from
the reference tree by oracle/Makefile in the build container and travel in oracle/_ref/; the index
is built on the spot (the 12-mer lookup table alone is 64 MiB)."""
import os
import subprocess

import pytest

import dropin_data

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REFDIR = os.path.join(ROOT, "oracle", "_ref")

pytestmark = pytest.mark.gpu


def _have():
    return all(os.path.exists(os.path.join(REFDIR, b)) for b in ("salt", "salt-idx", "salt_dropin"))


def _run(cmd, cwd, out, env=None):
    with open(out, "w") as f:
        p = subprocess.run(cmd, cwd=cwd, stdout=f, stderr=subprocess.PIPE, text=True, timeout=900,
                           env=dict(os.environ, **env) if env else None)
    assert p.returncode == 0, p.stderr[-2000:]
    return p.stderr


def _sam_body(path):
    return [ln for ln in open(path).read().split("\n") if not ln.startswith("@PG")]     # @PG carries date + command line


@pytest.mark.parametrize("flags,seed", [(["-d", "-r", "1", "-l", "100", "-n", "20", "-c", "-m", "500", "-t", "1"], "host"),     # run_se_test.sh:12
                                        (["-d", "-r", "1", "-l", "100", "-n", "20", "-c", "-m", "500", "-t", "4"], "host"),     # ... as shipped: 4 seeding threads
                                        (["-r", "1", "-l", "100", "-t", "1"], "host"),
                                        (["-d", "-r", "1", "-l", "100", "-n", "20", "-c", "-m", "500", "-t", "4"], "gpu"),      # seeding + locate on the device too
                                        (["-r", "1", "-l", "100", "-t", "1"], "gpu"),
                                        # ... and the SAM lines written by salt_sam_se instead of the reference's aln_samse
                                        (["-d", "-r", "1", "-l", "100", "-n", "20", "-c", "-m", "500", "-t", "4"], "gpu+sam"),
                                        (["-r", "1", "-l", "100", "-t", "2"], "host+sam")])
def test_se_sam_identical(tmp_path, flags, seed):
    if not _have():
        pytest.skip("oracle/_ref programs not built (reference tree absent at build time)")
    d = str(tmp_path)
    fa, sn, fq = dropin_data.write_inputs(d)
    _run([os.path.join(REFDIR, "salt-idx"), "-k", "19", "ref.fa", "snps.txt", "idx"], d, os.path.join(d, "idx.log"))
    _run([os.path.join(REFDIR, "salt")] + flags + ["idx", "reads.fq"], d, os.path.join(d, "ref.sam"))
    native = seed.endswith("+sam"); seed = seed.split("+")[0]
    err = _run([os.path.join(REFDIR, "salt_dropin")] + flags + ["idx", "reads.fq"], d, os.path.join(d, "gpu.sam"),
               env={"SALT_DROPIN_SEED": seed, "SALT_DROPIN_SAM": "native" if native else "reference"})
    assert "verification on libsalt_b200" in err
    assert ("seeding + locate on libsalt_b200" in err) == (seed == "gpu")
    assert ("SAM lines by salt_sam_se" in err) == native
    if seed == "host":
        assert "%s seeding threads" % flags[flags.index("-t") + 1] in err
    want, got = _sam_body(os.path.join(d, "ref.sam")), _sam_body(os.path.join(d, "gpu.sam"))
    assert len(want) == len(got) and len(want) > 6000
    for a, b in zip(want, got):
        assert a == b
    body = [ln.split("\t") for ln in want if ln and not ln.startswith("@")]
    if "-d" in flags:        # the MD/NM/XV tags and the XA alternates' CIGARs in that SAM came from the GPU
        import re
        m = re.search(r"MD/NM/XV tags prepared (\d+) printed (\d+), XA CIGARs prepared (\d+) printed (\d+)", err)
        n_md = sum(1 for f in body if any(x.startswith("MD:Z:") for x in f[11:]))
        n_xa_gapped = sum(sum(1 for alt in x[5:].split(";") if alt and ("I" in alt.split(",")[2] or "D" in alt.split(",")[2]))
                          for f in body for x in f[11:] if x.startswith("XA:Z:"))
        assert m and int(m.group(2)) == n_md and n_md >= 5000, (m and m.groups(), n_md)
        assert int(m.group(4)) == n_xa_gapped and int(m.group(3)) >= int(m.group(4)), (m.groups(), n_xa_gapped)
    assert sum(1 for f in body if f[1] == "4") >= 30                       # unmapped reads exist
    assert sum(1 for f in body if "I" in f[5] or "D" in f[5]) >= 300       # gapped CIGARs exist
    assert sum(1 for f in body if any(x.startswith("XA:") for x in f[11:])) >= 1


@pytest.mark.parametrize("flags", [["-d", "-p", "-e", "-l", "100", "-c", "-a", "350", "-b", "650", "-r", "5", "-t", "1"],    # run_pe_test.sh:14
                                   ["-d", "-p", "-e", "-l", "100", "-c", "-a", "350", "-b", "650", "-r", "5", "-t", "4", "GPUSEED"]])
def test_pe_sam_identical(tmp_path, flags):
    """paired-end: both mates verified on the GPU with the PE thresholds, then the reference's own pairing2 /
    mate rescue / alnpe_sam"""
    if not _have():
        pytest.skip("oracle/_ref programs not built (reference tree absent at build time)")
    d = str(tmp_path)
    seed_env = {"SALT_DROPIN_SEED": "gpu"} if "GPUSEED" in flags else {"SALT_DROPIN_SEED": "host"}
    flags = [f for f in flags if f != "GPUSEED"]
    dropin_data.write_pe_inputs(d)
    _run([os.path.join(REFDIR, "salt-idx"), "-k", "19", "ref.fa", "snps.txt", "idx"], d, os.path.join(d, "idx.log"))
    _run([os.path.join(REFDIR, "salt")] + flags + ["idx", "r1.fq", "r2.fq"], d, os.path.join(d, "ref.sam"))
    err = _run([os.path.join(REFDIR, "salt_dropin")] + flags + ["idx", "r1.fq", "r2.fq"], d, os.path.join(d, "gpu.sam"), env=seed_env)
    assert "verification on libsalt_b200" in err
    import re
    m = re.search(r"rescue windows recorded: (\d+), served by the reference's own ssw_align: (\d+)", err)
    assert m and int(m.group(1)) >= 200 and int(m.group(2)) <= int(m.group(1)) // 20, err[-500:]    # the rescues ran on the GPU
    # the host layer's re-staged pairing predicted every decision of the reference's pairing2 / pairing_singleton
    m = re.search(r"salt_pair_plan: (\d+) pairs checked .*?, (\d+) proper without rescue, (\d+) windows planned, (\d+) mismatches", err)
    assert m and int(m.group(1)) >= 2900 and int(m.group(2)) >= 2000 and int(m.group(3)) >= 200 and int(m.group(4)) == 0, err[-800:]
    # ... and salt_pair_apply reproduced the final state of both mates (primaries, rescued mates with their soft clips and
    # Smith-Waterman CIGARs, alternates promoted to primary)
    m = re.search(r"salt_pair_apply: (\d+) pairs checked .*?\((\d+) rescued mates, (\d+) alternates promoted.*?(\d+) not checked.*?(\d+) mismatches", err)
    assert m and int(m.group(1)) >= 2800 and int(m.group(2)) >= 100 and int(m.group(5)) == 0, err[-800:]
    want, got = _sam_body(os.path.join(d, "ref.sam")), _sam_body(os.path.join(d, "gpu.sam"))
    assert len(want) == len(got) and len(want) > 6000
    for a, b in zip(want, got):
        assert a == b
    # the re-staged flow: rescue windows scheduled from salt_pair_plan alone (no recording run of the reference's pairing)
    err2 = _run([os.path.join(REFDIR, "salt_dropin")] + flags + ["idx", "r1.fq", "r2.fq"], d, os.path.join(d, "gpu2.sam"),
                env=dict(seed_env, SALT_DROPIN_PLAN="1"))
    m2 = re.search(r"(\d+) pairs scheduled from the plan alone", err2)
    assert m2 and int(m2.group(1)) >= 2900
    assert _sam_body(os.path.join(d, "gpu2.sam")) == want
    # ... and with the reference's pairing out of the loop altogether: plan -> GPU batch -> salt_pair_apply -> query_t
    err3 = _run([os.path.join(REFDIR, "salt_dropin")] + flags + ["idx", "r1.fq", "r2.fq"], d, os.path.join(d, "gpu3.sam"),
                env=dict(seed_env, SALT_DROPIN_PLAN="2"))
    m3 = re.search(r"pairs finished by salt_pair_apply alone: (\d+), handed back to the reference's pairing: (\d+)", err3)
    assert m3 and int(m3.group(1)) >= 2900 and int(m3.group(2)) <= 30, err3[-600:]
    assert _sam_body(os.path.join(d, "gpu3.sam")) == want
    body = [ln.split("\t") for ln in want if ln and not ln.startswith("@")]
    assert sum(1 for f in body if int(f[1]) & 2) >= 4000                  # properly paired records
    assert sum(1 for f in body if "S" in f[5]) >= 20                      # soft-clipped = rescued by Smith-Waterman


def test_mixref_builder_matches_salt_idx(tmp_path, oracle):
    """row A: the SNP-aware reference built on the device from FASTA bases + SNP rows equals the PREFIX.ref file
    the reference's own indexer (salt-idx -> build_mixRef, Index_src/mixRef.c:93) writes"""
    import numpy as np
    from salt_b200 import api
    if not _have():
        pytest.skip("oracle/_ref programs not built (reference tree absent at build time)")
    d = str(tmp_path)
    fa, sn, fq = dropin_data.write_inputs(d, n_reads=10)
    _run([os.path.join(REFDIR, "salt-idx"), "-k", "19", "ref.fa", "snps.txt", "idx"], d, os.path.join(d, "idx.log"))
    raw = np.fromfile(os.path.join(d, "idx.ref"), np.uint32)
    l, words = int(raw[0]), raw[1:].copy()
    recs, name = [], None
    for ln in open(fa):
        ln = ln.strip()
        if ln.startswith(">"):
            name = ln[1:]; recs.append([name, []])
        else:
            recs[-1][1].append(ln)
    recs = [(n, "".join(p)) for n, p in recs]
    offs = {}
    tot = 0
    for n, sq in recs:
        offs[n] = tot; tot += len(sq)
    assert tot == l
    pos, mask = [], []
    for ln in open(sn):
        c, p1, al, rf = ln.rstrip("\n").split("\t")
        pos.append(offs[c] + int(p1) - 1); mask.append(oracle.lib.orc_allele_mask(al.encode()))
    eng = api.Engine.from_bases("".join(sq for _, sq in recs), np.array(pos, np.uint32), np.array(mask, np.uint8), device=0)
    got = eng.get_mixref()
    if l % 8:                                  # nibbles past l in the file's last word are uninitialised heap (realloc, mixRef.c:135)
        words[-1] &= np.uint32((1 << (4 * (l % 8))) - 1)
    assert len(got) == len(words) and np.array_equal(got, words)
    assert (np.bitwise_count(words) > 8).sum() > 100 if hasattr(np, "bitwise_count") else True      # SNP sites carry extra allele bits


C0 = os.path.join(REFDIR, "config0")


def _config0_index(d):
    """BASELINE configs[0] as shipped (Test/Run_test/run_test.sh): the bundled two-copy lambda genome, reads from the
    bundled wgsim with -S 11, the SNP table from the script's own awk line -- staged by oracle/Makefile under
    oracle/_ref/config0 together with the reference program's SAM (-t 4).  The index is built here (run_test.sh:32)."""
    if not (_have() and os.path.exists(os.path.join(C0, "pe.sam"))):
        pytest.skip("oracle/_ref/config0 not staged (reference tree absent at build time)")
    _run([os.path.join(REFDIR, "salt-idx"), "-k", "19", os.path.join(C0, "Genome.fa"), os.path.join(C0, "hapmap.txt"), "idx"],
         d, os.path.join(d, "idx.log"))


def _same_sam(want_path, got_path, min_lines):
    want, got = _sam_body(want_path), _sam_body(got_path)
    assert len(want) == len(got) and len(want) >= min_lines, (len(want), len(got))
    for i, (a, b) in enumerate(zip(want, got)):
        assert a == b, (i, a, b)
    return [ln.split("\t") for ln in want if ln and not ln.startswith("@")]


@pytest.mark.parametrize("threads,seed", [("4", "host"), ("1", "host"), ("4", "gpu"), ("4", "gpu+sam")])
def test_config0_se_sam_identical(tmp_path, threads, seed):
    """run_se_test.sh:12 flags on the bundled genome: every read ties between the two lambda copies (strand-1-wins and
    first-hit rules decide the primary), 5 % mutated reads against their own SNP table"""
    d = str(tmp_path)
    _config0_index(d)
    flags = ["-d", "-r", "1", "-l", "100", "-n", "20", "-c", "-m", "500", "-t", threads]
    native = seed.endswith("+sam"); seed = seed.split("+")[0]
    err = _run([os.path.join(REFDIR, "salt_dropin")] + flags + ["idx", os.path.join(C0, "Read1.fq")], d, os.path.join(d, "gpu.sam"),
               env={"SALT_DROPIN_SEED": seed, "SALT_DROPIN_SAM": "native" if native else "reference"})
    assert ("SAM lines by salt_sam_se" in err) == native
    assert "verification on libsalt_b200" in err
    assert ("seeding + locate on libsalt_b200" in err) == (seed == "gpu")
    body = _same_sam(os.path.join(C0, "se.sam"), os.path.join(d, "gpu.sam"), 20000)
    assert sum(1 for f in body if any(x.startswith("XA:") for x in f[11:])) >= 100


@pytest.mark.parametrize("threads,seed", [("4", "host"), ("4", "gpu")])
def test_config0_pe_sam_identical(tmp_path, threads, seed):
    """run_pe_test.sh:14 flags (what run_test.sh actually runs, :37); with seed == "gpu" the paired-end program's seeding +
    locate (alnse_seed_overlap + alnse_locate) run on the device as well"""
    d = str(tmp_path)
    _config0_index(d)
    flags = ["-d", "-p", "-e", "-l", "100", "-c", "-a", "350", "-b", "650", "-r", "5", "-t", threads]
    err = _run([os.path.join(REFDIR, "salt_dropin")] + flags + ["idx", os.path.join(C0, "Read1.fq"), os.path.join(C0, "Read2.fq")],
               d, os.path.join(d, "gpu.sam"), env={"SALT_DROPIN_PLAN": "2", "SALT_DROPIN_SEED": seed})
    assert "verification on libsalt_b200" in err
    if seed == "gpu":
        import re
        m = re.search(r"seeded on the GPU: (\d+) mates, (\d+) of them handed", err)
        assert m and int(m.group(1)) >= 39000 and int(m.group(2)) <= int(m.group(1)) // 10, err[-600:]
    body = _same_sam(os.path.join(C0, "pe.sam"), os.path.join(d, "gpu.sam"), 40000)
    assert sum(1 for f in body if int(f[1]) & 2) >= 30000


@pytest.mark.parametrize("mode", ["se-t1", "se-t4", "pe-t2"])
def test_level0_program_sam_identical(tmp_path, mode):
    """Level 0: the reference's UNMODIFIED objects minus editdistance.o / LandauVishkin.o / ssw.o, linked
    against libsalt_level0.so (oracle/_ref/salt_level0) -- every ed_mismatch / ed_diff / ed_diff_withcigar / ssw_* call of the
    reference's own loops is a batch-of-one GPU call, from several worker threads at once -- prints the reference's SAM"""
    if not (_have() and os.path.exists(os.path.join(REFDIR, "salt_level0"))):
        pytest.skip("oracle/_ref programs not built (reference tree absent at build time)")
    d = str(tmp_path)
    if mode.startswith("se"):
        dropin_data.write_inputs(d, glen=60_000, n_reads=500)
        files = ["reads.fq"]
        flags = ["-d", "-r", "5", "-l", "100", "-n", "20", "-c", "-m", "500", "-t", mode[-1]]
    else:
        dropin_data.write_pe_inputs(d, glen=60_000, n_pairs=250)
        files = ["r1.fq", "r2.fq"]
        flags = ["-d", "-p", "-e", "-l", "100", "-c", "-a", "350", "-b", "650", "-r", "5", "-t", mode[-1]]
    _run([os.path.join(REFDIR, "salt-idx"), "-k", "19", "ref.fa", "snps.txt", "idx"], d, os.path.join(d, "idx.log"))
    _run([os.path.join(REFDIR, "salt")] + flags + ["idx"] + files, d, os.path.join(d, "ref.sam"))
    _run([os.path.join(REFDIR, "salt_level0")] + flags + ["idx"] + files, d, os.path.join(d, "l0.sam"))
    body = _same_sam(os.path.join(d, "ref.sam"), os.path.join(d, "l0.sam"), 500)
    assert sum(1 for f in body if f[1] != "4") >= 400
    if mode.startswith("pe"):
        assert sum(1 for f in body if "S" in f[5]) >= 3          # rescued by ssw_align on the GPU
