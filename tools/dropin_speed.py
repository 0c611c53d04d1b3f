#!/usr/bin/env python
"""End-to-end wall time of the reference `salt` and of `salt_dropin` (same sources, verification stage and
paired-end rescues on the GPU) on one synthetic data set, single-end and paired-end, with the SAM
comparison.  Both programs seed on ONE host thread here (-t 1; the drop-in's driver is single-threaded),
so the difference is what the GPU removes from the per-read critical path; seeding itself (>90 % of the
reference's time) is untouched.  Prints one JSON object."""
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import dropin_data  # noqa: E402

REFDIR = os.path.join(ROOT, "oracle", "_ref")


def run(cmd, cwd, out):
    t0 = time.time()
    with open(out, "w") as f:
        p = subprocess.run(cmd, cwd=cwd, stdout=f, stderr=subprocess.PIPE, text=True)
    assert p.returncode == 0, p.stderr[-1500:]
    return time.time() - t0, p.stderr


def body(path):
    return [ln for ln in open(path).read().split("\n") if not ln.startswith("@PG")]


def main():
    glen = int(os.environ.get("GENOME", "5000000")); n = int(os.environ.get("READS", "200000"))
    res = {"genome_bp": glen, "reads": n}
    with tempfile.TemporaryDirectory() as d:
        dropin_data.write_inputs(d, glen=glen, n_reads=n)
        t, _ = run([os.path.join(REFDIR, "salt-idx"), "-k", "19", "ref.fa", "snps.txt", "idx"], d, os.path.join(d, "idx.log"))
        res["index_s"] = round(t, 2)
        flags = ["-d", "-r", "1", "-l", "100", "-n", "20", "-c", "-m", "500", "-t", "1"]
        t_ref, _ = run([os.path.join(REFDIR, "salt")] + flags + ["idx", "reads.fq"], d, os.path.join(d, "ref.sam"))
        t_gpu, _ = run([os.path.join(REFDIR, "salt_dropin")] + flags + ["idx", "reads.fq"], d, os.path.join(d, "gpu.sam"))
        same = body(os.path.join(d, "ref.sam")) == body(os.path.join(d, "gpu.sam"))
        res["se"] = {"reference_s": round(t_ref, 2), "dropin_s": round(t_gpu, 2), "reference_reads_per_s": round((n + 40) / t_ref),
                     "dropin_reads_per_s": round((n + 40) / t_gpu), "sam_identical": same}
        npairs = n // 2
        dropin_data.write_pe_inputs(d, glen=glen, n_pairs=npairs)
        run([os.path.join(REFDIR, "salt-idx"), "-k", "19", "ref.fa", "snps.txt", "idx"], d, os.path.join(d, "idx.log"))
        flags = ["-d", "-p", "-e", "-l", "100", "-c", "-a", "350", "-b", "650", "-r", "5", "-t", "1"]
        t_ref, _ = run([os.path.join(REFDIR, "salt")] + flags + ["idx", "r1.fq", "r2.fq"], d, os.path.join(d, "ref.sam"))
        t_gpu, err = run([os.path.join(REFDIR, "salt_dropin")] + flags + ["idx", "r1.fq", "r2.fq"], d, os.path.join(d, "gpu.sam"))
        same = body(os.path.join(d, "ref.sam")) == body(os.path.join(d, "gpu.sam"))
        res["pe"] = {"reference_s": round(t_ref, 2), "dropin_s": round(t_gpu, 2), "reference_reads_per_s": round(2 * npairs / t_ref),
                     "dropin_reads_per_s": round(2 * npairs / t_gpu), "sam_identical": same,
                     "rescue": [ln for ln in err.split("\n") if "rescue windows" in ln][-1:]}
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
