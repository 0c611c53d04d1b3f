#!/usr/bin/env python
"""Where salt_dropin's host time goes as -t grows (single-end, default seed spacing, seeding on the device): the program's own
wall-time breakdown line at -t 1, 2, 4, 8, 16, and at all threads with glibc's allocator told not to trim / to grow its
arenas in large steps (the SAM text phase is many small reallocs on many threads).  Prints one JSON object."""
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


def main():
    glen = int(os.environ.get("GENOME", "20000000")); n = int(os.environ.get("READS", "1000000"))
    res = {"genome_bp": glen, "reads": n, "rows": []}
    with tempfile.TemporaryDirectory() as d:
        dropin_data.write_inputs(d, glen=glen, n_reads=n)
        subprocess.run([os.path.join(REFDIR, "salt-idx"), "-k", "19", "ref.fa", "snps.txt", "idx"], cwd=d, check=True,
                       stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        flags = ["-d", "-l", "100", "-n", "20", "-c", "-m", "500"]
        cases = [(t, {}) for t in (1, 2, 4, 8, os.cpu_count() or 16)]
        cases.append((os.cpu_count() or 16, {"GLIBC_TUNABLES": "glibc.malloc.top_pad=67108864:glibc.malloc.trim_threshold=4294967295"}))
        cases.append((os.cpu_count() or 16, {"MALLOC_ARENA_MAX": "64"}))
        for prog in ("salt_dropin", "salt"):
            for t, extra in cases:
                env = dict(os.environ, SALT_DROPIN_SEED="gpu", **extra)
                t0 = time.time()
                with open(os.path.join(d, "out.sam"), "w") as f:
                    p = subprocess.run([os.path.join(REFDIR, prog)] + flags + ["-t", str(t), "idx", "reads.fq"], cwd=d, stdout=f,
                                       stderr=subprocess.PIPE, text=True, env=env)
                dt = time.time() - t0
                res["rows"].append({"program": prog, "t": t, "env": extra, "wall_s": round(dt, 2), "rc": p.returncode,
                                    "breakdown": [ln for ln in p.stderr.split("\n") if ln.startswith("[salt_dropin] ")][-3:-1]})
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
