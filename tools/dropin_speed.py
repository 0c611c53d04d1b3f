#!/usr/bin/env python
"""End-to-end wall time of the reference `salt` and of `salt_dropin` (same sources, verification stage and
paired-end rescues on the GPU) on one synthetic data set, single-end and paired-end, with the SAM
comparison, at -t 1 and at all host threads, with the drop-in seeding on the host (the reference's own functions on -t
workers) or on the device (SALT_DROPIN_SEED=gpu).  Prints one JSON object."""
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
    """SE with the flags of run_se_test.sh:12 (-r 1: a seed at every read position) and with the default seed spacing, at -t 1 and
    at all host threads: the reference, the drop-in seeding on the host (-t workers), the drop-in seeding on the device."""
    glen = int(os.environ.get("GENOME", "5000000")); n = int(os.environ.get("READS", "200000"))
    threads = os.cpu_count() or 1
    tlist = [threads] if os.environ.get("QUICK") else sorted({1, threads})      # QUICK=1: all host threads only
    # The boxes of this pool run without the driver's persistence mode: a process that starts while no other client holds the
    # GPU pays 2.5-3 s of device initialisation (measured: "GPU init + uploads" 0.4 s right after another CUDA process, 2.4-3.1 s
    # after a few idle seconds).  Holding one context open here for the whole run is what nvidia-persistenced does on a
    # production host; HOLD_CONTEXT=0 switches it off.  The reference program does not touch the GPU either way.
    held = None
    if os.environ.get("HOLD_CONTEXT", "1") != "0":
        try:
            import torch
            held = torch.zeros(1, device="cuda")
        except Exception:                               # noqa: BLE001
            held = None
    res = {"genome_bp": glen, "reads": n, "host_threads": threads, "cuda_context_held_open_by_this_script": held is not None}
    with tempfile.TemporaryDirectory() as d:
        dropin_data.write_inputs(d, glen=glen, n_reads=n)
        t, _ = run([os.path.join(REFDIR, "salt-idx"), "-k", "19", "ref.fa", "snps.txt", "idx"], d, os.path.join(d, "idx.log"))
        res["index_s"] = round(t, 2)
        res["se"] = []
        for name, base in (("run_se_test.sh:12 (-r 1)", ["-d", "-r", "1", "-l", "100", "-n", "20", "-c", "-m", "500"]),
                           ("default seed spacing", ["-d", "-l", "100", "-n", "20", "-c", "-m", "500"])):
            for t_ in tlist:
                flags = base + ["-t", str(t_)]
                row = {"flags": " ".join(flags), "what": name}
                t_ref, _ = run([os.path.join(REFDIR, "salt")] + flags + ["idx", "reads.fq"], d, os.path.join(d, "ref.sam"))
                row["reference_s"] = round(t_ref, 2); row["reference_reads_per_s"] = round((n + 40) / t_ref)
                want = body(os.path.join(d, "ref.sam"))
                for mode in ("host", "gpu", "gpu+sam"):          # +sam: SAM lines by the host layer's salt_sam_se instead of aln_samse
                    env = dict(os.environ, SALT_DROPIN_SEED=mode.split("+")[0], SALT_DROPIN_SAM="native" if mode.endswith("+sam") else "reference")
                    t0 = time.time()
                    with open(os.path.join(d, "gpu.sam"), "w") as f:
                        p = subprocess.run([os.path.join(REFDIR, "salt_dropin")] + flags + ["idx", "reads.fq"], cwd=d, stdout=f,
                                           stderr=subprocess.PIPE, text=True, env=env)
                    assert p.returncode == 0, p.stderr[-1500:]
                    dt = time.time() - t0
                    key = "dropin_%s_seeding" % mode.replace("+", "_")
                    row[key + "_s"] = round(dt, 2); row[key + "_reads_per_s"] = round((n + 40) / dt)
                    row[key + "_sam_identical"] = body(os.path.join(d, "gpu.sam")) == want
                    row[key + "_phases"] = [ln for ln in p.stderr.split("\n") if ln.startswith("[salt_dropin] ")][-2:]
                res["se"].append(row)
        if os.environ.get("SKIP_PE"):
            print(json.dumps(res, indent=1))
            return
        npairs = n // 2
        dropin_data.write_pe_inputs(d, glen=glen, n_pairs=npairs)
        run([os.path.join(REFDIR, "salt-idx"), "-k", "19", "ref.fa", "snps.txt", "idx"], d, os.path.join(d, "idx.log"))
        res["pe"] = []
        for t_ in tlist:
            flags = ["-d", "-p", "-e", "-l", "100", "-c", "-a", "350", "-b", "650", "-r", "5", "-t", str(t_)]
            t_ref, _ = run([os.path.join(REFDIR, "salt")] + flags + ["idx", "r1.fq", "r2.fq"], d, os.path.join(d, "ref.sam"))
            want = body(os.path.join(d, "ref.sam"))
            row = {"flags": " ".join(flags), "reference_s": round(t_ref, 2), "reference_reads_per_s": round(2 * npairs / t_ref)}
            for mode in ("host", "gpu"):
                env = dict(os.environ, SALT_DROPIN_SEED=mode)
                t0 = time.time()
                with open(os.path.join(d, "gpu.sam"), "w") as f:
                    p = subprocess.run([os.path.join(REFDIR, "salt_dropin")] + flags + ["idx", "r1.fq", "r2.fq"], cwd=d, stdout=f,
                                       stderr=subprocess.PIPE, text=True, env=env)
                assert p.returncode == 0, p.stderr[-1500:]
                dt = time.time() - t0
                key = "dropin_%s_seeding" % mode
                row[key + "_s"] = round(dt, 2); row[key + "_reads_per_s"] = round(2 * npairs / dt)
                row[key + "_sam_identical"] = body(os.path.join(d, "gpu.sam")) == want
                row[key + "_phases"] = [ln for ln in p.stderr.split("\n") if ln.startswith("[salt_dropin")][-3:]
            res["pe"].append(row)
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
