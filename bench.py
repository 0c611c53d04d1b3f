#!/usr/bin/env python
"""bench.py -- throughput of salt's SNP-aware verification hot path on B200.

Workload (BASELINE.json configs[1]): synthetic 50 Mbp genome + 1 % synthetic SNPs, 2 M single-end
100 bp reads per GPU, each with --cands candidate loci per strand (true locus + decoys).  One
"step" = one pass of the whole verification stage over the 2 M-read batch:
    ed_mismatch on every candidate fused with the acceptance scan -> Landau-Vishkin on the
    candidates of unmatched reads -> acceptance scan -> CIGARs for gapped primaries.

  value : reads/s with reads, candidate lists and outputs resident in HBM (CUDA events on the
          stream the kernels run on, max over ranks)
  e2e   : the same pass through the C ABI from pinned HOST buffers (salt_b200_set_reads +
          salt_b200_verify), H2D and D2H inside the timed region
  roofline / kernels : per-kernel device time from the library's own CUDA events
  sw / lv           : kernel-level sweeps of the mate-rescue Smith-Waterman and Landau-Vishkin
  cpu_baseline      : the reference's own C functions (oracle/_ref) on the host cores, bounded sample

`--impl reference` times only that CPU path (rank 0), K steps of the bounded sample.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "reads/s (SNP-aware verify stage: ed_mismatch + Landau-Vishkin + CIGAR)"
UNIT = "reads/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--reads", type=int, default=2_000_000)
    ap.add_argument("--genome", type=int, default=50_000_000)
    ap.add_argument("--read-len", type=int, default=100)
    ap.add_argument("--cands", type=int, default=8, help="candidate loci per read per strand")
    ap.add_argument("--snp-rate", type=float, default=0.01)
    ap.add_argument("--cpu-sample", type=int, default=400_000, help="reads in the bounded CPU-baseline sample")
    ap.add_argument("--sw-tasks", type=int, default=200_000)
    ap.add_argument("--skip-extras", action="store_true", help="skip the sw/lv kernel sweeps")
    ap.add_argument("--chunk", type=int, default=100_000, help="reads per pipeline chunk in the e2e leg (N_SEQS, aln.h:27)")
    ap.add_argument("--no-traffic-probe", action="store_true", help="skip the ncu subprocess that measures the dominant kernel's DRAM bytes")
    ap.add_argument("--traffic-probe-child", action="store_true", help=argparse.SUPPRESS)
    ap.add_argument("--probe-dir", default="", help=argparse.SUPPRESS)
    ap.add_argument("--pe-pairs", type=int, default=200_000, help="pairs in the paired-end pipeline leg (0 = skip)")
    ap.add_argument("--seed-reads", type=int, default=400_000, help="reads in the seeding + locate leg (0 = skip)")
    ap.add_argument("--seed-genome", type=int, default=10_000_000, help="genome of the seeding leg (its index is built by salt-idx)")
    ap.add_argument("--program-reads", type=int, default=400_000,
                    help="reads of the whole-program leg: salt_b200/salt_aln against the reference program, SAM compared (0 = skip)")
    ap.add_argument("--program-genome", type=int, default=5_000_000, help="genome of the whole-program leg (its index is built by salt-idx)")
    return ap.parse_args()


def make_workload(args, seed):
    from salt_b200 import synth
    t0 = time.time()
    if args.genome > 500_000_000:
        # configs[2]-sized reference: block-wise generator, the same genome on every rank (the reference is replicated),
        # reads differ per rank
        g = synth.BigGenome(args.genome, snp_rate=args.snp_rate, seed=11, n_records=24)
    else:
        g = synth.Genome(args.genome, snp_rate=args.snp_rate, seed=seed)
    n, L = args.reads, args.read_len
    reads = np.empty((n, L), np.uint8); pos = np.empty(n, np.uint32); strand = np.empty(n, np.uint8)
    chunk = 250_000
    for i in range(0, n, chunk):
        m = min(chunk, n - i)
        r, p, s = synth.sample_reads(g, m, L, seed=seed * 1000 + i // chunk, sub_rate=0.01, indel_frac=0.02)
        reads[i:i + m], pos[i:i + m], strand[i:i + m] = r, p, s
    offs0, loci0, offs1, loci1 = synth.make_candidates(g, pos, strand, L, per_strand=args.cands, seed=seed + 7)
    return dict(g=g, reads=reads, pos=pos, strand=strand, offs0=offs0, loci0=loci0, offs1=offs1, loci1=loci1,
                gen_s=time.time() - t0)


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region.  NVML in-process (a query takes ~0.1 ms, the
    timed region of the default run only a few ms), `nvidia-smi -lms` as the fallback."""

    REASONS = ((0x8, "hw_slowdown"), (0x40, "hw_thermal_slowdown"), (0x20, "sw_thermal_slowdown"), (0x4, "sw_power_cap"))

    def __init__(self, dev):
        self.dev = dev; self.rows = []; self.p = None; self.h = None; self.stop_flag = False; self.t = None

    def _nvml_handle(self):
        import pynvml
        pynvml.nvmlInit()
        try:
            import torch
            pr = torch.cuda.get_device_properties(self.dev)
            bus = "%08x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
            return pynvml, pynvml.nvmlDeviceGetHandleByPciBusId(bus.encode())
        except Exception:                                  # noqa: BLE001 -- older torch: fall back to the index
            return pynvml, pynvml.nvmlDeviceGetHandleByIndex(self.dev)

    def _sample_nvml(self):
        nv, h = self.nv, self.h
        try:
            reasons = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
        except Exception:                                  # noqa: BLE001
            reasons = 0
        self.rows.append((float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)), int(reasons)))

    def _loop(self):
        while not self.stop_flag:
            self._sample_nvml()
            time.sleep(0.0005)

    def start(self):
        try:
            self.nv, self.h = self._nvml_handle()
            self.max_mhz = float(self.nv.nvmlDeviceGetMaxClockInfo(self.h, self.nv.NVML_CLOCK_SM))
            self._sample_nvml()
            self.t = threading.Thread(target=self._loop, daemon=True); self.t.start()
            return
        except Exception:                                  # noqa: BLE001
            self.h = None
        q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(self.dev), "--query-gpu=" + q, "--format=csv,noheader,nounits",
                                       "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True); self.t.start()
        except Exception:
            self.p = None

    def _read(self):
        for line in self.p.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.h is not None:
            self.stop_flag = True
            self.t.join(timeout=1.0)
            self._sample_nvml()
            sm = [r[0] for r in self.rows]
            bits = 0
            for r in self.rows:
                bits |= r[1]
            busy = sorted(sm)[len(sm) // 2:]
            return {"sm_mhz": float(np.median(busy)), "sm_max_mhz": self.max_mhz, "samples": len(sm),
                    "reasons": [n for b, n in self.REASONS if bits & b], "source": "nvml"}
        if not self.p:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.p.terminate()
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 3 + i and r[3 + i].lower().startswith("active") for r in self.rows)]
        busy = sorted(sm)[len(sm) // 2:] if sm else []
        return {"sm_mhz": float(np.median(busy)) if busy else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": reasons, "source": "nvidia-smi"}


def cpu_reference_run(args, wl, n_sample, steps, warmup):
    """The reference's verification loop (oracle/_ref when present, else the port) on all host threads."""
    from oracle import orc
    o = orc.Oracle()
    ref = orc.Ref() if orc.ref_available() else None
    kind = "reference" if ref is not None else "port"
    cores = os.cpu_count() or 1
    n = min(n_sample, len(wl["reads"]))
    L = args.read_len
    codes = np.ascontiguousarray(wl["reads"][:n]).reshape(-1)
    roffs = (np.arange(n + 1, dtype=np.uint64) * L).astype(np.uint32)
    o0 = wl["offs0"][:n + 1].copy(); o1 = wl["offs1"][:n + 1].copy()
    l0 = wl["loci0"][:o0[-1]]; l1 = wl["loci1"][:o1[-1]]
    g = wl["g"]
    times = []
    for it in range(warmup + steps):
        sec, recs, a0, a1, cig = o.verify_batch(g.mixref, g.l, codes, roffs, o0, l0, o1, l1, 3, -1, n_threads=cores, ref=ref)
        if it >= warmup:
            times.append(sec)
    pairs = int(o0[-1]) + int(o1[-1])
    return dict(kind=kind, cores=cores, n=n, pairs=pairs, sec=float(np.mean(times)), times=times, recs=recs, acc0=a0, acc1=a1,
                sample="first %d of %d reads of the same workload (%d candidate pairs), verify stage with the reference's "
                       "running thresholds, %d host threads, gcc -O2" % (n, len(wl["reads"]), pairs, cores))


def traffic_probe_child(path):
    """Child of traffic_probe(): the device-resident verification stage on the workload the parent saved, three
    times, so that ncu (which wraps this process) can capture one warm launch of the dominant kernel."""
    import torch
    from salt_b200 import api
    z = np.load(os.path.join(path, "wl.npz"))
    mixref = np.load(os.path.join(path, "mixref.npy"), mmap_mode="r")
    eng = api.Engine(mixref, int(z["l"]), None, 0, device=0)
    n, L = z["reads"].shape
    eng.set_reads(z["reads"])
    dev = torch.device("cuda", 0)
    d = [torch.from_numpy(z[k]).to(dev) for k in ("offs0", "loci0", "offs1", "loci1")]
    n0, n1 = len(z["loci0"]), len(z["loci1"])
    d_rec = torch.empty(n * 16, dtype=torch.uint8, device=dev); d_acc = torch.empty(n0 + n1 + 16, dtype=torch.int8, device=dev)
    for _ in range(3):
        rc = eng.L.salt_b200_verify_dev(eng.h, d[0].data_ptr(), d[1].data_ptr(), n0, d[2].data_ptr(), d[3].data_ptr(), n1, 3, -1,
                                        d_rec.data_ptr(), d_acc.data_ptr(), d_acc.data_ptr() + n0, None, 0, None, None)
        assert rc == 0
    eng.sync()
    eng.close()


def traffic_probe(wl, timeout_s=420):
    """dram__bytes_read.sum + dram__bytes_write.sum of ONE warm nogap_fused launch on this run's workload, measured by
    an ncu subprocess wrapped around traffic_probe_child (ncu replays the kernel; the number never touches a timing).
    Returns (bytes, detail dict) or (None, reason)."""
    import shutil
    import tempfile
    ncu = shutil.which("ncu") or "/usr/local/cuda/bin/ncu"
    if not os.path.exists(ncu):
        return None, {"error": "ncu not found"}
    base = "/dev/shm" if os.path.isdir("/dev/shm") else None
    path = tempfile.mkdtemp(prefix="salt_bench_", dir=base)
    try:
        np.savez(os.path.join(path, "wl.npz"), reads=wl["reads"], offs0=wl["offs0"], loci0=wl["loci0"], offs1=wl["offs1"],
                 loci1=wl["loci1"], l=np.int64(wl["g"].l))
        np.save(os.path.join(path, "mixref.npy"), wl["g"].mixref)
        csv = os.path.join(path, "ncu.csv")
        metrics = "dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct," \
                  "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed,smsp__issue_active.avg.pct_of_peak_sustained_elapsed"
        cmd = [ncu, "--metrics", metrics, "--clock-control", "none", "-k", "regex:nogap_fused", "--launch-skip", "2",
               "--launch-count", "1", "--csv", "--log-file", csv, sys.executable, os.path.abspath(__file__),
               "--traffic-probe-child", "--probe-dir", path]
        env = dict(os.environ)
        for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK", "MASTER_ADDR", "MASTER_PORT"):
            env.pop(k, None)
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=timeout_s, env=env)
        if r.returncode != 0 or not os.path.exists(csv):
            return None, {"error": "ncu exit %d: %s" % (r.returncode, r.stdout[-300:])}
        vals = {}
        import csv as _csv
        rows = [row for row in _csv.reader(open(csv)) if len(row) > 5]
        hdr = next(i for i, row in enumerate(rows) if "Metric Name" in row)
        ci = {name: rows[hdr].index(name) for name in ("Metric Name", "Metric Unit", "Metric Value")}
        for row in rows[hdr + 1:]:
            v = float(row[ci["Metric Value"]].replace(",", ""))
            unit = row[ci["Metric Unit"]].lower()
            scale = {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "nsecond": 1e-9, "usecond": 1e-6, "msecond": 1e-3,
                     "ns": 1e-9, "us": 1e-6, "ms": 1e-3, "second": 1}.get(unit, 1)
            vals[row[ci["Metric Name"]]] = v * scale
        tot = vals.get("dram__bytes_read.sum", 0.0) + vals.get("dram__bytes_write.sum", 0.0)
        return (tot if tot > 0 else None), {"dram_bytes_read": vals.get("dram__bytes_read.sum"), "dram_bytes_write": vals.get("dram__bytes_write.sum"),
                                            "ncu_kernel_s": vals.get("gpu__time_duration.sum"),
                                            "l2_hit_pct": vals.get("lts__t_sector_hit_rate.pct"),
                                            "l1_lsu_wavefront_pct": vals.get("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed"),
                                            "issue_active_pct": vals.get("smsp__issue_active.avg.pct_of_peak_sustained_elapsed"),
                                            "how": "ncu subprocess on this run's workload, launch 3 of nogap_fused (warm), --clock-control none"}
    except Exception as ex:                                # noqa: BLE001 -- the probe is evidence, not the measurement
        return None, {"error": repr(ex)}
    finally:
        shutil.rmtree(path, ignore_errors=True)


_REAL_STDOUT = None


def _claim_stdout():
    """Everything libraries print to fd 1 (NCCL's version banner, for one) goes to stderr; the one JSON
    line is written to the real stdout by _emit()."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def _emit(obj):
    _REAL_STDOUT.write(json.dumps(obj) + "\n")
    _REAL_STDOUT.flush()


def _bind_to_gpu_numa_node(local):
    """Several ranks share the host: run this one (and first-touch its pinned buffers) on the NUMA node its
    GPU hangs off, so that every rank's H2D/D2H traffic stays on its own socket.  Best effort."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local)
        bus = pynvml.nvmlDeviceGetPciInfo(h).busId
        bus = (bus.decode() if isinstance(bus, bytes) else bus).lower()
        if len(bus.split(":")[0]) == 8:
            bus = bus[4:]                                  # sysfs uses a 4-digit PCI domain
        node = int(open("/sys/bus/pci/devices/%s/numa_node" % bus).read())
        if node < 0:
            return "unknown"
        cpus = set()
        for part in open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return "node%d (%d cpus)" % (node, len(cpus))
    except Exception as e:                                 # noqa: BLE001 -- affinity is an optimisation only
        return "unbound (%s)" % type(e).__name__
    return "unbound"


def main():
    args = parse()
    if args.traffic_probe_child:
        return traffic_probe_child(args.probe_dir)
    _claim_stdout()
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    which = "configs[1]" if args.genome <= 500_000_000 else "configs[2]-sized reference (24 records, snp144-like density), verify stage"
    cfg = {"workload": "%s: synthetic %d Mbp genome + %.2f%% SNPs, %d x %d bp SE reads per GPU, %d candidates/read/strand "
                       "(true locus + decoys), sub 1%%, 2%% reads with an indel" % (which, args.genome // 1_000_000, args.snp_rate * 100,
                                                                                  args.reads, args.read_len, args.cands),
           "nogap_T0": 3, "lv_T0": "l_seq/10", "parallelism": "reads sharded per GPU, reference replicated, no collective on the data path",
           "l2": "per-step inputs+outputs (~0.7 GB) exceed the 126 MB L2"}

    if args.impl == "reference":
        if rank != 0:
            return
        wl = make_workload(args, seed=11)
        r = cpu_reference_run(args, wl, args.cpu_sample, args.steps, args.warmup)
        val = r["n"] / r["sec"]
        _emit(({"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
                          "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["sec"] * 1e3,
                          "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
                          "config": cfg,
                          "cpu_baseline": {"value": val, "unit": UNIT, "cores": r["cores"], "kind": r["kind"], "sample": r["sample"]},
                          "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return

    import torch
    import torch.distributed as dist
    from salt_b200 import api
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the engine has no CPU path")
    torch.cuda.set_device(local)
    numa = None
    if world > 1:
        numa = _bind_to_gpu_numa_node(local)              # pinned host buffers next to this rank's GPU
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    wl = make_workload(args, seed=11 + rank)
    g = wl["g"]; n = args.reads; L = args.read_len
    eng = api.Engine(g.mixref, g.l, g.pac, g.l, device=local)
    # a real (non-default) stream: handle 0 would mean "library-owned stream" to salt_b200_set_stream
    stream = torch.cuda.Stream(dev)
    torch.cuda.set_stream(stream)
    assert stream.cuda_stream != 0
    eng.set_stream(stream.cuda_stream)
    lib, h = eng.L, eng.h

    # ---------------- pinned host buffers (e2e) and device-resident copies (value)
    def pin(a):
        t = torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
        return t
    h_codes = pin(wl["reads"].reshape(-1)); h_roffs = pin((np.arange(n + 1, dtype=np.uint64) * L).astype(np.uint32))
    h_o0, h_l0, h_o1, h_l1 = pin(wl["offs0"]), pin(wl["loci0"]), pin(wl["offs1"]), pin(wl["loci1"])
    n0, n1 = len(wl["loci0"]), len(wl["loci1"])
    h_rec = torch.empty(n * 16, dtype=torch.uint8).pin_memory()
    h_acc0 = torch.empty(n0, dtype=torch.int8).pin_memory(); h_acc1 = torch.empty(n1, dtype=torch.int8).pin_memory()
    h_cig = torch.zeros(n * 128, dtype=torch.uint8).pin_memory()
    reads_t = api.ReadsT(h_codes.data_ptr(), h_roffs.data_ptr(), n)
    cands = api.CandsT()
    cands.offs[0], cands.offs[1] = h_o0.data_ptr(), h_o1.data_ptr()
    cands.loci[0], cands.loci[1] = h_l0.data_ptr(), h_l1.data_ptr()

    def ck(rc):
        if rc != 0:
            raise RuntimeError(lib.salt_b200_last_error().decode())

    ck(lib.salt_b200_set_reads(h, C.byref(reads_t)))
    d_o0, d_l0, d_o1, d_l1 = (t.to(dev) for t in (h_o0, h_l0, h_o1, h_l1))
    d_rec = torch.empty(n * 16, dtype=torch.uint8, device=dev)
    d_acc = torch.empty(n0 + n1 + 16, dtype=torch.int8, device=dev)
    d_cig = torch.empty(n * 128, dtype=torch.uint8, device=dev)
    d_cigreads = torch.empty(n + 2, dtype=torch.int32, device=dev); d_cigcnt = torch.zeros(4, dtype=torch.int32, device=dev)

    def step_dev():
        ck(lib.salt_b200_verify_dev(h, d_o0.data_ptr(), d_l0.data_ptr(), n0, d_o1.data_ptr(), d_l1.data_ptr(), n1, 3, -1,
                                    d_rec.data_ptr(), d_acc.data_ptr(), d_acc.data_ptr() + n0,
                                    d_cig.data_ptr(), 128, d_cigreads.data_ptr(), d_cigcnt.data_ptr()))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # ---------------- value: device-resident pass
    for _ in range(args.warmup):
        step_dev()
    barrier()
    sampler = ClockSampler(local); sampler.start()
    eng.launch_count(reset=True)
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        step_dev()
    e1.record(stream)
    barrier()
    ms = e0.elapsed_time(e1)
    launches = eng.launch_count()
    clocks = sampler.stop()
    tmax = torch.tensor([ms], device=dev)
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    ms_step = float(tmax.item()) / args.steps
    value = world * n / (ms_step * 1e-3)

    # per-kernel times from the library's own events (separate, untimed-for-value pass)
    eng.profile(True)
    prof = {}
    for _ in range(3):
        step_dev()
        for k, v in eng.profile_read().items():
            prof.setdefault(k, []).append(v)
    eng.profile(False)
    kernels = {k: float(np.mean(v)) for k, v in prof.items()}
    rec_np = np.frombuffer(d_rec.cpu().numpy().tobytes(), api.VERIFY_DT)
    n_lv = int(rec_np["lv_ran"].sum())
    # candidate pairs of the reads that reached the gapped stage: each is one Landau-Vishkin problem
    lv_mask = rec_np["lv_ran"] == 1
    n_lv_pairs = int((np.diff(wl["offs0"].astype(np.int64))[lv_mask].sum() + np.diff(wl["offs1"].astype(np.int64))[lv_mask].sum()))
    n_gapped = int(d_cigcnt[0].item())

    # PCIe copy peaks beside the e2e number (pinned 256 MiB, best of 3), all ranks copying at the same time: the ranks
    # of one box share the host's memory / PCIe complex, so a rank-local peak would overstate what the pipeline can get
    def copy_peak(to_dev):
        nb = 256 << 20
        a = torch.empty(nb, dtype=torch.uint8).pin_memory(); b = torch.empty(nb, dtype=torch.uint8, device=dev)
        best = 0.0
        for _ in range(3):
            barrier()
            c0 = torch.cuda.Event(enable_timing=True); c1 = torch.cuda.Event(enable_timing=True)
            c0.record(stream)
            (b.copy_(a, non_blocking=True) if to_dev else a.copy_(b, non_blocking=True))
            c1.record(stream); torch.cuda.synchronize(dev)
            best = max(best, nb / (c0.elapsed_time(c1) * 1e-3) / 1e9)
        return best
    pcie_h2d, pcie_d2h = copy_peak(True), copy_peak(False)

    def copy_peak_bidir():
        """both directions at once (the pipeline uploads chunk c+1 while it downloads chunk c-1), all ranks together"""
        nb = 256 << 20
        a = torch.empty(nb, dtype=torch.uint8).pin_memory(); b = torch.empty(nb, dtype=torch.uint8, device=dev)
        a2 = torch.empty(nb, dtype=torch.uint8).pin_memory(); b2 = torch.empty(nb, dtype=torch.uint8, device=dev)
        s2 = torch.cuda.Stream(dev)
        best = (0.0, 0.0)
        for _ in range(3):
            barrier()
            u0 = torch.cuda.Event(enable_timing=True); u1 = torch.cuda.Event(enable_timing=True)
            d0 = torch.cuda.Event(enable_timing=True); d1 = torch.cuda.Event(enable_timing=True)
            u0.record(stream); b.copy_(a, non_blocking=True); u1.record(stream)
            with torch.cuda.stream(s2):
                d0.record(s2); a2.copy_(b2, non_blocking=True); d1.record(s2)
            torch.cuda.synchronize(dev)
            up = nb / (u0.elapsed_time(u1) * 1e-3) / 1e9; down = nb / (d0.elapsed_time(d1) * 1e-3) / 1e9
            if up + down > sum(best):
                best = (up, down)
        return best
    pcie_bi_h2d, pcie_bi_d2h = copy_peak_bidir()

    # ---------------- e2e: host buffers through the C ABI, the reference's chunk loop (alnse.c:1414-1440) through the
    # asynchronous slots.  Headline: the compact transport (2-bit bases + per-read counts, salt_packed_chunk_t);
    # beside it the plain format (one byte per base + three offset arrays) the reference's own structures map to.
    pk_bases, pk_npos = api.pack_bases(wl["reads"].reshape(-1), 2)
    h_bases = pin(pk_bases); h_npos = pin(pk_npos) if len(pk_npos) else None
    h_c0 = pin(np.diff(wl["offs0"].astype(np.int64)).astype(np.uint16)); h_c1 = pin(np.diff(wl["offs1"].astype(np.int64)).astype(np.uint16))
    pkc = api.PackedChunkT()
    pkc.n_reads = n; pkc.base_bits = 2; pkc.bases = h_bases.data_ptr(); pkc.base_start = 0; pkc.lens = None; pkc.l_seq = L
    pkc.n_pos = h_npos.data_ptr() if h_npos is not None else None; pkc.n_n = len(pk_npos); pkc.count_bits = 16
    pkc.n_cand[0], pkc.n_cand[1] = h_c0.data_ptr(), h_c1.data_ptr()
    pkc.loci[0], pkc.loci[1] = h_l0.data_ptr(), h_l1.data_ptr()

    def step_host_packed():
        ck(lib.salt_b200_verify_batch_packed(h, C.byref(pkc), args.chunk, 3, -1, h_rec.data_ptr(),
                                             h_acc0.data_ptr(), h_acc1.data_ptr(), h_cig.data_ptr(), 128))

    def step_host_plain():
        ck(lib.salt_b200_verify_batch(h, C.byref(reads_t), C.byref(cands), args.chunk, 3, -1, h_rec.data_ptr(),
                                      h_acc0.data_ptr(), h_acc1.data_ptr(), h_cig.data_ptr(), 128))

    def time_host(fn):
        for _ in range(max(1, args.warmup - 1)):
            fn()
        barrier()
        eng.launch_count(reset=True)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            fn()
        barrier()
        sec = (time.perf_counter() - t0) / args.steps
        te = torch.tensor([sec], device=dev)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        return float(te.item()), eng.launch_count() // args.steps

    plain_s, _ = time_host(step_host_plain)
    rec_plain = h_rec.numpy().copy()
    e2e_s, e2e_launches = time_host(step_host_packed)
    assert h_rec.numpy().tobytes() == rec_plain.tobytes(), "compact and plain transport disagree"
    acc_e2e = (h_acc0.numpy().copy(), h_acc1.numpy().copy())      # compared with the reference's own output below
    h2d_plain = h_codes.numel() + 4 * (h_roffs.numel() + h_o0.numel() + h_o1.numel() + n0 + n1)
    h2d = h_bases.numel() + 4 * len(pk_npos) + 2 * (h_c0.numel() + h_c1.numel()) + 4 * (n0 + n1)
    d2h = n * 16 + n0 + n1 + n_gapped * (128 + 4) + 4

    out = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
           "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8",
           "data": "synthetic", "config": cfg, "clocks": clocks, "gpu_launches": int(launches),
           "e2e": {"value": world * n / e2e_s, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                   "ms_per_step": e2e_s * 1e3, "chunk_reads": args.chunk, "slots": int(lib.salt_b200_n_slots()),
                   "gpu_launches_per_step": int(e2e_launches),
                   "transport": "salt_b200_verify_batch_packed: 2-bit bases, uint16 candidate counts, uint32 loci; pinned host buffers",
                   "h2d_bytes_per_read": h2d / n,
                   "pcie_gbs": {"h2d": h2d / e2e_s / 1e9, "d2h": d2h / e2e_s / 1e9, "h2d_copy_peak": pcie_h2d, "d2h_copy_peak": pcie_d2h,
                                "h2d_copy_peak_bidirectional": pcie_bi_h2d, "d2h_copy_peak_bidirectional": pcie_bi_d2h,
                                "copy_peak_how": "256 MiB pinned, best of 3, all %d ranks copying concurrently" % world},
                   "floor_ms": h2d / (pcie_h2d * 1e9) * 1e3,
                   "plain_format": {"value": world * n / plain_s, "ms_per_step": plain_s * 1e3, "h2d_bytes_per_step": int(h2d_plain),
                                    "transport": "salt_b200_verify_batch: one byte per base, three uint32 offset arrays"}},
           "pairs_per_step": int(n0 + n1), "pairs_per_s": world * (n0 + n1) / (ms_step * 1e-3),
           # SURVEY 8(d) whole-job figure: DP-cell equivalents of the step (L per ungapped pair, L*(L+4) per LV pair) / time
           "tcups_equivalent": world * ((n0 + n1) * L + n_lv_pairs * L * (L + 4)) / (ms_step * 1e-3) / 1e12,
           "lv_pairs_per_step": n_lv_pairs,
           "lv_reads_per_step": n_lv, "gapped_primaries_per_step": n_gapped, "kernels_ms": kernels, "numa": numa}

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm = float(peaks.get("hbm_gbs", 6650.0))
        # dominant kernel of the step; its algorithmic bytes per pair are stated in DESIGN.md §4
        dom = max(kernels, key=kernels.get) if kernels else "mismatch"
        bytes_per_pair = (L + 1) // 2 + 8 + 2          # window nibbles + pair descriptor + result (60 B at L=100)
        mm_ms = kernels.get("nogap_fused", float("nan"))
        ach = (n0 + n1) * bytes_per_pair / (mm_ms * 1e-3) / 1e9
        traffic, tdetail = (None, {"skipped": True})
        if not args.no_traffic_probe and world == 1:
            traffic, tdetail = traffic_probe(wl)
        out["roofline"] = {"kernel": "nogap_fused_kernel", "bound": "hbm", "achieved": ach, "peak": hbm, "unit": "GB/s",
                           "frac": ach / hbm, "traffic": traffic, "peak_source": "measured" if peaks else "fallback",
                           "algorithmic_bytes_per_pair": bytes_per_pair, "kernel_ms": mm_ms, "dominant_kernel_of_step": dom,
                           "traffic_probe": tdetail,
                           # what the DRAM actually moved in the kernel's own time (live event timing, not ncu's)
                           "dram_gbs": (traffic / (mm_ms * 1e-3) / 1e9) if traffic else None,
                           "dram_frac": (traffic / (mm_ms * 1e-3) / 1e9 / hbm) if traffic else None,
                           "note": "achieved/frac = ALGORITHMIC bytes / kernel time (HBM-equivalent); dram_gbs/dram_frac = "
                                   "measured DRAM bytes / kernel time.  When the reference fits the 126 MB L2 the kernel is "
                                   "L1/issue-bound and dram_frac is far below frac; at GRCh38 size the window gathers miss L2."}

    # ---------------- kernel-level extras: LV and SW sweeps (N=1 only)
    # the e2e leg left slot 0 holding its last chunk: the sweeps index the whole batch again
    ck(lib.salt_b200_set_reads(h, C.byref(reads_t)))
    if world == 1 and not args.skip_extras:
        out["lv"] = bench_lv(eng, lib, h, wl, args, dev, stream)
        out["sw"] = bench_sw(eng, lib, h, wl, args, dev, stream)
        out["sam_tail"] = bench_sam_tail(eng, lib, h, pkc, args, n, n0, n1, (h_rec, h_acc0, h_acc1, h_cig))
        if args.pe_pairs > 0:
            out["pe"] = bench_pe(eng, wl, args)
        if args.seed_reads > 0:
            out["seeding"] = bench_seeding(args)
        out["se_host_layer"] = bench_se_host_layer(eng, wl, args)
        if args.program_reads > 0:
            out["program"] = bench_program(args)

    if rank == 0 and world == 1:
        try:
            r = cpu_reference_run(args, wl, args.cpu_sample, 1, 0)
            out["cpu_baseline"] = {"value": r["n"] / r["sec"], "unit": UNIT, "cores": r["cores"], "kind": r["kind"],
                                   "sample": r["sample"], "pairs_per_s": r["pairs"] / r["sec"]}
            # the sample doubles as a bit-exact check of the e2e leg at this run's full size: primaries and per-candidate
            # results of the reference's own functions against what came back over the C ABI
            try:
                ns = r["n"]
                got = rec_plain.view(api.VERIFY_DT)[:ns]
                want = r["recs"]
                wpos = np.fromiter((w.pos for w in want), np.uint32, ns); wst = np.fromiter((w.strand for w in want), np.int64, ns)
                wnd = np.fromiter((w.n_diff for w in want), np.int64, ns); wgap = np.fromiter((w.is_gap for w in want), np.int64, ns)
                mapped = wpos != 0xFFFFFFFF
                same = (np.array_equal(got["pos"], wpos) and np.array_equal(got["strand"][mapped], wst[mapped])
                        and np.array_equal(got["n_diff"][mapped], wnd[mapped]) and np.array_equal(got["is_gap"][mapped], wgap[mapped]))
                o0e, o1e = int(wl["offs0"][ns]), int(wl["offs1"][ns])
                same_acc = np.array_equal(acc_e2e[0][:o0e], r["acc0"]) and np.array_equal(acc_e2e[1][:o1e], r["acc1"])
                out["cpu_baseline"]["gpu_e2e_identical_on_sample"] = bool(same and same_acc)
                out["cpu_baseline"]["sample_checked"] = "%d reads, %d candidate results, %d mapped, %d gapped" % (
                    ns, o0e + o1e, int(mapped.sum()), int((wgap[mapped] == 1).sum()))
            except Exception as ex:                      # noqa: BLE001
                out["cpu_baseline"]["gpu_e2e_identical_on_sample"] = "not compared: %r" % (ex,)
        except Exception as ex:   # the checker is optional for the bench line
            out["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 0, "kind": "port", "sample": "failed: %r" % (ex,)}
    if rank == 0:
        _emit(out)
    if world > 1:
        dist.destroy_process_group()


def bench_pe(eng, wl, args):
    from salt_b200 import api
    """Paired-end pipeline through the host layer from HOST buffers (alnpe_core re-staged, alnpe.c:530-615): per chunk of
    pairs salt_chunk_add_reads (pageable -> pinned queues), salt_chunk_submit with the PE thresholds 3 / 3, salt_chunk_wait,
    salt_chunk_pair = query_set_hits + pairing plans + one Smith-Waterman batch per rescue flavour + apply + MD/NM/XV of every
    mapped mate.  Chunks alternate between pipeline slots so that chunk k verifies while chunk k-1 pairs."""
    from salt_b200 import host_api, synth
    H = host_api.load()
    H.salt_host_set_threads(os.cpu_count() or 1)       # the reference pairs on its -t workers too
    g = wl["g"]; L = args.read_len; npairs = args.pe_pairs
    reads, pos, strand = synth.sample_pairs(g, npairs, L, seed=77, hard_frac=0.05, junk_frac=0.005)
    offs0, loci0, offs1, loci1 = synth.make_candidates(g, pos, strand, L, per_strand=args.cands, seed=78)
    n = 2 * npairs
    roffs = (np.arange(n + 1, dtype=np.uint64) * L).astype(np.uint32)
    cpairs = 50_000                                      # N_SEQS = 100000 reads per chunk (aln.h:27)
    n_slots = 2
    cap_c = 0
    for b in range(0, npairs, cpairs):
        r0, r1 = 2 * b, 2 * min(npairs, b + cpairs)
        cap_c = max(cap_c, int(offs0[r1]) - int(offs0[r0]), int(offs1[r1]) - int(offs1[r0]))
    chunks = [host_api.Chunk(H, 2 * cpairs + 8, (2 * cpairs + 8) * L, cap_c + 64) for _ in range(n_slots)]
    min_tlen, max_tlen = 250, 550                        # aln.c defaults (-a / -b)
    stats = []
    finals_keep = []
    # the caller's result buffers, one set per slot (a real caller keeps its query_t array)
    obufs = [((host_api.PairFinalT * cpairs)(), np.zeros(2 * cpairs, api.MDNM_OUT_DT), np.zeros((2 * cpairs, 64), np.uint8)) for _ in range(n_slots)]

    def run():
        stats.clear(); finals_keep.clear()
        pend = None
        k = 0
        for b in range(0, npairs, cpairs):
            m = min(cpairs, npairs - b)
            ch = chunks[k % n_slots]; slot = k % n_slots
            ch.reset()
            r0, r1 = 2 * b, 2 * (b + m)
            ch.add_reads(reads.reshape(-1), roffs[r0:r1 + 1], offs0[r0:r1 + 1], loci0, offs1[r0:r1 + 1], loci1)
            ch.submit(eng, slot, 3, 3)
            if pend is not None:
                pc_, ps_, pm_ = pend
                pc_.wait(eng, ps_)
                f, to, tm, st = pc_.pair(eng, ps_, pm_, min_tlen, max_tlen, g.l, md_stride=64, bufs=obufs[ps_])
                stats.append(st); finals_keep.append((f, to[:2 * pm_].copy()))
            pend = (ch, slot, m); k += 1
        pc_, ps_, pm_ = pend
        pc_.wait(eng, ps_)
        f, to, tm, st = pc_.pair(eng, ps_, pm_, min_tlen, max_tlen, g.l, md_stride=64, bufs=obufs[ps_])
        stats.append(st); finals_keep.append((f, to[:2 * pm_].copy()))
    run()
    t0 = time.perf_counter()
    reps = 2
    for _ in range(reps):
        run()
    sec = (time.perf_counter() - t0) / reps
    tot = {k: sum(getattr(s_, k) for s_ in stats) for k in ("pairs", "proper", "windows16", "windows5", "rescued", "promoted", "declined")}
    ms = {k: sum(getattr(s_, k) for s_ in stats) for k in ("ms_plan", "ms_ssw", "ms_apply", "ms_tail")}
    mapped = sum(int((to["md_len"] > 0).sum()) for _, to in finals_keep)
    cells = (tot["windows16"] + tot["windows5"]) * (8 * ((L + 7) // 8)) * 401
    res = {"pairs": npairs, "read_len": L, "reads_per_s": 2 * npairs / sec, "ms": sec * 1e3, "chunk_pairs": cpairs,
           "insert": "N(400,50)", "tlen_bounds": [min_tlen, max_tlen], "counts": tot, "stage_ms_host_wall": ms,
           "mapped_mates_with_tags": mapped, "rescue_frac_of_pairs": (tot["windows16"] + tot["windows5"]) / max(1, npairs),
           "sw_gcups_fwd_cells": cells / max(1e-9, ms["ms_ssw"] * 1e-3) / 1e9,
           "host_threads": os.cpu_count() or 1,
           "how": "wall time from pageable host arrays through salt_chunk_add_reads / _submit / _wait / salt_chunk_pair (two slots "
                  "alternating), results in host memory"}
    for ch in chunks:
        ch.close()
    # Two driver threads, each with its own handle on the same device (salt_b200_attach: own streams, slots and scratch, the
    # shared resident reference) and its own chunk queues, taking chunks alternately: one thread's GPU calls (rescue
    # Smith-Waterman, tags) run while the other's host threads plan / apply, and the next chunk is queued meanwhile.
    try:
        import threading
        engs = [eng, eng.attach()]
        wchunks = [[host_api.Chunk(H, 2 * cpairs + 8, (2 * cpairs + 8) * L, cap_c + 64) for _ in range(n_slots)] for _ in engs]
        wbufs = [[((host_api.PairFinalT * cpairs)(), np.zeros(2 * cpairs, api.MDNM_OUT_DT), np.zeros((2 * cpairs, 64), np.uint8))
                  for _ in range(n_slots)] for _ in engs]
        starts = list(range(0, npairs, cpairs))
        counts2 = [dict(), dict()]
        errs = []

        def drive(wi):
            try:
                e = engs[wi]; pend = None; k = 0; tot_w = {"pairs": 0, "rescued": 0, "mapped": 0}
                def finish(p):
                    pc_, ps_, pm_ = p
                    pc_.wait(e, ps_)
                    f, to, tm, st = pc_.pair(e, ps_, pm_, min_tlen, max_tlen, g.l, md_stride=64, bufs=wbufs[wi][ps_])
                    tot_w["pairs"] += st.pairs; tot_w["rescued"] += st.rescued; tot_w["mapped"] += int((to[:2 * pm_]["md_len"] > 0).sum())
                for b in starts[wi::len(engs)]:
                    m = min(cpairs, npairs - b)
                    ch = wchunks[wi][k % n_slots]; slot = k % n_slots
                    ch.reset()
                    r0, r1 = 2 * b, 2 * (b + m)
                    ch.add_reads(reads.reshape(-1), roffs[r0:r1 + 1], offs0[r0:r1 + 1], loci0, offs1[r0:r1 + 1], loci1)
                    ch.submit(e, slot, 3, 3)
                    if pend is not None:
                        finish(pend)
                    pend = (ch, slot, m); k += 1
                if pend is not None:
                    finish(pend)
                counts2[wi] = tot_w
            except Exception as ex:                      # noqa: BLE001
                errs.append(repr(ex))

        def run2():
            ths = [threading.Thread(target=drive, args=(wi,)) for wi in range(len(engs))]
            for t_ in ths: t_.start()
            for t_ in ths: t_.join()
        best = None
        for nthr in sorted({max(1, (os.cpu_count() or 1) // 2), os.cpu_count() or 1}):
            H.salt_host_set_threads(nthr)
            run2()
            t0 = time.perf_counter()
            for _ in range(reps):
                run2()
            sec2 = (time.perf_counter() - t0) / reps
            row = {"reads_per_s": 2 * npairs / sec2, "ms": sec2 * 1e3, "host_threads_per_driver": nthr,
                   "pairs": sum(c_.get("pairs", 0) for c_ in counts2), "rescued": sum(c_.get("rescued", 0) for c_ in counts2),
                   "mapped_mates_with_tags": sum(c_.get("mapped", 0) for c_ in counts2)}
            if best is None or row["reads_per_s"] > best["reads_per_s"]:
                best = row
        H.salt_host_set_threads(os.cpu_count() or 1)
        if errs:
            raise RuntimeError(errs[0])
        best["same_counts_as_one_driver"] = (best["pairs"] == tot["pairs"] and best["rescued"] == tot["rescued"]
                                             and best["mapped_mates_with_tags"] == mapped)
        best["how"] = "two driver threads, each with salt_b200_attach'ed handle and two chunk queues, chunks taken alternately"
        res["two_drivers"] = best
        for wc in wchunks:
            for ch in wc:
                ch.close()
        engs[1].close()
    except Exception as ex:                              # noqa: BLE001
        res["two_drivers"] = {"error": repr(ex)}
    # the reference's functions beside it (bounded sample): verification with the PE thresholds + ssw_align on as many windows
    try:
        from oracle import orc
        o = orc.Oracle(); ref = orc.Ref() if orc.ref_available() else None
        cores = os.cpu_count() or 1
        ns = min(n, 200_000)
        sec_v, *_ = o.verify_batch(g.mixref, g.l, np.ascontiguousarray(reads[:ns]).reshape(-1), roffs[:ns + 1], offs0[:ns + 1].copy(),
                                   loci0[:offs0[ns]], offs1[:ns + 1].copy(), loci1[:offs1[ns]], 3, 3, n_threads=cores, ref=ref)
        nw = max(1, int((tot["windows16"] + tot["windows5"]) * ns / n))
        nw = min(nw, 4000 * cores)
        rd = reads[1:2 * nw:2]
        rd = np.where(strand[1:2 * nw:2, None] == 1, synth.revcomp(rd), rd)
        st_ = np.maximum(0, pos[1:2 * nw:2].astype(np.int64) - 150); en_ = np.minimum(g.l - 1, st_ + 400)
        sec_w, _ = o.ssw_batch(g.mixref, np.ascontiguousarray(rd).reshape(-1), L, st_.astype(np.uint32), en_.astype(np.uint32),
                               o.score_mat2(), 3, 1, n_threads=cores, ref=ref)
        res["cpu"] = {"reads_per_s": ns / (sec_v + sec_w), "verify_s": sec_v, "ssw_s": sec_w, "windows": int(len(rd)), "cores": cores,
                      "kind": "reference" if ref else "port", "sample": "%d mates, %d rescue windows" % (ns, len(rd))}
    except Exception as ex:                                # noqa: BLE001
        res["cpu"] = {"error": repr(ex)}
    return res


def bench_se_host_layer(eng, wl, args):
    """The single-end chunk loop as a caller of the HOST LAYER runs it (include/salt_host.h; alnse_core re-staged,
    INTEGRATION.md section 2): per chunk salt_chunk_add_reads (pageable arrays -> the chunk's pinned queues: one byte per base,
    CSR lists), salt_chunk_submit, salt_chunk_wait, salt_chunk_results = query_set_hits + gen_mapq + query_gen_cigar of every
    read on the host threads.  Four chunk queues, one per pipeline slot.  Everything from the caller's arrays to the per-read
    query_t fields is inside the timed region."""
    from salt_b200 import host_api
    H = host_api.load()
    H.salt_host_set_threads(os.cpu_count() or 1)
    n, L = args.reads, args.read_len
    chunk = args.chunk
    n_slots = 4
    roffs = (np.arange(n + 1, dtype=np.uint64) * L).astype(np.uint32)
    o0, l0, o1, l1 = wl["offs0"], wl["loci0"], wl["offs1"], wl["loci1"]
    cap = 0
    for b in range(0, n, chunk):
        e = min(n, b + chunk)
        cap = max(cap, int(o0[e]) - int(o0[b]), int(o1[e]) - int(o1[b]))
    chunks = [host_api.Chunk(H, chunk + 8, (chunk + 8) * L, cap + 64) for _ in range(n_slots)]
    outs = [(host_api.ReadResultT * chunk)() for _ in range(n_slots)]
    codes = wl["reads"].reshape(-1)
    mapped = [0]

    def finish(k):
        ch = chunks[k % n_slots]
        ch.wait(eng, k % n_slots)
        rc = H.salt_chunk_results(ch.c, 5, outs[k % n_slots])
        assert rc == 0

    def run():
        k = 0
        for b in range(0, n, chunk):
            e = min(n, b + chunk)
            if k >= n_slots:
                finish(k - n_slots)
            ch = chunks[k % n_slots]
            ch.reset()
            ch.add_reads(codes, roffs[b:e + 1], o0[b:e + 1], l0, o1[b:e + 1], l1)
            ch.submit(eng, k % n_slots, 3, -1)
            k += 1
        for j in range(max(0, k - n_slots), k):
            finish(j)
    run()
    t0 = time.perf_counter()
    reps = 3
    for _ in range(reps):
        run()
    sec = (time.perf_counter() - t0) / reps
    last = outs[(((n + chunk - 1) // chunk) - 1) % n_slots]
    for ch in chunks:
        ch.close()
    return {"reads_per_s": n / sec, "ms": sec * 1e3, "chunk_reads": chunk, "host_threads": os.cpu_count() or 1,
            "last_read_pos": int(last[0].pos),
            "how": "pageable caller arrays -> salt_chunk_add_reads -> salt_chunk_submit / _wait -> salt_chunk_results (per-read query_t "
                   "fields on the host threads); plain transport (one byte per base), four chunk queues"}


def bench_seeding(args):
    """Row f1 beside the headline: single-end seeding + locate on the device, and the whole single-end stage from reads
    alone, against the reference's own functions on all host threads (tools/seed_bench.py at a size that fits the
    default run).  Needs the reference's indexer and seeding harness from oracle/_ref (input preparation and the CPU
    side); without them the block says so."""
    try:
        sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "tools"))
        import seed_bench
        import seed_cases
        if not seed_cases.have_ref():
            return {"unavailable": "oracle/_ref/salt-idx / libsaltref_seed.so not built (no reference tree at build time)"}
        a = argparse.Namespace(genome=args.seed_genome, reads=args.seed_reads, read_len=args.read_len, repeat_frac=0.10,
                               cpu_sample=min(args.seed_reads, 200_000), chunk=args.chunk, max_seed=50, max_locate=1000)
        return seed_bench.run(a)
    except Exception as ex:                                # noqa: BLE001 -- an extra, never the headline
        return {"error": repr(ex)}


def bench_program(args):
    """The aligner as a program (salt_b200/salt_aln: no reference code in the loop) against the reference program itself at all
    host threads, whole-process wall time, SAM compared (tools/aln_speed.py at a size that fits the default run).  This process
    holds a CUDA context while they run, as nvidia-persistenced would.  Needs oracle/_ref/salt and salt-idx (input preparation
    and the CPU side); without them the block says so."""
    try:
        sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "tools"))
        import aln_speed
        if not aln_speed.have_programs():
            return {"unavailable": "oracle/_ref/salt, salt-idx or salt_b200/salt_aln not built"}
        return aln_speed.run(args.program_genome, args.program_reads, args.program_reads // 4, os.cpu_count() or 1, None, hold_context=False)
    except Exception as ex:                                # noqa: BLE001 -- an extra, never the headline
        return {"error": repr(ex)}


def bench_lv(eng, lib, h, wl, args, dev, stream):
    """Landau-Vishkin kernel alone on decoy-heavy pair lists (every candidate of the first reads), k = 2..10."""
    import torch
    from salt_b200 import api
    n = min(len(wl["reads"]), 500_000)
    n0 = int(wl["offs0"][n]); n1 = int(wl["offs1"][n])
    rid0 = np.repeat(np.arange(n, dtype=np.uint32), np.diff(wl["offs0"][:n + 1].astype(np.int64)))
    rid1 = np.repeat(np.arange(n, dtype=np.uint32), np.diff(wl["offs1"][:n + 1].astype(np.int64)))
    pairs = np.concatenate([api.Engine.make_pairs(rid0, np.zeros(n0, np.uint32), wl["loci0"][:n0]),
                            api.Engine.make_pairs(rid1, np.ones(n1, np.uint32), wl["loci1"][:n1])])
    d_pairs = torch.from_numpy(pairs.view(np.uint8)).to(dev)
    d_out = torch.empty(len(pairs), dtype=torch.int8, device=dev)
    L = args.read_len
    res = {}
    for k in (2, 3, 5, 8, 10):
        for _ in range(2):
            lib.salt_b200_lv_dev(h, d_pairs.data_ptr(), len(pairs), k, d_out.data_ptr())
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(dev)
        e0.record(stream)
        for _ in range(3):
            lib.salt_b200_lv_dev(h, d_pairs.data_ptr(), len(pairs), k, d_out.data_ptr())
        e1.record(stream)
        torch.cuda.synchronize(dev)
        sec = e0.elapsed_time(e1) * 1e-3 / 3
        res["k%d" % k] = {"pairs_per_s": len(pairs) / sec, "gcups_equiv": len(pairs) * L * (L + 4) / sec / 1e9, "ms": sec * 1e3}
    # the two work mappings at the SE default k = L/10 (north star: warp per candidate, lanes over diagonals)
    for mapping, name in ((1, "warp_per_pair_k10"), (2, "thread_per_pair_k10")):
        lib.salt_b200_set_lv_mapping(h, mapping)
        lib.salt_b200_lv_dev(h, d_pairs.data_ptr(), len(pairs), 10, d_out.data_ptr())
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(dev)
        e0.record(stream)
        for _ in range(3):
            lib.salt_b200_lv_dev(h, d_pairs.data_ptr(), len(pairs), 10, d_out.data_ptr())
        e1.record(stream)
        torch.cuda.synchronize(dev)
        res[name + "_ms"] = e0.elapsed_time(e1) / 3
    lib.salt_b200_set_lv_mapping(h, 0)
    res["pairs"] = len(pairs)
    res["note"] = "GCUPS-equivalent = L*(L+4) DP cells per pair (SURVEY §8d); ~94% of pairs are decoys (worst case for LV)"
    return res


def bench_sam_tail(eng, lib, h, pkc, args, n, n0, n1, pins):
    """MD/NM/XV (sam_add_md_nm, sam.c:246-328) of every primary of the batch on the chunk pipeline: per chunk
    salt_b200_verify_submit_packed -> _verify_wait -> salt_b200_tail_primaries on the slot (positions, strands and CIGARs are
    already on the device; tags come back packed into pinned buffers).  Reported: the whole loop, and the same loop without the
    tail calls, so that the difference is what the SAM tail costs on top of the verification stage."""
    import torch
    from salt_b200 import api
    h_rec, h_acc0, h_acc1, h_cig = pins
    chunk = args.chunk
    L = args.read_len
    n_slots = int(lib.salt_b200_n_slots())
    h_out = torch.empty(n * 8, dtype=torch.uint8).pin_memory(); h_offs = torch.empty((chunk + 1) * n_slots, dtype=torch.int32).pin_memory()
    md_cap = chunk * 48
    h_md = torch.empty(md_cap * n_slots, dtype=torch.uint8).pin_memory()
    h_xv = torch.empty(n * 8, dtype=torch.int16).pin_memory()
    c0 = np.frombuffer(C.string_at(pkc.n_cand[0], 2 * n), np.uint16).astype(np.int64); c1 = np.frombuffer(C.string_at(pkc.n_cand[1], 2 * n), np.uint16).astype(np.int64)
    o0 = np.concatenate([[0], np.cumsum(c0)]); o1 = np.concatenate([[0], np.cumsum(c1)])
    views = []
    for b in range(0, n, chunk):
        m = min(chunk, n - b)
        v = api.PackedChunkT()
        v.n_reads = m; v.base_bits = 2; v.bases = pkc.bases; v.base_start = b * L; v.lens = None; v.l_seq = L
        v.n_pos = pkc.n_pos; v.n_n = pkc.n_n; v.count_bits = 16
        v.n_cand[0] = pkc.n_cand[0] + 2 * b; v.n_cand[1] = pkc.n_cand[1] + 2 * b
        v.loci[0] = pkc.loci[0] + 4 * int(o0[b]); v.loci[1] = pkc.loci[1] + 4 * int(o1[b])
        views.append((b, m, v))
    stat = {"md_bytes": 0, "nm": 0}

    def ck(rc):
        if rc != 0:
            raise RuntimeError(lib.salt_b200_last_error().decode())

    def finish(k, with_tail):
        b, m, v = views[k]; si = k % n_slots
        if with_tail:
            nb = C.c_size_t(0)
            ck(lib.salt_b200_tail_wait(h, si, C.byref(nb)))        # completes the verify of the slot as well
            stat["md_bytes"] += nb.value
        else:
            ck(lib.salt_b200_verify_wait(h, si))

    def run(with_tail):
        stat["md_bytes"] = 0
        for k, (b, m, v) in enumerate(views):
            si = k % n_slots
            if k >= n_slots:
                finish(k - n_slots, with_tail)
            ck(lib.salt_b200_verify_submit_packed(h, si, C.byref(v), 3, -1, h_rec.data_ptr() + 16 * b, h_acc0.data_ptr() + int(o0[b]),
                                                  h_acc1.data_ptr() + int(o1[b]), h_cig.data_ptr() + 128 * b, 128))
            if with_tail:                                          # queued right behind the verify, on the same stream
                ck(lib.salt_b200_tail_submit(h, si, h_out.data_ptr() + 8 * b, h_offs.data_ptr() + 4 * (chunk + 1) * si,
                                             h_md.data_ptr() + md_cap * si, md_cap, h_xv.data_ptr() + 16 * b, 8))
        for k in range(max(0, len(views) - n_slots), len(views)):
            finish(k, with_tail)
    res = {}
    for with_tail in (False, True):
        run(with_tail)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(3):
            run(with_tail)
        torch.cuda.synchronize()
        res[with_tail] = (time.perf_counter() - t0) / 3
    out = np.frombuffer(h_out.numpy().tobytes(), api.MDNM_OUT_DT)
    mapped = int((out["md_len"] > 0).sum())
    return {"alignments": mapped, "ms_verify_loop": res[False] * 1e3, "ms_verify_plus_tail_loop": res[True] * 1e3,
            "ms": (res[True] - res[False]) * 1e3, "alignments_per_s": mapped / max(1e-9, res[True] - res[False]),
            "nm_mean": float(out["nm"][out["md_len"] > 0].mean()) if mapped else 0.0, "md_overflow": int((out["md_len"] < 0).sum()),
            "md_bytes": int(stat["md_bytes"]), "d2h_bytes": int(stat["md_bytes"] + n * (8 + 16 + 4)),
            "note": "salt_b200_tail_submit right behind every salt_b200_verify_submit_packed, salt_b200_tail_wait per slot, pinned outputs; ms = loop with tails minus loop without"}


def bench_sw(eng, lib, h, wl, args, dev, stream):
    """Mate-rescue SSW (forward + reverse + banded traceback) on 401-wide windows around the true loci."""
    import torch
    from salt_b200 import api
    nt = min(args.sw_tasks, len(wl["reads"]))
    g = wl["g"]; L = args.read_len; W = 401
    rng = np.random.default_rng(5)
    start = np.maximum(0, wl["pos"][:nt].astype(np.int64) - rng.integers(0, W - L, nt))
    end = np.minimum(g.l - 1, start + W - 1)
    wins = np.zeros(nt, api.WIN_DT)
    wins["rs"] = (np.arange(nt, dtype=np.uint32) << 1) | wl["strand"][:nt]
    wins["start"] = start; wins["end"] = end
    d_w = torch.from_numpy(wins.view(np.uint8)).to(dev)
    d_out = torch.empty(nt * 28, dtype=torch.uint8, device=dev)
    d_cig = torch.empty(nt * 32, dtype=torch.int32, device=dev)
    mat = api.salt_score_mat2()
    lib.salt_b200_set_max_window(h, 408)

    def run():
        rc = lib.salt_b200_ssw_dev(h, d_w.data_ptr(), nt, 0, mat.ctypes.data, 16, 3, 1, 2, 0, 20, -1, d_out.data_ptr(), d_cig.data_ptr(), 32)
        if rc != 0:
            raise RuntimeError(lib.salt_b200_last_error().decode())
    for _ in range(2):
        run()
    torch.cuda.synchronize(dev)
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(3):
        run()
    e1.record(stream)
    torch.cuda.synchronize(dev)
    sec = e0.elapsed_time(e1) * 1e-3 / 3
    eng.profile(True); run(); st = {k: v for k, v in eng.profile_read().items() if k.startswith("ssw")}; eng.profile(False)
    cells = nt * (8 * ((L + 7) // 8)) * W          # forward cells, the convention of the CPU probe
    res = {"tasks": nt, "window": W, "read_len": L, "ms": sec * 1e3, "tasks_per_s": nt / sec,
           "gcups_fwd_cells_whole_pipeline": cells / sec / 1e9,
           "gcups_fwd_kernel": cells / (st.get("ssw_dp_fwd", float("nan")) * 1e-3) / 1e9, "stages_ms": st}
    # CPU beside it: the reference's ssw on a bounded sample
    try:
        from oracle import orc
        o = orc.Oracle(); ref = orc.Ref() if orc.ref_available() else None
        m = min(nt, 4000 * (os.cpu_count() or 1))
        rd = wl["reads"][:m]
        rd = np.where(wl["strand"][:m, None] == 1, __import__("salt_b200.synth", fromlist=["revcomp"]).revcomp(rd), rd)
        s, chk = o.ssw_batch(g.mixref, np.ascontiguousarray(rd).reshape(-1), L, wins["start"][:m].copy(), wins["end"][:m].copy(),
                             o.score_mat2(), 3, 1, n_threads=os.cpu_count() or 1, ref=ref)
        res["cpu"] = {"gcups_fwd_cells": m * (8 * ((L + 7) // 8)) * W / s / 1e9, "tasks_per_s": m / s, "cores": os.cpu_count(),
                      "kind": "reference" if ref else "port", "sample": "%d tasks" % m}
    except Exception as ex:
        res["cpu"] = {"error": repr(ex)}
    return res


if __name__ == "__main__":
    main()
