#!/usr/bin/env python
"""Per-SASS-instruction view of an .ncu-rep source page: executed counts and stall samples.
    python tools/ncu_hot.py rep [kernel-substring]  -> prints address index, instr, #executed (warp), samples"""
import csv
import subprocess
import sys

rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
i = 0
while i < len(rows):
    if rows[i] and rows[i][0] == "Kernel Name":
        name = rows[i][1]
        hdr = rows[i + 1]
        i += 2
        body = []
        while i < len(rows) and not (rows[i] and rows[i][0] == "Kernel Name"):
            body.append(rows[i]); i += 1
        if len(sys.argv) > 2 and sys.argv[2] not in name:
            continue
        ci = hdr.index("Instructions Executed"); si = hdr.index("# Samples"); src = hdr.index("Source")
        tot = sum(int(b[ci]) for b in body if len(b) > ci)
        tots = sum(int(b[si]) for b in body if len(b) > si)
        print("==", name[:100], "total warp instr", tot, "samples", tots)
        for k, b in enumerate(body):
            if len(b) > ci:
                print("%4d %-60s %10s %5.1f%% smp %5s %5.1f%%" % (k, b[src].strip()[:60], b[ci], 100.0 * int(b[ci]) / max(tot, 1), b[si], 100.0 * int(b[si]) / max(tots, 1)))
    else:
        i += 1
