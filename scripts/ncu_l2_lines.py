#!/usr/bin/env python3
"""Per-source-line L2 sector requests ("L2 Theoretical Sectors Global" of ncu's source page) of a kernel in an .ncu-rep:
which loads / stores make the memory traffic.  usage: ncu_l2_lines.py report.ncu-rep [launches_per_unit] [top_n]"""
import csv, io, subprocess, sys
rep = sys.argv[1]
units = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 30
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
hdr, lines = None, []
for r in csv.reader(io.StringIO(raw)):
    if r and r[0] == "Line No":
        hdr = r
        continue
    if hdr and len(r) == len(hdr) and r[0].isdigit():
        lines.append(r)
g, gi, op = hdr.index("L2 Theoretical Sectors Global"), hdr.index("L2 Theoretical Sectors Global Ideal"), hdr.index("Access Operation")
by = {}
for r in lines:
    k = (r[0], r[1].strip())
    a = by.setdefault(k, [0, 0])
    a[0] += int(r[g]); a[1] += int(r[gi])
tot = sum(v[0] for v in by.values()) or 1
print("L2 sector requests per unit: %.1f M = %.2f GB" % (tot / units / 1e6, tot * 32 / units / 1e9))
for (ln, s), (a, b) in sorted(by.items(), key=lambda kv: -kv[1][0])[:topn]:
    print("%5s %5.1f%% %8.1f M/unit = %6.2f GB (ideal %8.1f M)  %s" % (ln, 100 * a / tot, a / units / 1e6, a * 32 / units / 1e9, b / units / 1e6, s[:100]))
