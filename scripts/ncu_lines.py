#!/usr/bin/env python3
"""Per-source-line and per-region instruction counts (and stall samples, if sampled) of a kernel in an .ncu-rep.
usage: ncu_lines.py report.ncu-rep source_file [launches_per_unit] [top_n]
Regions are cut at marker comments of nec_kernels.cuh's engine."""
import csv, io, subprocess, sys
rep, src_path = sys.argv[1], sys.argv[2]
units = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
topn = int(sys.argv[4]) if len(sys.argv) > 4 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, lines = None, []
for r in rows:
    if r and r[0] == "Line No":
        hdr = r
        continue
    if hdr and len(r) == len(hdr) and r[0].isdigit():
        lines.append(r)
si, ii, ti = hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed")
tot_s, tot_i = sum(int(r[si]) for r in lines) or 1, sum(int(r[ii]) for r in lines) or 1
print("warp instructions per unit: %.1f M   (samples %d)" % (tot_i / units / 1e6, tot_s))
src = open(src_path).read().splitlines()
markers = [("reset", "---- reset: only the mask records"), ("seed", "---- seed level 0"), ("fwd level loop", "---- forward: level-synchronous"),
           ("fwd mark", "// mark: the level's vertices become"), ("fwd queue pass (rows, counters)", "for (uint32_t tile = qbeg + warp_t * 32; tile < qend;"),
           ("fwd push", "if (pull) continue;"), ("fwd pull scan", "---- pull scan: thread per vertex"), ("after fwd", "// here: levels 0..level-1 are non-empty"),
           ("bwd tile setup", "---- backward: levels depth .. 1"), ("bwd begin/finish entry", "// COMPACT LANES."), ("bwd consume group", "auto consume = [&]"),
           ("bwd staging pipeline", "// Flattened adjacency entries are consumed in chunks of 32."), ("bwd row loop", "for (uint32_t r = 0; r < cnt;) {"),
           ("bwd tile end", "if (idx < le && deg != 0) bslot[w] += my_total;"), ("debug / sigma", "// the sources themselves (level 0) get bk")]
bounds = []
for name, text in markers:
    for i, l in enumerate(src):
        if text in l:
            bounds.append((i + 1, name))
            break
bounds.sort()
agg = {}
for r in lines:
    ln, k = int(r[0]), "helpers (scans, shuffles, address setup)"
    for v, name in bounds:
        if ln >= v:
            k = name
    a = agg.setdefault(k, [0, 0])
    a[0] += int(r[si]); a[1] += int(r[ii])
for k, (s_, i_) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%-42s inst %5.1f%% (%7.1f M/unit)  samples %5.1f%%" % (k, 100 * i_ / tot_i, i_ / units / 1e6, 100 * s_ / tot_s))
print("--- top source lines by instructions")
for r in sorted(lines, key=lambda r: -int(r[ii]))[:topn]:
    print("%5s inst %5.1f%% (%6.1f M/unit) thr/inst %4.1f  %s" % (r[0], 100 * int(r[ii]) / tot_i, int(r[ii]) / units / 1e6, int(r[ti]) / max(1, int(r[ii])), r[1].strip()[:105]))
