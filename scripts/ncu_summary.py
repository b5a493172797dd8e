#!/usr/bin/env python3
"""Summarise an .ncu-rep: key raw metrics + SASS hot spots (stall samples, instruction share).
usage: ncu_summary.py report.ncu-rep [top_n]"""
import csv, subprocess, sys, io
rep = sys.argv[1]
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "lts__t_sectors.sum", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "launch__registers_per_thread", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "launch__grid_size", "launch__block_size", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__average_warp_latency_issue_stalled_long_scoreboard"]
for vals in rows[2:]:
    print("== kernel:", vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?")
    for i, h in enumerate(hdr):
        if h in want or any(h == w for w in want):
            print("  %-75s %-10s %s" % (h, units[i], vals[i]))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = rows[1]; data = [r for r in rows[2:] if len(r) == len(hdr)]
ix = {h: i for i, h in enumerate(hdr)}
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
ts = sum(int(r[ix["# Samples"]]) for r in data); ti = sum(int(r[ix["Instructions Executed"]]) for r in data)
print("total samples", ts, "warp instructions", ti, "sass lines", len(data))
tot = {c: sum(int(r[ix[c]]) for r in data) for c in stall_cols}
print("stall mix:", ", ".join("%s %.1f%%" % (c[6:], 100 * v / max(ts, 1)) for c, v in sorted(tot.items(), key=lambda kv: -kv[1])[:7]))
order = sorted(range(len(data)), key=lambda k: -int(data[k][ix["# Samples"]]))[:topn]
for k in sorted(order):
    r = data[k]
    top = max(stall_cols, key=lambda c: int(r[ix[c]]))
    print("%5d %-62s smp %5.1f%% inst %5.2f%% thr %4s %-14s L2sec %s" % (k, r[ix["Source"]].strip()[:62], 100 * int(r[ix["# Samples"]]) / max(ts, 1),
          100 * int(r[ix["Instructions Executed"]]) / max(ti, 1), r[ix["Avg. Threads Executed"]], top[6:], r[ix["L2 Theoretical Sectors Global"]]))
