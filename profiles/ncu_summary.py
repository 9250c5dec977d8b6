#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, no GPU needed): headline metrics + top stall lines of the source page.
usage: python profiles/ncu_summary.py gpurun_out/x.ncu-rep [nlines]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
ntop = int(sys.argv[2]) if len(sys.argv) > 2 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "smsp__inst_executed.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__inst_executed_op_shared_ld.sum", "sm__cycles_elapsed.max", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]
for r in rows[2:]:
    print("== kernel:", r[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?")
    for k in KEYS:
        if k in hdr:
            print("  %-90s %s %s" % (k, r[hdr.index(k)], units[hdr.index(k)]))
    for i, k in enumerate(hdr):
        if "issue_stalled" in k and k.endswith("per_issue_active.ratio") or k.endswith("_per_warp_active.pct") and "stalled" in k:
            try:
                if float(r[i]) > 0.05:
                    print("  %-90s %s" % (k, r[i]))
            except Exception:
                pass
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h = None
for i, r in enumerate(rows):
    if "Source" in r and any("Sampl" in c for c in r):
        h, body = r, rows[i + 1:]
        break
if h is None:
    sys.exit(0)
cs = h.index("Source")
print("== source-page columns:", h)
samp = [j for j, c in enumerate(h) if "Sampl" in c]
j = samp[0]
tot, acc = 0, []
for r in body:
    try:
        v = float(r[j])
    except Exception:
        continue
    tot += v
    acc.append((v, r[cs].strip(), r[0]))
acc.sort(reverse=True)
print("== top stall-sample lines (%s), total %d" % (h[j], tot))
for v, s, a in acc[:ntop]:
    print("  %6.2f%%  %s  | %s" % (100.0 * v / max(tot, 1), a, s[:120]))
