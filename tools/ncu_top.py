#!/usr/bin/env python
"""Summarise an ncu report: key raw metrics and the instructions with the most stall samples.
    python tools/ncu_top.py gpurun_out/x.ncu-rep [n]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
n = int(sys.argv[2]) if len(sys.argv) > 2 else 24
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
for krow in rows[2:]:
    d = {h: v for h, v in zip(rows[0], krow)}
    print("==", d.get("Kernel Name", "")[:100])
    for k in ["gpu__time_duration.sum", "sm__cycles_elapsed.avg", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
              "sm__issue_active.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.avg", "launch__registers_per_thread",
              "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
              "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
              "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__warps_active.avg.pct_of_peak_sustained_active"]:
        if k in d:
            print("  %-70s %s" % (k, d[k]))
    for k, v in d.items():
        if "issue_stalled" in k and k.endswith("per_issue_active.ratio"):
            try:
                if float(v) > 0.1:
                    print("  stall %-40s %s" % (k.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""), v))
            except ValueError:
                pass
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
ix = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[hi + 1:] if len(r) == len(hdr)]


def f(r, k):
    try:
        return float(r[ix[k]])
    except (ValueError, KeyError):
        return 0.0


print("total samples", sum(f(r, "# Samples") for r in data), "instructions", len(data))
for r in sorted(data, key=lambda r: -f(r, "# Samples"))[:n]:
    print(r[ix["Address"]][-5:], "%5d" % f(r, "# Samples"), "long %4d short %4d math %4d wait %4d bar %4d exec %8d" % (
        f(r, "stall_long_sb"), f(r, "stall_short_sb"), f(r, "stall_math"), f(r, "stall_wait"), f(r, "stall_barrier"),
        f(r, "Instructions Executed")), r[ix["Source"]][:90])
