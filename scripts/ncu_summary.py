#!/usr/bin/env python
"""One line per captured launch of an .ncu-rep (`ncu --set full`), the columns the roofline discussion uses.

    python scripts/ncu_summary.py X.ncu-rep out.csv "comment: the command that produced the capture"
"""
import csv
import subprocess
import sys

COLS = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__issue_active.avg.pct_of_peak_sustained_elapsed",
        "sm__cycles_elapsed.max", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__cluster_size",
        "sm__warps_active.avg.pct_of_peak_sustained_active"]
rep, out, comment = sys.argv[1], sys.argv[2], sys.argv[3]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, body = rows[0], rows[1], rows[2:]
idx = [hdr.index(c) for c in COLS if c in hdr]
with open(out, "w", newline="") as f:
    f.write("# " + comment + "\n")
    w = csv.writer(f)
    w.writerow([hdr[i] for i in idx])
    w.writerow([units[i] for i in idx])
    for r in body:
        w.writerow([r[i][:90] for i in idx])
print(out, len(body), "launches")
