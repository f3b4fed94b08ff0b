#!/usr/bin/env python
"""Print the key metrics of an .ncu-rep (raw page) -- usage: ncu_keys.py file.ncu-rep [regex]"""
import csv, re, subprocess, sys
rep = sys.argv[1]
pat = re.compile(sys.argv[2] if len(sys.argv) > 2 else
    r'^(Kernel Name|Grid Size|Block Size|gpu__time_duration.sum|dram__bytes_read.sum$|dram__bytes_write.sum$|gpu__dram_throughput.avg.pct|sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed|sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct|sm__warps_active.avg.pct|launch__registers_per_thread$|lts__t_sector_hit_rate.pct|sm__throughput.avg.pct|lts__t_bytes.sum$|lts__throughput.avg.pct|l1tex__throughput.avg.pct|sm__cycles_elapsed.max|sm__cycles_active.avg$|smsp__inst_executed.sum$|sm__inst_executed_pipe_uniform|smsp__warp_issue_stalled.*_per_warp_active.pct)')
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units, data = rows[0], rows[1], rows[2:]
for i, h in enumerate(hdr):
    if pat.match(h):
        print(f"{h} [{units[i]}]: " + " | ".join(r[i] for r in data))
