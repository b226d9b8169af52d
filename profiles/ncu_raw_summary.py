#!/usr/bin/env python
"""Per-kernel summary of a full ncu capture.
usage: ncu -i X.ncu-rep --page raw --csv | profiles/ncu_raw_summary.py > profiles/rN/ncu_<tag>_summary.csv
One row per captured launch: duration, DRAM bytes (roofline `traffic`), DRAM/L2/L1/SM throughput as % of peak,
issue utilisation, lanes per instruction, L1/L2 sector hit rates, occupancy, registers."""
import csv
import sys

COLS = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "smsp__inst_executed.sum", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "lts__t_sectors.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__grid_size", "launch__block_size"]
rows = list(csv.reader(sys.stdin))
h, units = rows[0], rows[1]
idx = [h.index(c) for c in COLS if c in h]
w = csv.writer(sys.stdout)
w.writerow([h[i] for i in idx])
w.writerow([units[i] for i in idx])
for r in rows[2:]:
    w.writerow([r[i] for i in idx])
