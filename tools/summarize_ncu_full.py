"""Reduce `ncu -i X.ncu-rep --page raw --csv` to the columns the profiles/ summaries keep (same set as round 1).
usage: ncu -i rep --page raw --csv | python tools/summarize_ncu_full.py > profiles/NAME.csv"""
import csv
import sys

KEEP_EXACT = ["ID", "Kernel Name", "Block Size", "Grid Size", "gpu__time_duration.sum", "launch__registers_per_thread",
              "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
              "launch__occupancy_limit_warps", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second",
              "dram__bytes_write.sum.per_second", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
              "lts__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
              "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
              "smsp__inst_executed.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__cycles_elapsed.avg.per_second"]
rows = list(csv.reader(sys.stdin))
hdr, units, data = rows[0], rows[1], rows[2:]
cols = [i for i, h in enumerate(hdr) if h in KEEP_EXACT or ("issue_stalled" in h and h.endswith("per_issue_active.ratio"))]
w = csv.writer(sys.stdout)
w.writerow([hdr[i] for i in cols])
w.writerow([units[i] for i in cols])
for d in data:
    w.writerow([d[i] for i in cols])
