"""Reduce an `ncu --set full` raw export (`ncu -i X.ncu-rep --page raw --csv > X_raw.csv`, tens of MB) to the columns profiles/README.md cites,
so that the evidence can be committed:  python tools/ncu_reduce.py X_raw.csv profiles/r2_x_ncu_full_part.csv"""
import csv
import sys

KEEP = ["ID", "Kernel Name", "Block Size", "Grid Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "lts__t_sector_hit_rate.pct",
        "sm__cycles_active.avg", "smsp__cycles_active.avg"]
rows = list(csv.reader(open(sys.argv[1])))
h = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
idx = [rows[h].index(k) for k in KEEP if k in rows[h]]
with open(sys.argv[2], "w", newline="") as f:
    w = csv.writer(f)
    for r in rows[h:]:
        if len(r) > max(idx):
            w.writerow([r[i] for i in idx])
print("wrote", sys.argv[2], len(rows) - h - 2, "launches")
