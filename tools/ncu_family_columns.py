"""Reduce an `ncu --page raw --csv` dump to the columns profiles/*_ncu_full_longconv_family_*.csv keep
(duration, DRAM bytes, issue/occupancy, stall reasons).  usage: python tools/ncu_family_columns.py raw.csv out.csv"""
import csv, sys
KEEP = ["ID", "Kernel Name", "Block Size", "Grid Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__occupancy_limit_shared_mem",
        "launch__occupancy_limit_registers", "smsp__inst_executed.sum", "lts__t_sector_hit_rate.pct",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]
rows = list(csv.reader(open(sys.argv[1], errors="replace")))
hdr = rows[0]
keep = [h for h in KEEP if h in hdr] + [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled_") and
                                         h.endswith("_per_issue_active.ratio") and "not_issued" not in h]
idx = [hdr.index(h) for h in keep]
with open(sys.argv[2], "w", newline="") as f:
    w = csv.writer(f)
    for r in rows:
        if len(r) >= len(hdr):
            w.writerow([r[i] for i in idx])
