"""Summarise an ncu raw-page CSV (ncu -i X.ncu-rep --page raw --csv) into the metrics we track."""
import csv
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__occupancy_limit",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "sm__inst_executed.sum", "smsp__issue_active.avg.pct",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
        "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_bytes.sum", "launch__shared_mem_per_block", "launch__grid_size", "launch__block_size",
        "smsp__average_warp_latency_issue_stalled", "smsp__average_warps_issue_stalled", "sm__throughput.avg.pct",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared", "smsp__inst_executed_op_shared", "lts__throughput.avg.pct", "l1tex__throughput.avg.pct",
        "sm__cycles_elapsed.avg ", "smsp__cycles_active.avg", "dram__cycles_active"]
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    print("=== kernel:", r[hdr.index("Kernel Name")][:70], "grid", r[hdr.index("Grid Size")], "block", r[hdr.index("Block Size")])
    for i, h in enumerate(hdr):
        if any(w in h for w in WANT) or (len(sys.argv) > 2 and sys.argv[2] in h):
            print(f"{h[:95]:95s} {units[i]:14s} {r[i]}")
