"""One line per kernel from an ncu raw-page CSV (ncu -i X.ncu-rep --page raw --csv > X.csv): the counters
DESIGN.md argues from.  Usage: python scripts/ncu_table.py X.csv"""
import csv
import sys

COLS = [("ms", "gpu__time_duration.sum"), ("rd_GB", "dram__bytes_read.sum"), ("wr_GB", "dram__bytes_write.sum"),
        ("dram%", "dram__throughput.avg.pct_of_peak_sustained_elapsed"), ("L2hit%", "lts__t_sector_hit_rate.pct"),
        ("L1hit%", "l1tex__t_sector_hit_rate.pct"), ("sect/req", None), ("warps%", "sm__warps_active.avg.pct_of_peak_sustained_active"),
        ("issue%", "smsp__issue_active.avg.pct_of_peak_sustained_active"), ("regs", "launch__registers_per_thread"),
        ("longSB", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio"),
        ("shortSB", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio"),
        ("lgthr", "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio"),
        ("L2thr%", "lts__throughput.avg.pct_of_peak_sustained_elapsed")]
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
ix = {h: i for i, h in enumerate(hdr)}


def val(r, name):
    if name not in ix:
        return float("nan")
    try:
        v = float(r[ix[name]].replace(",", ""))
    except ValueError:
        return float("nan")
    u = units[ix[name]]
    scale = {"Mbyte": 1e-3, "Kbyte": 1e-6, "byte": 1e-9, "Gbyte": 1.0, "us": 1e-3, "ms": 1.0, "s": 1e3, "ns": 1e-6}.get(u, 1.0)
    return v * scale


print(f"{'kernel':44s} " + " ".join(f"{c[0]:>8s}" for c in COLS))
for r in rows[2:]:
    name = r[ix["Kernel Name"]][:44]
    out = []
    for label, metric in COLS:
        if label == "sect/req":
            s, q = val(r, "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum"), val(r, "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum")
            out.append(s / q if q else float("nan"))
        else:
            out.append(val(r, metric))
    print(f"{name:44s} " + " ".join(f"{v:8.3f}" for v in out))
