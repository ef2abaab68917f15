"""ncu launch list (ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file X.csv <command>) as a table:
id, microseconds, grid, block, kernel - plus each kernel's share of the total.  Usage: python scripts/launch_list.py X.csv"""
import csv
import sys
from collections import OrderedDict

rows = list(csv.reader(open(sys.argv[1])))
h = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
H = rows[h]
ix = {k: H.index(k) for k in ("ID", "Kernel Name", "Grid Size", "Block Size", "Metric Name", "Metric Unit", "Metric Value")}
tot = OrderedDict()
print("# id  time_us  grid  block  kernel")
for r in rows[h + 1:]:
    if len(r) < len(H) or r[ix["Metric Name"]] != "gpu__time_duration.sum":
        continue
    v = float(r[ix["Metric Value"]].replace(",", ""))
    us = v / 1e3 if r[ix["Metric Unit"]] in ("ns", "nsecond") else (v if r[ix["Metric Unit"]] in ("us", "usecond") else v * 1e3)
    name = r[ix["Kernel Name"]].split("(")[0][:60]
    print(f"{int(r[ix['ID']]):4d} {us:10.1f}  {r[ix['Grid Size']]:>15s}  {r[ix['Block Size']]:>12s}  {name}")
    a = tot.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += us
s = sum(v[1] for v in tot.values())
print("# per kernel: launches, total us, share")
for k, (n, us) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print(f"# {n:5d} {us:12.1f} {100 * us / s:6.2f}%  {k}")
