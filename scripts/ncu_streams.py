"""Sectors per request of every global load / store / atomic instruction of a kernel, from the SASS source page of an ncu
report:  ncu -i X.ncu-rep --page source --csv --print-source sass [-k kernel] > X_src.csv ; python scripts/ncu_streams.py X_src.csv
(one warp-level request per executed instruction; 'sectors' = 32-byte sectors the request asks L2 for if it misses L1;
ideal = the minimum for the bytes requested).  This is the per-stream view the north star asks for: val / col_ind / x."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
kernel = ""
hdr = None
for r in rows:
    if len(r) >= 2 and r[0] == "Kernel Name":
        kernel = r[1][:70]
        print(f"=== {kernel}")
        hdr = None
        continue
    if r and r[0] == "Address":
        hdr = {h: i for i, h in enumerate(r)}
        continue
    if hdr is None or len(r) < len(hdr):
        continue
    src = r[hdr["Source"]].strip()
    try:
        ex = int(r[hdr["Instructions Executed"]] or 0)
        sect = int(r[hdr["L2 Theoretical Sectors Global"]] or 0)
        ideal = int(r[hdr["L2 Theoretical Sectors Global Ideal"]] or 0)
    except ValueError:
        continue
    if sect == 0 or ex == 0:
        continue
    op = " ".join(src.split()[:2]) if src.startswith("@") else src.split()[0]
    print(f"  {op:34s} executed {ex:11d}  sectors/request {sect / ex:6.2f}  (ideal {ideal / ex:6.2f})")
