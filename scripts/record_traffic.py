"""Writes profiles/traffic.json from an ncu raw-page CSV (ncu -i X.ncu-rep --page raw --csv): DRAM bytes read + written by
one launch of the kernel bench.py reports the roofline for, together with the SHA-256 of the sources that kernel is
compiled from.  bench.py only quotes the number while those sources are unchanged.
  python scripts/record_traffic.py X.csv <kernel-name-substring> <grid n> [label]"""
import csv
import hashlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SOURCES = ["arm-spmv_b200/csrc/csr_spmv.cu", "arm-spmv_b200/csrc/common.cuh"]


def sources_sha():
    h = hashlib.sha256()
    for p in SOURCES:
        h.update(open(os.path.join(ROOT, p), "rb").read())
    return h.hexdigest()


if __name__ == "__main__":
    path, needle, n = sys.argv[1], sys.argv[2], int(sys.argv[3])
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
    for r in rows[2:]:
        if needle in r[ix["Kernel Name"]]:
            tot = 0.0
            for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                tot += float(r[ix[m]].replace(",", "")) * scale[units[ix[m]]]
            out = {"kernel": r[ix["Kernel Name"]][:80], "grid": n, "dram_bytes_per_launch": int(tot), "sources": SOURCES,
                   "sources_sha256": sources_sha(), "capture": os.path.basename(path), "label": sys.argv[4] if len(sys.argv) > 4 else ""}
            json.dump(out, open(os.path.join(ROOT, "profiles", "traffic.json"), "w"), indent=1)
            print(out)
            break
    else:
        sys.exit(f"no kernel matching {needle!r} in {path}")
