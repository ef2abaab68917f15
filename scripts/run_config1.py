"""BASELINE configs[0]: 5-point Laplacian 1024^2 written as a .mtx file and run through the
reference's OWN unmodified driver twice: bin/main (main.cpp linked against this library, GPU) and
oracle/_ref/ref_main (main.cpp linked against the reference's sources, CPU).  Prints both outputs."""
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
threads = sys.argv[2] if len(sys.argv) > 2 else str(os.cpu_count())
path = f"/tmp/lap5_{n}.mtx"
t0 = time.time()
with open(path, "w") as f:
    nnz = 5 * n * n - 4 * n
    f.write("%%MatrixMarket matrix coordinate real general\n")
    f.write(f"{n * n} {n * n} {nnz}\n")
    buf = []
    for i in range(n):
        for j in range(n):
            r = i * n + j + 1
            if i > 0: buf.append(f"{r} {r - n} -1\n")
            if j > 0: buf.append(f"{r} {r - 1} -1\n")
            buf.append(f"{r} {r} 4\n")
            if j < n - 1: buf.append(f"{r} {r + 1} -1\n")
            if i < n - 1: buf.append(f"{r} {r + n} -1\n")
        if len(buf) > 200000:
            f.write("".join(buf)); buf = []
    f.write("".join(buf))
print(f"wrote {path} ({os.path.getsize(path) / 1e6:.0f} MB) in {time.time() - t0:.1f} s", flush=True)
for exe, label in ((os.path.join(ROOT, "bin", "main"), "bin/main = reference main.cpp + THIS library (GPU)"),
                   (os.path.join(ROOT, "oracle", "_ref", "ref_main"), "oracle/_ref/ref_main = reference main.cpp + reference sources (CPU)")):
    if not os.path.exists(exe):
        print("missing", exe); continue
    t0 = time.time()
    r = subprocess.run([exe, path, threads], capture_output=True, text=True, env=dict(os.environ, THSP_TRACE=os.environ.get("THSP_TRACE", "0")))
    print(f"--- {label}: argv = {path} {threads}, exit {r.returncode}, wall {time.time() - t0:.1f} s")
    print("\n".join(l for l in r.stdout.splitlines() if l.startswith("###")))
    if r.returncode != 0: print(r.stderr[-1000:])
    if os.environ.get("THSP_TRACE") == "1" and "bin/main" in exe:
        print("\n".join(l for l in r.stderr.splitlines() if l.startswith("[thsp]"))[:6000])
