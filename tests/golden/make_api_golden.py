"""Run the reference build of tests/cpp/api_check.cpp (oracle/_ref/api_check_ref) on small .mtx
inputs and commit its dumps as fixtures: tests/golden/api_<name>/.  Build container only."""
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, HERE)
import pyoracle  # noqa: E402


def write_mtx(path, nrow, ncol, ri, ci, va):
    with open(path, "w") as f:
        f.write("%%MatrixMarket matrix coordinate real general\n% generated for the arm-spmv parity tests\n")
        f.write(f"{nrow} {ncol} {len(va)}\n")
        for r, c, v in zip(ri, ci, va):
            f.write(f"{r + 1} {c + 1} {float(v)!r}\n")


def inputs():
    O = pyoracle.Oracle()
    ri, ci, va = O.gen_lap5_coo(12)
    yield "lap5_12", 144, 144, ri, ci, va
    rs = np.random.RandomState(21)
    n = 90
    ri = rs.randint(0, n, 700).astype(np.int32)
    ci = rs.randint(0, n, 700).astype(np.int32)
    ci[(ri == 0) & (ci == n - 1)] = 0
    d = np.flatnonzero(ri == ci)
    ci[d[n // 2:]] = (ci[d[n // 2:]] + 1) % (n - 1)
    yield "rand90", n, n, ri, ci, rs.uniform(-1, 1, 700)


def main():
    exe = os.path.join(ROOT, "oracle", "_ref", "api_check_ref")
    for name, nrow, ncol, ri, ci, va in inputs():
        d = os.path.join(HERE, "api_" + name)
        os.makedirs(d, exist_ok=True)
        mtx = os.path.join(d, "matrix.mtx")
        write_mtx(mtx, nrow, ncol, ri, ci, va)
        env = dict(os.environ, OMP_NUM_THREADS="1")
        subprocess.run([exe, mtx, d], check=True, env=env, stdout=subprocess.DEVNULL)
        print(name, sorted(os.listdir(d)))


if __name__ == "__main__":
    main()
