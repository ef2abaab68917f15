"""CPU: the number conversions of the GPU Matrix Market parser (csrc/mtx_number.h, compiled for the host) against
strtod / strtol - the functions behind the reference's fscanf("%d %d %lg") (src/data_io.cpp:85)."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_decimal_to_double_matches_strtod(tmp_path):
    exe = str(tmp_path / "number_check")
    src = os.path.join(ROOT, "tests", "cpp", "number_check.cpp")
    subprocess.run(["g++", "-O2", "-std=c++17", "-w", "-I", os.path.join(ROOT, "arm-spmv_b200", "csrc"), src, "-o", exe], check=True)
    r = subprocess.run([exe, "3000000", "11"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-500:]
    assert r.stdout.startswith("ok 3000000"), r.stdout[-500:]
