"""Drop-in boundary, CPU-side checks: bin/TH_sparse.a carries every symbol the reference's archive
exports, and (in the build container, where /root/reference exists) the reference's unmodified
main.cpp links against include/ + bin/TH_sparse.a with the reference's own link line."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ARCHIVE = os.path.join(ROOT, "bin", "TH_sparse.a")


def demangled_defined(path):
    out = subprocess.run(["nm", "-C", "--defined-only", path], capture_output=True, text=True).stdout
    names = set()
    for line in out.splitlines():
        parts = line.split(None, 2)
        if len(parts) == 3 and parts[1] in "TDBWVR":
            names.add(parts[2].strip())
    return names


@pytest.fixture(scope="module")
def archive():
    if not os.path.exists(ARCHIVE):
        subprocess.run(["make", "-C", os.path.join(ROOT, "arm-spmv_b200", "csrc", "host")], check=True, capture_output=True)
    return ARCHIVE


def test_archive_exports_every_reference_symbol(archive):
    """tests/golden/ref_archive_symbols.txt = `nm -C --defined-only` of the reference's bin/TH_sparse.a."""
    want = [l.strip() for l in open(os.path.join(ROOT, "tests", "golden", "ref_archive_symbols.txt")) if l.strip()]
    assert len(want) > 100
    have = demangled_defined(archive)
    missing = [s for s in want if s not in have]
    assert not missing, missing


def test_headers_cover_reference_interface():
    for h in ("matrix.h", "vector.h", "mat_vec.h", "vec_vec.h", "data_io.h", "mmio.h", "mytime.h", "numa_node.h", "thsp.h"):
        assert os.path.exists(os.path.join(ROOT, "include", h)), h


@pytest.mark.skipif(not os.path.exists("/root/reference/main.cpp"), reason="reference sources only exist in the build container")
def test_reference_main_links_unchanged(archive, tmp_path):
    """reference Makefile:11, verbatim except for the input/output paths."""
    exe = tmp_path / "main"
    env = dict(os.environ, LIBRARY_PATH=os.path.join(ROOT, "bin"))
    cmd = ["/usr/bin/g++", "-O2", "-lpthread", "-fopenmp", "-DUSE_OPENMP", "-I./include", "/root/reference/main.cpp", "-o", str(exe),
           "-I./include", "bin/TH_sparse.a", "-lm", "-lnuma"]
    r = subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    und = subprocess.run(["nm", "-u", str(exe)], capture_output=True, text=True).stdout
    assert "CSRMatrixMatVector" not in und   # resolved statically from the archive


@pytest.mark.skipif(not os.path.exists("/root/reference/include/matrix.h"), reason="reference headers only exist in the build container")
def test_api_exerciser_compiles_against_both_header_sets(tmp_path):
    """tests/cpp/api_check.cpp uses only the reference's public API: it must compile with either include dir."""
    src = os.path.join(ROOT, "tests", "cpp", "api_check.cpp")
    for inc in (os.path.join(ROOT, "include"), "/root/reference/include"):
        r = subprocess.run(["/usr/bin/g++", "-O0", "-fsyntax-only", "-I", inc, src], capture_output=True, text=True)
        assert r.returncode == 0, (inc, r.stderr[-1500:])
