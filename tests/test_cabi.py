"""The C-ABI library loads and exports exactly what include/thsp.h declares (CPU box, no compute)."""
import ctypes
import os
import re
import subprocess

import pytest


def test_library_exports_every_declared_symbol(thsp):
    lib = thsp.load()
    names = thsp.lib.declared_symbols()
    assert len(names) >= 60
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, f"declared in include/thsp.h but not exported: {missing}"


def test_no_undeclared_exports(thsp):
    out = subprocess.run(["nm", "-D", "--defined-only", thsp.lib.SO_PATH], capture_output=True, text=True).stdout
    exported = {l.split()[-1] for l in out.splitlines() if " T " in l and "thsp_" in l}
    declared = set(thsp.lib.declared_symbols())
    assert exported == declared, (exported - declared, declared - exported)


def test_version_and_host_only_calls(thsp):
    lib = thsp.load()
    assert b"sm_100a" in lib.thsp_version()
    s, c = ctypes.c_int64(), ctypes.c_int64()
    covered = 0
    for p in range(3):   # src/mat_vec.cpp:233,245-246
        assert lib.thsp_partition_rows(ctypes.c_int64(10), 3, p, ctypes.byref(s), ctypes.byref(c)) == 0
        assert s.value == covered
        covered += c.value
    assert covered == 10 and c.value == 4
    assert lib.thsp_partition_rows(ctypes.c_int64(10), 3, 3, ctypes.byref(s), ctypes.byref(c)) != 0
    assert lib.thsp_stencil27_nnz(4, 0, 64) == (3 * 4 - 2) ** 3
    assert lib.thsp_stencil27_nnz(256, 0, 256 ** 3) == 449455096
    assert lib.thsp_stencil27_nnz(512, 0, 512 ** 3) == 3609741304         # C5: > INT_MAX
    assert lib.thsp_lap5_nnz(1024) == 5238784                             # C1


def test_compute_fails_loudly_without_a_gpu(thsp):
    """No CPU fallback: on a box without CUDA every compute entry point reports an error."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("this box has a GPU")
    lib = thsp.load()
    rc = lib.thsp_fill_f64(ctypes.c_int64(4), ctypes.c_double(1.0), None, None)
    assert rc != 0 and b"no usable CUDA device" in lib.thsp_last_error()
    rc = lib.thsp_csr_spmv_f64(1, 1, 1, None, None, None, None, None, 1, None)
    assert rc != 0


def test_product_never_touches_the_oracle():
    """The package and the C-ABI sources must not import, link or call anything under oracle/."""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    pkg = os.path.join(root, "arm-spmv_b200")
    bad = []
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", "Makefile")):
                text = open(os.path.join(dp, f), errors="replace").read()
                for needle in ("pyoracle", "liboracle", "libref.so", "ref_shim", "import oracle", "from oracle"):
                    if needle in text:
                        bad.append((f, needle))
                if re.search(r"\boracle_\w+\s*\(", text):   # a call (comments may name the CPU twins)
                    bad.append((f, "oracle_*() call"))
    assert not bad, bad
