"""-m gpu: vec_vec.h / vector.h kernels.  Elementwise results are bit-identical to the reference
(same branch structure, unfused mul/add); reductions within 1e-12 relative."""
import numpy as np
import pytest

import cases as C
from conftest import load_golden
from gpu_util import assert_bits, host

pytestmark = pytest.mark.gpu


def test_vector_ops_match_golden(thsp, cuda):
    from arm_spmv_b200 import host as H
    g = load_golden("vec_ops")
    x, y, v = g["x"], g["y"], g["v"]
    X, Y = H.Vector(x), H.Vector(y)
    for i, (a, b) in enumerate(C.AXPBY_COEFFS):
        W = H.Vector(len(x)); H.vec_axpby(a, X, b, Y, W)
        assert_bits(host(W.values), g[f"axpby_{i}"], f"axpby {a},{b}")
    F = H.Vector(17); F.Fill(3.25); assert_bits(host(F.values), g["fill"], "fill")
    V = H.Vector(v); V.Scale(1.7); assert_bits(host(V.values), g["scale"], "scale")
    V = H.Vector(v); V.Shift(-0.3); assert_bits(host(V.values), g["shift"], "shift")
    V = H.Vector(len(v)); V.Copy(X); assert_bits(host(V.values), x, "copy")
    for i, a in enumerate(C.ADD_SCALED_COEFFS):
        V = H.Vector(v); V.AddScaled(a, X); assert_bits(host(V.values), g[f"add_scaled_{i}"], f"add_scaled {a}")
    for i, (a, b) in enumerate(C.ADD2_COEFFS):
        V = H.Vector(v); V.Add2Scaled(a, X, b, Y); assert_bits(host(V.values), g[f"add2_scaled_{i}"], f"add2 {a},{b}")
    d = H.vec_dot(X, Y)
    assert abs(d - float(g["dot"][0])) <= 1e-12 * float(np.sum(np.abs(x * y)))
    assert H.checkVector(X, H.Vector(x + 5e-7)) and not H.checkVector(X, H.Vector(x + 2e-6))
    assert not H.checkVector(X, H.Vector(x[:-1]))


@pytest.mark.parametrize("n", [0, 1, 31, 1000, 1 << 20, (1 << 22) + 3])
def test_dot_sizes_and_determinism(thsp, cuda, oracle, n):
    from arm_spmv_b200 import host as H
    x = oracle.gen_vector(n, 1) - 0.5; y = oracle.gen_vector(n, 2)
    X, Y = H.Vector(x), H.Vector(y)
    d1, d2 = H.vec_dot(X, Y), H.vec_dot(X, Y)
    assert d1 == d2                                    # fixed reduction tree
    assert abs(d1 - oracle.dot(x, y)) <= 1e-12 * max(float(np.sum(np.abs(x * y))), 1e-300)
    W = H.Vector(n); H.vec_axpby(0.3, X, -1.7, Y, W)
    assert_bits(host(W.values), oracle.axpby(0.3, x, -1.7, y), f"axpby n={n}")
