"""Generate tests/golden/*.npz from the UNMODIFIED reference (oracle/_ref/libref.so).

Run in the build container (needs /root/reference to have been compiled by oracle/Makefile):
    python tests/golden/make_golden.py
The fixtures pin the oracle and the CUDA library on boxes where the reference is absent.
SpMV outputs are produced with one OpenMP thread, the only setting in which the reference's
COO/CSC atomics have a defined order."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "..", "oracle"))
sys.path.insert(0, HERE)
import pyoracle  # noqa: E402
from cases import ADD2_COEFFS, ADD_SCALED_COEFFS, AXPBY_COEFFS, cases, vec_inputs  # noqa: E402


def main():
    pyoracle.build()
    R = pyoracle.Ref()
    R.set_threads(1)
    for name, c in cases().items():
        nrow, ncol, ri, ci, va, x, y0 = c["nrow"], c["ncol"], c["ri"], c["ci"], c["va"], c["x"], c["y0"]
        g = dict(nrow=nrow, ncol=ncol, ri=ri, ci=ci, va=va, x=x, y0=y0)
        rp, co, cv, dg = R.coo2csr(nrow, ncol, ri, ci, va)
        g.update(csr_row_ptr=rp, csr_col_ind=co, csr_values=cv, csr_diagonal=dg)
        cp, ro, cv2 = R.coo2csc(nrow, ncol, ri, ci, va)
        g.update(csc_col_ptr=cp, csc_row_ind=ro, csc_values=cv2)
        k, eco, eva, edg = R.coo2ell(nrow, ncol, ri, ci, va)
        g.update(ell_width=k, ell_col_ind=eco, ell_values=eva, ell_diagonal=edg)
        g["y_coo"] = R.coo_spmv(nrow, ncol, ri, ci, va, x, y0)
        g["y_csr"] = R.csr_spmv(nrow, ncol, rp, co, cv, x, y0)
        g["y_csc"] = R.csc_spmv(nrow, ncol, cp, ro, cv2, x, y0)
        g["y_ell"] = R.ell_spmv(nrow, ncol, k, eco, eva, x, y0)
        if not c.get("no_dia"):
            off, dv = R.csr2dia(nrow, ncol, rp, co, cv)
            g.update(dia_offsets=off, dia_values=dv)
            if ncol >= nrow:  # the reference's j < nrow guard reads x[j]; keep j inside x
                g["y_dia"] = R.dia_spmv(nrow, ncol, off, dv, x, y0)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **g)
        print("wrote", name, "nnz", len(va), "K", k)
    x, y, v = vec_inputs()
    g = dict(x=x, y=y, v=v, dot=np.array([R.dot(x, y)]))
    for i, (a, b) in enumerate(AXPBY_COEFFS):
        g[f"axpby_{i}"] = R.axpby(a, x, b, y)
    g["fill"] = R.fill(17, 3.25)
    g["scale"] = R.scale(1.7, v)
    g["shift"] = R.shift(-0.3, v)
    for i, a in enumerate(ADD_SCALED_COEFFS):
        g[f"add_scaled_{i}"] = R.add_scaled(a, x, v)
    for i, (a, b) in enumerate(ADD2_COEFFS):
        g[f"add2_scaled_{i}"] = R.add2_scaled(a, x, b, y, v)
    np.savez_compressed(os.path.join(HERE, "vec_ops.npz"), **g)
    print("wrote vec_ops")


if __name__ == "__main__":
    main()
