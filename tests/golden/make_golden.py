"""Generates tests/golden/*.npz: small input/output vectors for the sort-based path.

The reference is Julia and cannot run in this image, so these vectors come from the CPU oracle
(oracle/nl_oracle.cpp), and each is accepted only if the INDEPENDENT brute-force enumeration and
(where comparable) the legacy linked-list restatement agree with it.  They freeze today's verified
behaviour for regression purposes; they are not outputs of the reference itself.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import nl_oracle as O  # noqa: E402
from tests import util as U  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def cases():
    X, C, L = U.rand_config(150, seed=1001)
    yield "cubic_pbc_ttt_f64", X, C, (True, True, True), L * 0.25, np.float64
    X, C, L = U.rand_config(150, seed=1002, dtype=np.float32)
    yield "cubic_pbc_ttt_f32", X, C, (True, True, True), np.float32(L * 0.25), np.float32
    Xt = U.rand_in_cell(100, U.TRICLINIC, seed=1003)
    yield "triclinic_pbc_ttf_displaced_f64", U.displace_by_lattice(Xt, U.TRICLINIC, (True, True, False)), U.TRICLINIC, (True, True, False), 3.0, np.float64
    X, C = U.fcc(3.61, (2, 2, 2))
    yield "fcc_cu_2x2x2_rc5_f64", X, C, (True, True, True), 5.0, np.float64
    X, C, L = U.rand_config(40, seed=1004)
    yield "cutoff_0p6L_f64", X, C, (True, False, True), L * 0.6, np.float64
    C8 = np.eye(3) * 8.0
    yield "issue6_tiny_negative_frac_f64", np.array([[0.5, 7.7, 0.5], [0.5, -5e-17 * 8.0, 0.5]]), C8, (True, True, True), 1.5, np.float64


def main():
    for name, X, C, pbc, rc, dt in cases():
        r = O.sortbased(X, rc, C, pbc, dtype=dt)
        b = O.brute(X, rc, C, pbc, dtype=dt)
        assert np.array_equal(U.pair_tuples(r), U.pair_tuples(b)), name
        i, j, S, R = O.canonical(r["i"], r["j"], r["S"], r["R"])
        np.savez_compressed(os.path.join(HERE, name + ".npz"), X=np.asarray(X, dt), cell=np.asarray(C, dt), pbc=np.asarray(pbc),
                            cutoff=dt(rc), first=r["first"], i=i.astype(np.int32), j=j.astype(np.int32), S=S.astype(np.int32), R=R,
                            perm=r["perm"], cell_id=r["cell_id"], cell_offsets=r["cell_offsets"])
        print(name, r["npairs"])


if __name__ == "__main__":
    main()
