#!/usr/bin/env python
"""Golden binarisations of matrices WITH tied distances, produced by executing the reference's own
``csm_to_binary`` (``/root/reference/acoss/algorithms/utils/cross_recurrence.py:137-161``, ``.py_func``: numba 0.65
cannot type its argpartition call) under this container's numpy (2.3.x).

``np.argpartition`` (introselect) decides WHICH of several columns holding the NN-th smallest value of a row become
ones; the GPU kernel takes the lowest column indices.  The golden lets the GPU test say exactly where the two differ.

    python tests/golden/make_golden_knn_ties.py        -> tests/golden/knn_ties_golden.npz
"""
import importlib.util
import os
import sys

import numpy as np

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
sys.dont_write_bytecode = True


def cases():
    """name -> (D float64, kappa).  Ties come the way they do in practice: float32 features with repeated blocks
    (the reference stores float32 block features), so equal distances are exactly equal after widening."""
    out = {}
    rng = np.random.default_rng(31)
    # 1. euclidean CSM of float32 features where the second song repeats blocks -> duplicated COLUMNS
    X = rng.random((40, 24)).astype(np.float32)
    Yb = rng.random((12, 24)).astype(np.float32)
    Y = Yb[rng.integers(0, 12, size=60)]                     # every column value appears ~5 times
    D = np.sqrt(np.maximum(0.0, (X.astype(np.float64) ** 2).sum(1)[:, None] + (Y.astype(np.float64) ** 2).sum(1)[None, :]
                           - 2.0 * X.astype(np.float64) @ Y.astype(np.float64).T))
    for j in range(60):                                      # make duplicates exact (BLAS summation order may differ per column)
        first = int(np.nonzero((Y == Y[j]).all(1))[0][0])
        D[:, j] = D[:, first]
    out["dupcols_k01"] = (D, 0.1)
    out["dupcols_k7"] = (D, 7)
    # 2. quantised distances (many ties everywhere)
    Q = np.round(rng.random((33, 75)) * 8.0) / 8.0
    out["quant_k01"] = (Q, 0.1)
    out["quant_k02"] = (Q, 0.2)
    # 3. all-equal rows mixed with ordinary rows
    M = rng.random((20, 50))
    M[::4] = 0.5
    out["mixed_k01"] = (M, 0.1)
    # 4. ties only ABOVE / only BELOW the boundary (must be identical to numpy whatever the tie rule)
    T = rng.random((16, 40))
    T[:, :3] = -1.0                                          # three smallest tied, NN = 4 -> boundary not tied
    out["ties_below_k01"] = (T, 0.1)
    return out


def main():
    spec = importlib.util.spec_from_file_location("ref_cross_recurrence", os.path.join(REF, "acoss/algorithms/utils/cross_recurrence.py"))
    cr = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(cr)
    to_bin = cr.csm_to_binary.py_func
    out = {"numpy_version": np.array(np.__version__)}
    for name, (D, kappa) in cases().items():
        out["D_" + name] = D
        out["kappa_" + name] = np.float64(kappa)
        out["B_" + name] = np.asarray(to_bin(D, kappa)).astype(np.uint8)
    np.savez_compressed(os.path.join(HERE, "knn_ties_golden.npz"), **out)
    print("wrote", os.path.join(HERE, "knn_ties_golden.npz"), "numpy", np.__version__)


if __name__ == "__main__":
    main()
