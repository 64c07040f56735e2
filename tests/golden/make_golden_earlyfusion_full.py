#!/usr/bin/env python
"""Golden vectors for the FULL EarlyFusion pair scoring, made by EXECUTING THE REFERENCE'S OWN
``EarlyFusion.similarity`` method (and ``getWCSM``) in the build container.

Run from the repo root (needs /root/reference, numba, scipy):

    python tests/golden/make_golden_earlyfusion_full.py

Writes tests/golden/earlyfusion_full_golden.npz.  /root/reference does not exist on the GPU box, so the
tests only read the committed outputs; inputs are re-derived from seeds (acoss_b200.synthetic.ef_dataset).

What is executed (file:line under /root/reference):
  acoss/algorithms/earlyfusion_traile.py:157-198     EarlyFusion.similarity (unmodified method body), on an
                                                     instance built with __new__ whose block-feature cache
                                                     (all_block_feats, :92-95) is pre-filled
  acoss/algorithms/utils/similarity_fusion.py:38-54  getWCSM
  acoss/algorithms/utils/cross_recurrence.py         get_csm / get_csm_blocked_oti / csm_to_binary / get_oti as
                                                     ``.py_func`` (numba 0.65 cannot type them);
                                                     get_csm_cosine and smith_waterman_constrained jitted
Two flavours of input: float64 block features (the reference then computes in float64: pins the oracle
bit for bit) and float32 block features (what load_features really stores; the reference then runs
float32 BLAS, so its CSMs carry ~1e-7 relative noise — recorded to document the tie tolerance).
"""
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
sys.dont_write_bytecode = True

from make_golden import install_shims  # noqa: E402
from acoss_b200 import synthetic  # noqa: E402

# (name, cliques, n_blocks, seed, dims, dtype, kappa, K)
CASES = [
    ("f64_small", [2, 2, 1], 70, 301, dict(mfccs=60, ssms=45, chromas=48), "float64", 0.1, 10),
    ("f64_full", [2, 1], 90, 302, None, "float64", 0.1, 10),
    ("f64_k5", [3], 55, 303, dict(mfccs=40, ssms=28, chromas=24), "float64", 0.15, 5),
    ("f32_full", [2, 1], 90, 302, None, "float32", 0.1, 10),
    ("f64_count", [2, 1, 1], 45, 304, dict(mfccs=30, ssms=21, chromas=36), "float64", 6, 4),      # kappa >= 1: a count
    ("f32_small", [3, 1], 60, 305, dict(mfccs=50, ssms=36, chromas=60), "float32", 0.2, 12),
    ("f64_ragged", [2, 2], 80, 306, dict(mfccs=24, ssms=15, chromas=12), "float64", 0.05, 8),
]
JITTER = {"f64_ragged": 0.6}


def pack(B):
    return np.packbits(np.asarray(B, dtype=np.uint8), axis=1, bitorder="little")


def main():
    install_shims()
    cwd = os.getcwd()
    tmp = tempfile.mkdtemp()
    os.chdir(tmp)                                     # importing acoss writes log files into CWD
    out = {}
    try:
        import acoss.algorithms.earlyfusion_traile as eft
        import acoss.algorithms.utils.cross_recurrence as cr
        from acoss.algorithms.utils.similarity_fusion import getWCSM
        # numba 0.65 cannot compile these in nopython mode: run the same source as plain Python
        cr.get_oti = cr.get_oti.py_func
        eft.get_csm = cr.get_csm.py_func
        eft.get_csm_blocked_oti = cr.get_csm_blocked_oti.py_func
        eft.csm_to_binary = cr.csm_to_binary.py_func

        # getWCSM known answers
        r = np.random.default_rng(77)
        D = r.random((37, 53)) * 3.0
        out["wcsm_D"] = D
        out["wcsm_k10"] = getWCSM(D.copy(), 10, 10)
        out["wcsm_k3_7"] = getWCSM(D.copy(), 3, 7)

        names = []
        for name, cliques, nb, seed, dims, dtype, kappa, K in CASES:
            feats = synthetic.ef_dataset(cliques, nb, seed, dims=dims, dtype=np.dtype(dtype),
                                         jitter=JITTER.get(name, 0.15))
            n = len(feats)
            ef = eft.EarlyFusion.__new__(eft.EarlyFusion)
            ef.name, ef.shortname, ef.cachedir, ef.chroma_type = "EarlyFusionTraile", "golden", tmp, "hpcp"
            ef.kappa, ef.K, ef.log_times = kappa, K, False
            ef.all_block_feats = {i: {k: v for k, v in f.items() if k != "label"} for i, f in enumerate(feats)}
            ef.Ds = {s: np.zeros((n, n), dtype=np.float32) for s in ("mfccs", "ssms", "chromas", "early")}
            i, j = np.triu_indices(n, k=1)
            idxs = np.stack([i, j], axis=1)
            ef.similarity(idxs)
            for s in ef.Ds:
                out["%s_D_%s" % (name, s)] = ef.Ds[s]
            # matrices of the first pair, recomputed with the same reference functions the method calls
            f1, f2 = ef.all_block_feats[0], ef.all_block_feats[1]
            C = {"mfccs": eft.get_csm(f1["mfccs"], f2["mfccs"]), "ssms": eft.get_csm(f1["ssms"], f2["ssms"]),
                 "chromas": eft.get_csm_blocked_oti(f1["chromas"], f2["chromas"], f1["chroma_med"], f2["chroma_med"],
                                                    cr.get_csm_cosine)}
            W = np.zeros_like(C["mfccs"])
            for s in C:
                W += getWCSM(C[s], K, K)
            E = np.exp(-W)
            out["%s_oti01" % name] = np.int64(cr.get_oti(f1["chroma_med"], f2["chroma_med"]))
            for s in C:
                out["%s_bin01_%s" % (name, s)] = pack(eft.csm_to_binary(C[s], kappa))
                out["%s_csmsum01_%s" % (name, s)] = np.float64(np.sum(C[s], dtype=np.float64))
            out["%s_bin01_early" % name] = pack(eft.csm_to_binary(E, kappa))
            out["%s_early01" % name] = E.astype(np.float64) if name != "f64_full" else E[::7, ::5].astype(np.float64)
            names.append(name)
        out["cases"] = np.array(names)
    finally:
        os.chdir(cwd)
    np.savez_compressed(os.path.join(HERE, "earlyfusion_full_golden.npz"), **out)
    print("wrote earlyfusion_full_golden.npz:", len(out), "arrays",
          os.path.getsize(os.path.join(HERE, "earlyfusion_full_golden.npz")), "bytes")


if __name__ == "__main__":
    main()
