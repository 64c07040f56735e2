"""numpy restatement of the reference's score-matrix tail (TEST INFRASTRUCTURE).

PINNED against the reference's own ``CoverAlgorithm.getEvalStatistics`` executed in this
container (``tests/golden/make_golden.py`` -> ``tests/golden/evalstats_golden.json``).

Reference sources followed (file:line under /root/reference):
  acoss/algorithms/algorithm_template.py:205-290   getEvalStatistics
  acoss/algorithms/algorithm_template.py:189-191   Ds += Ds.T  (symmetric fill)
  acoss/algorithms/rqa_serra09.py:71-83            normalize_by_length
"""
from __future__ import annotations

import warnings
import numpy as np

__all__ = ["symmetrize", "normalize_by_length", "eval_statistics"]


def symmetrize(D: np.ndarray) -> np.ndarray:
    """all_pairwise(symmetric=True) tail: D += D.T (algorithm_template.py:189-191)."""
    D = np.array(D, dtype=np.float32)
    return (D + D.T).astype(np.float32)


def normalize_by_length(D: np.ndarray, n_frames) -> np.ndarray:
    """Ds[i, j] /= sqrt(n_frames_j) for all i, j (rqa_serra09.py:71-83).  The divisor is a
    float64 np.sqrt of an int; the float32 memmap cell is divided in float64 and stored
    back as float32."""
    D = np.array(D, dtype=np.float32)
    fac = np.sqrt(np.asarray(n_frames, dtype=np.int64))          # float64
    return (D.astype(np.float64) / fac[None, :]).astype(np.float32)


def eval_statistics(D: np.ndarray, cliques: dict, topsidx=(1, 10, 100, 1000)):
    """MR, MRR, MDR, MAP, Top-k exactly as getEvalStatistics computes them.

    ``cliques`` is the reference's ``self.cliques``: {label: set(indices)}; dict order and
    set iteration order are used the same way the reference uses them."""
    D = np.array(D, dtype=np.float32)
    N = D.shape[0]
    cl = [list(cliques[s]) for s in cliques]
    Ks = np.array([len(c) for c in cl])
    order = np.argsort(-Ks)
    Ks = Ks[order]
    cl = [cl[i] for i in order]
    idx = np.array([x for c in cl for x in c], dtype=int)
    D = D[idx, :][:, idx]
    np.fill_diagonal(D, -np.inf)
    srt = np.argsort(-D, 1)
    ranks = np.nan * np.ones(N)
    allmap = np.nan * np.ones(N)
    startidx, kidx = 0, 0
    for i in range(N):
        if i >= startidx + Ks[kidx]:
            startidx += Ks[kidx]
            kidx += 1
            if Ks[kidx] < 2:
                break
        diff = srt[i] - startidx
        iranks = (np.nonzero((diff >= 0) & (diff < Ks[kidx]))[0] + 1)[:-1]
        if len(iranks) == 0:
            warnings.warn("Recalling 0 songs for clique of size %i at song index %i" % (Ks[kidx], i))
            break
        ranks[i] = iranks[0]
        P = np.array([float(j) / float(r) for (j, r) in zip(range(1, Ks[kidx]), iranks)])
        allmap[i] = np.mean(P)
    MAP = np.nanmean(allmap)
    ranks = ranks[np.isnan(ranks) == 0]
    MR = np.mean(ranks)
    MRR = 1.0 / N * (np.sum(1.0 / ranks))
    MDR = np.median(ranks)
    tops = np.array([np.sum(ranks <= t) for t in topsidx], dtype=np.float64)
    return MR, MRR, MDR, MAP, tops, ranks
