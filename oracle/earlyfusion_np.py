"""numpy restatement of the reference's in-tree (EarlyFusion flavour) CSM / binarise /
Smith-Waterman utilities (TEST INFRASTRUCTURE).

PINNED: every function here is checked in ``tests/test_oracle_earlyfusion.py`` against
golden vectors produced by executing the reference's own code in this container
(``tests/golden/make_golden.py`` -> ``tests/golden/earlyfusion_golden.npz``).

Reference sources followed (file:line under /root/reference):
  acoss/algorithms/utils/cross_recurrence.py:31-48    get_csm (Euclidean)
  acoss/algorithms/utils/cross_recurrence.py:54-73    get_csm_cosine
  acoss/algorithms/utils/cross_recurrence.py:76-103   get_oti
  acoss/algorithms/utils/cross_recurrence.py:106-134  get_csm_blocked_oti
  acoss/algorithms/utils/cross_recurrence.py:137-161  csm_to_binary
  acoss/algorithms/utils/alignment_tools.py:8-46      delta_func / match / smith_waterman_constrained
  acoss/algorithms/utils/similarity_fusion.py:38-54   getWCSM
  acoss/algorithms/earlyfusion_traile.py:157-198      EarlyFusion.similarity (the four scores of one pair)
The last two are pinned by ``tests/golden/make_golden_earlyfusion_full.py`` ->
``tests/golden/earlyfusion_full_golden.npz`` (the reference's own ``EarlyFusion.similarity`` method executed
on seeded block features).
"""
from __future__ import annotations

import numpy as np

__all__ = ["get_oti", "get_csm", "get_csm_cosine", "get_csm_blocked_oti", "csm_to_binary",
           "nneighbs", "smith_waterman_constrained", "smith_waterman_constrained_x10",
           "get_wcsm", "early_fusion_csm", "similarity_pair"]


def get_oti(C1, C2) -> int:
    """argmax_i sum(roll(C1, i) * C2): rolls the FIRST song, first max wins
    (cross_recurrence.py:94-103)."""
    C1 = np.asarray(C1, dtype=np.float64)
    C2 = np.asarray(C2, dtype=np.float64)
    n = len(C1)
    scores = np.zeros(n)
    for i in range(n):
        scores[i] = np.sum(np.roll(C1, i) * C2)
    return int(np.argmax(scores))


def get_csm(X, Y):
    """sqrt(max(0, |x|^2 + |y|^2 - 2 X Y^T))  (cross_recurrence.py:45-48)."""
    X = np.asarray(X, dtype=np.float64)
    Y = np.asarray(Y, dtype=np.float64)
    C = np.sum(X ** 2, 1)[:, None] + np.sum(Y ** 2, 1)[None, :] - 2 * X.dot(Y.T)
    C[C < 0] = 0
    return np.sqrt(C)


def get_csm_cosine(X, Y):
    """1 - Xhat Yhat^T, zero norms replaced by 1 (cross_recurrence.py:67-73)."""
    X = np.asarray(X, dtype=np.float64)
    Y = np.asarray(Y, dtype=np.float64)
    xn = np.sqrt(np.sum(X ** 2, 1)); xn[xn == 0] = 1
    yn = np.sqrt(np.sum(Y ** 2, 1)); yn[yn == 0] = 1
    return 1 - (X / xn[:, None]).dot((Y / yn[:, None]).T)


def get_csm_blocked_oti(X, Y, C1, C2, csm_fn=get_csm_cosine):
    """Roll the chroma axis of every block of X by get_oti(C1, C2), then csm_fn
    (cross_recurrence.py:128-134)."""
    nb = len(C1)
    per = int(X.shape[1] / nb)
    oti = get_oti(C1, C2)
    X1 = np.reshape(X, (X.shape[0], per, nb))
    X1 = np.roll(X1, oti, axis=2)
    X1 = np.reshape(X1, [X.shape[0], per * nb])
    return csm_fn(X1, Y)


def nneighbs(kappa, n_cols: int) -> int:
    """Neighbour count rule of csm_to_binary (cross_recurrence.py:151-155): banker's
    rounding through np.round for kappa < 1, kappa itself otherwise."""
    if kappa < 1:
        return int(np.round(kappa * n_cols))
    return int(kappa)


def csm_to_binary(D, kappa):
    """Row-only k-NN binarisation (cross_recurrence.py:137-161).  kappa == 0 -> all ones.
    Exactly NN ones per row; ties at the NN-th value are broken by numpy's introselect
    (np.argpartition), which this restatement calls too."""
    D = np.asarray(D)
    if kappa == 0:
        return np.ones_like(D)
    nn = nneighbs(kappa, D.shape[1])
    J = np.argpartition(D, nn, 1)[:, 0:nn]
    out = np.zeros(D.shape, dtype=np.uint8)
    np.put_along_axis(out, J, 1, axis=1)
    return out


def smith_waterman_constrained(B) -> float:
    """Row-vectorised float64 restatement of alignment_tools.py:26-46.

    S[i][j] = max(0, d1, d2, d3), i,j >= 3, with mv = +1/-1 from B[i-1][j-1] and
      d1 = S[i-1][j-1] + mv + delta(B[i-2][j-2])
      d2 = S[i-2][j-1] + mv + delta(B[i-3][j-2])
      d3 = S[i-1][j-2] + mv + delta(B[i-2][j-3])
    delta(a) = 0 if a > 0 else -0.7 (the gap_opening branch of delta_func, :11-12, is
    unreachable).  Returns 0.0 when M < 4 or N < 4; IOError on non-binary input (:23).
    """
    B = np.asarray(B)
    M, N = B.shape
    if N < 4 or M < 4:
        return 0.0
    # the reference only inspects B[i-1][j-1] for i in 3..M-1, j in 3..N-1
    seen = B[2:M - 1, 2:N - 1]
    if not np.isin(seen, (0, 1)).all():
        raise IOError("Non-binary elements found in input")
    S = np.zeros((M, N))
    best = 0.0
    delta = np.where(B > 0, 0.0, -0.7)
    for i in range(3, M):
        mv = np.where(B[i - 1, 2:N - 1] == 1, 1.0, -1.0)
        d1 = S[i - 1, 2:N - 1] + mv + delta[i - 2, 1:N - 2]
        d2 = S[i - 2, 2:N - 1] + mv + delta[i - 3, 1:N - 2]
        d3 = S[i - 1, 1:N - 2] + mv + delta[i - 2, 0:N - 3]
        row = np.maximum(np.maximum(d1, d2), np.maximum(d3, 0.0))
        S[i, 3:] = row
        m = row.max()
        if m > best:
            best = float(m)
    return best


def smith_waterman_constrained_x10(B) -> int:
    """Exact integer (x10) form of the same recurrence: +10/-10 match, -7 gap.
    score_x10 / 10 equals the float64 result to ~1e-13 (checked against the reference)."""
    B = np.asarray(B)
    M, N = B.shape
    if N < 4 or M < 4:
        return 0
    S = np.zeros((M, N), dtype=np.int64)
    delta = np.where(B > 0, 0, -7).astype(np.int64)
    best = 0
    for i in range(3, M):
        mv = np.where(B[i - 1, 2:N - 1] == 1, 10, -10)
        d1 = S[i - 1, 2:N - 1] + mv + delta[i - 2, 1:N - 2]
        d2 = S[i - 2, 2:N - 1] + mv + delta[i - 3, 1:N - 2]
        d3 = S[i - 1, 1:N - 2] + mv + delta[i - 2, 0:N - 3]
        row = np.maximum(np.maximum(d1, d2), np.maximum(d3, 0))
        S[i, 3:] = row
        best = max(best, int(row.max()))
    return best


def get_wcsm(CSMAB, k1, k2, Mu=0.5):
    """Exponentially weighted cross-similarity of a cross-dissimilarity matrix
    (similarity_fusion.py:38-54): Eps = (mean of the k2 smallest of the row + mean of the k1 smallest
    of the column + d) / 3, W = exp(-d^2 / (2 (Mu Eps)^2)).  np.partition raises ValueError when
    k2 >= columns or k1 >= rows, as the reference does."""
    CSMAB = np.asarray(CSMAB)
    m1 = np.mean(np.partition(CSMAB, k2, 1)[:, 0:k2], 1)
    m2 = np.mean(np.partition(CSMAB, k1, 0)[0:k1, :], 0)
    Eps = m1[:, None] + m2[None, :] + CSMAB
    Eps /= 3
    return np.exp(-CSMAB ** 2 / (2 * (Mu * Eps) ** 2))


def early_fusion_csm(csms, K):
    """exp(-(sum of getWCSM(C, K, K) over the CSMs, in the order given)) — the "distance" matrix the
    early score binarises (earlyfusion_traile.py:178-183)."""
    total = np.zeros_like(csms[0])
    for C in csms:
        total += get_wcsm(C, K, K)
    return np.exp(-total)


def similarity_pair(f1, f2, kappa=0.1, K=10, want_matrices=False):
    """The four scores EarlyFusion.similarity stores for one pair (earlyfusion_traile.py:166-183):
    'mfccs' / 'ssms' Euclidean CSMs, 'chromas' blocked-OTI cosine CSM, 'early' = SW of the binarised
    exp(-sum of the three WCSMs).  f1 / f2: block-feature dictionaries of load_features."""
    csms = {
        "mfccs": get_csm(f1["mfccs"], f2["mfccs"]),
        "ssms": get_csm(f1["ssms"], f2["ssms"]),
        "chromas": get_csm_blocked_oti(f1["chromas"], f2["chromas"], f1["chroma_med"], f2["chroma_med"],
                                       get_csm_cosine),
    }
    mats = dict(csms)
    mats["early"] = early_fusion_csm([csms["mfccs"], csms["ssms"], csms["chromas"]], K)
    bins = {s: csm_to_binary(mats[s], kappa) for s in mats}
    scores = {s: smith_waterman_constrained(bins[s]) for s in mats}
    if want_matrices:
        return scores, mats, bins
    return scores
