"""numpy restatement of the reference's Serra09 pair score (TEST INFRASTRUCTURE).

PARITY UNPINNED: restates essentia ``ChromaCrossSimilarity`` +
``CoverSongSimilarity`` (third-party, unpinned ``'essentia'`` extra,
``/root/reference/setup.py:53``; absent from this image) exactly as SURVEY.md
Appendix A describes them, anchored on the reference call site
``/root/reference/acoss/algorithms/rqa_serra09.py:55-69``:

    crp_algo = ChromaCrossSimilarity(frameStackSize=m, frameStackStride=tau,
                                     binarizePercentile=kappa, oti=oti)   # :60-63
    alignment_algo = CoverSongSimilarity(alignmentType='serra09',
                                         distanceType='symmetric')       # :64
    csm = crp_algo(query, reference)                                     # :66
    _, score = alignment_algo(csm)                                       # :67

Every ambiguous point of the restatement (F1..F8 in SURVEY.md App. A) is an
explicit keyword switch whose default is the App. A value.

All arithmetic is float32 unless stated (essentia ``Real`` = float); dot
products multiply in float32 and accumulate sequentially in float64 (F3:
``std::inner_product(..., 0.0)``), narrowing to float32 on return.
"""
from __future__ import annotations

import numpy as np

F32 = np.float32

__all__ = [
    "global_chroma", "oti_index", "rotate_reference", "stack_frames",
    "dot_rows_f32_f64", "pairwise_distance", "percentile", "thresholds",
    "binarize", "chroma_cross_similarity", "qmax", "dmax", "dmax_bruteforce", "serra09_pair",
    "kappa_f32", "Serra09Error",
]


class Serra09Error(RuntimeError):
    """Stands in for essentia's ``EssentiaException`` (surfaces as RuntimeError)."""


def kappa_f32(kappa) -> np.float32:
    """binarizePercentile as essentia sees it: Real(kappa)*100 then /100. (App. A4).

    The float32 round trip gives back float32(kappa) for kappa=0.095 (checked in tests).
    """
    q = F32(F32(kappa) * F32(100))
    return F32(np.float64(q) / 100.0)


# --------------------------------------------------------------------------- A1
def global_chroma(frames: np.ndarray) -> np.ndarray:
    """Sequential float32 frame sum divided by its max (skip if max == 0). App. A1."""
    frames = np.ascontiguousarray(frames, dtype=F32)
    if frames.shape[0] == 0:
        raise Serra09Error("empty input")
    # cumsum is a strictly sequential float32 accumulation along axis 0
    g = np.cumsum(frames, axis=0, dtype=F32)[-1].astype(F32)
    mx = g.max()
    if mx != 0:
        g = (g / mx).astype(F32)
    return g


def _dot_f32_f64(a: np.ndarray, b: np.ndarray) -> np.float32:
    """essentia dotProduct: float32 products, sequential float64 accumulate (F3)."""
    acc = np.float64(0.0)
    p = (a.astype(F32) * b.astype(F32)).astype(F32)
    for v in p:
        acc = acc + np.float64(v)
    return F32(acc)


def oti_index(query: np.ndarray, reference: np.ndarray, noti: int = 12) -> int:
    """argmax_{s=0..noti} dot(g_q, rotR(g_r, s)), first max wins (F6). App. A1."""
    gq = global_chroma(query)
    gr = global_chroma(reference)
    best, best_s = None, 0
    for s in range(noti + 1):
        v = _dot_f32_f64(gq, np.roll(gr, s))
        if best is None or v > best:
            best, best_s = v, s
    return int(best_s)


def rotate_reference(reference: np.ndarray, oti: int) -> np.ndarray:
    """Every reference frame rotated right by ``oti`` bins (np.roll(x, oti))."""
    return np.roll(np.asarray(reference, dtype=F32), oti, axis=1)


# --------------------------------------------------------------------------- A2
def stack_frames(frames: np.ndarray, m: int = 9, tau: int = 1, *, drop_one: bool = True) -> np.ndarray:
    """Time-delay embedding.  F4: essentia emits n - m*tau rows (one fewer than natural)."""
    frames = np.ascontiguousarray(frames, dtype=F32)
    n = frames.shape[0]
    if m == 1:
        return frames
    incr = m * tau
    if n < incr + 1:
        raise Serra09Error("stackChromaFrames: not enough frames (%d) for m*tau=%d" % (n, incr))
    rows = n - incr if drop_one else n - (m - 1) * tau
    idx = np.arange(rows)[:, None] + tau * np.arange(m)[None, :]
    return frames[idx].reshape(rows, m * frames.shape[1])


# --------------------------------------------------------------------------- A3
def dot_rows_f32_f64(a: np.ndarray, b: np.ndarray, *, f64_accumulate: bool = True) -> np.ndarray:
    """All-pairs dotProduct(a[i], b[j]) with float32 products and a sequential
    accumulator (float64 by default, F3), narrowed to float32."""
    a = np.ascontiguousarray(a, dtype=F32)
    b = np.ascontiguousarray(b, dtype=F32)
    acc_t = np.float64 if f64_accumulate else F32
    acc = np.zeros((a.shape[0], b.shape[0]), dtype=acc_t)
    for k in range(a.shape[1]):
        p = a[:, k][:, None] * b[:, k][None, :]        # float32 product, one rounding
        acc += p.astype(acc_t)
    return acc.astype(F32)


def _self_dot(a: np.ndarray, f64_accumulate: bool = True) -> np.ndarray:
    acc_t = np.float64 if f64_accumulate else F32
    acc = np.zeros(a.shape[0], dtype=acc_t)
    for k in range(a.shape[1]):
        acc += (a[:, k] * a[:, k]).astype(acc_t)
    return acc.astype(F32)


def pairwise_distance(qs: np.ndarray, rs: np.ndarray, *, f64_accumulate: bool = True,
                      return_sq: bool = False):
    """d[i][j] = sqrtf( f32(f32(aa - 2ab) + bb) ).  Negative ``item`` -> NaN (F7)."""
    aa = _self_dot(qs, f64_accumulate)
    bb = _self_dot(rs, f64_accumulate)
    ab = dot_rows_f32_f64(qs, rs, f64_accumulate=f64_accumulate)
    item = (aa[:, None] - F32(2) * ab).astype(F32)
    item = (item + bb[None, :]).astype(F32)
    with np.errstate(invalid="ignore"):
        d = np.sqrt(item).astype(F32)
    if return_sq:
        return d, item
    return d


# --------------------------------------------------------------------------- A4
def percentile_k(length: int, q: np.float32) -> np.float32:
    """Fractional rank: (L-1)*q in float32 if L>1 else L*q."""
    if length > 1:
        return F32(F32(length - 1) * F32(q))
    return F32(F32(length) * F32(q))


def percentile(sorted_vals: np.ndarray, q: np.float32, *, integer_guard: bool = False) -> np.float32:
    """essentia percentile on an ascending-sorted float32 vector.

    F1: no floor==ceil guard => integer k gives thr = 0.  ``integer_guard=True``
    is the alternative (thr = s[k]).
    """
    s = np.asarray(sorted_vals, dtype=F32)
    L = s.shape[0]
    k = percentile_k(L, q)
    fk = F32(np.floor(k))
    ck = F32(np.ceil(k))
    if integer_guard and fk == ck:
        return F32(s[int(fk)])
    d0 = F32(s[int(fk)] * F32(ck - k))
    d1 = F32(s[int(ck)] * F32(k - fk))
    return F32(d0 + d1)


def thresholds(d: np.ndarray, kappa, *, integer_guard: bool = False):
    """thrQ[i] from row i, thrR[j] from column j of the float32 distance matrix."""
    q = kappa_f32(kappa)
    rows_sorted = np.sort(d, axis=1)
    cols_sorted = np.sort(d, axis=0)
    thr_q = np.array([percentile(rows_sorted[i], q, integer_guard=integer_guard)
                      for i in range(d.shape[0])], dtype=F32)
    thr_r = np.array([percentile(cols_sorted[:, j], q, integer_guard=integer_guard)
                      for j in range(d.shape[1])], dtype=F32)
    return thr_q, thr_r


# --------------------------------------------------------------------------- A5
def binarize(d: np.ndarray, thr_q: np.ndarray, thr_r: np.ndarray, *, strict: bool = False) -> np.ndarray:
    """csm = heaviside(thrQ[i]-d) * heaviside(thrR[j]-d), heaviside(x)=1 iff x>=0 (F2; ``strict=True`` is the
    alternative x>0, i.e. d == thr does not count as similar).
    Shape (M', N') = (query, reference) (F8).  uint8 {0,1}; NaN distances raise (F7)."""
    if np.isnan(d).any():
        raise Serra09Error("NaN distance (negative squared distance, F7): non-binary CRP")
    if strict:
        sx = (thr_q[:, None] - d) > 0
        sy = (thr_r[None, :] - d) > 0
    else:
        sx = (thr_q[:, None] - d) >= 0
        sy = (thr_r[None, :] - d) >= 0
    return (sx & sy).astype(np.uint8)


def chroma_cross_similarity(query, reference, *, m=9, tau=1, kappa=0.095, oti=True, noti=12,
                            f64_accumulate=True, integer_guard=False, drop_one=True, strict=False,
                            return_debug=False):
    """essentia ChromaCrossSimilarity (otiBinary=False, streaming=False). App. A1-A5."""
    query = np.ascontiguousarray(query, dtype=F32)
    reference = np.ascontiguousarray(reference, dtype=F32)
    if query.shape[0] == 0 or reference.shape[0] == 0:
        raise Serra09Error("empty input")
    s = 0
    if oti:
        s = oti_index(query, reference, noti)
        reference = rotate_reference(reference, s)
    qs = stack_frames(query, m, tau, drop_one=drop_one)
    rs = stack_frames(reference, m, tau, drop_one=drop_one)
    if qs.shape[0] < 2 or rs.shape[0] < 2:
        # F9: essentia's percentile() reads sorted[ceil(L*q)] = sorted[1] for a 1-element vector,
        # i.e. out of bounds (undefined behaviour).  The restatement refuses such inputs.
        raise Serra09Error("fewer than 2 stacked frames: essentia percentile() is undefined (F9)")
    d = pairwise_distance(qs, rs, f64_accumulate=f64_accumulate)
    thr_q, thr_r = thresholds(d, kappa, integer_guard=integer_guard)
    crp = binarize(d, thr_q, thr_r, strict=strict)
    if return_debug:
        return crp, dict(oti=s, d=d, thr_q=thr_q, thr_r=thr_r)
    return crp


# --------------------------------------------------------------------------- A6
def qmax(crp: np.ndarray, gamma_o: float = 0.5, gamma_e: float = 0.5, *, return_matrix=False,
         asymmetric: bool = False):
    """essentia CoverSongSimilarity alignmentType='serra09', distanceType='symmetric'
    (F5: returns max(Q), un-normalised; ``asymmetric=True`` is the other distanceType as App. A6 records it,
    float32 sqrt(N') / max(Q) with N' the reference axis).  Row-vectorised float32 restatement of

        Q[i][j] = max3(Q[i-1][j-1], Q[i-2][j-1], Q[i-1][j-2]) + 1              if crp[i][j]==1
                = max(0, Q[p] - gamma(crp[p]) for the same three p)            otherwise
    for i,j >= 2, gamma(1)=gamma_o (disOnset), gamma(0)=gamma_e (disExtension).
    """
    c = np.asarray(crp)
    if not np.isin(c, (0, 1)).all():
        raise Serra09Error("Non-binary elements found in input")
    c = c.astype(bool)
    M, N = c.shape
    Q = np.zeros((M, N), dtype=F32)
    go, ge = F32(gamma_o), F32(gamma_e)
    for i in range(2, M):
        p1 = Q[i - 1, 1:N - 1]; b1 = c[i - 1, 1:N - 1]
        p2 = Q[i - 2, 1:N - 1]; b2 = c[i - 2, 1:N - 1]
        p3 = Q[i - 1, 0:N - 2]; b3 = c[i - 1, 0:N - 2]
        hit = (np.maximum(np.maximum(p1, p2), p3) + F32(1)).astype(F32)
        g1 = np.where(b1, go, ge); g2 = np.where(b2, go, ge); g3 = np.where(b3, go, ge)
        miss = np.maximum(np.maximum(p1 - g1, p2 - g2), np.maximum(p3 - g3, F32(0))).astype(F32)
        Q[i, 2:] = np.where(c[i, 2:], hit, miss)
    score = F32(Q.max()) if Q.size else F32(0)
    if asymmetric:
        with np.errstate(divide="ignore"):
            score = F32(np.sqrt(F32(N)) / score)
    if return_matrix:
        return score, Q
    return score


def dmax(crp: np.ndarray, gamma_o: float = 0.5, gamma_e: float = 0.5, *, bonus: bool = True,
         return_matrix=False):
    """essentia CoverSongSimilarity alignmentType='chen17' (Dmax, Chen et al. 2017), 'symmetric'
    distance: the second score ChenFusion stores per pair (``latefusion_chen.py:69-73``).
    UNPINNED restatement (essentia is absent).  For i, j >= 3:

        P1 = D[i-1][j-1]
        P2 = D[i-2][j-1] + c[i-1][j]                     P3 = D[i-1][j-2] + c[i][j-1]
        P4 = D[i-3][j-1] + c[i-2][j] + c[i-1][j]         P5 = D[i-1][j-3] + c[i][j-2] + c[i][j-1]
        D[i][j] = max(P1..P5) + 1                                         if c[i][j] == 1
                = max(0, Pk - gamma(c at the predecessor cell of Pk))     otherwise

    **F10** ``bonus``: the ``+ c[..]`` terms of P2..P5 (Chen's bridging of one / two skipped cells)
    are how upstream essentia is recollected; ``bonus=False`` drops them (plain five-predecessor
    form).  float32 arithmetic, operations left to right as written."""
    c = np.asarray(crp)
    if not np.isin(c, (0, 1)).all():
        raise Serra09Error("Non-binary elements found in input")
    c = c.astype(bool)
    M, N = c.shape
    Q = np.zeros((M, N), dtype=F32)
    go, ge = F32(gamma_o), F32(gamma_e)
    cf = c.astype(F32)
    for i in range(3, M):
        sl = lambda r, dj: (r, slice(3 - dj, N - dj))          # cells (r, j - dj) for j = 3..N-1
        P = [Q[sl(i - 1, 1)].copy(), Q[sl(i - 2, 1)].copy(), Q[sl(i - 1, 2)].copy(),
             Q[sl(i - 3, 1)].copy(), Q[sl(i - 1, 3)].copy()]
        if bonus:
            P[1] = (P[1] + cf[sl(i - 1, 0)]).astype(F32)
            P[2] = (P[2] + cf[sl(i, 1)]).astype(F32)
            P[3] = ((P[3] + cf[sl(i - 2, 0)]).astype(F32) + cf[sl(i - 1, 0)]).astype(F32)
            P[4] = ((P[4] + cf[sl(i, 2)]).astype(F32) + cf[sl(i, 1)]).astype(F32)
        B = [c[sl(i - 1, 1)], c[sl(i - 2, 1)], c[sl(i - 1, 2)], c[sl(i - 3, 1)], c[sl(i - 1, 3)]]
        hit = (np.maximum.reduce(P) + F32(1)).astype(F32)
        pens = [(p - np.where(b, go, ge)).astype(F32) for p, b in zip(P, B)]
        miss = np.maximum(np.maximum.reduce(pens), F32(0)).astype(F32)
        Q[i, 3:] = np.where(c[i, 3:], hit, miss)
    score = F32(Q.max()) if Q.size else F32(0)
    if return_matrix:
        return score, Q
    return score


def dmax_bruteforce(crp: np.ndarray, gamma_o: float = 0.5, gamma_e: float = 0.5, *, bonus: bool = True):
    """Cell-by-cell definition of :func:`dmax` (pure Python, small inputs only)."""
    c = np.asarray(crp).astype(np.int64)
    M, N = c.shape
    D = np.zeros((M, N), dtype=F32)
    g = lambda v: F32(gamma_o) if v else F32(gamma_e)
    for i in range(3, M):
        for j in range(3, N):
            k = F32(1) if bonus else F32(0)
            P = [D[i - 1, j - 1],
                 F32(D[i - 2, j - 1] + k * F32(c[i - 1, j])),
                 F32(D[i - 1, j - 2] + k * F32(c[i, j - 1])),
                 F32(F32(D[i - 3, j - 1] + k * F32(c[i - 2, j])) + k * F32(c[i - 1, j])),
                 F32(F32(D[i - 1, j - 3] + k * F32(c[i, j - 2])) + k * F32(c[i, j - 1]))]
            if c[i, j] == 1:
                D[i, j] = F32(max(P) + F32(1))
            else:
                G = [g(c[i - 1, j - 1]), g(c[i - 2, j - 1]), g(c[i - 1, j - 2]), g(c[i - 3, j - 1]), g(c[i - 1, j - 3])]
                D[i, j] = max([F32(0)] + [F32(p - q) for p, q in zip(P, G)])
    return F32(D.max()) if D.size else F32(0)


def serra09_pair(query, reference, *, m=9, tau=1, kappa=0.095, oti=True, gamma_o=0.5, gamma_e=0.5,
                 **flags) -> np.float32:
    """One ``Serra09.similarity`` pair: the value written to ``Ds['main'][i][j]``
    (``rqa_serra09.py:66-69``)."""
    asym = bool(flags.pop("asymmetric", False))
    crp = chroma_cross_similarity(query, reference, m=m, tau=tau, kappa=kappa, oti=oti, **flags)
    return qmax(crp, gamma_o, gamma_e, asymmetric=asym)
