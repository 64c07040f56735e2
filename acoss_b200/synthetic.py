"""Synthetic HPCP datasets of the shapes BASELINE.json names.

Frames are generated directly at post-``load_features`` resolution (``downsample_fac=1``): float32,
12 bins, non-negative, per-frame max-normalised to 1 like essentia's HPCP.  A clique is one seeded
chord-template random walk; each cover of it is the base sequence circularly shifted by a random
number of bins (exercises OTI), linearly time-warped, plus |N(0, 0.1)| noise, renormalised.
Singleton tracks are independent walks.  Clique layouts follow the reference's annotation files
(acoss/data/covers80_annotations.csv, acoss/data/da-tacos_benchmark_subset.csv).
"""
from __future__ import annotations

import numpy as np

__all__ = ["CONFIGS", "make_dataset", "config_dataset", "all_pairs_upper", "pair_cells"]

# name -> (clique sizes, nominal length L, seed)   
CONFIGS = {
    "C1": dict(cliques=[2] * 80, L=2000, seed=20241),                       # covers80-shaped, 160 tracks
    "C3": dict(cliques=[13] * 66 + [1] * 142, L=2000, seed=20243),          # 1 000-track Da-TACOS slice
    "C4": dict(cliques=[13] * 1000 + [1] * 2000, L=2000, seed=20244),       # 15 000 tracks
    "C4s": dict(cliques=[13] * 1000 + [1] * 2000, L=500, seed=20244),       # same, realistic x40-downsampled length
    "C5": dict(cliques=[13] * 38 + [1] * 6, L=8000, seed=20245),            # 500 tracks x ~8k frames
    "tiny": dict(cliques=[3] * 4 + [1] * 4, L=120, seed=20240),             # 16 tracks, unit tests
}

_TEMPLATES = None


def _chord_templates() -> np.ndarray:
    global _TEMPLATES
    if _TEMPLATES is None:
        t = []
        for root in range(12):
            for third in (4, 3):                      # major / minor triads
                v = np.full(12, 0.05)
                v[root] = 1.0
                v[(root + third) % 12] = 0.7
                v[(root + 7) % 12] = 0.8
                v[(root + 10) % 12] += 0.1            # a little seventh colour
                t.append(v)
        _TEMPLATES = np.asarray(t, dtype=np.float64)
    return _TEMPLATES


def _walk(rng: np.random.Generator, n: int) -> np.ndarray:
    """Chord-template random walk, n frames, float64 un-normalised."""
    T = _chord_templates()
    dwell = rng.geometric(1.0 / 8.0, size=n // 2 + 2)
    chords = np.empty(len(dwell), dtype=np.int64)
    c = rng.integers(0, 24)
    for i in range(len(dwell)):
        chords[i] = c
        step = rng.choice([14, 10, 7 * 2, 5 * 2, 1, 23, 3])   # fifths / fourths / neighbours (in template index space)
        c = (c + step) % 24
    idx = np.repeat(chords, dwell)[:n]
    if len(idx) < n:
        idx = np.concatenate([idx, np.full(n - len(idx), idx[-1])])
    X = T[idx] * rng.uniform(0.6, 1.0, size=(n, 1))
    # smooth chord changes a little and add broadband energy
    X[1:] = 0.7 * X[1:] + 0.3 * X[:-1]
    X += 0.15 * rng.random((n, 12))
    return X


def _normalise(X: np.ndarray) -> np.ndarray:
    X = np.maximum(X, 0.0)
    mx = X.max(axis=1, keepdims=True)
    mx[mx == 0] = 1.0
    return (X / mx).astype(np.float32)


def _cover(rng: np.random.Generator, base: np.ndarray, n: int) -> np.ndarray:
    nb = base.shape[0]
    pos = np.linspace(0, nb - 1, n)
    i0 = np.floor(pos).astype(np.int64)
    i1 = np.minimum(i0 + 1, nb - 1)
    w = (pos - i0)[:, None]
    X = base[i0] * (1 - w) + base[i1] * w
    X = np.roll(X, int(rng.integers(0, 12)), axis=1)
    X = X + np.abs(rng.normal(0.0, 0.1, size=X.shape))
    return X


def make_dataset(cliques, L: int, seed: int, length_jitter: float = 0.1):
    """-> (tracks: list of (n,12) float32, labels: list[int]) ; lengths ~ U{(1-j)L .. (1+j)L}."""
    rng = np.random.default_rng(seed)
    tracks, labels = [], []
    lo, hi = int(round((1 - length_jitter) * L)), int(round((1 + length_jitter) * L))
    for cid, size in enumerate(cliques):
        base = _walk(rng, int(rng.integers(lo, hi + 1)))
        for k in range(size):
            n = int(rng.integers(lo, hi + 1))
            X = _cover(rng, base, n) if size > 1 else base[:n] if n <= len(base) else _cover(rng, base, n)
            tracks.append(_normalise(X))
            labels.append(cid)
    return tracks, labels


def config_dataset(name: str, max_tracks: int | None = None):
    cfg = CONFIGS[name]
    cl = list(cfg["cliques"])
    if max_tracks is not None:
        out, tot = [], 0
        for s in cl:
            if tot + s > max_tracks:
                break
            out.append(s)
            tot += s
        cl = out
    return make_dataset(cl, cfg["L"], cfg["seed"])


def all_pairs_upper(n: int) -> np.ndarray:
    """(i, j) with i < j in the order itertools.combinations yields them
    (algorithm_template.py:168-169), as an int32 (n(n-1)/2, 2) array."""
    i, j = np.triu_indices(n, k=1)
    return np.stack([i, j], axis=1).astype(np.int32)


def pair_cells(lengths, pairs, incr: int = 9) -> int:
    """sum over pairs of M' * N' = (n_q - m*tau)(n_r - m*tau)  (GCUPS numerator)."""
    ln = np.asarray(lengths, dtype=np.int64) - incr
    p = np.asarray(pairs)
    return int((ln[p[:, 0]] * ln[p[:, 1]]).sum())


# ---------------------------------------------------------------------------------------------
# EarlyFusion block features (what EarlyFusion.load_features returns, earlyfusion_traile.py:66-155)
# ---------------------------------------------------------------------------------------------
EF_DIMS = dict(mfccs=20 * 50, ssms=50 * 49 // 2, chromas=12 * 40)     # reference defaults: 1000, 1225, 480


def ef_dataset(cliques, n_blocks: int, seed: int, dims=None, dtype=np.float32, jitter: float = 0.15):
    """Synthetic beat-synchronous block features, one dict per track with the reference's keys
    ('mfccs', 'ssms', 'chromas', 'chroma_med', 'label').  A clique is one smooth latent walk; a cover
    is the walk linearly time-warped plus noise, its chroma blocks rolled by a random number of bins
    (exercises the blocked OTI).  Row counts ~ U{(1-j) n_blocks .. (1+j) n_blocks}.  dtype float32 is
    what the reference stores; float64 features exercise the same arithmetic without input rounding."""
    dims = dict(EF_DIMS if dims is None else dims)
    if dims["chromas"] % 12:
        raise ValueError("chroma block dimension must be a multiple of 12")
    rng = np.random.default_rng(seed)
    L = 8
    proj = {k: rng.normal(0.0, 1.0, size=(L, d)) / np.sqrt(L) for k, d in dims.items()}
    lo, hi = max(8, int(round((1 - jitter) * n_blocks))), int(round((1 + jitter) * n_blocks))
    feats = []
    for cid, size in enumerate(cliques):
        nb = int(rng.integers(lo, hi + 1))
        z = np.cumsum(rng.normal(0.0, 0.35, size=(nb, L)), axis=0)
        z -= z.mean(0)
        for _ in range(size):
            n = int(rng.integers(lo, hi + 1))
            pos = np.linspace(0, nb - 1, n)
            i0 = np.floor(pos).astype(np.int64)
            i1 = np.minimum(i0 + 1, nb - 1)
            w = (pos - i0)[:, None]
            zz = z[i0] * (1 - w) + z[i1] * w if size > 1 else z[np.minimum(np.arange(n), nb - 1)]
            f = {}
            for k in ("mfccs", "ssms"):
                f[k] = (np.tanh(zz @ proj[k]) + rng.normal(0.0, 0.25, size=(n, dims[k]))).astype(dtype)
            c = np.abs(np.tanh(zz @ proj["chromas"])) + 0.2 * rng.random((n, dims["chromas"]))
            shift = int(rng.integers(0, 12)) if size > 1 else 0
            c = np.roll(c.reshape(n, -1, 12), shift, axis=2).reshape(n, -1).astype(dtype)
            f["chromas"] = c
            f["chroma_med"] = np.median(c.reshape(-1, 12), axis=0)
            f["label"] = "w%d" % cid
            feats.append(f)
    return feats
