"""Serra09 oracle self-consistency.  PARITY UNPINNED: essentia cannot be executed here, so
these tests pin the oracle to (a) SURVEY.md Appendix B2's restatement known answers, (b) a
brute-force per-cell definition, (c) the plain-C port used as CPU baseline."""
import hashlib

import numpy as np
import pytest

from oracle import serra09_c as oc
from oracle import serra09_np as o

F32 = np.float32


def hp(rng, n):
    X = rng.random((n, 12)).astype(F32)
    return (X / X.max(1, keepdims=True)).astype(F32)


B2 = [  # seed, nq, nr, oti, shape, d.sum, thrQ0, thrR0, ones, row0 ones pre-AND, sha1, qmax
    (100, 60, 50, 2, (51, 41), 9303.165169, 4.0898261, 3.9494464, 119, 4, "fb67617394da", 11.0),
    (101, 45, 80, 2, (36, 71), 11523.283125, 4.1037178, 4.1747255, 181, 7, "b3f008ee2613", 15.5),
    (102, 30, 30, 2, (21, 21), 1924.135206, 4.0345597, 4.2414074, 28, 2, "0358209a63e8", 5.0),
]


@pytest.mark.parametrize("case", B2)
def test_b2_restatement_kats(case):
    seed, nq, nr, oti, shape, dsum, tq0, tr0, ones, row0, sha, qm = case
    rng = np.random.default_rng(seed)
    Q = hp(rng, nq); R = hp(rng, nr)
    crp, dbg = o.chroma_cross_similarity(Q, R, return_debug=True)
    assert dbg["oti"] == oti and crp.shape == shape
    assert float(dbg["d"].astype(np.float64).sum()) == pytest.approx(dsum, abs=5e-6)
    assert dbg["thr_q"][0] == F32(tq0) and dbg["thr_r"][0] == F32(tr0)
    assert int(crp.sum()) == ones
    assert int(((dbg["thr_q"][0] - dbg["d"][0]) >= 0).sum()) == row0
    assert hashlib.sha1(np.packbits(crp, axis=1).tobytes()).hexdigest()[:12] == sha
    assert float(o.qmax(crp)) == qm
    assert float(o.serra09_pair(Q, R)) == qm


def test_planted_cover():
    rng = np.random.default_rng(103)
    Q = hp(rng, 70)
    R = np.roll(Q, 3, axis=1) + F32(0.01) * rng.random((70, 12)).astype(F32)
    R = (R / R.max(1, keepdims=True)).astype(F32)
    crp, dbg = o.chroma_cross_similarity(Q, R, return_debug=True)
    assert dbg["oti"] == 9 and int(crp.sum()) == 234 and int(np.trace(crp)) == 61
    assert float(o.qmax(crp)) == 59.0


def test_kappa_roundtrip_and_f1():
    assert o.kappa_f32(0.095) == F32(0.095)
    s = np.sort(np.random.default_rng(0).random(201).astype(F32))
    assert o.percentile(s, o.kappa_f32(0.095)) == 0.0            # F1: k = 19.0 exactly
    assert o.percentile(s, o.kappa_f32(0.095), integer_guard=True) == s[19]
    ks = [L for L in range(2, 8300) if float(o.percentile_k(L, F32(0.095))).is_integer()]
    assert ks[:3] == [201, 401, 601] and len(ks) == 41
    assert abs(float(o.percentile_k(1991, F32(0.095))) - 189.05) < 1e-3


def test_bruteforce_definition():
    """Per-cell scalar evaluation in essentia's order == vectorised restatement."""
    rng = np.random.default_rng(7)
    Q = hp(rng, 31); R = hp(rng, 27)
    crp, dbg = o.chroma_cross_similarity(Q, R, return_debug=True)
    Rr = np.roll(R, dbg["oti"], axis=1)
    M, N = crp.shape
    d = np.zeros((M, N), F32)

    def dot(a, b):
        acc = np.float64(0)
        for x, y in zip(a, b):
            acc += np.float64(F32(x * y))
        return F32(acc)
    for i in range(M):
        a = Q[i:i + 9].ravel()
        for j in range(N):
            b = Rr[j:j + 9].ravel()
            item = F32(F32(dot(a, a) - F32(2) * dot(a, b)) + dot(b, b))
            d[i, j] = np.sqrt(item)
    assert np.array_equal(d, dbg["d"])
    q = o.kappa_f32(0.095)
    for i in range(M):
        assert o.percentile(np.sort(d[i]), q) == dbg["thr_q"][i]
    # brute-force Qmax, cell by cell
    Qm = np.zeros((M, N), F32)
    for i in range(2, M):
        for j in range(2, N):
            p = [(i - 1, j - 1), (i - 2, j - 1), (i - 1, j - 2)]
            if crp[i, j] == 1:
                Qm[i, j] = max(Qm[x] for x in p) + F32(1)
            else:
                Qm[i, j] = max([F32(0)] + [Qm[x] - F32(0.5) for x in p])
    assert float(Qm.max()) == float(o.qmax(crp))


@pytest.mark.parametrize("seed,nq,nr", [(100, 60, 50), (101, 45, 80), (5, 210, 333), (6, 19, 11),
                                        (8, 11, 11), (9, 150, 410)])
def test_c_port_matches_numpy(seed, nq, nr):
    rng = np.random.default_rng(seed)
    Q = hp(rng, nq); R = hp(rng, nr)
    crp, dbg = o.chroma_cross_similarity(Q, R, return_debug=True)
    s, d = oc.pair(Q, R, want_debug=True)
    assert d["oti"] == dbg["oti"]
    assert np.array_equal(d["d"], dbg["d"])
    assert np.array_equal(d["thr_q"], dbg["thr_q"]) and np.array_equal(d["thr_r"], dbg["thr_r"])
    assert np.array_equal(d["crp"], crp)
    assert s == float(o.qmax(crp))
    assert oc.pair(Q, R, oc.params(hoist_norms=True)) == s
    assert oc.oti(Q, R) == dbg["oti"]


def test_gamma_variants_and_qmax_c():
    rng = np.random.default_rng(3)
    for go, ge in [(0.5, 0.5), (0.5, 0.7), (1.0, 0.25)]:
        c = (rng.random((57, 83)) < 0.12).astype(np.uint8)
        assert float(o.qmax(c, go, ge)) == pytest.approx(oc.qmax(c, go, ge), abs=0)
    c = (rng.random((40, 40)) < 0.1).astype(np.uint8)
    assert float(o.qmax(c)) == float(o.qmax(c.T))     # equal penalties: transpose invariant (F8)


def test_errors_and_edges():
    rng = np.random.default_rng(4)
    with pytest.raises(o.Serra09Error):
        o.chroma_cross_similarity(hp(rng, 9), hp(rng, 30))      # n < m*tau + 1
    with pytest.raises(RuntimeError):
        oc.pair(hp(rng, 9), hp(rng, 30))
    with pytest.raises(o.Serra09Error):
        o.chroma_cross_similarity(np.zeros((0, 12), F32), hp(rng, 30))
    with pytest.raises(o.Serra09Error):
        o.qmax(np.full((5, 5), 2))
    with pytest.raises(o.Serra09Error):                          # F9: 1 stacked frame
        o.chroma_cross_similarity(hp(rng, 10), hp(rng, 11))
    crp = o.chroma_cross_similarity(hp(rng, 11), hp(rng, 12))
    assert crp.shape == (2, 3) and float(o.qmax(crp)) == 0.0
    # all-zero frames: distances 0, thresholds 0, CRP all ones
    z = np.zeros((20, 12), F32)
    crp = o.chroma_cross_similarity(z, z)
    assert crp.all() and float(o.qmax(crp)) == 9.0


def test_c_batch_threads():
    rng = np.random.default_rng(11)
    lens = [40, 55, 33, 61]
    tracks = [hp(rng, n) for n in lens]
    frames = np.concatenate(tracks)
    offs = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    pr = np.array([(i, j) for i in range(4) for j in range(i + 1, 4)], np.int32)
    got = oc.pairs(frames, offs, pr, nthreads=3)
    want = [float(o.serra09_pair(tracks[i], tracks[j])) for i, j in pr]
    assert list(got) == want


def test_sw_c_port(golden_dir):
    import os
    g = np.load(os.path.join(golden_dir, "earlyfusion_golden.npz"))
    for (seed, m, n, p), (s, _) in zip(g["sw_meta"][:6], g["sw_scores"][:6]):
        B = (np.random.default_rng(int(seed)).random((int(m), int(n))) < p).astype(np.uint8)
        assert oc.sw_constrained(B) == pytest.approx(s, abs=1e-9)


def test_dmax_restatement_self_consistency():
    """Row-vectorised Dmax == cell-by-cell definition == plain-C port, with and without Chen's bridging
    terms (F10), default and general gammas; known answers on structured matrices."""
    from oracle import serra09_c as oc
    from oracle import serra09_np as o
    rng = np.random.default_rng(3)
    for shape, p in [((40, 50), 0.15), ((30, 30), 0.3), ((5, 60), 0.2), ((64, 64), 0.1), ((3, 3), 0.5), ((4, 7), 0.9)]:
        c = (rng.random(shape) < p).astype(np.uint8)
        for bonus in (True, False):
            for g in ((0.5, 0.5), (0.5, 0.7), (0.3, 0.9)):
                a = float(o.dmax(c, *g, bonus=bonus))
                assert a == float(o.dmax_bruteforce(c, *g, bonus=bonus)) == oc.dmax(c, *g, bonus=bonus)
    assert float(o.dmax(np.eye(50, dtype=np.uint8))) == 47.0          # the diagonal alone: M - 3
    assert float(o.dmax(np.ones((50, 50), np.uint8))) == 71.0        # bridging terms add 3 per 2 rows
    assert float(o.dmax(np.ones((50, 50), np.uint8), bonus=False)) == 47.0
    assert float(o.dmax(np.ones((3, 9), np.uint8))) == 0.0           # DP starts at (3, 3)
    with pytest.raises(o.Serra09Error):
        o.dmax(np.full((5, 5), 2))


def test_chen_pairs_c_matches_numpy():
    from oracle import serra09_c as oc
    from oracle import serra09_np as o
    rng = np.random.default_rng(8)
    Q, R = hp(rng, 70), hp(rng, 55)
    frames = np.concatenate([Q, R]); offs = np.array([0, 70, 125], np.int64)
    q, d = oc.chen_pairs(frames, offs, np.array([[0, 1]], np.int32))
    crp = o.chroma_cross_similarity(Q, R)
    assert float(q[0]) == float(o.qmax(crp)) and float(d[0]) == float(o.dmax(crp))
