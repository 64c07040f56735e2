"""GPU parity tests (run with -m gpu on the B200 box): the CUDA path, called through the C ABI
(acoss_b200.Engine -> libacoss_b200.so), against the CPU oracle on the same seeded inputs."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

F32 = np.float32


def hp(rng, n):
    X = rng.random((n, 12)).astype(F32)
    return (X / X.max(1, keepdims=True)).astype(F32)


@pytest.fixture(scope="module")
def eng():
    from acoss_b200 import Engine
    e = Engine(0)
    yield e
    e.close()


@pytest.fixture(scope="module")
def golden(golden_dir):
    return np.load(os.path.join(golden_dir, "earlyfusion_golden.npz"))


def _set(eng, tracks):
    from acoss_b200 import pack_tracks
    frames, offs = pack_tracks(tracks)
    eng.set_tracks(frames, offs)


# ------------------------------------------------------------------------------------------- K1
def test_oti_exact(eng):
    from oracle import serra09_np as o
    rng = np.random.default_rng(1)
    tracks = [hp(rng, int(n)) for n in rng.integers(20, 400, size=24)]
    # planted transpositions
    for s in range(12):
        tracks.append(np.roll(tracks[s], s, axis=1))
    _set(eng, tracks)
    pairs = np.array([(i, j) for i in range(len(tracks)) for j in range(len(tracks)) if i != j], np.int32)
    got = eng.oti_pairs(pairs)
    want = np.array([o.oti_index(tracks[i], tracks[j]) for i, j in pairs])
    assert np.array_equal(got, want)
    for s in range(12):                     # rotR(reference, oti) undoes the planted roll
        assert eng.oti_pairs([(s, 24 + s)])[0] == (12 - s) % 12


# ------------------------------------------------------------------------------------------- K2 + K3, one pair
PAIR_CASES = [(100, 60, 50), (101, 45, 80), (102, 30, 30), (5, 210, 333), (6, 19, 11), (8, 11, 11),
              (9, 150, 410), (12, 300, 157), (13, 700, 1100)]


@pytest.mark.parametrize("path", ["exact", "auto"])
@pytest.mark.parametrize("seed,nq,nr", PAIR_CASES)
def test_pair_bit_exact(eng, seed, nq, nr, path):
    """OTI exact, thresholds bit-identical, CRP bit-exact, score identical (integer DP)."""
    from acoss_b200 import default_params
    from acoss_b200._lib import CRP_AUTO, CRP_EXACT
    from oracle import serra09_c as oc
    rng = np.random.default_rng(seed)
    Q = hp(rng, nq); R = hp(rng, nr)
    _set(eng, [Q, R])
    p = default_params(crp_path=CRP_EXACT if path == "exact" else CRP_AUTO)
    got = eng.dump_pair(0, 1, p)
    score, dbg = oc.pair(Q, R, want_debug=True)
    assert got["oti"] == dbg["oti"]
    assert np.array_equal(got["thr_q"], dbg["thr_q"])
    assert np.array_equal(got["thr_r"], dbg["thr_r"])
    mism = np.argwhere(got["crp"] != dbg["crp"])
    assert len(mism) == 0, "CRP mismatches at %s" % mism[:10]
    assert got["score"] == score
    assert eng.score_pairs([(0, 1)], p)[0] == score


def test_planted_cover_and_f1(eng):
    from oracle import serra09_c as oc
    rng = np.random.default_rng(103)
    Q = hp(rng, 70)
    R = np.roll(Q, 3, axis=1) + F32(0.01) * rng.random((70, 12)).astype(F32)
    R = (R / R.max(1, keepdims=True)).astype(F32)
    Z = hp(rng, 210)                              # n' = 201: F1 quirk, thresholds 0 on that axis
    _set(eng, [Q, R, Z])
    d = eng.dump_pair(0, 1)
    assert d["oti"] == 9 and int(d["crp"].sum()) == 234 and d["score"] == 59.0
    d = eng.dump_pair(0, 2)
    s, dbg = oc.pair(Q, Z, want_debug=True)
    assert np.array_equal(d["crp"], dbg["crp"]) and d["score"] == s
    assert (dbg["thr_q"] == 0).all()             # rows have L = 201 entries -> k = 19.0 -> thr 0


def test_all_zero_and_duplicate_frames(eng):
    from oracle import serra09_c as oc
    rng = np.random.default_rng(21)
    z = np.zeros((25, 12), F32)
    dup = np.repeat(hp(rng, 8), 5, axis=0)       # runs of identical frames -> tied distances
    mix = np.concatenate([hp(rng, 30), z[:12], hp(rng, 20)])
    _set(eng, [z, dup, mix])
    tracks = [z, dup, mix]
    for a, b in [(0, 0), (1, 1), (2, 2), (1, 2), (2, 1)]:
        try:
            s, dbg = oc.pair(tracks[a], tracks[b], want_debug=True)
        except RuntimeError:
            continue                              # NaN distance (F7) cases are covered below
        d = eng.dump_pair(a, b)
        assert np.array_equal(d["crp"], dbg["crp"]) and d["score"] == s
    d = eng.dump_pair(0, 0)
    assert d["crp"].all() and d["score"] == 14.0  # all-zero frames: all ones, Qmax = M' - 2


def test_errors(eng):
    from acoss_b200 import AcossError
    from acoss_b200._lib import E_TOO_SHORT
    rng = np.random.default_rng(4)
    _set(eng, [hp(rng, 10), hp(rng, 30)])         # 10 frames -> 1 stacked frame (F9)
    with pytest.raises(AcossError) as ei:
        eng.score_pairs([(0, 1)])
    assert ei.value.code == E_TOO_SHORT
    _set(eng, [hp(rng, 11), hp(rng, 12)])
    d = eng.dump_pair(0, 1)
    assert d["crp"].shape == (2, 3) and d["score"] == 0.0


def test_score_pairs_batch(eng):
    """A whole tiny dataset through acoss_score_pairs vs the C oracle, both CRP paths."""
    from acoss_b200 import default_params, synthetic
    from acoss_b200._lib import CRP_EXACT
    from oracle import serra09_c as oc
    tracks, labels = synthetic.config_dataset("tiny")
    _set(eng, tracks)
    pairs = synthetic.all_pairs_upper(len(tracks))
    from acoss_b200 import pack_tracks
    frames, offs = pack_tracks(tracks)
    want = oc.pairs(frames, offs, pairs, nthreads=8)
    got = eng.score_pairs(pairs)
    assert np.array_equal(got, want)
    got = eng.score_pairs(pairs, default_params(crp_path=CRP_EXACT))
    assert np.array_equal(got, want)
    assert want.max() > 20                        # covers actually align


# ------------------------------------------------------------------------------------------- K3 alone
def test_sw_golden(eng, golden):
    from acoss_b200.engine import ALIGN_SW
    mats, want = [], []
    for (seed, m, n, p), (s, st) in zip(golden["sw_meta"], golden["sw_scores"]):
        B = (np.random.default_rng(int(seed)).random((int(m), int(n))) < p).astype(np.uint8)
        mats += [B, np.ascontiguousarray(B.T)]
        want += [s, st]
    mats += [np.eye(50, dtype=np.uint8), np.ones((50, 50), np.uint8), np.ones((3, 10), np.uint8)]
    want += [float(golden["sw_eye50"]), float(golden["sw_ones50"]), 0.0]
    got = eng.dp_bytes(mats, ALIGN_SW)
    assert np.allclose(got, want, rtol=1e-5, atol=0)      # north-star: scores within 1e-5 relative
    assert np.array_equal(np.round(got * 10), np.round(np.array(want) * 10))


def test_sw_pipeline_golden(eng, golden):
    from acoss_b200.engine import ALIGN_SW
    for seed in (10, 11):
        got = eng.dp_bytes([golden["pipe%d_bin" % seed]], ALIGN_SW)[0]
        assert got == pytest.approx(float(golden["pipe%d_score" % seed]), rel=1e-5)


@pytest.mark.parametrize("shape", [(64, 64), (300, 1030), (130, 2100), (90, 3100), (70, 4200), (40, 9000),
                                   (2500, 70), (3, 3), (4, 4), (2, 50), (50, 2), (5, 33), (33, 34)])
def test_qmax_and_sw_shapes(eng, shape):
    """Strip boundaries (1024/2048/3072 columns), multi-strip halos, degenerate sizes."""
    from acoss_b200.engine import ALIGN_QMAX, ALIGN_SW
    from oracle import earlyfusion_np as ef
    from oracle import serra09_c as oc
    rng = np.random.default_rng(sum(shape))
    mats = [(rng.random(shape) < p).astype(np.uint8) for p in (0.05, 0.12, 0.4)]
    got = eng.dp_bytes(mats, ALIGN_QMAX)
    want = [oc.qmax(m) for m in mats]
    assert list(got) == want
    got = eng.dp_bytes(mats, ALIGN_SW)
    want = [ef.smith_waterman_constrained_x10(m) / 10.0 for m in mats]
    assert np.allclose(got, want, rtol=1e-6, atol=0)


def test_qmax_general_gamma_scalar_path(eng):
    from acoss_b200.engine import ALIGN_QMAX
    from oracle import serra09_c as oc
    rng = np.random.default_rng(77)
    mats = [(rng.random(s) < 0.12).astype(np.uint8) for s in [(57, 83), (200, 300), (90, 700)]]
    for go, ge in [(0.5, 0.7), (1.0, 0.25), (0.3, 0.3)]:
        got = eng.dp_bytes(mats, ALIGN_QMAX, go, ge)
        want = [oc.qmax(m, go, ge) for m in mats]
        assert list(got) == want


def test_sw_large_uses_scalar_path(eng):
    from acoss_b200.engine import ALIGN_SW
    from oracle import earlyfusion_np as ef
    B = np.eye(3400, dtype=np.uint8)              # score 3397 > int16/10 range
    assert eng.dp_bytes([B], ALIGN_SW)[0] == pytest.approx(ef.smith_waterman_constrained_x10(B) / 10.0, rel=1e-6)


def test_nonbinary_raises(eng):
    from acoss_b200 import AcossError
    from acoss_b200._lib import E_NONBINARY
    from acoss_b200.engine import ALIGN_SW
    with pytest.raises(AcossError) as ei:
        eng.dp_bytes([np.full((5, 5), 2)], ALIGN_SW)
    assert ei.value.code == E_NONBINARY


@pytest.mark.parametrize("shape", [(64, 64), (300, 1030), (130, 600), (70, 2100), (900, 70), (3, 3), (4, 4),
                                   (4, 50), (50, 4), (5, 33), (259, 259), (260, 261)])
def test_dmax_shapes(eng, shape):
    """Dmax (chen17): strip boundaries at 256 columns, multi-strip halos, degenerate sizes, F10 switch,
    general gammas — float32 results identical to the oracle."""
    from acoss_b200.engine import ALIGN_DMAX, ALIGN_DMAX_PLAIN
    from oracle import serra09_c as oc
    rng = np.random.default_rng(sum(shape) + 1)
    mats = [(rng.random(shape) < p).astype(np.uint8) for p in (0.05, 0.12, 0.4)]
    assert list(eng.dp_bytes(mats, ALIGN_DMAX)) == [oc.dmax(m) for m in mats]
    assert list(eng.dp_bytes(mats, ALIGN_DMAX_PLAIN)) == [oc.dmax(m, bonus=False) for m in mats]
    assert list(eng.dp_bytes(mats, ALIGN_DMAX, 0.5, 0.7)) == [oc.dmax(m, 0.5, 0.7) for m in mats]


def test_dmax_known_answers(eng):
    from acoss_b200.engine import ALIGN_DMAX
    got = eng.dp_bytes([np.eye(50, dtype=np.uint8), np.ones((50, 50), np.uint8)], ALIGN_DMAX)
    assert list(got) == [47.0, 71.0]              # oracle/serra09_np.py::dmax (tests/test_oracle_serra09.py)


def test_chen_pairs_both_scores(eng):
    """acoss_score_pairs_chen: Qmax and Dmax of the same CRPs equal the oracle's, and Qmax equals the
    Serra09 entry point."""
    from acoss_b200 import pack_tracks
    from oracle import serra09_c as oc
    rng = np.random.default_rng(31)
    tracks = [hp(rng, n) for n in (120, 333, 280, 64, 410)]
    tracks.append(np.roll(tracks[1], 4, axis=1))
    frames, offs = pack_tracks(tracks)
    eng.set_tracks(frames, offs)
    pairs = np.array([(i, j) for i in range(len(tracks)) for j in range(i + 1, len(tracks))], np.int32)
    q, d = eng.score_pairs_chen(pairs)
    wq, wd = oc.chen_pairs(frames, offs, pairs, nthreads=8)
    assert np.array_equal(q, wq) and np.array_equal(d, wd)
    assert np.array_equal(q, eng.score_pairs(pairs))


# ------------------------------------------------------------------------------------------- C2: SW over emitted CRPs
@pytest.mark.parametrize("path", ["auto", "exact"])
def test_sw_over_emitted_crps(eng, path):
    """BASELINE.json configs[1]: Smith-Waterman (alignment_tools.py:26-46) over the binary CRPs the Serra09
    stage emits, as one batched pair-pipeline call (align = ACOSS_ALIGN_SW) — against the oracle's SW of the
    oracle's CRP of every pair; ragged lengths incl. CRPs with fewer than 4 rows / columns (score 0.0)."""
    from acoss_b200 import default_params, synthetic
    from acoss_b200._lib import CRP_AUTO, CRP_EXACT
    from acoss_b200.engine import ALIGN_SW
    from oracle import earlyfusion_np as ef
    from oracle import serra09_c as oc
    tracks, _ = synthetic.config_dataset("tiny")
    rng = np.random.default_rng(3)
    tracks = list(tracks[:9]) + [hp(rng, n) for n in (11, 12, 13, 40, 1040)]
    _set(eng, tracks)
    pairs = np.array([(i, j) for i in range(len(tracks)) for j in range(len(tracks)) if i != j], np.int32)
    p = default_params(align=ALIGN_SW, crp_path=CRP_EXACT if path == "exact" else CRP_AUTO)
    got = eng.score_pairs(pairs, p)
    want = np.empty(len(pairs), np.float32)
    for k, (i, j) in enumerate(pairs):
        _, dbg = oc.pair(tracks[i], tracks[j], want_debug=True)
        want[k] = ef.smith_waterman_constrained_x10(dbg["crp"]) / 10.0
    assert np.array_equal(got, want)
    assert (want > 0).any() and (want == 0).any()
    # the debug dump of a pair is unaffected by the SW trim: full CRP, SW score
    d = eng.dump_pair(0, 1, p)
    _, dbg = oc.pair(tracks[0], tracks[1], want_debug=True)
    assert np.array_equal(d["crp"], dbg["crp"])
    assert d["score"] == np.float32(ef.smith_waterman_constrained_x10(dbg["crp"]) / 10.0)
