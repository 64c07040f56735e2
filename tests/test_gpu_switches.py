"""GPU: the switches F2-F5 of acoss_params (the points where the restatement of essentia is
uncertain) against the oracle carrying the same switches.  With a switch set the library scores the pair on the
exact CRP path; thresholds, every CRP bit and the score must equal the oracle's."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _hp(rng, n):
    X = rng.random((n, 12)).astype(np.float32)
    return (X / X.max(1, keepdims=True)).astype(np.float32)


def _tracks(seed, lens):
    rng = np.random.default_rng(seed)
    tracks = [_hp(rng, n) for n in lens]
    # a planted cover (exercises OTI and long diagonals) and a duplicated block of frames (distance ties)
    tracks[1] = (np.roll(tracks[0][:lens[1]], 3, axis=1) + np.float32(0.02) * rng.random((lens[1], 12)).astype(np.float32))
    tracks[1] = (tracks[1] / tracks[1].max(1, keepdims=True)).astype(np.float32)
    tracks[2][30:45] = tracks[2][5:20]
    return tracks


CASES = [
    ("f2_strict", dict(f2_strict=1), dict(strict=True)),
    ("f2_strict+guard", dict(f2_strict=1, integer_guard=1), dict(strict=True, integer_guard=True)),
    ("f3_float_acc", dict(f3_float_acc=1), dict(f64_accumulate=False)),
    ("f4_keep_last", dict(f4_keep_last=1), dict(drop_one=False)),
    ("f5_asymmetric", dict(f5_asymmetric=1), dict(asymmetric=True)),
    ("all", dict(f2_strict=1, f3_float_acc=1, f4_keep_last=1, f5_asymmetric=1, integer_guard=1),
     dict(strict=True, f64_accumulate=False, drop_one=False, asymmetric=True, integer_guard=True)),
]


@pytest.mark.parametrize("name,gpu_kw,oracle_kw", CASES, ids=[c[0] for c in CASES])
def test_switch_matches_oracle(name, gpu_kw, oracle_kw):
    from acoss_b200 import Engine, default_params, pack_tracks
    from oracle import serra09_np as o
    # 210 / 209 frames: 201 stacked frames under F4's default / alternative => integer rank (k = 19.0), where
    # thr is an element of the line (with the guard) and F2 decides whether that element counts
    lens = [210, 209, 64, 97, 51]
    tracks = _tracks(11, lens)
    frames, offs = pack_tracks(tracks)
    pairs = np.array([(0, 1), (1, 0), (0, 2), (2, 3), (3, 4), (4, 0), (1, 3)], dtype=np.int32)
    ck = {k: v for k, v in oracle_kw.items() if k != "asymmetric"}
    with Engine(0) as eng:
        eng.set_tracks(frames, offs)
        got = eng.score_pairs(pairs, default_params(**gpu_kw))
        assert eng.last_stats()["fallback_pairs"] == 0
        for k, (i, j) in enumerate(pairs):
            want = o.serra09_pair(tracks[i], tracks[j], **oracle_kw)
            assert got[k] == np.float32(want), (name, i, j, got[k], want)
            crp, dbg = o.chroma_cross_similarity(tracks[i], tracks[j], return_debug=True, **ck)
            d = eng.dump_pair(int(i), int(j), default_params(**gpu_kw))
            assert d["oti"] == dbg["oti"]
            assert np.array_equal(d["thr_q"], dbg["thr_q"]) and np.array_equal(d["thr_r"], dbg["thr_r"]), name
            assert d["crp"].shape == crp.shape and np.array_equal(d["crp"], crp), name


def test_switches_change_something():
    """The switches are not no-ops: on the integer-rank pair F2 (with the guard) removes exactly the threshold
    cells, F4 adds a row and a column, F5 rescales."""
    from acoss_b200 import Engine, default_params, pack_tracks
    tracks = _tracks(11, [210, 210, 64])
    frames, offs = pack_tracks(tracks)
    with Engine(0) as eng:
        eng.set_tracks(frames, offs)
        base = eng.dump_pair(0, 1, default_params(integer_guard=1))
        strict = eng.dump_pair(0, 1, default_params(integer_guard=1, f2_strict=1))
        assert strict["crp"].sum() < base["crp"].sum() and not (strict["crp"] & ~base["crp"]).any()
        keep = eng.dump_pair(0, 1, default_params(f4_keep_last=1))
        assert keep["crp"].shape == (base["crp"].shape[0] + 1, base["crp"].shape[1] + 1)
        s0 = eng.score_pairs(np.array([[0, 1]], np.int32))[0]
        s5 = eng.score_pairs(np.array([[0, 1]], np.int32), default_params(f5_asymmetric=1))[0]
        with np.errstate(divide="ignore"):
            assert s5 == np.float32(np.sqrt(np.float32(201)) / s0)
