"""GPU: the two implementations of the K2 sweeps give the same bits.

The fast CRP path has its histogram / emit sweeps twice: on the tensor cores (k2_tc.inl: tcgen05.mma kind::i8 over 24-bit
byte-limb planes, items from TMEM) and as FFMA2 chains on the CUDA cores (k2_fast.cu).  The library picks one by the length
of the longer side; ACOSS_K2_SWEEPS=tc|ffma forces one.  Their fixed-point items differ (different quantisation, both
inside the EPS bound), the outputs must not: thresholds, every CRP bit and the score equal each other, the exact path and
the C oracle, at short, ordinary and long tracks, with ragged lengths that exercise partial blocks and strips."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    from acoss_b200 import Engine
    e = Engine(0)
    yield e
    e.close()


def _both(monkeypatch, eng, pairs, dump=()):
    out = {}
    for mode in ("tc", "ffma"):
        monkeypatch.setenv("ACOSS_K2_SWEEPS", mode)
        s = eng.score_pairs(pairs)
        st = eng.last_stats()
        out[mode] = (s, st["fallback_pairs"], [eng.dump_pair(int(q), int(r)) for q, r in dump])
    monkeypatch.delenv("ACOSS_K2_SWEEPS")
    return out


@pytest.mark.parametrize("config,ntracks,npairs", [("C3", 26, 96), ("C4s", 39, 256), ("C5", 13, 12)])
def test_tensor_and_ffma_sweeps_identical(monkeypatch, eng, config, ntracks, npairs):
    from acoss_b200 import default_params, pack_tracks, synthetic
    from acoss_b200._lib import CRP_EXACT
    tracks, labels = synthetic.config_dataset(config, max_tracks=ntracks)
    frames, offs = pack_tracks(tracks)
    eng.set_tracks(frames, offs)
    pairs = synthetic.all_pairs_upper(len(tracks))
    pairs = pairs[np.random.default_rng(11).permutation(len(pairs))[:npairs]]
    dump = [tuple(p) for p in pairs[:3]]
    out = _both(monkeypatch, eng, pairs, dump)
    exact = eng.score_pairs(pairs, default_params(crp_path=CRP_EXACT))
    assert np.array_equal(out["tc"][0], exact) and np.array_equal(out["ffma"][0], exact)
    assert out["tc"][1] <= 2 and out["ffma"][1] <= 2                       # ordinary pairs: (almost) none needs the exact path
    for a, b in zip(out["tc"][2], out["ffma"][2]):
        assert np.array_equal(a["crp"], b["crp"]) and np.array_equal(a["thr_q"], b["thr_q"]) and np.array_equal(a["thr_r"], b["thr_r"])


def test_tensor_sweeps_against_oracle_ragged(monkeypatch, eng):
    """Lengths around the block (64 windows) and strip (128 lines) sizes of the tensor sweeps, against the C oracle."""
    from acoss_b200 import pack_tracks, synthetic
    from oracle import serra09_c as oc
    lens = [73, 74, 137, 138, 200, 265, 266, 329, 521, 1033]               # windows = len - 9: 64, 65, 128, 129, ..., 1024
    base, _ = synthetic.config_dataset("C3", max_tracks=13)               # slices of ~2k-frame synthetic tracks
    tracks = [np.ascontiguousarray(base[i % len(base)][7 * i:7 * i + n]) for i, n in enumerate(lens)]
    frames, offs = pack_tracks(tracks)
    eng.set_tracks(frames, offs)
    pairs = np.array([(i, j) for i in range(len(lens)) for j in range(len(lens)) if i != j], dtype=np.int32)
    monkeypatch.setenv("ACOSS_K2_SWEEPS", "tc")
    got = eng.score_pairs(pairs)
    dumps = {(int(q), int(r)): eng.dump_pair(int(q), int(r)) for q, r in pairs[::7]}
    monkeypatch.delenv("ACOSS_K2_SWEEPS")
    want = oc.pairs(frames, offs, pairs, oc.params(hoist_norms=True), nthreads=8)
    assert np.array_equal(got, want)
    op = oc.params(hoist_norms=True)
    for (q, r), d in dumps.items():
        s, dbg = oc.pair(tracks[q], tracks[r], op, want_debug=True)
        assert d["score"] == s and np.array_equal(d["crp"], dbg["crp"])
        assert np.array_equal(d["thr_q"], dbg["thr_q"]) and np.array_equal(d["thr_r"], dbg["thr_r"])


def test_tensor_sweeps_feature_scales(monkeypatch, eng):
    """Features far from HPCP's [0, 1]: the quantisation exponent follows the largest feature; tiny and large scales score
    like the exact path.  Negative features switch the whole fast path off (exact path), as before."""
    from acoss_b200 import default_params, pack_tracks, synthetic
    from acoss_b200._lib import CRP_EXACT
    base, _ = synthetic.config_dataset("C3", max_tracks=13)
    base = [t[:1100] for t in base[:6]]
    pairs = synthetic.all_pairs_upper(len(base))
    monkeypatch.setenv("ACOSS_K2_SWEEPS", "tc")
    for scale in (1.0 / 4096, 0.37, 3.0, 1000.0):
        tracks = [np.ascontiguousarray(t * np.float32(scale)) for t in base]
        frames, offs = pack_tracks(tracks)
        eng.set_tracks(frames, offs)
        fast = eng.score_pairs(pairs)
        st = eng.last_stats()
        exact = eng.score_pairs(pairs, default_params(crp_path=CRP_EXACT))
        assert np.array_equal(fast, exact)
        assert st["fallback_pairs"] <= 1
    tracks = [np.ascontiguousarray(t - np.float32(0.01)) for t in base]    # a few negative values
    frames, offs = pack_tracks(tracks)
    eng.set_tracks(frames, offs)
    fast = eng.score_pairs(pairs)
    exact = eng.score_pairs(pairs, default_params(crp_path=CRP_EXACT))
    assert np.array_equal(fast, exact)
    monkeypatch.delenv("ACOSS_K2_SWEEPS")
