"""GPU: oracle parity at the shapes BASELINE.json's configs name.

C1 (covers80-shaped, ~2k frames): 8 pairs with OTI, thresholds, every CRP bit and the score equal to the C oracle.
C5 (long tracks, ~8k frames): 4 pairs the same way (strips, multi-level brackets, DP strips), and fast == exact on 64
pairs.  C4s (Da-TACOS-shaped cliques, ~500 frames = a 4-minute song after the x40 median downsampling): a 78-track
set through Serra09.all_pairwise -> normalize_by_length -> getEvalStatistics with metrics identical to the
oracle-fed sequence."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    from acoss_b200 import Engine
    e = Engine(0)
    yield e
    e.close()


def _check_pairs_bit_exact(eng, tracks, pairs):
    """OTI, both threshold vectors, every CRP bit and the score of each pair against the C oracle (one oracle
    thread per pair; hoist_norms only skips essentia's redundant re-evaluation of the norms, same values)."""
    from concurrent.futures import ThreadPoolExecutor
    from acoss_b200 import pack_tracks
    from oracle import serra09_c as oc
    frames, offs = pack_tracks(tracks)
    eng.set_tracks(frames, offs)
    pairs = np.asarray(pairs, dtype=np.int32)
    got = eng.score_pairs(pairs)
    assert eng.last_stats()["fallback_pairs"] == 0           # these are ordinary pairs: the fast path scores them
    op = oc.params(hoist_norms=True)
    with ThreadPoolExecutor(max_workers=len(pairs)) as ex:
        ref = list(ex.map(lambda qr: oc.pair(tracks[qr[0]], tracks[qr[1]], op, want_debug=True), pairs))
    for k, (q, r) in enumerate(pairs):
        s, dbg = ref[k]
        d = eng.dump_pair(int(q), int(r))
        assert got[k] == s and d["score"] == s and d["oti"] == dbg["oti"]
        assert np.array_equal(d["thr_q"], dbg["thr_q"]) and np.array_equal(d["thr_r"], dbg["thr_r"])
        assert np.array_equal(d["crp"], dbg["crp"])


def test_c1_shaped_pairs_bit_exact(eng):
    from acoss_b200 import synthetic
    tracks, labels = synthetic.config_dataset("C1", max_tracks=8)     # 4 cliques of 2, ~2k frames
    # both members of a clique (covers), and non-covers in both orders
    _check_pairs_bit_exact(eng, tracks, [(0, 1), (2, 3), (4, 5), (6, 7), (0, 2), (3, 0), (5, 6), (7, 1)])


def test_c5_shaped_pairs_bit_exact(eng):
    from acoss_b200 import synthetic
    tracks, labels = synthetic.config_dataset("C5", max_tracks=13)    # one clique of 13, ~8k frames
    assert min(len(t) for t in tracks) > 7000
    _check_pairs_bit_exact(eng, tracks, [(0, 1), (5, 2), (3, 12), (11, 7)])


def test_c5_fast_equals_exact_64_pairs(eng):
    from acoss_b200 import default_params, pack_tracks, synthetic
    from acoss_b200._lib import CRP_EXACT
    tracks, labels = synthetic.config_dataset("C5", max_tracks=26)
    frames, offs = pack_tracks(tracks)
    eng.set_tracks(frames, offs)
    pairs = synthetic.all_pairs_upper(len(tracks))
    pairs = pairs[np.random.default_rng(5).permutation(len(pairs))[:64]]
    fast = eng.score_pairs(pairs)
    st = eng.last_stats()
    exact = eng.score_pairs(pairs, default_params(crp_path=CRP_EXACT))
    assert np.array_equal(fast, exact)
    assert st["fallback_pairs"] <= 3


def test_c4s_shaped_set_identical_metrics(tmp_path, monkeypatch):
    monkeypatch.chdir(tmp_path)
    from acoss_b200 import pack_tracks, synthetic
    from acoss_b200.serra09 import Serra09
    from oracle import evalstats_np as ev
    from oracle import serra09_c as oc
    tracks, labels = synthetic.config_dataset("C4s", max_tracks=78)   # 6 cliques of 13, ~500 frames
    assert len(tracks) >= 64
    feats = [dict(hpcp=t, label="w%d" % l) for t, l in zip(tracks, labels)]
    s = Serra09(None, None, features=feats, downsample_fac=1, shortname="c4s")
    s.all_pairwise(parallel=1, n_cores=4, symmetric=True)
    raw = np.array(s.Ds["main"])
    s.normalize_by_length()
    got = s.getEvalStatistics("main", topsidx=[1, 10, 100])
    frames, offs = pack_tracks(tracks)
    pairs = synthetic.all_pairs_upper(len(tracks))
    sc = oc.pairs(frames, offs, pairs, oc.params(hoist_norms=True), nthreads=16)
    D = np.zeros((len(tracks), len(tracks)), np.float32)
    D[pairs[:, 0], pairs[:, 1]] = sc
    D = ev.symmetrize(D)
    assert np.array_equal(raw, D)
    Dn = ev.normalize_by_length(D, [len(t) for t in tracks])
    assert np.array_equal(np.array(s.Ds["main"]), Dn)
    want = ev.eval_statistics(Dn, s.cliques, [1, 10, 100])
    assert got[:4] == want[:4] and list(got[4]) == list(want[4])
    s.cleanup_memmap()
    s.close()
