"""GPU: the drop-in plugins end to end, against the oracle-driven reference pipeline."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture()
def workdir(tmp_path, monkeypatch):
    monkeypatch.chdir(tmp_path)
    return tmp_path


def test_serra09_dropin_identical_metrics(workdir):
    """all_pairwise -> normalize_by_length -> getEvalStatistics (the coverid.benchmark sequence,
    coverid.py:57-70) on the GPU plugin == the same sequence fed with CPU-oracle scores."""
    from acoss_b200 import pack_tracks, synthetic
    from acoss_b200.serra09 import Serra09
    from oracle import evalstats_np as ev
    from oracle import serra09_c as oc
    tracks, labels = synthetic.config_dataset("tiny")
    feats = [dict(hpcp=t, label="w%d" % l) for t, l in zip(tracks, labels)]
    s = Serra09(None, None, features=feats, downsample_fac=1, shortname="tiny")
    s.all_pairwise(parallel=1, n_cores=4, symmetric=True)
    raw = np.array(s.Ds["main"])
    s.normalize_by_length()
    got = s.getEvalStatistics("main", topsidx=[1, 10])
    # oracle pipeline
    frames, offs = pack_tracks(tracks)
    pairs = synthetic.all_pairs_upper(len(tracks))
    sc = oc.pairs(frames, offs, pairs, nthreads=8)
    D = np.zeros((len(tracks), len(tracks)), np.float32)
    D[pairs[:, 0], pairs[:, 1]] = sc
    D = ev.symmetrize(D)
    assert np.array_equal(raw, D)
    Dn = ev.normalize_by_length(D, [len(t) for t in tracks])
    assert np.array_equal(np.array(s.Ds["main"]), Dn)
    want = ev.eval_statistics(Dn, s.cliques, [1, 10])
    assert got[:4] == want[:4] and list(got[4]) == list(want[4])
    assert got[3] > 0.9                                     # covers are found (MAP)
    s.cleanup_memmap()
    s.close()


def test_serra09_single_pair_calls(workdir):
    """The reference's serial mode calls similarity(np.array([[i, j]])) once per pair."""
    from acoss_b200 import synthetic
    from acoss_b200.serra09 import Serra09
    from oracle import serra09_c as oc
    tracks, labels = synthetic.config_dataset("tiny")
    feats = [dict(hpcp=t, label=str(l)) for t, l in zip(tracks, labels)]
    s = Serra09(None, None, features=feats, downsample_fac=1)
    for i, j in [(0, 1), (2, 9), (15, 3)]:
        s.similarity(np.array([[i, j]]))
        assert s.Ds["main"][i][j] == oc.pair(tracks[i], tracks[j])
    s.close()


def test_serra09_downsample_onramp(workdir):
    """load_features median-downsamples raw HPCP by 40 before upload (rqa_serra09.py:51)."""
    from acoss_b200.serra09 import Serra09
    from oracle import serra09_c as oc
    from oracle.onramp_np import median_sync
    rng = np.random.default_rng(5)
    raw = [rng.random((int(n), 12)).astype(np.float32) for n in (2000, 2400, 1810)]
    s = Serra09(None, None, features=[dict(hpcp=r, label="a") for r in raw])
    s.similarity(np.array([[0, 1], [0, 2], [1, 2]]))
    ds = [median_sync(r, 40) for r in raw]
    for i, j in [(0, 1), (0, 2), (1, 2)]:
        assert s.Ds["main"][i][j] == oc.pair(ds[i], ds[j])
    s.close()


def test_knn_sw_matches_reference_pipeline(golden_dir):
    from acoss_b200 import Engine
    from acoss_b200.earlyfusion import sw_of_csms
    from oracle import earlyfusion_np as ef
    g = np.load(os.path.join(golden_dir, "earlyfusion_golden.npz"))
    with Engine(0) as eng:
        mats = [g["pipe10_csm"], g["pipe11_csm"], g["pipe10_euclid"], g["bin_D"]]
        scores, bins = sw_of_csms(eng, mats, 0.1, want_bits=True)
        assert np.array_equal(bins[0], g["pipe10_bin"]) and np.array_equal(bins[1], g["pipe11_bin"])
        assert np.array_equal(bins[3], g["bin_D_k01"])
        assert scores[0] == pytest.approx(float(g["pipe10_score"]), rel=1e-5)
        assert scores[1] == pytest.approx(float(g["pipe11_score"]), rel=1e-5)
        assert scores[2] == pytest.approx(float(g["pipe10_euclid_score"]), rel=1e-5)
        # kappa >= 1 -> count; kappa == 0 -> all ones; ties -> lowest column first
        s3, b3 = sw_of_csms(eng, [g["bin_D79"]], 3, want_bits=True)
        assert np.array_equal(b3[0], g["bin_D79_k3"])
        s0, b0 = sw_of_csms(eng, [g["bin_D"]], 0, want_bits=True)
        assert b0[0].all() and s0[0] == pytest.approx(ef.smith_waterman_constrained(np.ones_like(g["bin_D"])), rel=1e-6)
        st, bt = sw_of_csms(eng, [np.ones((3, 10))], 0.3, want_bits=True)
        assert np.array_equal(bt[0], g["bin_alleq_k03"])
        rng = np.random.default_rng(9)
        big = [rng.random((int(a), int(b))) for a, b in [(300, 257), (64, 1500), (513, 33)]]
        sc = sw_of_csms(eng, big, 0.1)
        want = [ef.smith_waterman_constrained(ef.csm_to_binary(D, 0.1)) for D in big]
        assert np.allclose(sc, want, rtol=1e-5, atol=0)


def test_earlyfusion_plugin(workdir):
    from acoss_b200.earlyfusion import EarlyFusion
    from oracle import earlyfusion_np as ef
    rng = np.random.default_rng(12)
    feats = []
    for i in range(5):
        nb = int(rng.integers(60, 110))
        feats.append(dict(mfccs=rng.random((nb, 50)), ssms=rng.random((nb, 40)), chromas=rng.random((nb, 48)),
                          chroma_med=rng.random(12), label=str(i // 2)))
    e = EarlyFusion(None, None, features=feats)
    idxs = np.array([(0, 1), (0, 2), (3, 4), (1, 4)])
    e.similarity(idxs)
    for i, j in idxs:
        f1, f2 = feats[i], feats[j]
        want = dict(
            mfccs=ef.smith_waterman_constrained(ef.csm_to_binary(ef.get_csm(f1["mfccs"], f2["mfccs"]), 0.1)),
            ssms=ef.smith_waterman_constrained(ef.csm_to_binary(ef.get_csm(f1["ssms"], f2["ssms"]), 0.1)),
            chromas=ef.smith_waterman_constrained(ef.csm_to_binary(
                ef.get_csm_blocked_oti(f1["chromas"], f2["chromas"], f1["chroma_med"], f2["chroma_med"]), 0.1)))
        for k in want:
            assert e.Ds[k][i, j] == pytest.approx(want[k], rel=1e-5)
    e.close()


def test_chenfusion_dropin(workdir):
    """ChenFusion.all_pairwise -> normalize_by_length -> getEvalStatistics per key, against the same
    sequence fed with the CPU oracle's Qmax / Dmax scores."""
    from acoss_b200 import pack_tracks, synthetic
    from acoss_b200.chenfusion import ChenFusion
    from oracle import evalstats_np as ev
    from oracle import serra09_c as oc
    tracks, labels = synthetic.config_dataset("tiny")
    feats = [dict(hpcp=t, label="w%d" % l) for t, l in zip(tracks, labels)]
    c = ChenFusion(None, None, features=feats, downsample_fac=1, shortname="tinychen")
    c.all_pairwise(parallel=0, n_cores=1, symmetric=True)
    frames, offs = pack_tracks(tracks)
    pairs = synthetic.all_pairs_upper(len(tracks))
    wq, wd = oc.chen_pairs(frames, offs, pairs, nthreads=8)
    for key, w in (("qmax", wq), ("dmax", wd)):
        D = np.zeros((len(tracks), len(tracks)), np.float32)
        D[pairs[:, 0], pairs[:, 1]] = w
        assert np.array_equal(np.array(c.Ds[key]), ev.symmetrize(D))
    c.normalize_by_length()
    lens = np.array([len(t) for t in tracks], np.float64)
    with np.errstate(divide="ignore"):
        want = (np.sqrt(lens)[None, :] / ev.symmetrize(D).astype(np.float64)).astype(np.float32)
    assert np.array_equal(np.array(c.Ds["dmax"]), want)
    c.cleanup_memmap()
    c.close()


def test_gpu_median_onramp_bit_exact():
    """acoss_set_tracks_raw == librosa.util.sync(..., aggregate=np.median) restated by median_sync (np.median on
    float32 blocks): full, short-last and single-frame blocks, odd and even block sizes, ties, several factors."""
    from acoss_b200 import Engine, pack_tracks
    from oracle.onramp_np import median_sync
    rng = np.random.default_rng(12)
    raws = [rng.random((int(n), 12)).astype(np.float32) for n in (2000, 2401, 1810, 41, 40, 39, 1, 81, 517)]
    raws.append(np.round(rng.random((333, 12)) * 4).astype(np.float32) / np.float32(4))       # many ties
    frames, offs = pack_tracks(raws)
    with Engine(0) as eng:
        for fac in (40, 7, 2, 1, 128):
            out_off = eng.set_tracks_raw(frames, offs, fac)
            got = eng.get_tracks()
            want = [median_sync(r, fac) for r in raws]
            assert list(np.diff(out_off)) == [len(w) for w in want]
            assert np.array_equal(got, np.concatenate(want)), fac
        from acoss_b200 import AcossError
        with pytest.raises(AcossError):
            eng.set_tracks_raw(frames, offs, 129)


def test_serra09_gpu_onramp_equals_host_onramp(workdir):
    """The plugin's GPU on-ramp: load_features (before or after scoring) returns exactly the oracle's block
    medians, and the scores equal the oracle's on those downsampled tracks."""
    from acoss_b200.serra09 import Serra09
    from oracle import serra09_c as oc
    from oracle.onramp_np import median_sync
    rng = np.random.default_rng(6)
    raw = [rng.random((int(n), 12)).astype(np.float32) for n in (2000, 2400, 1810, 2222)]
    idx = np.array([[0, 1], [0, 2], [1, 2], [2, 3]])
    a = Serra09(None, None, features=[dict(hpcp=r, label="a") for r in raw], shortname="gpuramp")
    f2 = a.load_features(2)                                   # first touch: the GPU downsamples every song
    assert f2.dtype == np.float32 and np.array_equal(f2, median_sync(raw[2], 40))
    a.similarity(idx)
    ds = [median_sync(r, 40) for r in raw]
    for i, j in idx:
        assert a.Ds["main"][i][j] == oc.pair(ds[i], ds[j])
    for i, r in enumerate(raw):
        assert np.array_equal(a.load_features(i), ds[i])
    a.close()
