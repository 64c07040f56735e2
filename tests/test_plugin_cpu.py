"""Host-side mirror of the reference plugin API (no GPU needed)."""
import json
import os

import numpy as np
import pytest

from acoss_b200.algorithm_template import CoverAlgorithm, create_dataset_filepaths, eval_statistics
from acoss_b200.serra09 import Serra09
from oracle.onramp_np import median_sync
from acoss_b200 import earlyfusion as efp
from oracle import earlyfusion_np as ef
from oracle import evalstats_np as ev


@pytest.fixture()
def workdir(tmp_path, monkeypatch):
    monkeypatch.chdir(tmp_path)
    return tmp_path


def _feats(n, rng, labels=None):
    return [dict(hpcp=rng.random((int(rng.integers(30, 60)), 12)).astype(np.float32),
                 label=str(labels[i] if labels else i // 2)) for i in range(n)]


def test_eval_matches_reference_golden(golden_dir):
    with open(os.path.join(golden_dir, "evalstats_golden.json")) as f:
        g = json.load(f)
    for c in g["eval_cases"]:
        N, labels = c["N"], c["labels"]
        D = np.random.default_rng(c["seed"]).random((N, N)).astype(np.float32)
        lab = np.asarray(labels)
        D = D + np.float32(c["boost"]) * (lab[:, None] == lab[None, :]).astype(np.float32)
        cliques = {}
        for i, l in enumerate(labels):
            cliques.setdefault(str(l), set()).add(i)
        for rb in (2048, 5):                                   # row blocking must not matter
            MR, MRR, MDR, MAP, tops, _ = eval_statistics(D, cliques, c["topsidx"], row_block=rb)
            assert (MR, MRR, MDR, MAP) == (c["MR"], c["MRR"], c["MDR"], c["MAP"])
            assert list(tops) == c["tops"]
        want = ev.eval_statistics(D, cliques, c["topsidx"])
        assert want[:4] == (MR, MRR, MDR, MAP)


def test_get_eval_statistics_side_effects(workdir):
    rng = np.random.default_rng(0)
    alg = CoverAlgorithm(None, name="T", shortname="s", features=_feats(8, rng))
    assert alg.N == 8 and alg.Ds["main"].shape == (8, 8) and alg.Ds["main"].dtype == np.float32
    assert os.path.exists("cache/T_s_main_dmat")
    alg.get_all_clique_ids()
    assert alg.cliques == {"0": {0, 1}, "1": {2, 3}, "2": {4, 5}, "3": {6, 7}}
    alg.Ds["main"][:] = rng.random((8, 8)).astype(np.float32)
    out = alg.getEvalStatistics("main", topsidx=[1, 10])
    assert len(out) == 5
    rows = open("results_s_T.csv").read().strip().splitlines()
    assert rows[0] == "name, MR, MRR, MDR, MAP,Top-1,Top-10" and rows[1].startswith("T_main,")
    alg.getEvalStatistics("main", topsidx=[1, 10])
    assert len(open("results_s_T.csv").read().strip().splitlines()) == 3
    alg.cleanup_memmap()
    assert not os.path.exists("cache/T_s_main_dmat")


def test_base_all_pairwise_and_pairs(workdir):
    rng = np.random.default_rng(1)
    alg = CoverAlgorithm(None, name="T", shortname="s", features=_feats(6, rng))
    from itertools import combinations, permutations
    assert [tuple(p) for p in alg._pair_array(True)] == list(combinations(range(6), 2))
    assert [tuple(p) for p in alg._pair_array(False)] == list(permutations(range(6), 2))
    alg.all_pairwise(symmetric=True)
    assert float(np.abs(alg.Ds["main"]).sum()) == 0.0
    alg2 = CoverAlgorithm(None, name="T", shortname="s", features=_feats(6, rng))
    alg2.all_pairwise(precomputed=True)
    assert alg2.Ds["main"].shape == (6, 6)


def test_create_dataset_filepaths(tmp_path):
    p = tmp_path / "d.csv"
    p.write_text("work_id,track_id\nW1,a\nW1,b\nW2,c\n")
    assert create_dataset_filepaths(str(p), "root/", ".h5") == ["root/W1/a.h5", "root/W1/b.h5", "root/W2/c.h5"]
    bad = tmp_path / "bad.csv"
    bad.write_text("work,track_id\nW1,a\n")
    with pytest.raises(IOError):
        create_dataset_filepaths(str(bad), "root/")


def test_npz_feature_files(workdir):
    os.makedirs("feat/W1"); os.makedirs("feat/W2")
    rng = np.random.default_rng(2)
    open("d.csv", "w").write("work_id,track_id\nW1,a\nW1,b\nW2,c\n")
    for w, t in (("W1", "a"), ("W1", "b"), ("W2", "c")):
        np.savez("feat/%s/%s.npz" % (w, t), hpcp=rng.random((90, 12)).astype(np.float32), label=w)
    s = Serra09("d.csv", "feat/", downsample_fac=1)
    f = s.load_features(1)
    assert f.shape == (90, 12) and f.dtype == np.float32
    assert s.cliques == {"W1": {1}}
    assert s.load_features(1) is f                       # cached
    # downsample_fac > 1: the block medians are computed on the GPU only — without a device the call raises
    import torch
    if not torch.cuda.is_available():
        from acoss_b200 import AcossError
        s40 = Serra09("d.csv", "feat/", downsample_fac=40, shortname="fac40")
        with pytest.raises(AcossError):
            s40.load_features(1)


def test_median_sync_definition():
    rng = np.random.default_rng(3)
    for n in (1, 39, 40, 41, 80, 517, 2001):
        X = rng.random((n, 12)).astype(np.float32)
        got = median_sync(X, 40)
        want = np.stack([np.median(X[k:k + 40], axis=0) for k in range(0, n, 40)])
        assert got.dtype == np.float32 and np.array_equal(got, want)
    assert np.array_equal(median_sync(X, 1), X)


def test_serra09_signature_and_no_fallback(workdir):
    import inspect
    import torch
    sig = inspect.signature(Serra09.__init__)
    ref_order = ["self", "dataset_csv", "datapath", "chroma_type", "shortname", "oti", "kappa", "tau", "m",
                 "downsample_fac"]
    assert list(sig.parameters)[:len(ref_order)] == ref_order
    d = {k: v.default for k, v in sig.parameters.items()}
    assert (d["chroma_type"], d["shortname"], d["oti"], d["kappa"], d["tau"], d["m"], d["downsample_fac"]) == \
        ("hpcp", "benchmark", True, 0.095, 1, 9, 40)
    rng = np.random.default_rng(4)
    s = Serra09(None, None, features=_feats(4, rng), downsample_fac=1)
    assert s.name == "Serra09" and list(s.Ds) == ["main"]
    if not torch.cuda.is_available():
        from acoss_b200 import AcossError
        with pytest.raises(AcossError):                    # no CPU fallback
            s.similarity(np.array([[0, 1]]))
    # normalize_by_length is host-side and must match the reference golden
    with open(os.path.join(os.path.dirname(__file__), "golden", "evalstats_golden.json")) as f:
        n = json.load(f)["normalize"]
    s2 = Serra09(None, None, downsample_fac=1,
                 features=[dict(hpcp=np.zeros((k, 12), np.float32), label="x") for k in n["n_frames"]])
    s2.Ds["main"][:] = np.random.default_rng(n["seed"]).random((9, 9)).astype(np.float32) * np.float32(n["scale"])
    s2.normalize_by_length()
    assert np.array_equal(np.asarray(s2.Ds["main"]), np.array(n["out"], dtype=np.float32))


def test_earlyfusion_host_helpers(golden_dir):
    assert efp.nneighbs(0.1, 25) == 2 and efp.nneighbs(0.1, 35) == 4 and efp.nneighbs(3, 9) == 3
    assert efp.nneighbs(0, 9) == -1
    # no host implementation of the CSM stages in the product package (the CUDA path is the only path)
    for name in ("get_oti", "csm_euclidean", "csm_cosine", "csm_blocked_oti"):
        assert not hasattr(efp, name)
    sig = __import__("inspect").signature(efp.EarlyFusion.__init__)
    assert list(sig.parameters)[1:13] == ["dataset_csv", "datapath", "chroma_type", "shortname", "blocksize",
                                          "mfccs_per_block", "ssm_res", "chromas_per_block", "kappa", "K",
                                          "niters", "log_times"]


def test_chenfusion_signature_and_normalize(workdir):
    """ChenFusion mirror (latefusion_chen.py:18-91): constructor signature, score keys, algorithm name,
    and normalize_by_length against the golden the reference class produced (zero scores -> inf)."""
    import inspect
    from acoss_b200.chenfusion import ChenFusion
    sig = inspect.signature(ChenFusion.__init__)
    assert list(sig.parameters)[:10] == ["self", "dataset_csv", "datapath", "chroma_type", "shortname", "oti",
                                         "kappa", "tau", "m", "downsample_fac"]
    with open(os.path.join(os.path.dirname(__file__), "golden", "evalstats_golden.json")) as f:
        n = json.load(f)["normalize_chen"]
    c = ChenFusion(None, None, downsample_fac=1,
                   features=[dict(hpcp=np.zeros((k, 12), np.float32), label="x") for k in n["n_frames"]])
    assert c.name == "LateFusionChen" and list(c.Ds) == ["qmax", "dmax"]
    Dq = np.floor(np.random.default_rng(n["seeds"][0]).random((9, 9)) * 60).astype(np.float32) * np.float32(0.5)
    Dd = np.floor(np.random.default_rng(n["seeds"][1]).random((9, 9)) * 90).astype(np.float32) * np.float32(0.5)
    np.fill_diagonal(Dq, 0); np.fill_diagonal(Dd, 0)
    c.Ds["qmax"][:] = Dq; c.Ds["dmax"][:] = Dd
    c.normalize_by_length()
    for key in ("qmax", "dmax"):
        want = np.array([[np.inf if x == "inf" else x for x in row] for row in n[key]], dtype=np.float32)
        assert np.array_equal(np.asarray(c.Ds[key]), want)
    with pytest.raises(NotImplementedError):               # SNF late fusion is outside the hot path
        c.do_late_fusion()


def test_earlyfusion_load_features_lookup_order(workdir):
    """EarlyFusion.load_features (earlyfusion_traile.py:84-97): memory cache, then <cacheprefix>_<i> cache file,
    then the song's own dictionary; the clique side effect happens in every case; missing blocks are an error."""
    rng = np.random.default_rng(0)
    def blocks(nb):
        return dict(mfccs=rng.random((nb, 8)).astype(np.float32), ssms=rng.random((nb, 6)).astype(np.float32),
                    chromas=rng.random((nb, 12)).astype(np.float32), chroma_med=rng.random(12).astype(np.float32))
    feats = [dict(blocks(20), label="a"), dict(blocks(22), label="a"), dict(hpcp=np.zeros((5, 12)), label="b")]
    e = efp.EarlyFusion(None, None, features=feats, shortname="lf")
    f0 = e.load_features(0)
    assert f0 is e.load_features(0) and set(("mfccs", "ssms", "chromas", "chroma_med")) <= set(f0)
    e.load_features(1)
    assert e.cliques == {"a": {0, 1}}
    with pytest.raises(KeyError):
        e.load_features(2)
    # file-backed: raw feature file without blocks + the reference-style per-song cache file next to the score matrices
    os.makedirs("data/w1", exist_ok=True)
    with open("ds.csv", "w") as f:
        f.write("work_id,track_id\nw1,t1\n")
    np.savez("data/w1/t1.npz", hpcp=np.zeros((5, 12), np.float32), label="w1")
    e2 = efp.EarlyFusion("ds.csv", "data/", shortname="lf2")
    b = blocks(9)
    np.savez("%s_0.npz" % e2.get_cacheprefix(), **b)
    got = e2.load_features(0)
    assert np.array_equal(got["mfccs"], b["mfccs"]) and e2.cliques == {"w1": {0}}
