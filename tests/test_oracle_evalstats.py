"""Oracle (oracle/evalstats_np.py) vs the reference's getEvalStatistics / normalize_by_length
outputs recorded by tests/golden/make_golden.py."""
import json
import os

import numpy as np
import pytest

from oracle import evalstats_np as ev


@pytest.fixture(scope="module")
def g(golden_dir):
    with open(os.path.join(golden_dir, "evalstats_golden.json")) as f:
        return json.load(f)


def _case_inputs(c):
    N, labels = c["N"], c["labels"]
    D = np.random.default_rng(c["seed"]).random((N, N)).astype(np.float32)
    lab = np.asarray(labels)
    D = D + np.float32(c["boost"]) * (lab[:, None] == lab[None, :]).astype(np.float32)
    cliques = {}
    for i, l in enumerate(labels):
        cliques.setdefault(str(l), set()).add(i)
    return D, cliques


def test_eval_cases(g):
    assert len(g["eval_cases"]) >= 4
    for c in g["eval_cases"]:
        D, cliques = _case_inputs(c)
        MR, MRR, MDR, MAP, tops, _ = ev.eval_statistics(D, cliques, c["topsidx"])
        assert MR == pytest.approx(c["MR"], rel=1e-12)
        assert MRR == pytest.approx(c["MRR"], rel=1e-12)
        assert MDR == pytest.approx(c["MDR"], rel=1e-12)
        assert MAP == pytest.approx(c["MAP"], rel=1e-12)
        assert list(tops) == c["tops"]


def test_survey_kat(g):
    """SURVEY.md Appendix B evaluation KAT."""
    c = g["eval_cases"][0]
    assert c["MR"] == pytest.approx(1.777777778, abs=1e-9)
    assert c["MRR"] == pytest.approx(0.6208333333, abs=1e-9)
    assert c["MDR"] == 1 and c["MAP"] == pytest.approx(0.7430335097, abs=1e-9)
    assert c["tops"] == [7.0, 9.0]


def test_normalize_by_length(g):
    n = g["normalize"]
    D = np.random.default_rng(n["seed"]).random((9, 9)).astype(np.float32) * np.float32(n["scale"])
    out = ev.normalize_by_length(D, n["n_frames"])
    assert np.array_equal(out, np.array(n["out"], dtype=np.float32))
