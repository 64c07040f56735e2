#!/usr/bin/env python
"""Generate golden vectors by EXECUTING THE REFERENCE'S OWN CODE in the build container.

Run from the repo root (needs /root/reference, numba, scipy):

    python tests/golden/make_golden.py

Writes tests/golden/earlyfusion_golden.npz and tests/golden/evalstats_golden.json.
/root/reference does not exist on the GPU box, so the tests only read the committed outputs.

What is executed (file:line under /root/reference):
  acoss/algorithms/utils/alignment_tools.py:26-46     smith_waterman_constrained (numba jit)
  acoss/algorithms/utils/cross_recurrence.py          get_oti / get_csm / get_csm_cosine /
                                                      get_csm_blocked_oti / csm_to_binary (.py_func
                                                      where numba 0.65 cannot type them)
  acoss/algorithms/algorithm_template.py:205-290      CoverAlgorithm.getEvalStatistics
  acoss/algorithms/rqa_serra09.py:71-83               Serra09.normalize_by_length
  acoss/algorithms/latefusion_chen.py:75-85           ChenFusion.normalize_by_length
The Serra09 essentia calls themselves cannot be executed (essentia absent): no golden for them.
"""
import importlib.util
import json
import os
import sys
import tempfile
import types

import numpy as np

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
sys.dont_write_bytecode = True


def load_by_path(name, rel):
    spec = importlib.util.spec_from_file_location(name, os.path.join(REF, rel))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def install_shims():
    """stub modules so the untouched reference package imports."""
    def mk(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m
    dd = mk("deepdish"); io = mk("deepdish.io", load=lambda *a, **k: {}, save=lambda *a, **k: None)
    dd.io = io; dd.load = io.load; dd.save = io.save

    class Bar:
        def __init__(self, *a, **k): pass
        def next(self): pass
        def finish(self): pass
    pr = mk("progress"); pr.bar = mk("progress.bar", Bar=Bar)
    es = mk("essentia", Pool=object, array=np.array, run=lambda *a: None)
    es.standard = mk("essentia.standard", ChromaCrossSimilarity=object, CoverSongSimilarity=object)
    lr = mk("librosa")
    lr.util = mk("librosa.util", sync=None, normalize=None)
    lr.filters = mk("librosa.filters", get_window=None)
    sys.path.insert(0, REF)


def earlyfusion_golden():
    at = load_by_path("ref_alignment_tools", "acoss/algorithms/utils/alignment_tools.py")
    cr = load_by_path("ref_cross_recurrence", "acoss/algorithms/utils/cross_recurrence.py")
    get_oti = cr.get_oti.py_func
    cr.get_oti = get_oti                              # get_csm_blocked_oti calls the dispatcher (:130)
    get_csm = cr.get_csm.py_func
    get_csm_cosine = cr.get_csm_cosine               # jits fine
    blocked = cr.get_csm_blocked_oti.py_func
    to_bin = cr.csm_to_binary.py_func
    sw = at.smith_waterman_constrained               # jits fine

    out = {}
    # --- smith_waterman_constrained known answers: inputs re-derivable from (seed, shape, p)
    sw_cases = [(0, 64, 64, .1), (1, 100, 80, .1), (2, 300, 257, .095), (3, 512, 512, .2), (4, 33, 1000, .1),
                (5, 400, 400, .1), (6, 4, 4, .5), (7, 5, 70, .3), (8, 129, 65, .15), (9, 700, 613, .095),
                (10, 96, 2050, .1), (11, 37, 41, .9)]
    meta, scores = [], []
    for seed, m, n, p in sw_cases:
        B = (np.random.default_rng(seed).random((m, n)) < p).astype(np.uint8)
        meta.append([seed, m, n, p])
        scores.append([float(sw(B)), float(sw(np.ascontiguousarray(B.T)))])
    out["sw_meta"] = np.array(meta, dtype=np.float64)
    out["sw_scores"] = np.array(scores, dtype=np.float64)
    out["sw_eye50"] = np.float64(sw(np.eye(50)))
    out["sw_ones50"] = np.float64(sw(np.ones((50, 50))))
    out["sw_ones3x10"] = np.float64(sw(np.ones((3, 10))))
    try:
        sw(np.full((5, 5), 2))
        out["sw_nonbinary_raises"] = np.int64(0)
    except (IOError, OSError):
        out["sw_nonbinary_raises"] = np.int64(1)

    # --- get_oti
    r = np.random.default_rng(42)
    c = r.random(12)
    out["oti_roll3"] = np.int64(get_oti(c, np.roll(c, 3)))
    out["oti_rollm3"] = np.int64(get_oti(c, np.roll(c, -3)))
    oc1 = r.random((40, 12)); oc2 = r.random((40, 12))
    out["oti_c1"] = oc1; out["oti_c2"] = oc2
    out["oti_vals"] = np.array([get_oti(a, b) for a, b in zip(oc1, oc2)], dtype=np.int64)

    # --- csm_to_binary
    D79 = r.random((7, 9))
    out["bin_D79"] = D79
    out["bin_D79_k01"] = to_bin(D79, 0.1)
    out["bin_D79_k3"] = to_bin(D79, 3)
    out["bin_alleq_k03"] = to_bin(np.ones((3, 10)), 0.3)
    D = r.random((60, 45))
    out["bin_D"] = D
    out["bin_D_k01"] = to_bin(D, 0.1)
    out["bin_D_k0"] = to_bin(D, 0)
    out["bin_round"] = np.array([int(np.round(0.1 * 25)), int(np.round(0.1 * 35))], dtype=np.int64)

    # --- CSMs + pipeline (SURVEY App. B pipeline KAT: seeds 10, 11)
    for seed in (10, 11):
        rr = np.random.default_rng(seed)
        X = rr.random((120, 48)); Y = rr.random((90, 48)); c1 = rr.random(12); c2 = rr.random(12)
        Dc = blocked(X, Y, c1, c2, get_csm_cosine)
        Bc = to_bin(Dc, 0.1)
        out["pipe%d_oti" % seed] = np.int64(get_oti(c1, c2))
        out["pipe%d_csm_sum" % seed] = np.float64(Dc.sum())
        out["pipe%d_csm" % seed] = Dc
        out["pipe%d_ones" % seed] = np.int64(Bc.sum())
        out["pipe%d_bin" % seed] = Bc
        out["pipe%d_score" % seed] = np.float64(sw(Bc))
        De = get_csm(X, Y)
        out["pipe%d_euclid" % seed] = De
        out["pipe%d_euclid_score" % seed] = np.float64(sw(to_bin(De, 0.1)))
    np.savez_compressed(os.path.join(HERE, "earlyfusion_golden.npz"), **out)
    print("wrote earlyfusion_golden.npz:", len(out), "arrays")


def evalstats_golden():
    install_shims()
    cwd = os.getcwd()
    tmp = tempfile.mkdtemp()
    os.chdir(tmp)                                     # importing acoss writes log files into CWD
    try:
        from acoss.algorithms.algorithm_template import CoverAlgorithm
        from acoss.algorithms.rqa_serra09 import Serra09
        cases = []

        def run(N, labels, seed, boost, topsidx):
            D = np.random.default_rng(seed).random((N, N)).astype(np.float32)
            lab = np.asarray(labels)
            D = D + np.float32(boost) * (lab[:, None] == lab[None, :]).astype(np.float32)
            alg = CoverAlgorithm.__new__(CoverAlgorithm)
            alg.name, alg.shortname = "golden", "kat"
            alg.Ds = {"main": D}
            alg.cliques = {}
            for i, l in enumerate(labels):
                alg.cliques.setdefault(str(l), set()).add(i)
            MR, MRR, MDR, MAP, tops = alg.getEvalStatistics("main", topsidx=list(topsidx))
            cases.append(dict(N=N, labels=[int(x) for x in labels], seed=seed, boost=boost,
                              topsidx=list(topsidx), MR=float(MR), MRR=float(MRR), MDR=float(MDR),
                              MAP=float(MAP), tops=[float(t) for t in tops]))

        run(12, [0, 0, 0, 1, 1, 2, 2, 2, 2, 3, 4, 5], 0, 0.5, (1, 10))          # SURVEY App. B KAT
        run(40, [i // 4 for i in range(32)] + list(range(100, 108)), 1, 0.3, (1, 10, 100))
        run(60, [i // 13 for i in range(52)] + list(range(200, 208)), 2, 0.1, (1, 10, 100, 1000))
        run(30, [i // 2 for i in range(30)], 3, 0.05, (1, 10))

        # normalize_by_length through the reference class
        n_frames = [int(x) for x in np.random.default_rng(7).integers(50, 400, size=9)]
        s = Serra09.__new__(Serra09)
        D = np.random.default_rng(8).random((9, 9)).astype(np.float32) * np.float32(40)
        s.Ds = {"main": D.copy()}
        s.all_feats = {i: np.zeros((n, 12), dtype=np.float32) for i, n in enumerate(n_frames)}
        s.normalize_by_length()
        norm = dict(n_frames=n_frames, seed=8, scale=40.0,
                    out=[[float(x) for x in row] for row in s.Ds["main"]])
        # ChenFusion.normalize_by_length (latefusion_chen.py:75-85): norm_fac / score, zero scores -> inf
        from acoss.algorithms.latefusion_chen import ChenFusion
        cf = ChenFusion.__new__(ChenFusion)
        Dq = np.floor(np.random.default_rng(9).random((9, 9)) * 60).astype(np.float32) * np.float32(0.5)
        Dd = np.floor(np.random.default_rng(10).random((9, 9)) * 90).astype(np.float32) * np.float32(0.5)
        np.fill_diagonal(Dq, 0); np.fill_diagonal(Dd, 0)
        cf.Ds = {"qmax": Dq.copy(), "dmax": Dd.copy()}
        cf.filepaths = [None] * 9
        cf.all_feats = {i: np.zeros((n, 12), dtype=np.float32) for i, n in enumerate(n_frames)}
        with np.errstate(divide="ignore"):
            cf.normalize_by_length()
        tolist = lambda M: [[("inf" if np.isinf(x) else float(x)) for x in row] for row in M]
        norm_chen = dict(n_frames=n_frames, seeds=[9, 10], qmax=tolist(cf.Ds["qmax"]), dmax=tolist(cf.Ds["dmax"]))
    finally:
        os.chdir(cwd)
    with open(os.path.join(HERE, "evalstats_golden.json"), "w") as f:
        json.dump(dict(eval_cases=cases, normalize=norm, normalize_chen=norm_chen), f, indent=1)
    print("wrote evalstats_golden.json:", len(cases), "eval cases")


if __name__ == "__main__":
    earlyfusion_golden()
    evalstats_golden()
