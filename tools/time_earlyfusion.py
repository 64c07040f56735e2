#!/usr/bin/env python
"""Throughput of the full EarlyFusion pair scoring (acoss_ef_score_pairs) on one B200, with the CPU oracle
(numpy, the reference's own arithmetic: BLAS GEMM + argpartition + the row-vectorised Smith-Waterman) timed
on a bounded sample of the same pairs and checked for identical scores.

    python tools/time_earlyfusion.py [--tracks 160] [--blocks 400] [--pairs 4096] [--cpu-pairs 8] [--out f.json]

Workload: covers80-shaped (BASELINE.json configs[1]: "EarlyFusionTralie Smith-Waterman on the same
covers80-shaped ... (shared CSM/DP path)"): 80 cliques x 2 tracks, ~400 beat-synchronous blocks per track,
the reference's block dimensions 1000 / 1225 / 480, float32 features.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--tracks", type=int, default=160)
    ap.add_argument("--blocks", type=int, default=400)
    ap.add_argument("--pairs", type=int, default=4096)
    ap.add_argument("--cpu-pairs", type=int, default=8)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    from acoss_b200 import Engine, synthetic
    from oracle import earlyfusion_np as ef
    t0 = time.time()
    feats = synthetic.ef_dataset([2] * (a.tracks // 2), a.blocks, 20242)
    t_gen = time.time() - t0
    n = len(feats)
    i, j = np.triu_indices(n, k=1)
    allp = np.stack([i, j], axis=1).astype(np.int32)
    rng = np.random.default_rng(1)
    sel = np.sort(rng.choice(len(allp), size=min(a.pairs, len(allp)), replace=False))
    pairs = np.ascontiguousarray(allp[sel])
    nb = np.array([f["mfccs"].shape[0] for f in feats], dtype=np.int64)
    cells = int((nb[pairs[:, 0]] * nb[pairs[:, 1]]).sum())
    d = {k: feats[0][k].shape[1] for k in ("mfccs", "ssms", "chromas")}
    flop = 2.0 * cells * sum(d.values())
    with Engine(0) as eng:
        t0 = time.time()
        eng.ef_set_tracks(feats)
        t_up = time.time() - t0
        eng.ef_score_pairs(pairs[:64])                                   # warm-up (allocations, first launches)
        eng.ef_score_pairs(pairs)
        eng.set_profiling(True)
        times = []
        for _ in range(a.reps):
            t0 = time.time()
            got = eng.ef_score_pairs(pairs)
            times.append(time.time() - t0)
        ms = eng.ef_stage_ms()
        st = eng.ef_last_stats()
        eng.set_profiling(False)
    best = min(times)
    dev_ms = sum(ms.values()) / a.reps
    # CPU oracle on a bounded sample
    csel = rng.choice(len(pairs), size=min(a.cpu_pairs, len(pairs)), replace=False)
    t0 = time.time()
    same = True
    for k in csel:
        w = ef.similarity_pair(feats[pairs[k, 0]], feats[pairs[k, 1]])
        same &= all(got[q, k] == np.float32(w[s]) for q, s in enumerate(("mfccs", "ssms", "chromas", "early")))
    t_cpu = time.time() - t0
    out = dict(
        workload="EarlyFusion pair scoring, %d tracks x ~%d blocks, dims %s, float32 features" % (n, a.blocks, d),
        pairs=int(len(pairs)), cells=cells, pairs_per_s=len(pairs) / best, wall_s=best,
        device_ms_per_call={k: v / a.reps for k, v in ms.items()}, device_pairs_per_s=len(pairs) / (dev_ms * 1e-3),
        csm_tflops_f64=flop / (ms["csm"] / a.reps * 1e-3) / 1e12,
        gcups_sw=4 * cells / (ms["sw"] / a.reps * 1e-3) / 1e9,
        launches=st["launches"], chunks=st["chunks"], upload_s=t_up, upload_bytes=eng.ef_h2d_bytes, gen_s=t_gen,
        cpu_oracle=dict(pairs=int(len(csel)), pairs_per_s=len(csel) / t_cpu, threads=os.cpu_count(),
                        kind="port (numpy oracle, BLAS threads)", identical_scores=bool(same)),
    )
    print(json.dumps(out))
    if a.out:
        os.makedirs(os.path.dirname(a.out) or ".", exist_ok=True)
        with open(a.out, "w") as f:
            json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
