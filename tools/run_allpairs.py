#!/usr/bin/env python
"""Full all-pairs Serra09 run of a synthetic BASELINE config on 1..8 GPUs (one process per GPU).

    python tools/run_allpairs.py --config C4s
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 \
        tools/run_allpairs.py --config C4

The coverid.benchmark sequence (coverid.py:57-70): all_pairwise(symmetric) -> normalize_by_length ->
getEvalStatistics, with the pair list sharded across ranks by DP cells (acoss_b200/distributed.py), one NCCL
all_gather of the score slices, and the evaluation on rank 0.  A random sample of pairs is re-scored by the CPU
oracle (rank 0) as the parity check at sizes the oracle cannot cover in full.  Prints one JSON line.
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
    ap.add_argument("--config", default="C4s")
    ap.add_argument("--max-tracks", type=int, default=None)
    ap.add_argument("--oracle-sample", type=int, default=256, help="pairs re-scored by the CPU oracle (0 = skip)")
    ap.add_argument("--tile", type=int, default=1 << 18, help="pairs per engine call")
    ap.add_argument("--no-eval", action="store_true")
    ap.add_argument("--align", default="qmax", choices=["qmax", "sw"],
                    help="alignment over the CRPs: Serra09's Qmax, or smith_waterman_constrained (BASELINE configs[1], C2)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))

    import torch
    import torch.distributed as dist
    from acoss_b200 import pack_tracks, synthetic
    from acoss_b200.distributed import gather_scores, shard_bounds
    from acoss_b200.serra09 import Serra09
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    t0 = time.perf_counter()
    tracks, labels = synthetic.config_dataset(args.config, max_tracks=args.max_tracks)
    t_gen = time.perf_counter() - t0
    N = len(tracks)
    lens = np.array([len(t) for t in tracks], dtype=np.int64)
    i, j = np.triu_indices(N, k=1)
    pairs = np.stack([i, j], axis=1).astype(np.int32)
    del i, j
    cells = (lens[pairs[:, 0]] - 9) * (lens[pairs[:, 1]] - 9)
    bounds = shard_bounds(cells, world)
    mine = pairs[bounds[rank]:bounds[rank + 1]]

    import contextlib
    with contextlib.redirect_stdout(sys.stderr):
        alg = Serra09(None, None, features=[dict(hpcp=t, label=str(l)) for t, l in zip(tracks, labels)],
                      downsample_fac=1, shortname="%s_r%d" % (args.config, rank), device=local,
                      cachedir="/tmp/acoss_allpairs_%d" % rank) if rank == 0 else None
    from acoss_b200 import Engine, default_params
    eng = Engine(local)
    frames, offs = pack_tracks(tracks)
    t0 = time.perf_counter()
    eng.set_tracks(frames, offs)
    t_upload = time.perf_counter() - t0

    from acoss_b200.engine import ALIGN_QMAX, ALIGN_SW
    run_params = default_params(align=ALIGN_SW if args.align == "sw" else ALIGN_QMAX)

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    sync_all()
    t0 = time.perf_counter()
    parts, fallbacks = [], 0
    for k in range(0, len(mine), args.tile):
        parts.append(eng.score_pairs(mine[k:k + args.tile], run_params))
        fallbacks += eng.last_stats()["fallback_pairs"]
    local_scores = np.concatenate(parts) if parts else np.zeros(0, np.float32)
    sync_all()
    t_score = time.perf_counter() - t0
    t0 = time.perf_counter()
    full = gather_scores(torch.from_numpy(local_scores).cuda(local), bounds, rank, world)
    sync_all()
    t_gather = time.perf_counter() - t0
    full = full.cpu().numpy()
    tt = torch.tensor([t_score, t_gather, float(fallbacks)], dtype=torch.float64, device="cuda:%d" % local)
    if world > 1:
        mx = tt.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = tt.clone(); dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        t_score, t_gather, fallbacks = float(mx[0]), float(mx[1]), int(sm[2])
    if rank == 0:
        out = dict(config=args.config, align=args.align, n_gpus=world, tracks=N, pairs=int(len(pairs)), cells=int(cells.sum()),
                   mean_frames=float(lens.mean()), t_generate_s=t_gen, t_upload_s=t_upload, t_score_s=t_score,
                   t_gather_s=t_gather, pairs_per_s=len(pairs) / t_score, gcups=float(cells.sum()) / t_score / 1e9,
                   fallback_pairs=fallbacks, shard_pairs=[int(x) for x in np.diff(bounds)])
        if args.oracle_sample:
            from oracle import serra09_c as oc
            sel = np.random.default_rng(3).permutation(len(pairs))[:args.oracle_sample]
            t0 = time.perf_counter()
            if args.align == "sw":                             # oracle CRP of each sampled pair -> oracle SW
                from oracle import earlyfusion_np as ef
                want = np.array([ef.smith_waterman_constrained_x10(
                    oc.pair(tracks[a], tracks[b], want_debug=True)[1]["crp"]) / 10.0 for a, b in pairs[sel]], np.float32)
            else:
                want = oc.pairs(frames, offs, pairs[sel], nthreads=os.cpu_count() or 1)
            out["oracle_sample"] = dict(pairs=int(len(sel)), identical=bool(np.array_equal(want, full[sel])),
                                        cpu_s=time.perf_counter() - t0, cores=os.cpu_count())
        if not args.no_eval:
            t0 = time.perf_counter()
            D = alg.Ds["main"]
            D[pairs[:, 0], pairs[:, 1]] = full
            D += D.T                                           # all_pairwise(symmetric=True), algorithm_template.py:189-191
            alg.cliques = {}
            for idx, l in enumerate(labels):
                alg.cliques.setdefault(str(l), set()).add(idx)
            for idx in range(N):
                alg.all_feats[idx] = tracks[idx]
            alg.normalize_by_length()
            out["t_fill_normalize_s"] = time.perf_counter() - t0
            t0 = time.perf_counter()
            with contextlib.redirect_stdout(sys.stderr):
                MR, MRR, MDR, MAP, tops = alg.getEvalStatistics("main")
            out["t_eval_s"] = time.perf_counter() - t0
            out["eval"] = dict(MR1=float(MR), MRR=float(MRR), MDR=float(MDR), MAP=float(MAP), tops=[float(x) for x in tops])
            alg.cleanup_memmap()
        print(json.dumps(out), flush=True)
    eng.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
