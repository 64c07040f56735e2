#!/usr/bin/env python
"""Full all-pairs EarlyFusion run of a covers80-shaped synthetic set on 1..8 GPUs (one process per GPU).

    python tools/run_allpairs_ef.py --tracks 160 --blocks 400
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 \
        tools/run_allpairs_ef.py

The sequence of the reference's __main__ (earlyfusion_traile.py:269-281): all_pairwise(symmetric=True) over the
drop-in plugin -> getEvalStatistics per score type, with the pair list sharded across ranks by cross-similarity
cells (acoss_b200/distributed.py), ONE NCCL all_gather for the four score rows, evaluation on rank 0.  A random
sample of pairs is re-scored by the CPU oracle on rank 0.  Prints one JSON line.
"""
import argparse
import contextlib
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
    ap.add_argument("--oracle-sample", type=int, default=12)
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    import torch
    import torch.distributed as dist
    from acoss_b200 import synthetic
    from acoss_b200.distributed import all_pairwise_distributed
    from acoss_b200.earlyfusion import EarlyFusion
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    t0 = time.perf_counter()
    feats = synthetic.ef_dataset([2] * (args.tracks // 2), args.blocks, 20242)
    t_gen = time.perf_counter() - t0
    with contextlib.redirect_stdout(sys.stderr):
        alg = EarlyFusion(None, None, features=feats, shortname="ef_r%d" % rank, device=local,
                          cachedir="/tmp/acoss_allpairs_ef_%d" % rank)
    t0 = time.perf_counter()
    alg._ensure_resident()
    t_upload = time.perf_counter() - t0
    alg.score_pairs(np.array([[0, 1], [2, 3]]))                    # warm-up: first launches, allocations

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    sync_all()
    t0 = time.perf_counter()
    bounds = all_pairwise_distributed(alg, symmetric=True)
    sync_all()
    t_all = time.perf_counter() - t0
    tt = torch.tensor([t_all], dtype=torch.float64, device="cuda:%d" % local)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    t_all = float(tt[0])
    if rank == 0:
        pairs = alg._pair_array(True)
        w = alg.pair_weights(pairs)
        out = dict(workload="EarlyFusion all-pairs, %d tracks x ~%d blocks, dims %s" % (alg.N, args.blocks, synthetic.EF_DIMS),
                   n_gpus=world, pairs=int(len(pairs)), cells=int(w.sum()), t_generate_s=t_gen, t_upload_s=t_upload,
                   t_all_pairwise_s=t_all, pairs_per_s=len(pairs) / t_all, shard_pairs=[int(x) for x in np.diff(bounds)])
        if args.oracle_sample:
            from oracle import earlyfusion_np as ef
            sel = np.random.default_rng(3).permutation(len(pairs))[:args.oracle_sample]
            t0 = time.perf_counter()
            same = True
            for k in sel:
                i, j = pairs[k]
                want = ef.similarity_pair(feats[i], feats[j], alg.kappa, alg.K)
                same &= all(alg.Ds[s][i, j] == np.float32(want[s]) and alg.Ds[s][j, i] == np.float32(want[s]) for s in want)
            out["oracle_sample"] = dict(pairs=int(len(sel)), identical=bool(same), cpu_s=time.perf_counter() - t0,
                                        cores=os.cpu_count())
        ev = {}
        with contextlib.redirect_stdout(sys.stderr):
            for s in alg.Ds:
                MR, MRR, MDR, MAP, tops = alg.getEvalStatistics(s)
                ev[s] = dict(MR1=float(MR), MRR=float(MRR), MDR=float(MDR), MAP=float(MAP))
        out["eval"] = ev
        print(json.dumps(out), flush=True)
    alg.cleanup_memmap()
    alg.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
