import time, numpy as np, sys
sys.path.insert(0, '.')
from acoss_b200 import Engine, pack_tracks, synthetic, default_params
tracks, labels = synthetic.config_dataset("C1", max_tracks=40)
frames, offs = pack_tracks(tracks)
pairs = synthetic.all_pairs_upper(len(tracks))
cells = synthetic.pair_cells([len(t) for t in tracks], pairs)
with Engine(0) as eng:
    eng.set_tracks(frames, offs)
    for rep in range(2):
        t = time.time(); s = eng.score_pairs(pairs); dt = time.time() - t
        print("exact path: %d pairs %.3fs -> %.1f pairs/s, %.2f GCUPS" % (len(pairs), dt, len(pairs)/dt, cells/dt/1e9), eng.last_stats())
    from oracle import serra09_c as oc
    idx = np.arange(0, len(pairs), max(1, len(pairs)//16))[:16]
    t = time.time(); want = oc.pairs(frames, offs, pairs[idx], nthreads=16); print("oracle 16 pairs", time.time()-t)
    print("match", np.array_equal(want, s[idx]), want[:8], s[idx][:8])
    # DP only timing
    rng = np.random.default_rng(0)
    mats = [(rng.random((1991, 1991)) < 0.05).astype(np.uint8) for _ in range(64)]
    from acoss_b200.engine import ALIGN_QMAX, ALIGN_SW
    for mode in (ALIGN_QMAX, ALIGN_SW):
        t = time.time(); r = eng.dp_bytes(mats * 8, mode); dt = time.time() - t
        print("dp mode", mode, "512 mats 1991^2 in %.3fs (incl H2D)" % dt, r[:3])
