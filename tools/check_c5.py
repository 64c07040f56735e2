import time, numpy as np, sys
sys.path.insert(0, '.')
from acoss_b200 import Engine, pack_tracks, synthetic, default_params
from acoss_b200._lib import CRP_EXACT
t=time.time(); tracks, labels = synthetic.config_dataset("C5", max_tracks=26); print("gen", time.time()-t, [len(x) for x in tracks][:6])
frames, offs = pack_tracks(tracks)
pairs = synthetic.all_pairs_upper(len(tracks))
cells = synthetic.pair_cells([len(t) for t in tracks], pairs)
with Engine(0) as eng:
    eng.set_tracks(frames, offs)
    t=time.time(); se = eng.score_pairs(pairs, default_params(crp_path=CRP_EXACT)); print("exact %.2fs"%(time.time()-t))
    for rep in range(2):
        eng.set_profiling(True)
        t = time.time(); s = eng.score_pairs(pairs); dt = time.time() - t
        print("fast: %d pairs %.3fs -> %.1f pairs/s, %.2f GCUPS" % (len(pairs), dt, len(pairs)/dt, cells/dt/1e9), eng.last_stats(), eng.stage_ms())
    print("fast == exact:", np.array_equal(s, se), int((s != se).sum()))
    from oracle import serra09_c as oc
    idx = np.random.default_rng(0).permutation(len(pairs))[:16]
    t=time.time(); want = oc.pairs(frames, offs, pairs[idx], nthreads=16); print("oracle 16 pairs %.1fs" % (time.time()-t))
    print("gpu == oracle:", np.array_equal(want, s[idx]), want[:6], s[idx][:6])
