import time, numpy as np, sys
sys.path.insert(0, '.')
from acoss_b200 import Engine, pack_tracks, synthetic, default_params
from acoss_b200._lib import CRP_EXACT
tracks, labels = synthetic.config_dataset("C1", max_tracks=64)
frames, offs = pack_tracks(tracks)
pairs = synthetic.all_pairs_upper(len(tracks))
cells = synthetic.pair_cells([len(t) for t in tracks], pairs)
with Engine(0) as eng:
    eng.set_tracks(frames, offs)
    se = eng.score_pairs(pairs, default_params(crp_path=CRP_EXACT))
    for rep in range(3):
        eng.set_profiling(True)
        t = time.time(); s = eng.score_pairs(pairs); dt = time.time() - t
        print("fast path: %d pairs %.3fs -> %.1f pairs/s, %.2f GCUPS" % (len(pairs), dt, len(pairs)/dt, cells/dt/1e9), eng.last_stats(), eng.stage_ms())
    print("fast == exact:", np.array_equal(s, se), "mismatches", int((s != se).sum()))
    bad = np.nonzero(s != se)[0][:5]
    for b in bad: print(pairs[b], s[b], se[b], [len(tracks[i]) for i in pairs[b]])
