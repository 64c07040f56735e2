import numpy as np, sys
sys.path.insert(0, '.')
from acoss_b200 import Engine, pack_tracks, synthetic, default_params
tracks, labels = synthetic.config_dataset("C1", max_tracks=64)
frames, offs = pack_tracks(tracks)
pairs = synthetic.all_pairs_upper(len(tracks))
lens = np.array([len(t) for t in tracks])
with Engine(0) as eng:
    eng.set_tracks(frames, offs)
    s = eng.score_pairs(pairs); print(eng.last_stats())
    # find which pairs fall back: score one by one in small groups
    bad = []
    for k in range(0, len(pairs), 32):
        eng.score_pairs(pairs[k:k+32]); st = eng.last_stats()
        if st['fallback_pairs']:
            for kk in range(k, min(k+32, len(pairs))):
                eng.score_pairs(pairs[kk:kk+1]); st2 = eng.last_stats()
                if st2['fallback_pairs']: bad.append((kk, st2['status_or']))
    for kk, r in bad[:40]: print(pairs[kk], lens[pairs[kk]] - 9, "reason", r, "score", s[kk])
