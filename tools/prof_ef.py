"""Small EarlyFusion scoring run for ncu (one warm-up call, one profiled-size call)."""
import sys
import numpy as np
sys.path.insert(0, '.')
from acoss_b200 import Engine, synthetic
npairs = int(sys.argv[1]) if len(sys.argv) > 1 else 256
feats = synthetic.ef_dataset([2] * 16, 400, 20242)
i, j = np.triu_indices(len(feats), k=1)
pairs = np.stack([i, j], axis=1).astype(np.int32)[:npairs]
with Engine(0) as eng:
    eng.ef_set_tracks(feats)
    for rep in range(2):
        s = eng.ef_score_pairs(pairs)
print(len(pairs), s[:, :3])
