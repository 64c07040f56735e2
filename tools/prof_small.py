import numpy as np, sys
sys.path.insert(0, '.')
from acoss_b200 import Engine, pack_tracks, synthetic
tracks, labels = synthetic.config_dataset("C1", max_tracks=34)
frames, offs = pack_tracks(tracks)
pairs = synthetic.all_pairs_upper(len(tracks))
with Engine(0) as eng:
    eng.set_tracks(frames, offs)
    for rep in range(2):
        s = eng.score_pairs(pairs)
print(len(pairs), s[:4])
