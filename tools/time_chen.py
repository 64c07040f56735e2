"""GPU box: ChenFusion (Qmax + Dmax over one CRP) against Serra09 (Qmax) on a C3 slice, per-stage milliseconds."""
import sys, time
import numpy as np
sys.path.insert(0, '.')
from acoss_b200 import Engine, pack_tracks, synthetic

tracks, labels = synthetic.config_dataset("C3", max_tracks=120)
frames, offs = pack_tracks(tracks)
pairs = synthetic.all_pairs_upper(len(tracks))
cells = synthetic.pair_cells([len(t) for t in tracks], pairs)
with Engine(0) as eng:
    eng.set_tracks(frames, offs)
    eng.score_pairs(pairs[:512]); eng.score_pairs_chen(pairs[:512])
    for name, fn in (("serra09", eng.score_pairs), ("chen", eng.score_pairs_chen)):
        eng.set_profiling(True)
        t = time.time(); fn(pairs); dt = time.time() - t
        print("%-8s %d pairs %.3fs -> %.0f pairs/s, %.1f GCUPS" % (name, len(pairs), dt, len(pairs) / dt, cells / dt / 1e9), eng.stage_ms())
