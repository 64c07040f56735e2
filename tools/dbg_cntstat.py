import sys, numpy as np
sys.path.insert(0, '.')
from acoss_b200 import Engine, pack_tracks, synthetic
import ctypes as C
for cfg, mt, npairs in (("C3", 400, 4096), ("C4s", 1000, 32768)):
    tracks, labels = synthetic.config_dataset(cfg, max_tracks=mt)
    frames, offs = pack_tracks(tracks)
    pairs = synthetic.all_pairs_upper(len(tracks))
    pairs = pairs[np.random.default_rng(7).permutation(len(pairs))[:npairs]].astype(np.int32)
    with Engine(0) as eng:
        eng.set_tracks(frames, offs)
        eng.score_pairs(pairs)
        d = np.zeros(32, dtype=np.int64)
        eng._lib.acoss_debug_counters(eng._ctx, d.ctypes.data)
        lines, s1, s2, mx, big = d[26], d[27], d[28] * 64, d[29], d[30]
        print(cfg, "lines", lines, "mean cnt %.1f" % (s1 / lines), "mean cnt^2 %.0f" % (s2 / lines), "max", mx, "frac>64 %.4f" % (big / lines))
