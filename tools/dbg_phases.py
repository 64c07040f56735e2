#!/usr/bin/env python
"""Cycle budget of a tensor histogram CTA (variant build with -DTC_PHASES: tools/build_variant.sh phases -DTC_PHASES).
   ACOSS_B200_LIB=acoss_b200/csrc/variants/lib_phases.so python tools/dbg_phases.py [--config C4s]"""
import argparse, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from acoss_b200 import Engine, pack_tracks, synthetic
from acoss_b200._lib import check
ap = argparse.ArgumentParser(); ap.add_argument("--config", default="C3"); ap.add_argument("--pairs", type=int, default=2048)
a = ap.parse_args()
tracks, _ = synthetic.config_dataset(a.config, max_tracks=400)
frames, offs = pack_tracks(tracks)
e = Engine(0); e.set_tracks(frames, offs)
pairs = synthetic.all_pairs_upper(len(tracks)); pairs = pairs[np.random.default_rng(1).permutation(len(pairs))[:a.pairs]].astype(np.int32)
e.score_pairs(pairs)
d = np.zeros(32, dtype=np.int64); check(e._lib.acoss_debug_counters(e._ctx, d.ctypes.data))
# orientation 0 (columns) uses dbg[20..23, 26..28]; the orientation-1 offsets overlap other counters and are ignored here
n = max(1, d[27])
names = ["begin (alloc, barriers, zero)", "to first block consumed", "rest of the sweep", "bar + scan + decision", "level 2 / exit (tc_end)"]
vals = [d[20], d[21], d[22], d[23], d[26]]
print("CTAs %d, blocks per CTA %.1f" % (n, d[28] / n))
for nm, v in zip(names, vals):
    print("  %-32s %8.0f cycles per CTA" % (nm, 64.0 * v / n))
