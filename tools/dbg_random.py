"""Debug helper (GPU box): the track set of tests/test_gpu_properties.py::test_random_shapes_vs_oracle;
prints the fallback reason bits of every pair the fast CRP path hands to the exact path."""
import collections
import sys

import numpy as np

sys.path.insert(0, '.')
from acoss_b200 import Engine, pack_tracks

F32 = np.float32


def hp(rng, n):
    X = rng.random((n, 12)).astype(F32)
    return (X / X.max(1, keepdims=True)).astype(F32)


rng = np.random.default_rng(2024)
lens = list(rng.integers(11, 700, size=34)) + [11, 12, 13, 210, 410, 610, 209, 211, 137, 64, 65, 66, 129]
tracks = [hp(rng, int(n)) for n in lens]
tracks.append(np.roll(tracks[3], 5, axis=1))
tracks.append((tracks[5] + F32(0.05) * rng.random(tracks[5].shape).astype(F32)).astype(F32))
tracks.append(np.repeat(hp(rng, 40), 6, axis=0))
frames, offs = pack_tracks(tracks)
n = len(tracks)
pairs = np.stack([rng.integers(0, n, 220), rng.integers(0, n, 220)], 1).astype(np.int32)
L = np.array([len(t) for t in tracks])
with Engine(0) as eng:
    eng.set_tracks(frames, offs)
    eng.score_pairs(pairs)
    print("batch:", eng.last_stats(), eng.debug_counters())
    reasons = collections.Counter()
    for k in range(len(pairs)):
        eng.score_pairs(pairs[k:k + 1])
        st = eng.last_stats()
        if st['fallback_pairs']:
            reasons[st['status_or']] += 1
            print(" pair", pairs[k], "M',N'", L[pairs[k]] - 9, "status 0x%x" % st['status_or'], eng.debug_counters())
    print("reasons", {hex(k): v for k, v in reasons.items()})
