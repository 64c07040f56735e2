"""Debug helper (GPU box): fast vs exact CRP path on a small synthetic set; prints mismatching pairs and
the fallback reason bits of every pair that fell back."""
import sys, time, collections
import numpy as np
sys.path.insert(0, '.')
from acoss_b200 import Engine, pack_tracks, synthetic, default_params
from acoss_b200._lib import CRP_EXACT
name = sys.argv[1] if len(sys.argv) > 1 else "tiny"
mt = int(sys.argv[2]) if len(sys.argv) > 2 else None
tracks, labels = synthetic.config_dataset(name, max_tracks=mt)
frames, offs = pack_tracks(tracks)
pairs = synthetic.all_pairs_upper(len(tracks))
lens = np.array([len(t) for t in tracks])
print("tracks", len(tracks), "pairs", len(pairs), "lens", lens.min(), lens.max())
with Engine(0) as eng:
    eng.set_tracks(frames, offs)
    se = eng.score_pairs(pairs, default_params(crp_path=CRP_EXACT))
    eng.set_profiling(True)
    t = time.time(); s = eng.score_pairs(pairs); dt = time.time() - t
    print("fast: %.3fs %.1f pairs/s" % (dt, len(pairs) / dt), eng.last_stats(), eng.stage_ms())
    print("levels (strips, live, miss, left):", eng.debug_counters())
    bad = np.nonzero(s != se)[0]
    print("mismatches", len(bad))
    reasons = collections.Counter()
    nfb = 0
    for k in range(len(pairs)):
        if len(pairs) > 400 and k not in bad[:20]:
            continue
        eng.score_pairs(pairs[k:k + 1]); st = eng.last_stats()
        if st['fallback_pairs'] or k in bad:
            nfb += 1
            reasons[st['status_or']] += 1
            if k in bad or nfb < 12:
                print(" pair", pairs[k], "M',N'", lens[pairs[k]] - 9, "status", st['status_or'], "fast", s[k], "exact", se[k])
    print("reasons", dict(reasons))
