"""Dump the float64 matrices of one pair with one generation of the CSM kernel (ACOSS_EF_CSM=1|2|3) to a file,
or compare two such files bit for bit.  The superseded generations are compiled only into a library built with
ACOSS_NVCC_EXTRA=-DACOSS_EF_GENERATIONS (python acoss_b200/csrc/build.py); the shipped library carries the DMMA kernel and the
generic DFMA kernel it falls back to for chroma blocks wider than 2048."""
import sys
import numpy as np
sys.path.insert(0, '.')
if sys.argv[1] == "dump":
    from acoss_b200 import Engine, synthetic
    feats = synthetic.ef_dataset([2, 1], 300, 512, jitter=0.2)
    with Engine(0) as eng:
        eng.ef_set_tracks(feats)
        d = eng.ef_dump_pair(0, 2, 0.1, 10)
    np.save(sys.argv[2], d["csms"])
else:
    a, b = np.load(sys.argv[2]), np.load(sys.argv[3])
    for k, name in enumerate(("mfccs", "ssms", "chromas", "early")):
        diff = np.abs(a[k] - b[k])
        print(name, "bit-identical" if np.array_equal(a[k], b[k]) else "max abs diff %.3g (rel %.3g), %d of %d cells differ"
              % (diff.max(), (diff / np.abs(a[k]).clip(1e-300)).max(), int((a[k] != b[k]).sum()), a[k].size))
