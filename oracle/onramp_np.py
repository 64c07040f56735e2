"""numpy restatement of the median aggregation inside ``Serra09.load_features`` (TEST INFRASTRUCTURE).

Reference followed: /root/reference/acoss/algorithms/rqa_serra09.py:47-53 —
``librosa.util.sync(chroma.T, np.arange(0, n, downsample_fac), aggregate=np.median).T`` (librosa 0.6.1 pads the
boundaries with 0 and n, so the last block may be short).  librosa is absent from the image, so this is a
restatement of its documented behaviour; ``tests/test_plugin_cpu.py::test_median_sync_definition`` checks it against
the block-by-block ``np.median`` definition.  Only tests may import this module (the product computes the medians
on the GPU: acoss_set_tracks_raw / k0_onramp.cu)."""
from __future__ import annotations

import numpy as np

__all__ = ["median_sync"]


def median_sync(chroma: np.ndarray, fac: int) -> np.ndarray:
    """Per-bin median of the blocks [fac*k, min(fac*k+fac, n)); output dtype = input dtype."""
    chroma = np.asarray(chroma)
    n = chroma.shape[0]
    if fac <= 1:
        return chroma.copy()
    nblk = (n + fac - 1) // fac
    out = np.empty((nblk, chroma.shape[1]), dtype=chroma.dtype)
    full = n // fac
    if full:
        out[:full] = np.median(chroma[:full * fac].reshape(full, fac, -1), axis=1)
    if nblk > full:
        out[full] = np.median(chroma[full * fac:], axis=0)
    return out
