"""CPU oracle for the acoss all-pairs scoring hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import, link or execute it, and only as the
checker / reported CPU baseline.  The product path (``acoss_b200``) never
imports this package and fails loudly when its CUDA library is missing.

PARITY STATUS
-------------
* Serra09 flavour (``serra09_np`` / ``serra09_c.c``): **parity unpinned**.
  The arithmetic lives in essentia (``ChromaCrossSimilarity`` /
  ``CoverSongSimilarity``), an unpinned, un-vendored dependency of the reference
  (``/root/reference/setup.py:53``) that is absent from this image.  The oracle
  restates essentia's published algorithm (flags F1-F8 are
  explicit switches) and is anchored on the reference's call sites
  (``acoss/algorithms/rqa_serra09.py:55-69``).
* EarlyFusion flavour (``earlyfusion_np``): **pinned** against the reference's
  own in-tree code executed in this container (``tests/golden/make_golden.py``
  -> ``tests/golden/*.npz|json``).
  The full pair scoring (``similarity_pair``, ``get_wcsm``) is pinned by the
  reference's unmodified ``EarlyFusion.similarity`` method executed here
  (``tests/golden/make_golden_earlyfusion_full.py``).
* Evaluation tail (``evalstats_np``): **pinned** against the reference's
  ``CoverAlgorithm.getEvalStatistics`` executed in this container.
* Feature on-ramp (``onramp_np.median_sync``): **parity unpinned** — restates
  ``librosa.util.sync(..., aggregate=np.median)`` (librosa is absent from the image),
  checked against the block-by-block ``np.median`` definition.
"""
