"""INTEGRATION.md section B, executable: the binding a maintainer of the reference adds to
``acoss/algorithms/rqa_serra09.py`` so that ``Serra09.similarity`` runs on the C ABI of ``libacoss_b200.so``.

``bind_serra09(RefSerra09)`` returns a subclass of the class it is given — the reference's own, unmodified
``acoss.algorithms.rqa_serra09.Serra09`` (on its own ``CoverAlgorithm`` base) or this package's mirror — that
overrides ONLY ``similarity`` (``rqa_serra09.py:55-69``) and adds the ``_gpu`` helper.  Everything else (``__init__``,
``load_features``, ``all_pairwise``, ``normalize_by_length``, ``getEvalStatistics``, ``cleanup_memmap``, the ``Ds``
memmaps, ``cliques``) stays the base class's code.  Only ctypes and numpy are used, no other module of this package,
exactly like the stub printed in INTEGRATION.md.

``lib`` is the ctypes handle of the library (default: the in-tree ``libacoss_b200.so``).
``tests/test_reference_class.py`` runs this binding under the reference's real classes.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

__all__ = ["bind_serra09", "load_library"]

_HERE = os.path.dirname(os.path.abspath(__file__))


class _Params(C.Structure):                      # struct acoss_params (include/acoss_b200.h)
    _fields_ = [("m", C.c_int32), ("tau", C.c_int32), ("kappa", C.c_float), ("oti", C.c_int32),
                ("noti", C.c_int32), ("gamma_o", C.c_float), ("gamma_e", C.c_float),
                ("align", C.c_int32), ("integer_guard", C.c_int32), ("crp_path", C.c_int32),
                ("f2_strict", C.c_int32), ("f3_float_acc", C.c_int32), ("f4_keep_last", C.c_int32),
                ("f5_asymmetric", C.c_int32)]


def load_library(path: str | None = None):
    lib = C.CDLL(path or os.path.join(_HERE, "csrc", "libacoss_b200.so"))
    lib.acoss_last_error.restype = C.c_char_p
    return lib


def bind_serra09(RefSerra09, lib=None, device: int = 0):
    """Subclass of ``RefSerra09`` whose ``similarity(idxs)`` calls ``acoss_score_pairs`` (one batched call per
    ``idxs``; the reference's serial ``all_pairwise`` passes one pair at a time, which works and is merely slow)."""
    _lib = lib if lib is not None else load_library()
    _lib.acoss_last_error.restype = C.c_char_p

    def _check(rc):
        if rc != 0:
            raise RuntimeError(_lib.acoss_last_error().decode())   # essentia raised RuntimeError too

    class Serra09(RefSerra09):
        def _gpu(self):
            if not hasattr(self, "_ctx"):
                self._ctx = C.c_void_p()
                _check(_lib.acoss_create(C.byref(self._ctx), device))
                tracks = [np.ascontiguousarray(self.load_features(i), np.float32) for i in range(self.N)]
                offsets = np.concatenate([[0], np.cumsum([len(t) for t in tracks])]).astype(np.int64)
                frames = np.ascontiguousarray(np.concatenate(tracks), np.float32)
                # replaces the per-process feature cache self.all_feats (rqa_serra09.py:44-53)
                _check(_lib.acoss_set_tracks(self._ctx, frames.ctypes.data_as(C.c_void_p),
                                             offsets.ctypes.data_as(C.c_void_p), self.N, 0))
            return self._ctx

        def similarity(self, idxs):                  # replaces rqa_serra09.py:55-69
            pairs = np.ascontiguousarray(idxs, np.int32).reshape(-1, 2)
            scores = np.empty(len(pairs), np.float32)
            p = _Params()
            _lib.acoss_default_params(C.byref(p))
            p.m, p.tau, p.kappa, p.oti = int(self.m), int(self.tau), float(self.kappa), int(bool(self.oti))
            _check(_lib.acoss_score_pairs(self._gpu(), pairs.ctypes.data_as(C.c_void_p), C.c_int64(len(pairs)),
                                          C.byref(p), scores.ctypes.data_as(C.c_void_p)))
            for key in self.Ds.keys():
                self.Ds[key][pairs[:, 0], pairs[:, 1]] = scores

        def release(self):
            if hasattr(self, "_ctx"):
                _lib.acoss_destroy(self._ctx)
                del self._ctx

    Serra09.__name__ = "Serra09"
    return Serra09
