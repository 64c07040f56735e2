"""ctypes wrapper around the plain-C oracle (TEST INFRASTRUCTURE / CPU baseline only)."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import build as _build

_LIB = None


class OracleParams(C.Structure):
    _fields_ = [("m", C.c_int), ("tau", C.c_int), ("kappa", C.c_float), ("oti", C.c_int),
                ("noti", C.c_int), ("gamma_o", C.c_float), ("gamma_e", C.c_float),
                ("integer_guard", C.c_int), ("hoist_norms", C.c_int)]


def params(m=9, tau=1, kappa=0.095, oti=True, noti=12, gamma_o=0.5, gamma_e=0.5,
           integer_guard=False, hoist_norms=False) -> OracleParams:
    return OracleParams(m, tau, float(np.float32(kappa)), int(bool(oti)), noti, gamma_o, gamma_e,
                        int(integer_guard), int(hoist_norms))


def lib():
    global _LIB
    if _LIB is None:
        path = _build.OUT
        if not os.path.exists(path) or os.path.getmtime(path) < os.path.getmtime(_build.SRC):
            _build.build()
        L = C.CDLL(path)
        fp = C.POINTER(C.c_float)
        L.oracle_oti.restype = C.c_int
        L.oracle_oti.argtypes = [fp, C.c_int, fp, C.c_int, C.c_int]
        L.oracle_qmax.restype = C.c_float
        L.oracle_qmax.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_float]
        L.oracle_dmax.restype = C.c_float
        L.oracle_dmax.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_float, C.c_int]
        L.oracle_chen_pairs.restype = C.c_int
        L.oracle_chen_pairs.argtypes = [fp, C.c_void_p, C.c_void_p, C.c_int64,
                                        C.POINTER(OracleParams), C.c_int, fp, fp]
        L.oracle_sw_constrained.restype = C.c_double
        L.oracle_sw_constrained.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.oracle_serra09_pair.restype = C.c_int
        L.oracle_serra09_pair.argtypes = [fp, C.c_int, fp, C.c_int, C.POINTER(OracleParams), fp,
                                          C.POINTER(C.c_int), C.c_void_p, C.c_void_p, C.c_void_p,
                                          C.c_void_p]
        L.oracle_serra09_pairs.restype = C.c_int
        L.oracle_serra09_pairs.argtypes = [fp, C.c_void_p, C.c_void_p, C.c_int64,
                                           C.POINTER(OracleParams), C.c_int, fp]
        L.oracle_sw_batch.restype = C.c_int
        L.oracle_sw_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_void_p]
        _LIB = L
    return _LIB


def _fp(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def oti(q, r, noti=12) -> int:
    q = np.ascontiguousarray(q, np.float32); r = np.ascontiguousarray(r, np.float32)
    return int(lib().oracle_oti(_fp(q), len(q), _fp(r), len(r), noti))


def qmax(crp, gamma_o=0.5, gamma_e=0.5) -> float:
    c = np.ascontiguousarray(crp, np.uint8)
    return float(lib().oracle_qmax(c.ctypes.data, c.shape[0], c.shape[1], gamma_o, gamma_e))


def dmax(crp, gamma_o=0.5, gamma_e=0.5, bonus=True) -> float:
    c = np.ascontiguousarray(crp, np.uint8)
    return float(lib().oracle_dmax(c.ctypes.data, c.shape[0], c.shape[1], gamma_o, gamma_e, int(bonus)))


def sw_constrained(B) -> float:
    b = np.ascontiguousarray(B, np.uint8)
    return float(lib().oracle_sw_constrained(b.ctypes.data, b.shape[0], b.shape[1]))


def pair(q, r, p: OracleParams | None = None, want_debug=False):
    """Returns score, or (score, dict(oti, crp, thr_q, thr_r, d)) with want_debug."""
    p = p or params()
    q = np.ascontiguousarray(q, np.float32); r = np.ascontiguousarray(r, np.float32)
    M = len(q) - p.m * p.tau if p.m > 1 else len(q)
    N = len(r) - p.m * p.tau if p.m > 1 else len(r)
    score = C.c_float(0); o = C.c_int(0)
    if want_debug and M > 0 and N > 0:
        crp = np.zeros((M, N), np.uint8); tq = np.zeros(M, np.float32); tr = np.zeros(N, np.float32)
        d = np.zeros((M, N), np.float32)
        rc = lib().oracle_serra09_pair(_fp(q), len(q), _fp(r), len(r), C.byref(p), C.byref(score),
                                       C.byref(o), crp.ctypes.data, tq.ctypes.data, tr.ctypes.data,
                                       d.ctypes.data)
        if rc == -1:
            raise RuntimeError("oracle: empty or too-short input")
        dbg = dict(oti=o.value, crp=crp, thr_q=tq, thr_r=tr, d=d, rc=rc)
        if rc == -2:
            raise RuntimeError("oracle: NaN distance (F7)")
        return float(score.value), dbg
    rc = lib().oracle_serra09_pair(_fp(q), len(q), _fp(r), len(r), C.byref(p), C.byref(score),
                                   C.byref(o), None, None, None, None)
    if rc == -1:
        raise RuntimeError("oracle: empty or too-short input")
    if rc == -2:
        raise RuntimeError("oracle: NaN distance (F7)")
    return float(score.value)


def pairs(frames, offsets, pair_idx, p: OracleParams | None = None, nthreads=1):
    """Batched multi-threaded driver: the timed 'reference CPU path' (kind='port')."""
    p = p or params()
    frames = np.ascontiguousarray(frames, np.float32)
    offsets = np.ascontiguousarray(offsets, np.int64)
    pair_idx = np.ascontiguousarray(pair_idx, np.int32)
    out = np.zeros(len(pair_idx), np.float32)
    rc = lib().oracle_serra09_pairs(_fp(frames), offsets.ctypes.data, pair_idx.ctypes.data,
                                    len(pair_idx), C.byref(p), int(nthreads), _fp(out))
    if rc != 0:
        raise RuntimeError("oracle: pair batch failed rc=%d" % rc)
    return out


def chen_pairs(frames, offsets, pair_idx, p: OracleParams | None = None, nthreads=1):
    """ChenFusion.similarity over a batch: (qmax, dmax) score arrays of the same CRPs."""
    p = p or params()
    frames = np.ascontiguousarray(frames, np.float32)
    offsets = np.ascontiguousarray(offsets, np.int64)
    pair_idx = np.ascontiguousarray(pair_idx, np.int32)
    q = np.zeros(len(pair_idx), np.float32); d = np.zeros(len(pair_idx), np.float32)
    rc = lib().oracle_chen_pairs(_fp(frames), offsets.ctypes.data, pair_idx.ctypes.data, len(pair_idx),
                                 C.byref(p), int(nthreads), _fp(q), _fp(d))
    if rc != 0:
        raise RuntimeError("oracle: pair batch failed rc=%d" % rc)
    return q, d


def sw_batch(mats, nthreads=1):
    shapes = np.array([m.shape for m in mats], np.int32)
    sizes = np.array([m.size for m in mats], np.int64)
    off = np.concatenate([[0], np.cumsum(sizes)[:-1]]).astype(np.int64)
    buf = np.concatenate([np.ascontiguousarray(m, np.uint8).ravel() for m in mats])
    out = np.zeros(len(mats), np.float64)
    lib().oracle_sw_batch(buf.ctypes.data, off.ctypes.data, shapes.ctypes.data, len(mats), nthreads,
                          out.ctypes.data)
    return out
