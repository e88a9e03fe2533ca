"""ctypes loader of oracle/knn_brute.c (TEST INFRASTRUCTURE ONLY)."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "_build", "libknn_brute.so")
_lib = None


def _load():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB):
            subprocess.run(["make", "-s", "-C", _HERE], check=True)
        _lib = C.CDLL(_LIB)
        _lib.knn_brute.restype = C.c_int
        _lib.knn_brute.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_void_p]
    return _lib


def knn_brute(points, queries, k):
    """(dist, idx, d2), each (nq,k), ascending by (d2, index) -- same contract as
    reference_port.knn_bruteforce / knn_canonical."""
    points = np.ascontiguousarray(points, dtype=np.float64)
    queries = np.ascontiguousarray(queries, dtype=np.float64)
    nq = len(queries)
    idx = np.empty((nq, k), dtype=np.int64)
    d2 = np.empty((nq, k), dtype=np.float64)
    rc = _load().knn_brute(points.ctypes.data, len(points), queries.ctypes.data, nq, k, idx.ctypes.data, d2.ctypes.data)
    if rc != 0:
        raise IndexError(f"k={k} exceeds the number of particles {len(points)}")
    return np.sqrt(d2), idx, d2
