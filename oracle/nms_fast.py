"""ctypes wrapper around oracle/nms_c.c (TEST INFRASTRUCTURE ONLY)."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build():
    subprocess.run(["make", "-s", "-C", _HERE], check=True)


def _lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "liboracle_nms.so")
        if not os.path.exists(path):
            build()
        _LIB = ctypes.CDLL(path)
        _LIB.oracle_nms.restype = ctypes.c_int
    return _LIB


def nms_pick_order_c(boxes, scores, classes, iou_threshold=0.5):
    b = np.ascontiguousarray(np.asarray(boxes, dtype=np.float64).reshape(-1, 4))
    s = np.ascontiguousarray(np.asarray(scores, dtype=np.float64))
    c = np.ascontiguousarray(np.asarray(classes, dtype=np.float64))
    n = len(s)
    out = np.empty(max(n, 1), dtype=np.int32)
    P = ctypes.c_void_p
    k = _lib().oracle_nms(P(b.ctypes.data), P(s.ctypes.data), P(c.ctypes.data),
                          ctypes.c_int32(n), ctypes.c_double(iou_threshold), P(out.ctypes.data))
    return out[:k].copy()
