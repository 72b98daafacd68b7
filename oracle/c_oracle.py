"""ctypes wrapper around oracle/liboracle.so (plain-C oracle).  TEST INFRASTRUCTURE ONLY."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE, "liboracle.so"])


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(path):
            build()
        _LIB = ctypes.CDLL(path)
        _LIB.oracle_num_threads.restype = ctypes.c_int
    return _LIB


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p) if a is not None else None


def encode(squares, black):
    squares = np.ascontiguousarray(squares, np.int8).reshape(-1, 64)
    black = np.ascontiguousarray(black, np.uint8)
    n = squares.shape[0]
    planes = np.empty((n, 8, 8, 4), np.int8)
    lib().oracle_encode(_p(squares), _p(black), ctypes.c_int64(n), _p(planes))
    return planes


def legal_mask(squares, black):
    squares = np.ascontiguousarray(squares, np.int8).reshape(-1, 64)
    black = np.ascontiguousarray(black, np.uint8)
    n = squares.shape[0]
    mask = np.empty((n, 155), np.uint8)
    lib().oracle_legal_mask(_p(squares), _p(black), ctypes.c_int64(n), _p(mask))
    return mask.astype(bool)


def forward(planes, w, final_k=3, want_h5=False):
    planes = np.ascontiguousarray(planes, np.float32).reshape(-1, 8, 8, 4)
    n = planes.shape[0]
    logits = np.empty((n, 155), np.float32)
    probs = np.empty((n, 155), np.float32)
    h5 = np.empty((n, 64), np.float32) if want_h5 else None
    args = [_p(planes), ctypes.c_int64(n)]
    keep = []
    for i in range(1, 6):
        for k in ("weights", "bias"):
            a = np.ascontiguousarray(w["hidden_layer/%d/%s" % (i, k)], np.float32)
            keep.append(a)
            args.append(_p(a))
    args.append(ctypes.c_int(final_k))
    for k in ("weights", "bias"):
        a = np.ascontiguousarray(w["output_layer/" + k], np.float32)
        keep.append(a)
        args.append(_p(a))
    args += [_p(logits), _p(probs), _p(h5)]
    lib().oracle_forward(*args)
    return (logits, probs, h5) if want_h5 else (logits, probs)


def rank(probs, mask):
    probs = np.ascontiguousarray(probs, np.float32)
    mask = np.ascontiguousarray(mask, np.uint8)
    n = probs.shape[0]
    idx = np.empty((n, 48), np.uint8)
    pri = np.empty((n, 48), np.float32)
    cnt = np.empty(n, np.uint8)
    lib().oracle_rank(_p(probs), _p(mask), ctypes.c_int64(n), _p(idx), _p(pri), _p(cnt))
    return idx, pri, cnt


def num_threads():
    return lib().oracle_num_threads()


def set_threads(n):
    lib().oracle_set_threads(ctypes.c_int(int(n)))
    try:                                    # numpy's OpenBLAS pool, when threadpoolctl is around
        from threadpoolctl import threadpool_limits
        threadpool_limits(limits=int(n))
    except Exception:  # noqa: BLE001
        pass
