"""ctypes front-end for oracle/tip_oracle.c (TEST INFRASTRUCTURE ONLY — see np_oracle.py).

`build()` compiles oracle/libtiporacle.so with gcc; `lib()` loads it.  The C port is used
(a) to validate the claim that the CUDA re-rank's summation order is NumPy's, (b) as the
strictly sequential restatement of scipy-1.4.1's KDE loop nest, and (c) as the
multi-threaded "port" CPU baseline in bench.py.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libtiporacle.so")
_SRC = os.path.join(_HERE, "tip_oracle.c")
_lib = None


def build(force: bool = False) -> str:
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(_SRC):
        cmd = ["gcc", "-O3", "-mavx2", "-fopenmp", "-ffp-contract=off", "-fno-fast-math", "-shared", "-fPIC",
               _SRC, "-o", _SO, "-lm"]
        subprocess.run(cmd, check=True)
    return _SO


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_SO)
    return _lib


def _p(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


def max_threads() -> int:
    return int(lib().oracle_max_threads())


def pairwise_sumsq(x: np.ndarray, y: np.ndarray = None):
    x = np.ascontiguousarray(x)
    l = lib()
    if x.dtype == np.float32:
        f = l.oracle_pairwise_sumsq_f32
        f.restype = C.c_float
    else:
        f = l.oracle_pairwise_sumsq_f64
        f.restype = C.c_double
    yy = None if y is None else np.ascontiguousarray(y, dtype=x.dtype)
    return x.dtype.type(f(_p(x), None if yy is None else _p(yy), C.c_int64(x.shape[0])))


def dsa(train, train_pred, test, test_pred, threads: int = 0):
    """Brute-force DSA stage-1/stage-2 in the input dtype (float32 or float64)."""
    train = np.ascontiguousarray(train)
    test = np.ascontiguousarray(test, dtype=train.dtype)
    tp = np.ascontiguousarray(train_pred, dtype=np.int64)
    sp = np.ascontiguousarray(test_pred, dtype=np.int64)
    n, d = train.shape
    m = test.shape[0]
    da = np.empty(m, dtype=train.dtype)
    db = np.empty(m, dtype=train.dtype)
    ia = np.empty(m, dtype=np.int64)
    nthreads = threads or max_threads()
    by_rows = m * 4 < nthreads and n >= 4096          # few inputs, many train rows: parallel over train rows
    if by_rows:
        f = lib().oracle_dsa_rows_f32 if train.dtype == np.float32 else lib().oracle_dsa_rows_f64
    else:
        f = lib().oracle_dsa_f32 if train.dtype == np.float32 else lib().oracle_dsa_f64
    f(_p(train), _p(tp), C.c_int64(n), C.c_int64(d), _p(test), _p(sp), C.c_int64(m),
      _p(da), _p(db), _p(ia), C.c_int(threads or max_threads()))
    with np.errstate(divide="ignore", invalid="ignore"):
        score = (da / db).astype(np.float64)
    return {"dsa": score, "dist_a": da, "dist_b": db, "idx_a": ia}


def kde_eval(points_w, xi_w, weight: float, norm: float, threads: int = 0) -> np.ndarray:
    pw = np.ascontiguousarray(points_w, dtype=np.float64)
    xw = np.ascontiguousarray(xi_w, dtype=np.float64)
    out = np.empty(xw.shape[0], dtype=np.float64)
    lib().oracle_kde_eval(_p(pw), C.c_int64(pw.shape[0]), _p(xw), C.c_int64(xw.shape[0]),
                          C.c_int64(pw.shape[1]), C.c_double(weight), C.c_double(norm), _p(out),
                          C.c_int(threads or max_threads()))
    return out


def deepgini(p, threads: int = 0) -> Tuple[np.ndarray, np.ndarray]:
    p = np.ascontiguousarray(p)
    n, c = p.shape
    pred = np.empty(n, dtype=np.int64)
    gini = np.empty(n, dtype=p.dtype)
    f = lib().oracle_deepgini_f32 if p.dtype == np.float32 else lib().oracle_deepgini_f64
    f(_p(p), C.c_int64(n), C.c_int64(c), _p(pred), _p(gini), C.c_int(threads or max_threads()))
    return pred, gini


def kmnc(act, thresh, threads: int = 0) -> Tuple[np.ndarray, np.ndarray]:
    """thresh: (sections+1) x d array exactly as NumPy built it (same dtype as act)."""
    act = np.ascontiguousarray(act)
    thresh = np.ascontiguousarray(thresh, dtype=act.dtype)
    n, d = act.shape
    k = thresh.shape[0] - 1
    bucket = np.empty((n, d), dtype=np.int32)
    score = np.empty(n, dtype=np.int64)
    f = lib().oracle_kmnc_f32 if act.dtype == np.float32 else lib().oracle_kmnc_f64
    f(_p(act), C.c_int64(n), C.c_int64(d), _p(thresh), C.c_int64(k), _p(bucket), _p(score),
      C.c_int(threads or max_threads()))
    return bucket, score
