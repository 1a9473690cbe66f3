"""Range-finder helpers -- mirror of reference src/lora_helpers.rs (QB1 :17, RF1 :37, tsog1 :58, Orth :131,
Stabilizer :144).  Names keep the reference's capitalisation."""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import check
from . import runtime


def Orth(X, return_r=False):
    """`Orth(X)` = thin Q of a QR factorisation with R_ii >= 0 (reference src/lora_helpers.rs:131-133)."""
    lib = _lib.load()
    X = runtime.as_f(X)
    rows, cols = X.shape
    p = min(rows, cols)
    Q = np.empty((rows, p), dtype=np.float64, order="F")
    R = np.empty((p, cols), dtype=np.float64, order="F") if return_r else None
    qc = C.c_int64(0)
    check(lib.rnla_orth(runtime.ptr(X), rows, cols, runtime.ptr(Q), runtime.ptr(R) if return_r else None, C.byref(qc)))
    return (Q, R) if return_r else Q


def Stabilizer(X):
    """`Stabilizer(X)` = unit-lower-trapezoidal L of the full-pivot LU, permutations dropped
    (reference src/lora_helpers.rs:144-146)."""
    lib = _lib.load()
    X = runtime.as_f(X)
    rows, cols = X.shape
    mn = min(rows, cols)
    L = np.empty((rows, mn), dtype=np.float64, order="F")
    lc = C.c_int64(0)
    check(lib.rnla_stabilizer(runtime.ptr(X), rows, cols, runtime.ptr(L), C.byref(lc)))
    return L


def tsog1(A, k, num_passes, passes_per_stab):
    """`tsog1(A, k, num_passes, passes_per_stab)` -> S (n x k) (reference src/lora_helpers.rs:58-105).
    Runs the monograph algorithm or the reference's literal statements according to the `mode` option."""
    lib = _lib.load()
    A = runtime.as_f(A)
    m, n = A.shape
    S = np.empty((n, max(int(k), 0)), dtype=np.float64, order="F")
    check(lib.rnla_tsog1(runtime.ptr(A), m, n, int(k), int(num_passes), int(passes_per_stab), runtime.ptr(S)))
    return S


def RF1(A, k):
    """`RF1(A, k)` -> Q with min(k, m, n) orthonormal columns (reference src/lora_helpers.rs:37-44)."""
    lib = _lib.load()
    A = runtime.as_f(A)
    m, n = A.shape
    l = max(min(int(k), m, n), 0)
    Q = np.empty((m, l), dtype=np.float64, order="F")
    qc = C.c_int64(0)
    check(lib.rnla_rf1(runtime.ptr(A), m, n, int(k), runtime.ptr(Q), C.byref(qc)))
    return Q


def QB1(A, k, epsilon):
    """`QB1(A, k, epsilon)` -> (Q, B = Q^T A); epsilon is ignored as in the reference (src/lora_helpers.rs:17-23)."""
    lib = _lib.load()
    A = runtime.as_f(A)
    m, n = A.shape
    l = max(min(int(k), m, n), 0)
    Q = np.empty((m, l), dtype=np.float64, order="F")
    B = np.empty((l, n), dtype=np.float64, order="F")
    qc = C.c_int64(0)
    check(lib.rnla_qb1(runtime.ptr(A), m, n, int(k), float(epsilon), runtime.ptr(Q), runtime.ptr(B), C.byref(qc)))
    return Q, B
