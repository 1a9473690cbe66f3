"""Pivoted decompositions -- mirror of reference src/pivot_decompositions.rs: `lupp` (:21-86), `qrcp` (:105-180) and
`economic_qrcp` (:196-269).  Householder reflections with pivoting on exactly recomputed trailing column norms and row-pivoted
Gaussian elimination in the reference's operation order, on the GPU (csrc/pivot.cu)."""
import numpy as np

from . import _lib
from ._lib import check
from . import runtime


def _qrcp(a, steps, qcols):
    lib = _lib.load()
    a = runtime.as_f(a)
    m, n = a.shape
    q = np.empty((m, qcols), dtype=np.float64, order="F")
    r = np.empty((m, n), dtype=np.float64, order="F")
    p = np.zeros(max(n, 1), dtype=np.int64)
    check(lib.rnla_qrcp(runtime.ptr(a), m, n, int(steps), int(qcols), runtime.ptr(q), runtime.ptr(r), runtime.ptr(p)))
    return q, r, [int(v) for v in p[:n]]


def qrcp(a):
    """`qrcp(a) -> (q, r, p)`: q m x m, r m x n, p the column permutation (reference :105-180)."""
    a = runtime.as_f(a)
    m, n = a.shape
    return _qrcp(a, min(m, n), m)


def economic_qrcp(a, k):
    """`economic_qrcp(a, k) -> (q_eco m x k, r_eco k x n, p)` (reference :196-269).  The reference asserts
    "k must be <= min(m,n)" and "k must be positive" (:200-201): raised as `InvalidParameters`."""
    q, r, p = _qrcp(a, k, k)
    return q, np.asfortranarray(r[:k, :]), p


def lupp(matrix):
    """`lupp(matrix) -> Result<(l, u, p), Box<dyn Error>>` (reference :21-86): `NotSquare` / `SingularMatrix` as the reference;
    l, u, p carry the same bits as the reference's arithmetic (first-maximum pivot, separately rounded multiply and subtract)."""
    a = runtime.as_f(matrix)
    rows, cols = a.shape
    n = max(rows, 1)
    l = np.empty((n, n), dtype=np.float64, order="F")
    u = np.empty((n, n), dtype=np.float64, order="F")
    p = np.zeros(n, dtype=np.int64)
    check(_lib.load().rnla_lupp(runtime.ptr(a), rows, cols, runtime.ptr(l), runtime.ptr(u), runtime.ptr(p)))
    return l, u, [int(v) for v in p]
