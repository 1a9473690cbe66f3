"""Column-pivoted QR -- mirror of reference src/pivot_decompositions.rs: `qrcp` (:105-180) and `economic_qrcp` (:196-269).
Householder reflections with pivoting on exactly recomputed trailing column norms, on the GPU (csrc/pivot.cu)."""
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
