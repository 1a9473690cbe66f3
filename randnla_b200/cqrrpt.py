"""CholeskyQR with randomised pivoting for tall matrices -- mirror of reference src/cqrrpt.rs (`sap_chol_qrcp` :27-58)."""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import check
from . import runtime
from .sketch_and_precondition import SKETCH_DENSE, SKETCH_SASO_BLOCK


def sap_chol_qrcp(a, d, kind=SKETCH_DENSE, zeta=8, width=0):
    """`sap_chol_qrcp(a, d) -> (q, r, j)` with a[:, j[:k]] ~ q r[:, :k]..., i.e. a[:, j] = q r (reference :27-58):
    q m x k orthonormal, r k x n upper trapezoidal, j the column permutation, k the numerical rank of the sketch
    (|R_ii| > 1e-10, :37-43).  The reference panics with "d must satisfy n <= d << m" (:29): raised as `InvalidParameters`.
    kind / zeta / width choose the sketching operator (the reference's own dense Gaussian is the default)."""
    lib = _lib.load()
    a = runtime.as_f(a)
    m, n = a.shape
    q = np.empty((m, n), dtype=np.float64, order="F")
    r = np.empty(n * n, dtype=np.float64)
    j = np.zeros(max(n, 1), dtype=np.int64)
    k = C.c_int64(0)
    dist = width if kind == SKETCH_SASO_BLOCK else runtime.GAUSSIAN
    check(lib.rnla_sap_chol_qrcp(runtime.ptr(a), m, n, int(d), kind, dist, zeta, runtime.ptr(q), runtime.ptr(r), runtime.ptr(j),
                                 C.byref(k)))
    k = int(k.value)
    return np.asfortranarray(q[:, :k]), r[:k * n].reshape((k, n), order="F"), [int(v) for v in j[:n]]
