"""Interpolative and CUR decompositions -- mirror of reference src/id.rs: `cur` (:34-71), `two_sided_id_randomised`
(:94-101), `two_sided_id` (:118-129), `cur_randomised` (:154-193), `osid_randomised` (:217-249), `osid_qrcp` (:272-318).
Index vectors come back as Python lists of ints (the reference's Vec<usize>)."""
import numpy as np

from . import _lib
from ._lib import check
from . import runtime
from .sketch import MatrixAttribute


def _idx(k):
    return np.zeros(max(int(k), 1), dtype=np.int64)


def osid_qrcp(y, k, attr):
    """`osid_qrcp(y, k, attr) -> (x, j)`: Column: y ~ y[:, j] x (x k x w); Row: y ~ x y[j, :] (x l x k)."""
    lib = _lib.load()
    y = runtime.as_f(y)
    l, w = y.shape
    k = int(k)
    x = np.empty((k, w) if attr == MatrixAttribute.Column else (l, k), dtype=np.float64, order="F")
    j = _idx(k)
    check(lib.rnla_osid_qrcp(runtime.ptr(y), l, w, k, int(attr), runtime.ptr(x), runtime.ptr(j)))
    return x, [int(v) for v in j[:k]]


def osid_randomised(a, k, attr):
    """`osid_randomised(a, k, attr) -> (x, j)`: the ID of a sketch (Column: S a with a k x m Gaussian S; Row:
    a tsog1(a, k, 2, 1)^T, which needs a.ncols() == k exactly as in the reference, :230-233)."""
    lib = _lib.load()
    a = runtime.as_f(a)
    m, n = a.shape
    k = int(k)
    x = np.empty((k, n) if attr == MatrixAttribute.Column else (m, k), dtype=np.float64, order="F")
    j = _idx(k)
    check(lib.rnla_osid_randomised(runtime.ptr(a), m, n, k, int(attr), runtime.ptr(x), runtime.ptr(j)))
    return x, [int(v) for v in j[:k]]


def _two_sided(a, k, randomised):
    lib = _lib.load()
    a = runtime.as_f(a)
    m, n = a.shape
    k = int(k)
    z = np.empty((m, k), dtype=np.float64, order="F")
    x = np.empty((k, n), dtype=np.float64, order="F")
    i, j = _idx(k), _idx(k)
    check(lib.rnla_two_sided_id(runtime.ptr(a), m, n, k, randomised, runtime.ptr(z), runtime.ptr(i), runtime.ptr(j), runtime.ptr(x)))
    return z, [int(v) for v in i[:k]], [int(v) for v in j[:k]], x


def two_sided_id(a, k):
    """`two_sided_id(a, k) -> (z, i, j, x)`, a ~ z a[i, j] x (reference :118-129)."""
    return _two_sided(a, k, 0)


def two_sided_id_randomised(a, k):
    """`two_sided_id_randomised(a, k) -> (z, i, j, x)` (reference :94-101)."""
    return _two_sided(a, k, 1)


def _cur(a, k, randomised):
    lib = _lib.load()
    a = runtime.as_f(a)
    m, n = a.shape
    k = int(k)
    u = np.empty((k, k), dtype=np.float64, order="F")
    i, j = _idx(k), _idx(k)
    check(lib.rnla_cur(runtime.ptr(a), m, n, k, randomised, runtime.ptr(j), runtime.ptr(u), runtime.ptr(i)))
    return [int(v) for v in j[:k]], u, [int(v) for v in i[:k]]


def cur(a, k):
    """`cur(a, k) -> (j, u, i)`, a ~ a[:, j] u a[i, :] (reference :34-71)."""
    return _cur(a, k, 0)


def cur_randomised(a, k):
    """`cur_randomised(a, k) -> (j, u, i)` (reference :154-193)."""
    return _cur(a, k, 1)
