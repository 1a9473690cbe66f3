"""Sketch-and-solve least squares -- mirror of reference src/sketch_and_solve.rs: `sketched_least_squares_qr` (:24-33) and
`sketched_least_squares_svd` (:54-66).  The sketch has rows/4 rows, as in the reference."""
import numpy as np

from . import _lib
from ._lib import check
from . import runtime
from .sketch_and_precondition import SKETCH_DENSE, SKETCH_SASO_BLOCK


def _solve(fn, a, b, kind, zeta, width):
    a = runtime.as_f(a)
    b = runtime.as_f(b)
    m, n = a.shape
    x = np.empty((n, 1), dtype=np.float64, order="F")
    dist = width if kind == SKETCH_SASO_BLOCK else runtime.GAUSSIAN
    check(fn(runtime.ptr(a), m, n, runtime.ptr(b), kind, dist, zeta, runtime.ptr(x)))
    return x


def sketched_least_squares_qr(a, b, kind=SKETCH_DENSE, zeta=8, width=0):
    """`sketched_least_squares_qr(a, b) -> x` (reference :24-33): QR of the sketch, back-substitution with
    `solve_upper_triangular_system`'s zero-pivot rule (src/solvers.rs:22-41)."""
    return _solve(_lib.load().rnla_sketched_least_squares_qr, a, b, kind, zeta, width)


def sketched_least_squares_svd(a, b, kind=SKETCH_DENSE, zeta=8, width=0):
    """`sketched_least_squares_svd(a, b) -> x` (reference :54-66): SVD of the sketch, x = V Sigma^-1 U^T b_sk."""
    return _solve(_lib.load().rnla_sketched_least_squares_svd, a, b, kind, zeta, width)
