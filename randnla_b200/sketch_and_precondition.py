"""Sketch step of the sketch-and-precondition solvers -- mirror of the first lines of
`blendenpik_overdetermined` (reference src/sketch_and_precondition.rs:26-52), `lsrn_overdetermined`
(:82-107) and `sketch_saddle_point_precondition` (:150-176): argument validation, the sketch dimension
rule and A_sk = S A (and b_sk = S b).  The preconditioned CGLS iteration that follows in the reference is a
"next" row of SURVEY.md §8(f) and is not part of this path."""
import numpy as np

from . import _lib
from ._lib import check
from . import runtime
import ctypes as C

from .errors import InvalidParameters, NotOverdetermined

SKETCH_DENSE, SKETCH_SASO, SKETCH_SASO_BLOCK = 0, 1, 2


def _validate(a, epsilon, l, sampling_factor):
    m, n = a.shape
    if m < n:  # :29-33
        raise NotOverdetermined(f"Need more columns than rows, found {m} rows and {n} columns")
    if sampling_factor < 1.0:  # :34-38
        raise InvalidParameters(f"Sampling factor must be greater than 1, current input is {sampling_factor}")
    if epsilon <= 0.0:  # :39-43
        raise InvalidParameters(f"Epsilon must be positive, current input is {epsilon}")
    if l == 0:  # :44-48
        raise InvalidParameters(f"Number of iterations must be positive, current input is {l}")


def sketch_dim(m, n, sampling_factor, saddle=False):
    """d of reference :49 / :105 (saddle=False) or :172 (saddle=True)."""
    return int(_lib.load().rnla_sketch_dim(int(m), int(n), float(sampling_factor), 1 if saddle else 0))


def sketch_apply(a, b=None, d=None, kind=SKETCH_DENSE, dist=runtime.GAUSSIAN, seed=0, zeta=8, width=0):
    """A_sk = S a (d x n) and b_sk = S b for a dense i.i.d. or sparse-sign S (d x m).
    For SKETCH_SASO_BLOCK `width` is the block width w (0 = default), carried in the C ABI's `dist` slot."""
    lib = _lib.load()
    if kind == SKETCH_SASO_BLOCK:
        dist = width
    a = runtime.as_f(a)
    m, n = a.shape
    a_sk = np.empty((d, n), dtype=np.float64, order="F")
    if b is not None:
        b = runtime.as_f(b)
        b_sk = np.empty((d, b.shape[1]), dtype=np.float64, order="F")
        check(lib.rnla_sketch_apply(kind, dist, seed, d, zeta, runtime.ptr(a), m, n, runtime.ptr(b), b.shape[1],
                                    runtime.ptr(a_sk), runtime.ptr(b_sk)))
        return a_sk, b_sk
    check(lib.rnla_sketch_apply(kind, dist, seed, d, zeta, runtime.ptr(a), m, n, None, 0, runtime.ptr(a_sk), None))
    return a_sk


def blendenpik_sketch(a, b, epsilon, l, sampling_factor, kind=SKETCH_DENSE, zeta=8):
    """Validation + sketch step of `blendenpik_overdetermined` (reference :26-52) -> (a_sk, b_sk)."""
    a = runtime.as_f(a)
    _validate(a, epsilon, l, sampling_factor)
    d = sketch_dim(a.shape[0], a.shape[1], sampling_factor)
    return sketch_apply(a, b, d, kind=kind, zeta=zeta, seed=runtime.get_options().seed)


def lsrn_sketch(a, b, epsilon, l, sampling_factor, kind=SKETCH_DENSE, zeta=8):
    """Validation + sketch step of `lsrn_overdetermined` (reference :82-107) -> a_sk."""
    a = runtime.as_f(a)
    _validate(a, epsilon, l, sampling_factor)
    d = sketch_dim(a.shape[0], a.shape[1], sampling_factor)
    return sketch_apply(a, None, d, kind=kind, zeta=zeta, seed=runtime.get_options().seed)


def saddle_point_sketch(a, b, c, mu, epsilon, l, sampling_factor, kind=SKETCH_DENSE, zeta=8):
    """Validation + sketch step of `sketch_saddle_point_precondition` (reference :150-176) -> a_sk."""
    a = runtime.as_f(a)
    _validate(a, epsilon, l, sampling_factor)
    d = sketch_dim(a.shape[0], a.shape[1], sampling_factor, saddle=True)
    return sketch_apply(a, None, d, kind=kind, zeta=zeta, seed=runtime.get_options().seed)


def blendenpik_overdetermined(a, b, epsilon, l, sampling_factor, kind=SKETCH_DENSE, zeta=8, width=0, info=None):
    """`blendenpik_overdetermined` end to end (reference :26-59): sketch, QR of the sketch, CGLS on A R^-1 in operator
    form, x = R^-1 z.  Same validation, same errors.  kind / zeta / width choose the sketch operator (the reference's own
    is the dense Gaussian, the default here).  `info`, if a dict, receives the CGLS iteration count and convergence flag."""
    lib = _lib.load()
    a = runtime.as_f(a)
    b = runtime.as_f(b)
    m, n = a.shape
    x = np.empty((n, 1), dtype=np.float64, order="F")
    it = C.c_int64(0); conv = C.c_int32(0)
    dist = width if kind == SKETCH_SASO_BLOCK else runtime.GAUSSIAN
    check(lib.rnla_blendenpik_overdetermined(runtime.ptr(a), m, n, runtime.ptr(b), float(epsilon), int(l), float(sampling_factor),
                                             kind, dist, zeta, runtime.ptr(x), C.byref(it), C.byref(conv)))
    if info is not None:
        info["iterations"] = int(it.value); info["converged"] = bool(conv.value)
    return x


def lsrn_overdetermined(a, b, epsilon, l, sampling_factor, kind=SKETCH_DENSE, zeta=8, width=0, info=None):
    """`lsrn_overdetermined` end to end (reference :82-119): sketch, SVD of the sketch, N = V Sigma^-1, CGLS on A N in
    operator form from y = 0, x = N y.  n <= 1024 (size of the on-device SVD core)."""
    lib = _lib.load()
    a = runtime.as_f(a)
    b = runtime.as_f(b)
    m, n = a.shape
    x = np.empty((n, 1), dtype=np.float64, order="F")
    it = C.c_int64(0); conv = C.c_int32(0)
    dist = width if kind == SKETCH_SASO_BLOCK else runtime.GAUSSIAN
    check(lib.rnla_lsrn_overdetermined(runtime.ptr(a), m, n, runtime.ptr(b), float(epsilon), int(l), float(sampling_factor),
                                       kind, dist, zeta, runtime.ptr(x), C.byref(it), C.byref(conv)))
    if info is not None:
        info["iterations"] = int(it.value); info["converged"] = bool(conv.value)
    return x


def sketch_saddle_point_precondition(a, b, c, mu, epsilon, l, sampling_factor, info=None):
    """`sketch_saddle_point_precondition(a, b, c, mu, epsilon, l, sampling_factor) -> (x, y)` end to end (reference :150-216):
    min ||a x - b||^2 + mu ||x||^2 + 2 <c, x> through the SVD of a dense Gaussian sketch and CGLS in operator form;
    y = b - a x.  `c` may be None / empty (`c.is_empty()`, :195).  n <= 1024 (size of the on-device SVD core)."""
    lib = _lib.load()
    a = runtime.as_f(a)
    b = runtime.as_f(b)
    m, n = a.shape
    cc = None if c is None or np.size(c) == 0 else runtime.as_f(c)
    x = np.empty((n, 1), dtype=np.float64, order="F")
    y = np.empty((m, 1), dtype=np.float64, order="F")
    it = C.c_int64(0); conv = C.c_int32(0)
    check(lib.rnla_sketch_saddle_point_precondition(runtime.ptr(a), m, n, runtime.ptr(b), runtime.ptr(cc) if cc is not None else None,
                                                    float(mu), float(epsilon), int(l), float(sampling_factor), runtime.ptr(x),
                                                    runtime.ptr(y), C.byref(it), C.byref(conv)))
    if info is not None:
        info["iterations"] = int(it.value); info["converged"] = bool(conv.value)
    return x, y
