"""Conjugate-gradient solvers -- mirror of reference src/cg.rs: `cgls` (:18-61), `conjugate_grad` (:77-112) and
`verify_solution` (:115-117).  The iterations run in librnla.so on the GPU (A streamed once or twice per iteration by the
matrix-vector kernels of the sketch-and-precondition drivers); this module only marshals buffers and prints what the
reference prints."""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import check
from . import runtime


def cgls(a, b, tolerance, num_iterations, x=None, info=None):
    """`cgls(a, b, tolerance, num_iterations, x) -> x` (reference :18-61).  `info`, if a dict, receives the iteration count and
    the convergence flag the reference only prints."""
    a = runtime.as_f(a)
    m, n = a.shape
    b = runtime.as_f(np.asarray(b, dtype=np.float64).reshape(-1, 1))
    if b.shape[0] != m:
        raise ValueError(f"cgls: a has {m} rows, b has {b.shape[0]}")                  # the reference panics inside nalgebra
    x0 = None if x is None else runtime.as_f(np.asarray(x, dtype=np.float64).reshape(-1, 1))
    if x0 is not None and x0.shape[0] != n:
        raise ValueError(f"cgls: a has {n} columns, x has {x0.shape[0]} entries")
    out = np.empty((n, 1), dtype=np.float64, order="F")
    it = C.c_int64(0)
    conv = C.c_int32(0)
    check(_lib.load().rnla_cgls(runtime.ptr(a), m, n, runtime.ptr(b), float(tolerance), int(num_iterations),
                                runtime.ptr(x0) if x0 is not None else None, runtime.ptr(out), C.byref(it), C.byref(conv)))
    if conv.value:
        print(f"CGLS converged after {it.value} iterations")                            # :46
    else:
        print(f"CGLS failed to converged after {int(num_iterations)} iterations")       # :58
    if info is not None:
        info["iterations"] = int(it.value)
        info["converged"] = bool(conv.value)
    return out


def conjugate_grad(a, b, x=None, info=None):
    """`conjugate_grad(a, b, x) -> Result<x, RandNLAError>` (reference :77-112): `NotPositiveSemiDefinite` from the eigenvalue
    check (run for n <= 512, see include/rnla.h); initial guess of ones when `x` is None (:88)."""
    a = runtime.as_f(a)
    n = a.shape[0]
    b = runtime.as_f(np.asarray(b, dtype=np.float64).reshape(-1, 1))
    if a.shape[1] != n or b.shape[0] != n:
        raise ValueError("conjugate_grad: a must be n x n and b of length n")           # the reference panics inside nalgebra
    x0 = None if x is None else runtime.as_f(np.asarray(x, dtype=np.float64).reshape(-1, 1))
    out = np.empty(n, dtype=np.float64)
    it = C.c_int64(0)
    conv = C.c_int32(0)
    check(_lib.load().rnla_conjugate_grad(runtime.ptr(a), n, runtime.ptr(b), runtime.ptr(x0) if x0 is not None else None,
                                          runtime.ptr(out), C.byref(it), C.byref(conv)))
    if conv.value:
        print(f"Converged after {it.value} iterations")                                 # :101
    if info is not None:
        info["iterations"] = int(it.value)
        info["converged"] = bool(conv.value)
    return out


def verify_solution(a, b, x):
    """`verify_solution(a, b, x) -> f64` = ||a x - b|| (reference :115-117)."""
    a = runtime.as_f(a)
    m, n = a.shape
    b = runtime.as_f(np.asarray(b, dtype=np.float64).reshape(-1, 1))
    x = runtime.as_f(np.asarray(x, dtype=np.float64).reshape(-1, 1))
    if b.shape[0] != m or x.shape[0] != n:
        raise ValueError("verify_solution: shapes do not conform")
    out = C.c_double(0.0)
    check(_lib.load().rnla_verify_solution(runtime.ptr(a), m, n, runtime.ptr(b), runtime.ptr(x), C.byref(out)))
    return out.value
