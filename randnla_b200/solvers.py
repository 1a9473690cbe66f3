"""Iterative least-squares solver -- mirror of reference src/solvers.rs: `lsqr` (:115-278, the reference's translation of
scipy 1.14.1 sparse.linalg.lsqr).  The Golub-Kahan bidiagonalisation (two streamed passes over A per iteration) runs in
librnla.so on the GPU; this module only marshals buffers."""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import check, LsqrResult
from . import runtime


def lsqr(a, b, damp=0.0, atol=1e-6, btol=1e-6, conlim=1e8, iter_lim=None, calc_var=False, x0=None):
    """`lsqr(a, b, damp, atol, btol, conlim, iter_lim, calc_var, x0)` (reference :115-278) ->
    (x, istop, itn, r1norm, r2norm, anorm, acond, arnorms, xnorm, var), the reference's return order: `arnorms` is the
    history of the ||A^T r|| estimates (one per iteration), where scipy returns only the last."""
    a = runtime.as_f(a)
    m, n = a.shape
    b = runtime.as_f(np.asarray(b, dtype=np.float64).reshape(-1, 1))
    if b.shape[0] != m:
        raise ValueError(f"lsqr: a has {m} rows, b has {b.shape[0]}")        # the reference panics inside nalgebra
    lim = -1 if iter_lim is None else int(iter_lim)
    cap = 2 * n if iter_lim is None else max(int(iter_lim), 1)
    x = np.empty((n, 1), dtype=np.float64, order="F")
    var = np.zeros(n, dtype=np.float64)
    hist = np.zeros(max(cap, 1), dtype=np.float64)
    x0f = None
    if x0 is not None:
        x0f = runtime.as_f(np.asarray(x0, dtype=np.float64).reshape(-1, 1))
        if x0f.shape[0] != n:
            raise ValueError(f"lsqr: a has {n} columns, x0 has {x0f.shape[0]} entries")
    res = LsqrResult()
    check(_lib.load().rnla_lsqr(runtime.ptr(a), m, n, runtime.ptr(b), float(damp), float(atol), float(btol), float(conlim), lim,
                                1 if calc_var else 0, runtime.ptr(x0f) if x0f is not None else None, runtime.ptr(x), C.byref(res),
                                runtime.ptr(hist), hist.size, runtime.ptr(var)))
    nh = min(int(res.n_arnorms), hist.size)
    return (x, int(res.istop), int(res.itn), res.r1norm, res.r2norm, res.anorm, res.acond, hist[:nh].copy(), res.xnorm, var)
