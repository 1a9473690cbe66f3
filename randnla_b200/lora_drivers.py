"""Drivers -- mirror of reference src/lora_drivers.rs (rand_svd :30-69, rand_evd1 :87-151, rand_evd2 :167-224).

Host `numpy` matrices go through the host-buffer C entry points (copies inside the call); the `*_dev`
variants take column-major torch CUDA tensors and leave everything on the device."""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import check
from . import runtime


def rand_svd(A, k, epsilon, s):
    """`rand_svd(A, k, epsilon, s) -> (U, S, Vt)` (reference src/lora_drivers.rs:30-69).

    U is m x r, S a dense r x r diagonal matrix, and the third value is V^T (r x n), r = min(k, Q.ncols()).
    Raises InvalidParameters for k == 0, epsilon <= 0, s == 0 (:31-45)."""
    lib = _lib.load()
    A = runtime.as_f(A)
    m, n = A.shape
    k, s = int(k), int(s)
    print("Running RSVD")  # reference :47
    kk = max(k, 1)
    r_cap = max(min(kk, m, n), 1)
    U = np.empty((m, r_cap), dtype=np.float64, order="F")
    S = np.empty((r_cap, r_cap), dtype=np.float64, order="F")
    Vt = np.empty((r_cap, n), dtype=np.float64, order="F")
    r = C.c_int64(0)
    check(lib.rnla_rand_svd(runtime.ptr(A), m, n, k, float(epsilon), s, runtime.ptr(U), runtime.ptr(S), runtime.ptr(Vt), C.byref(r)))
    assert r.value == r_cap
    return U, S, Vt


def rand_evd1(A, k, epsilon, s):
    """`rand_evd1(A, k, epsilon, s) -> (V, lambda)` (reference src/lora_drivers.rs:87-151); lambda sorted by
    decreasing absolute value; NotHermitian when A != A^T exactly (:106)."""
    lib = _lib.load()
    A = runtime.as_f(A)
    if A.shape[0] != A.shape[1]:
        from .errors import NotHermitian
        raise NotHermitian("Input matrix is not Hermitian")
    n = A.shape[0]
    k, s = int(k), int(s)
    print("Running REVD1")  # reference :112
    r_cap = max(min(max(k, 1), n), 1)
    V = np.empty((n, r_cap), dtype=np.float64, order="F")
    lam = np.empty(r_cap, dtype=np.float64)
    r = C.c_int64(0)
    check(lib.rnla_rand_evd1(runtime.ptr(A), n, k, float(epsilon), s, runtime.ptr(V), runtime.ptr(lam), C.byref(r)))
    return V[:, :r.value], lam[:r.value].tolist()


def rand_evd2(A, k, s):
    """`rand_evd2(A, k, s) -> (V, lambda)` (reference src/lora_drivers.rs:167-224), Nystrom with shift."""
    lib = _lib.load()
    A = runtime.as_f(A)
    if A.shape[0] != A.shape[1]:
        from .errors import NotSquare
        raise NotSquare("rand_evd2 needs a square matrix")
    n = A.shape[0]
    k, s = int(k), int(s)
    print("Running REVD2")  # reference :175
    r_cap = max(min(max(k, 1), n), 1)
    V = np.empty((n, r_cap), dtype=np.float64, order="F")
    lam = np.empty(r_cap, dtype=np.float64)
    r = C.c_int64(0)
    check(lib.rnla_rand_evd2(runtime.ptr(A), n, k, s, runtime.ptr(V), runtime.ptr(lam), C.byref(r)))
    return np.asfortranarray(V[:, :r.value]), lam[:r.value].tolist()


# ---- device-resident variants -------------------------------------------------------------------
def rand_svd_dev(dA, k, s, opts=None, n=None):
    """Device-resident rand_svd.  `dA` is the LOCAL row shard (column-major torch CUDA tensor, m_local x n).
    Returns (dU m_local x r, dSigma r, dVt r x n) as torch tensors.  Like the C entry point it returns once the work is
    enqueued on the library stream: `runtime.synchronize()` (or `runtime.use_torch_stream()` beforehand) before torch reads them."""
    import torch
    lib = _lib.load()
    pA, lda = runtime.dev_ptr_ld(dA)
    m_local, n = dA.shape
    kk = int(k)
    dU = runtime.empty_colmajor(m_local, kk)
    dS = torch.empty(kk, dtype=torch.float64, device=dA.device)
    dVt = runtime.empty_colmajor(kk, n)
    pU, ldu = runtime.dev_ptr_ld(dU)
    pVt, ldvt = runtime.dev_ptr_ld(dVt)
    r = C.c_int64(0)
    check(lib.rnla_rand_svd_dev(pA, lda, m_local, n, kk, int(s), C.byref(opts) if opts is not None else None,
                                pU, ldu, C.c_void_p(dS.data_ptr()), pVt, ldvt, C.byref(r)))
    rv = r.value
    return dU[:, :rv], dS[:rv], dVt[:rv, :]


def rand_evd1_dev(dA, k, s, opts=None):
    import torch
    lib = _lib.load()
    pA, lda = runtime.dev_ptr_ld(dA)
    n = dA.shape[1]
    kk = int(k)
    dV = runtime.empty_colmajor(dA.shape[0], kk)
    dL = torch.empty(kk, dtype=torch.float64, device=dA.device)
    pV, ldv = runtime.dev_ptr_ld(dV)
    r = C.c_int64(0)
    check(lib.rnla_rand_evd1_dev(pA, lda, n, kk, int(s), C.byref(opts) if opts is not None else None,
                                 pV, ldv, C.c_void_p(dL.data_ptr()), C.byref(r)))
    return dV[:, :r.value], dL[:r.value]


def rand_evd2_dev(dA, k, s, opts=None):
    import torch
    lib = _lib.load()
    pA, lda = runtime.dev_ptr_ld(dA)
    n = dA.shape[1]
    kk = int(k)
    dV = runtime.empty_colmajor(dA.shape[0], kk)
    dL = torch.empty(kk, dtype=torch.float64, device=dA.device)
    pV, ldv = runtime.dev_ptr_ld(dV)
    r = C.c_int64(0)
    check(lib.rnla_rand_evd2_dev(pA, lda, n, kk, int(s), C.byref(opts) if opts is not None else None,
                                 pV, ldv, C.c_void_p(dL.data_ptr()), C.byref(r)))
    return dV[:, :r.value], dL[:r.value]
