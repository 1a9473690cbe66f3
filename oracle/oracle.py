"""ctypes wrapper of oracle/liboracle.so -- the CPU restatement of the reference path.

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs.  Nothing under randnla_b200/ imports this module."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "liboracle.so")
_lib = None

i32, i64, u32, u64, f64, P = C.c_int, C.c_int64, C.c_uint32, C.c_uint64, C.c_double, C.c_void_p


class Opts(C.Structure):
    _fields_ = [("mode", C.c_int), ("dist", C.c_int), ("seed", u64), ("num_passes", C.c_int),
                ("passes_per_stab", C.c_int), ("omega_n", P), ("omega_m", P), ("skip_psd_check", C.c_int)]


def build():
    subprocess.run(["make", "-s", "-C", _HERE], check=True)


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        build()
    lib = C.CDLL(LIB_PATH)
    lib.orc_threefry_rng_u64.restype = u64
    lib.orc_threefry_rng_u64.argtypes = [u64, u64]
    lib.orc_gauss_from_u32.restype = C.c_float
    lib.orc_gauss_from_u32.argtypes = [u32]
    lib.orc_sketch_dim.restype = i64
    lib.orc_sketch_dim.argtypes = [i64, i64, f64, C.c_int]
    lib.orc_get_threads.restype = C.c_int
    _lib = lib
    return lib


def F(a):
    a = np.asarray(a, dtype=np.float64)
    if a.ndim == 1:
        a = a.reshape(-1, 1)
    return np.asfortranarray(a)


def p(a):
    return C.c_void_p(a.ctypes.data)


MODE_INTENDED, MODE_LITERAL = 0, 1


def make_opts(mode=MODE_INTENDED, dist=0, seed=0, num_passes=0, passes_per_stab=0, omega_n=None, omega_m=None, skip_psd_check=False):
    o = Opts(mode, dist, seed, num_passes, passes_per_stab, None, None, int(bool(skip_psd_check)))
    keep = []
    if omega_n is not None:
        on = F(omega_n); keep.append(on); o.omega_n = on.ctypes.data
    if omega_m is not None:
        om = F(omega_m); keep.append(om); o.omega_m = om.ctypes.data
    o._keep = keep
    return o


def philox4x32_10(ctr, key):
    lib = load()
    ctr = np.ascontiguousarray(ctr, dtype=np.uint32).reshape(-1, 4)
    key = np.ascontiguousarray(key, dtype=np.uint32).reshape(-1, 2)
    out = np.empty_like(ctr)
    for i in range(ctr.shape[0]):
        k = key[i if key.shape[0] > 1 else 0]
        lib.orc_philox4x32_10(p(ctr[i]), p(np.ascontiguousarray(k)), C.c_void_p(out[i].ctypes.data))
    return out


def threefry2x64_20(ctr, key):
    lib = load()
    ctr = np.ascontiguousarray(ctr, dtype=np.uint64).reshape(-1, 2)
    key = np.ascontiguousarray(key, dtype=np.uint64).reshape(-1, 2)
    out = np.empty_like(ctr)
    for i in range(ctr.shape[0]):
        k = key[i if key.shape[0] > 1 else 0]
        lib.orc_threefry2x64_20(p(ctr[i]), p(np.ascontiguousarray(k)), C.c_void_p(out[i].ctypes.data))
    return out


def seed_from_u64(state):
    key = np.zeros(2, dtype=np.uint64)
    load().orc_seed_from_u64(u64(state), p(key))
    return key


def threefry_rng_u64(seed, t):
    return int(load().orc_threefry_rng_u64(seed, t))


def omega_fill(dist, rows, cols, seed=0, stream=0, row_off=0):
    out = np.empty((rows, cols), dtype=np.float64, order="F")
    load().orc_omega_fill(C.c_int(dist), u64(seed), u32(stream), i64(rows), i64(cols), i64(row_off), p(out), i64(max(rows, 1)))
    return out


def gauss_from_u32(k):
    return float(load().orc_gauss_from_u32(u32(k)))


def sketching_operator_ref(dist, rows, cols, seed=0):
    out = np.empty((rows, cols), dtype=np.float64, order="F")
    rc = load().orc_sketching_operator_ref(C.c_int(dist), u64(seed), i64(rows), i64(cols), p(out))
    if rc != 0:
        raise NotImplementedError("the reference's Gaussian stream needs rand_distr's ziggurat tables (not in the tree)")
    return out


def gemm_nn(A, B):
    A, B = F(A), F(B)
    m, K = A.shape; N = B.shape[1]
    Cm = np.empty((m, N), dtype=np.float64, order="F")
    load().orc_gemm_nn(p(A), i64(m), i64(m), i64(K), p(B), i64(K), i64(N), p(Cm), i64(m))
    return Cm


def gemm_tn(A, Q):
    A, Q = F(A), F(Q)
    m, n = A.shape; N = Q.shape[1]
    Z = np.empty((n, N), dtype=np.float64, order="F")
    load().orc_gemm_tn(p(A), i64(m), i64(m), i64(n), p(Q), i64(m), i64(N), p(Z), i64(n))
    return Z


def qr(X):
    X = F(X); rows, cols = X.shape; pp = min(rows, cols)
    Q = np.empty((rows, pp), dtype=np.float64, order="F"); R = np.empty((pp, cols), dtype=np.float64, order="F")
    load().orc_qr(p(X), i64(rows), i64(cols), p(Q), p(R))
    return Q, R


def Orth(X):
    return qr(X)[0]


def haar_sample(rows, cols, attr, seed=0):
    """src/sketch.rs:45-85; attr 0 = Row, 1 = Column"""
    out = np.empty((rows, cols), dtype=np.float64, order="F")
    rc = load().orc_haar_sample(i64(rows), i64(cols), C.c_int(attr), u64(seed), p(out))
    if rc:
        raise OracleError(rc)
    return out


def Stabilizer(X):
    X = F(X); rows, cols = X.shape
    L = np.empty((rows, min(rows, cols)), dtype=np.float64, order="F")
    load().orc_stabilizer(p(X), i64(rows), i64(cols), p(L))
    return L


def svd(M):
    M = F(M); rows, cols = M.shape; pp = min(rows, cols)
    U = np.empty((rows, pp), order="F"); s = np.empty(pp); Vt = np.empty((pp, cols), order="F")
    rc = load().orc_svd(p(M), i64(rows), i64(cols), p(U), p(s), p(Vt))
    assert rc == 0
    return U, s, Vt


def symmetric_eigen(A):
    A = F(A); n = A.shape[0]
    W = np.empty((n, n), order="F"); lam = np.empty(n)
    rc = load().orc_symmetric_eigen(p(A), i64(n), p(W), p(lam))
    assert rc == 0
    return lam, W


def tsog1(A, k, num_passes, passes_per_stab, opts=None):
    A = F(A); m, n = A.shape
    o = opts or make_opts()
    S = np.empty((n, k), dtype=np.float64, order="F")
    load().orc_tsog1(p(A), i64(m), i64(n), i64(k), C.c_int(num_passes), C.c_int(passes_per_stab), C.byref(o), p(S))
    return S


def RF1(A, k, opts=None):
    A = F(A); m, n = A.shape
    o = opts or make_opts()
    Q = np.empty((m, k), dtype=np.float64, order="F")
    load().orc_rf1(p(A), i64(m), i64(n), i64(k), C.byref(o), p(Q))
    return Q


def QB1(A, k, epsilon=0.0, opts=None):
    A = F(A); m, n = A.shape
    o = opts or make_opts()
    Q = np.empty((m, k), dtype=np.float64, order="F"); B = np.empty((k, n), dtype=np.float64, order="F")
    load().orc_qb1(p(A), i64(m), i64(n), i64(k), C.byref(o), p(Q), p(B))
    return Q, B


class OracleError(Exception):
    def __init__(self, code):
        self.code = code
        super().__init__(f"oracle status {code}")


def rand_svd(A, k, epsilon, s, opts=None):
    A = F(A); m, n = A.shape
    o = opts or make_opts()
    cap = max(min(max(k, 1), m, n), 1)
    U = np.empty((m, cap), order="F"); S = np.empty((cap, cap), order="F"); Vt = np.empty((cap, n), order="F")
    r = i64(0)
    rc = load().orc_rand_svd(p(A), i64(m), i64(n), i64(k), f64(epsilon), i64(s), C.byref(o), p(U), p(S), p(Vt), C.byref(r))
    if rc:
        raise OracleError(rc)
    return U, S, Vt


def rand_evd1(A, k, epsilon, s, opts=None):
    A = F(A); n = A.shape[0]
    o = opts or make_opts()
    cap = max(min(max(k, 1), n), 1)
    V = np.empty((n, cap), order="F"); lam = np.empty(cap)
    r = i64(0)
    rc = load().orc_rand_evd1(p(A), i64(n), i64(k), f64(epsilon), i64(s), C.byref(o), p(V), p(lam), C.byref(r))
    if rc:
        raise OracleError(rc)
    return V[:, :r.value], lam[:r.value]


def rand_evd2(A, k, s, opts=None):
    A = F(A); n = A.shape[0]
    o = opts or make_opts()
    cap = max(min(max(k, 1), n), 1)
    V = np.empty((n, cap), order="F"); lam = np.empty(cap)
    r = i64(0)
    rc = load().orc_rand_evd2(p(A), i64(n), i64(k), i64(s), C.byref(o), p(V), p(lam), C.byref(r))
    if rc:
        raise OracleError(rc)
    return V[:, :r.value], lam[:r.value]


def sketch_dim(m, n, sf, saddle=False):
    return int(load().orc_sketch_dim(m, n, sf, 1 if saddle else 0))


def sketch_apply_dense(A, d, dist=0, seed=0):
    A = F(A); m, n = A.shape
    out = np.empty((d, n), order="F")
    load().orc_sketch_apply_dense(C.c_int(dist), u64(seed), i64(d), p(A), i64(m), i64(n), p(out))
    return out


def sketch_apply_saso(A, d, zeta=8, seed=0):
    A = F(A); m, n = A.shape
    out = np.empty((d, n), order="F")
    load().orc_sketch_apply_saso(u64(seed), i64(d), C.c_int(zeta), p(A), i64(m), i64(n), p(out))
    return out


def sketch_apply_saso_block(A, d, zeta=8, seed=0, row_off=0, width=0):
    A = F(A); m, n = A.shape
    out = np.empty((d, n), order="F")
    w = width if width else min(zeta, 4)
    load().orc_sketch_apply_saso_block(u64(seed), i64(d), C.c_int(zeta), C.c_int(w), p(A), i64(m), i64(n), i64(row_off), p(out))
    return out


def cgls(a, b, tolerance, num_iterations, x0=None):
    """src/cg.rs:18-61 -> (x, iterations, converged)"""
    a = F(a); m, n = a.shape
    b = F(np.asarray(b, dtype=np.float64).reshape(-1, 1))
    x = F(np.zeros((n, 1)) if x0 is None else np.array(x0, dtype=np.float64).reshape(-1, 1))
    conv = C.c_int(0)
    lib = load()
    lib.orc_cgls.restype = i64
    it = lib.orc_cgls(p(a), i64(m), i64(n), p(b), C.c_double(tolerance), i64(num_iterations), p(x), C.byref(conv))
    return x, int(it), bool(conv.value)


def blendenpik(A, b, epsilon, l, sampling_factor, kind=0, dist_or_width=0, zeta=8, seed=0):
    """src/sketch_and_precondition.rs:26-59 -> (x, iterations, converged); raises ValueError(code) on the reference's errors"""
    A = F(A); m, n = A.shape
    b = F(np.asarray(b, dtype=np.float64).reshape(-1, 1))
    x = np.zeros((n, 1), order="F")
    it = i64(0); conv = C.c_int(0)
    rc = load().orc_blendenpik(p(A), i64(m), i64(n), p(b), C.c_double(epsilon), i64(l), C.c_double(sampling_factor), C.c_int(kind),
                               C.c_int(dist_or_width), C.c_int(zeta), u64(seed), p(x), C.byref(it), C.byref(conv))
    if rc:
        raise ValueError(rc)
    return x, int(it.value), bool(conv.value)


def lsrn(A, b, epsilon, l, sampling_factor, kind=0, dist_or_width=0, zeta=8, seed=0):
    """src/sketch_and_precondition.rs:82-119 -> (x, iterations, converged)"""
    A = F(A); m, n = A.shape
    b = F(np.asarray(b, dtype=np.float64).reshape(-1, 1))
    x = np.zeros((n, 1), order="F")
    it = i64(0); conv = C.c_int(0)
    rc = load().orc_lsrn(p(A), i64(m), i64(n), p(b), C.c_double(epsilon), i64(l), C.c_double(sampling_factor), C.c_int(kind),
                         C.c_int(dist_or_width), C.c_int(zeta), u64(seed), p(x), C.byref(it), C.byref(conv))
    if rc:
        raise ValueError(rc)
    return x, int(it.value), bool(conv.value)


# ---- next rows: oracle_next.c ---------------------------------------------------
def _ivec(n):
    return np.zeros(max(int(n), 1), dtype=np.int64)


def qrcp(A, steps=None, want_q=True):
    """src/pivot_decompositions.rs:105-180 (steps None = min(m, n)) / economic_qrcp :196-269 (steps = k)
    -> (Q m x m or None, R m x n work matrix, perm)"""
    A = F(A); m, n = A.shape
    steps = min(m, n) if steps is None else int(steps)
    R = np.empty((m, n), order="F"); Q = np.empty((m, m), order="F") if want_q else None
    perm = _ivec(n)
    load().orc_qrcp_steps(p(A), i64(m), i64(n), i64(steps), p(R), p(Q) if want_q else None, p(perm))
    return Q, R, perm[:n]


def economic_qrcp(A, k):
    Q, R, perm = qrcp(A, k, True)
    return Q[:, :k].copy(order="F"), R[:k, :].copy(order="F"), perm


def sap_chol_qrcp(A, d, kind=0, dist_or_width=0, zeta=8, seed=0):
    """src/cqrrpt.rs:27-58 -> (Q m x k, R k x n, J)"""
    A = F(A); m, n = A.shape
    Q = np.zeros((m, n), order="F"); R = np.zeros(n * n); J = _ivec(n); k = i64(0)
    rc = load().orc_sap_chol_qrcp(p(A), i64(m), i64(n), i64(d), C.c_int(kind), C.c_int(dist_or_width), C.c_int(zeta), u64(seed),
                                  p(Q), p(R), p(J), C.byref(k))
    if rc:
        raise ValueError(rc)
    k = int(k.value)
    return Q[:, :k].copy(order="F"), R[:k * n].reshape((k, n), order="F").copy(order="F"), J[:n]


def sketched_least_squares(which, A, b, kind=0, dist_or_width=0, zeta=8, seed=0):
    """src/sketch_and_solve.rs: which 0 = QR (:24-33), 1 = SVD (:54-66)"""
    A = F(A); m, n = A.shape
    b = F(np.asarray(b, dtype=np.float64).reshape(-1, 1))
    x = np.zeros((n, 1), order="F")
    rc = load().orc_sketched_least_squares(C.c_int(which), p(A), i64(m), i64(n), p(b), C.c_int(kind), C.c_int(dist_or_width),
                                           C.c_int(zeta), u64(seed), p(x))
    if rc:
        raise ValueError(rc)
    return x


ROW, COLUMN = 0, 1


def osid_qrcp(Y, k, attr):
    """src/id.rs:272-318 -> (X, J)"""
    Y = F(Y); l, w = Y.shape
    X = np.zeros((k, w) if attr == COLUMN else (l, k), order="F"); J = _ivec(k)
    rc = load().orc_osid_qrcp(p(Y), i64(l), i64(w), i64(k), C.c_int(attr), p(X), p(J))
    if rc:
        raise ValueError(rc)
    return X, J[:k]


def osid_randomised(A, k, attr, opts=None):
    """src/id.rs:217-249 -> (X, J)"""
    A = F(A); m, n = A.shape
    o = opts if opts is not None else make_opts()
    X = np.zeros((k, n) if attr == COLUMN else (m, k), order="F"); J = _ivec(k)
    rc = load().orc_osid_randomised(p(A), i64(m), i64(n), i64(k), C.c_int(attr), C.byref(o), p(X), p(J))
    if rc:
        raise ValueError(rc)
    return X, J[:k]


def two_sided_id(A, k, randomised=False, opts=None):
    """src/id.rs:118-129 / :94-101 -> (Z, I, J, X)"""
    A = F(A); m, n = A.shape
    o = opts if opts is not None else make_opts()
    Z = np.zeros((m, k), order="F"); X = np.zeros((k, n), order="F"); I = _ivec(k); J = _ivec(k)
    rc = load().orc_two_sided_id(C.c_int(1 if randomised else 0), p(A), i64(m), i64(n), i64(k), C.byref(o), p(Z), p(I), p(J), p(X))
    if rc:
        raise ValueError(rc)
    return Z, I[:k], J[:k], X


def cur(A, k, randomised=False, opts=None):
    """src/id.rs:34-71 / :154-193 -> (J, U, I)"""
    A = F(A); m, n = A.shape
    o = opts if opts is not None else make_opts()
    U = np.zeros((k, k), order="F"); I = _ivec(k); J = _ivec(k)
    rc = load().orc_cur(C.c_int(1 if randomised else 0), p(A), i64(m), i64(n), i64(k), C.byref(o), p(J), p(U), p(I))
    if rc:
        raise ValueError(rc)
    return J[:k], U, I[:k]


def saddle_point(A, b, c, mu, epsilon, l, sampling_factor, dist=0, seed=0):
    """src/sketch_and_precondition.rs:150-216 -> (x, y, iterations, converged)"""
    A = F(A); m, n = A.shape
    b = F(np.asarray(b, dtype=np.float64).reshape(-1, 1))
    cc = None if c is None or np.size(c) == 0 else F(np.asarray(c, dtype=np.float64).reshape(-1, 1))
    x = np.zeros((n, 1), order="F"); y = np.zeros((m, 1), order="F")
    it = i64(0); conv = C.c_int(0)
    rc = load().orc_saddle_point(p(A), i64(m), i64(n), p(b), p(cc) if cc is not None else None, C.c_double(mu), C.c_double(epsilon),
                                 i64(l), C.c_double(sampling_factor), C.c_int(dist), u64(seed), p(x), p(y), C.byref(it), C.byref(conv))
    if rc:
        raise ValueError(rc)
    return x, y, int(it.value), bool(conv.value)


def lsqr(a, b, damp=0.0, atol=1e-6, btol=1e-6, conlim=1e8, iter_lim=None, calc_var=False, x0=None):
    """src/solvers.rs:115-278 -> (x, istop, itn, r1norm, r2norm, anorm, acond, arnorms, xnorm, var), scipy's return order
    with the arnorm history in place of the last arnorm, as the reference returns it"""
    a = F(a); m, n = a.shape
    b = F(np.asarray(b, dtype=np.float64).reshape(-1, 1))
    lim = 2 * n if iter_lim is None else int(iter_lim)
    x = np.zeros((n, 1), order="F"); var = np.zeros(n); hist = np.zeros(max(lim, 1))
    x0f = None if x0 is None else F(np.asarray(x0, dtype=np.float64).reshape(-1, 1))
    istop = i64(0); itn = i64(0)
    r1 = C.c_double(0); r2 = C.c_double(0); an = C.c_double(0); ac = C.c_double(0); xn = C.c_double(0)
    lib = load()
    lib.orc_lsqr.restype = i64
    nh = lib.orc_lsqr(p(a), i64(m), i64(n), p(b), C.c_double(damp), C.c_double(atol), C.c_double(btol), C.c_double(conlim), i64(lim),
                      C.c_int(1 if calc_var else 0), p(x0f) if x0f is not None else None, p(x), C.byref(istop), C.byref(itn),
                      C.byref(r1), C.byref(r2), C.byref(an), C.byref(ac), p(hist), C.byref(xn), p(var))
    return x, int(istop.value), int(itn.value), r1.value, r2.value, an.value, ac.value, hist[:int(nh)].copy(), xn.value, var


def conjugate_grad(a, b, x0=None):
    """src/cg.rs:77-112 -> (x, iterations, converged); raises ValueError(9) for NotPositiveSemiDefinite"""
    a = F(a); n = a.shape[0]
    b = F(np.asarray(b, dtype=np.float64).reshape(-1, 1))
    x = F(np.ones((n, 1)) if x0 is None else np.array(x0, dtype=np.float64).reshape(-1, 1))
    it = i64(0); conv = C.c_int(0)
    rc = load().orc_conjugate_grad(p(a), i64(n), p(b), p(x), C.byref(it), C.byref(conv))
    if rc:
        raise ValueError(rc)
    return x[:, 0].copy(), int(it.value), bool(conv.value)


def verify_solution(a, b, x):
    """src/cg.rs:115-117"""
    return float(np.linalg.norm(np.asarray(a) @ np.asarray(x).reshape(-1) - np.asarray(b).reshape(-1)))


def lupp(a):
    """src/pivot_decompositions.rs:21-86 -> (l, u, p); raises ValueError(5) for a non-square input, ValueError(6) for a singular one"""
    a = F(a)
    if a.shape[0] != a.shape[1]:
        raise ValueError(5)
    n = a.shape[0]
    L = np.zeros((n, n), order="F"); U = np.zeros((n, n), order="F"); perm = np.zeros(n, dtype=np.int64)
    rc = load().orc_lupp(p(a), i64(n), p(L), p(U), p(perm))
    if rc:
        raise ValueError(rc)
    return L, U, perm


def set_threads(n):
    load().orc_set_threads(C.c_int(n))


def get_threads():
    return int(load().orc_get_threads())
