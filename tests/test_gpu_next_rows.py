"""Parity of the rows after the hot path (CQRRPT, sketch-and-solve, ID / CUR, saddle point) with the
CPU oracle's restatement (oracle/oracle_next.c), through the C ABI, on a B200.  The test bodies follow the reference's own
tests (src/pivot_decompositions.rs:318-395, src/cqrrpt.rs:72-133, src/sketch_and_solve.rs:80-158, src/id.rs:329-546,
src/sketch_and_precondition.rs:277-337), with the assertions the reference only prints turned into bounds.

Both sides draw the sketching operators from the same Philox map, so pivots, ranks and factors are comparable entry for
entry.  Tolerances: factors and solutions to 1e-9 relative (f64, different summation orders); index vectors exact."""
import numpy as np
import pytest

from conftest import rank_k_matrix, random_matrix

pytestmark = pytest.mark.gpu


def perm_matrix_t(p):
    """src/test_assist.rs permutation_vector_to_transpose_matrix: P[i, p[i]] = 1"""
    n = len(p)
    P = np.zeros((n, n))
    P[np.arange(n), p] = 1.0
    return P


# ------------------------------------------------------------------------------------------- pivot_decompositions
@pytest.mark.parametrize("m,n", [(137, 23), (500, 29), (40, 40), (3000, 64), (20, 55)])
def test_qrcp_matches_oracle_and_reference_properties(rb, orc, m, n):
    """src/pivot_decompositions.rs:320-348 test_qrcp: orthonormal q, upper-triangular r, q r p^T = a; plus entry-wise
    agreement with the oracle (same pivots, same reflectors)"""
    from randnla_b200 import pivot_decompositions as pd
    A = random_matrix(m, n, seed=m + n)
    q, r, p = pd.qrcp(A)
    Qo, Ro, po = orc.qrcp(A)
    assert p == [int(v) for v in po]
    assert sorted(p) == list(range(n))
    scale = np.abs(A).max()
    assert np.abs(r - Ro).max() <= 1e-12 * scale * np.sqrt(m)
    assert np.abs(q - Qo).max() <= 1e-12 * np.sqrt(m)
    assert np.abs(q.T @ q - np.eye(m)).max() < 1e-13 * m
    assert np.abs(np.tril(r, -1)).max() <= 1e-13 * scale * np.sqrt(m)
    assert np.abs(q @ r @ perm_matrix_t(p) - A).max() <= 1e-12 * scale * np.sqrt(m)
    d = np.abs(np.diag(r))
    assert (d[:-1] >= d[1:] * (1 - 1e-12)).all()          # rank revealing: non-increasing diagonal


@pytest.mark.parametrize("m,n,k", [(300, 25, 12), (64, 200, 30), (2500, 40, 40), (9, 7, 1)])
def test_economic_qrcp(rb, orc, m, n, k):
    """src/pivot_decompositions.rs:372-386 test_qrcp_economical + oracle parity; :389-395 bad input"""
    from randnla_b200 import pivot_decompositions as pd
    from randnla_b200.errors import InvalidParameters
    A = random_matrix(m, n, seed=k)
    q, r, p = pd.economic_qrcp(A, k)
    Qo, Ro, po = orc.economic_qrcp(A, k)
    assert q.shape == (m, k) and r.shape == (k, n)
    assert p == [int(v) for v in po]
    assert np.abs(r - Ro).max() <= 1e-12 * np.abs(A).max() * np.sqrt(m)
    assert np.abs(q - Qo).max() <= 1e-12 * np.sqrt(m)
    assert np.abs(q.T @ q - np.eye(k)).max() < 1e-13 * m
    assert np.abs(np.tril(r[:, :k], -1)).max() <= 1e-13 * np.abs(A).max() * np.sqrt(m)
    # the leading k columns are reproduced exactly by the truncated factorisation
    assert np.abs(q @ r[:, :k] - A[:, p[:k]]).max() <= 1e-12 * np.abs(A).max() * np.sqrt(m)
    with pytest.raises(InvalidParameters):
        pd.economic_qrcp(A, 0)
    with pytest.raises(InvalidParameters):
        pd.economic_qrcp(A, min(m, n) + 1)


def test_qrcp_rank_deficient_and_zero(rb, orc):
    """zero trailing columns: |x| == 0 skips the reflection (:143); a zero matrix returns q = I, r = 0, p = identity"""
    from randnla_b200 import pivot_decompositions as pd
    A = rank_k_matrix(60, 20, 5, seed=3)
    q, r, p = pd.qrcp(A)
    assert np.abs(q @ r @ perm_matrix_t(p) - A).max() <= 1e-11 * np.abs(A).max()
    assert np.abs(np.diag(r))[5:].max() <= 1e-10 * np.abs(np.diag(r))[0]
    Z = np.zeros((12, 5), order="F")
    q, r, p = pd.qrcp(Z)
    assert p == list(range(5)) and np.array_equal(q, np.eye(12)) and not r.any()


# ------------------------------------------------------------------------------------------------------- cqrrpt
@pytest.mark.parametrize("m,n,d,kind", [(100, 10, 37, 0), (4000, 60, 240, 0), (20000, 130, 520, 2), (6000, 300, 600, 0)])
def test_sap_chol_qrcp(rb, orc, m, n, d, kind):
    """src/cqrrpt.rs:73-112 test_cqrrpt: unit, mutually orthogonal columns of q; upper-triangular r; q r p^T = a
    (the reference compares against deterministic qrcp at 1e-4); plus parity with the oracle on the same sketch"""
    from randnla_b200 import cqrrpt
    A = np.asfortranarray(np.random.default_rng(d).uniform(-1, 1, (m, n)))
    q, r, j = cqrrpt.sap_chol_qrcp(A, d, kind=kind)
    Qo, Ro, Jo = orc.sap_chol_qrcp(A, d, kind=kind)
    assert q.shape == (m, n) and r.shape == (n, n)
    assert j == [int(v) for v in Jo]
    assert np.abs(q.T @ q - np.eye(n)).max() < 1e-13 * n
    assert np.abs(np.tril(r, -1)).max() <= 1e-12 * np.abs(r).max()
    assert np.abs(q @ r @ perm_matrix_t(j) - A).max() <= 1e-12 * np.sqrt(m)
    assert np.abs(r - Ro).max() <= 1e-10 * np.abs(Ro).max()
    assert np.abs(q - Qo).max() <= 1e-10


def test_sap_chol_qrcp_rank_deficient_and_errors(rb, orc):
    """numerical rank k < n (:37-43): q has k columns, a[:, j[:k]] = q r[:, :k]; the reference's panics (:114-133)"""
    from randnla_b200 import cqrrpt
    from randnla_b200.errors import InvalidParameters
    A = rank_k_matrix(500, 30, 11, seed=5)
    q, r, j = cqrrpt.sap_chol_qrcp(A, 90)
    Qo, Ro, Jo = orc.sap_chol_qrcp(A, 90)
    assert q.shape == (500, 11) and r.shape == (11, 30) and Qo.shape == q.shape
    assert j[:11] == [int(v) for v in Jo[:11]]
    assert np.abs(q.T @ q - np.eye(11)).max() < 1e-12
    assert np.abs(q @ r - A[:, j]).max() <= 1e-9 * np.abs(A).max()
    data = random_matrix(100, 10, seed=1)
    with pytest.raises(InvalidParameters, match="d must satisfy"):
        cqrrpt.sap_chol_qrcp(data, 5)                     # test_cqrrpt_wrong_d
    with pytest.raises(InvalidParameters, match="d must satisfy"):
        cqrrpt.sap_chol_qrcp(data.T.copy(), 50)           # test_cqrrpt_wide_matrix


# --------------------------------------------------------------------------------------------- sketch_and_solve
@pytest.mark.parametrize("kind", [0, 2])
@pytest.mark.parametrize("m,n", [(480, 25), (16000, 200)])
def test_sketched_least_squares(rb, orc, m, n, kind):
    """src/sketch_and_solve.rs:81-158: noisy data from a hypothesis; the sketched residual is within a small factor of the
    optimal one (the reference prints both); parity with the oracle on the same sketch"""
    from randnla_b200 import sketch_and_solve as ss
    from randnla_b200.errors import InvalidDimensions
    rng = np.random.default_rng(m)
    hyp = rng.uniform(-100, 100, (n, 1))
    A = rng.standard_normal((m, n))
    y = A @ hyp
    A = np.asfortranarray(A + 0.01 * rng.standard_normal((m, n)))
    xl = np.linalg.lstsq(A, y, rcond=None)[0]
    res_opt = np.linalg.norm(A @ xl - y)
    for which, fn in ((0, ss.sketched_least_squares_qr), (1, ss.sketched_least_squares_svd)):
        x = fn(A, y, kind=kind)
        xo = orc.sketched_least_squares(which, A, y, kind=kind)
        assert np.linalg.norm(x - xo) <= 1e-9 * np.linalg.norm(xo)
        assert np.linalg.norm(A @ x - y) <= 2.0 * res_opt      # (1 + eps) distortion with d = m / 4 >= 4.8 n
    with pytest.raises(InvalidDimensions):
        ss.sketched_least_squares_qr(random_matrix(40, 20, seed=1), random_matrix(40, 1, seed=2))


def test_sketched_least_squares_zero_pivot_rule(rb, orc):
    """src/solvers.rs:22-41, :57-69: a zero column gives a zero pivot / zero singular value; that unknown stays 0.
    (The zero column is the last one: for an interior zero column the reference's answer depends on which unit vector its
    Householder QR happens to put into q, which no other QR reproduces.)"""
    from randnla_b200 import sketch_and_solve as ss
    rng = np.random.default_rng(9)
    A = rng.standard_normal((400, 6)); A[:, 5] = 0.0
    A = np.asfortranarray(A)
    b = rng.standard_normal((400, 1))
    for which, fn in ((0, ss.sketched_least_squares_qr), (1, ss.sketched_least_squares_svd)):
        x = fn(A, b)
        xo = orc.sketched_least_squares(which, A, b)
        assert x[5, 0] == 0.0 and np.isfinite(x).all()
        assert np.linalg.norm(x - xo) <= 1e-9 * np.linalg.norm(xo)


# ----------------------------------------------------------------------------------------------------------- id
@pytest.mark.parametrize("m,n,k", [(104, 107, 60), (300, 90, 45), (70, 400, 33)])
def test_one_sided_id(rb, orc, m, n, k):
    """src/id.rs:463-490 test_one_sided_id on rank_k_matrix: the ID of an exactly rank-k matrix is exact"""
    from randnla_b200 import id as rid
    from randnla_b200.sketch import MatrixAttribute as MA
    A = rank_k_matrix(m, n, k, seed=k)
    nrm = np.linalg.norm(A)
    x, j = rid.osid_qrcp(A, k, MA.Column)
    xo, jo = orc.osid_qrcp(A, k, orc.COLUMN)
    assert j == [int(v) for v in jo] and x.shape == (k, n)
    assert np.linalg.norm(A[:, j] @ x - A) <= 1e-10 * nrm
    assert np.abs(x - xo).max() <= 1e-8 * max(1.0, np.abs(xo).max())
    assert np.array_equal(x[:, j], np.eye(k))
    x, i = rid.osid_qrcp(A, k, MA.Row)
    xo, io = orc.osid_qrcp(A, k, orc.ROW)
    assert i == [int(v) for v in io] and x.shape == (m, k)
    assert np.linalg.norm(x @ A[i, :] - A) <= 1e-10 * nrm
    x, j = rid.osid_randomised(A, k, MA.Column)
    xo, jo = orc.osid_randomised(A, k, orc.COLUMN)
    assert j == [int(v) for v in jo]
    assert np.linalg.norm(A[:, j] @ x - A) <= 1e-8 * nrm
    assert np.abs(x - xo).max() <= 1e-7 * max(1.0, np.abs(xo).max())


@pytest.mark.parametrize("mode", [0, 1])
@pytest.mark.parametrize("m,n,k", [(106, 101, 52), (500, 64, 40)])
def test_two_sided_id(rb, orc, m, n, k, mode):
    """src/id.rs:492-517 test_two_sided_id, deterministic and randomised; in both tsog1 modes (the Row step of the randomised
    variant calls tsog1(a, k, 2, 1), :230)"""
    from randnla_b200 import id as rid, runtime as rt
    A = rank_k_matrix(m, n, k, seed=k + 1)
    nrm = np.linalg.norm(A)
    z, i, j, x = rid.two_sided_id(A, k)
    zo, io, jo, xo = orc.two_sided_id(A, k, False)
    assert i == [int(v) for v in io] and j == [int(v) for v in jo]
    assert np.linalg.norm(z @ A[np.ix_(i, j)] @ x - A) <= 1e-9 * nrm
    with rt.options(mode=mode):
        z, i, j, x = rid.two_sided_id_randomised(A, k)
    zo, io, jo, xo = orc.two_sided_id(A, k, True, orc.make_opts(mode=mode))
    assert j == [int(v) for v in jo]
    assert np.linalg.norm(z @ A[np.ix_(i, j)] @ x - A) <= 1e-7 * nrm
    if mode == 1:
        # literal tsog1 is reproduced statement for statement (bit-identical Stabilizer): same row pivots, same Z.  In the
        # intended mode the two sides stabilise differently (CholeskyQR vs Householder: same range, different basis), so the
        # row pivots of a S^T may legitimately differ; the decomposition is still exact.
        assert i == [int(v) for v in io]
        assert np.abs(z - zo).max() <= 1e-6 * max(1.0, np.abs(zo).max())


@pytest.mark.parametrize("m,n,k", [(108, 103, 55), (90, 240, 31), (2000, 150, 64)])
def test_cur(rb, orc, m, n, k):
    """src/id.rs:519-546 test_cur, tall and wide, deterministic and randomised"""
    from randnla_b200 import id as rid
    A = rank_k_matrix(m, n, k, seed=k + 2)
    nrm = np.linalg.norm(A)
    for randomised, fn in ((False, rid.cur), (True, rid.cur_randomised)):
        j, u, i = fn(A, k)
        jo, uo, io = orc.cur(A, k, randomised)
        assert i == [int(v) for v in io] and j == [int(v) for v in jo]
        assert u.shape == (k, k)
        assert np.linalg.norm(A[:, j] @ u @ A[i, :] - A) <= 1e-7 * nrm
        assert np.abs(u - uo).max() <= 1e-6 * max(1.0, np.abs(uo).max())


def test_id_reference_panics(rb):
    """src/id.rs:329-461: k = 0 and k > min(m, n) panic in the reference -> InvalidParameters with the same text"""
    from randnla_b200 import id as rid
    from randnla_b200.sketch import MatrixAttribute as MA
    from randnla_b200.errors import InvalidParameters, InvalidDimensions
    data = np.asfortranarray(np.random.default_rng(0).uniform(-1, 1, (100, 10)))
    for k, msg in ((0, "k must be positive"), (11, "k must be <= min")):
        for call in (lambda: rid.osid_qrcp(data, k, MA.Column), lambda: rid.osid_randomised(data, k, MA.Column),
                     lambda: rid.cur(data, k), lambda: rid.cur_randomised(data, k),
                     lambda: rid.two_sided_id(data, k), lambda: rid.two_sided_id_randomised(data, k)):
            with pytest.raises(InvalidParameters, match=msg):
                call()
    with pytest.raises(InvalidDimensions):
        rid.osid_randomised(data, 4, MA.Row)              # a * tsog1(a, k, 2, 1)^T does not conform (:230-233)


def test_randomised_id_on_a_large_lowrank_matrix(rb):
    """the randomised column ID at a size where only the k x n sketch is ever pivoted: 200 000 x 2 000, rank 40"""
    import torch
    import ctypes as C
    from randnla_b200 import runtime as rt, _lib
    lib = _lib.load()
    m, n, k = 200_000, 2_000, 40
    g = torch.Generator(device="cuda").manual_seed(1)
    L = torch.randn((m, k), generator=g, device="cuda", dtype=torch.float64)
    Rm = torch.randn((k, n), generator=g, device="cuda", dtype=torch.float64)
    A = rt.empty_colmajor(m, n); A.copy_(L @ Rm)
    X = rt.empty_colmajor(k, n)
    J = torch.zeros(k, dtype=torch.int64, device="cuda")
    pa, lda = rt.dev_ptr_ld(A); px, ldx = rt.dev_ptr_ld(X)
    torch.cuda.synchronize()
    _lib.check(lib.rnla_osid_randomised_dev(pa, lda, m, n, k, 1, None, px, ldx, C.c_void_p(J.data_ptr())))
    torch.cuda.synchronize()
    err = torch.linalg.norm(A[:, J] @ X - A) / torch.linalg.norm(A)
    assert float(err) < 1e-9
    assert len(set(J.tolist())) == k


# ------------------------------------------------------------------------------------------------- saddle point
@pytest.mark.parametrize("mu,with_c", [(0.0, False), (0.0, True), (3.7, False), (3.7, True)])
@pytest.mark.parametrize("m,n", [(100, 10), (20000, 200)])
def test_saddle_point(rb, orc, m, n, mu, with_c):
    """src/sketch_and_precondition.rs:278-337 test_saddle_point (sampling factor 1.5 as there) + parity with the oracle.
    With mu = 0 the solution is that of the normal equations A^T A x = A^T b - c."""
    from randnla_b200 import sketch_and_precondition as sp
    rng = np.random.default_rng(m + n)
    A = np.asfortranarray(rng.standard_normal((m, n)))
    b = A @ rng.uniform(-100, 100, (n, 1)) + 1e-2 * rng.standard_normal((m, 1))
    c = rng.uniform(-10, 10, (n, 1)) if with_c else None
    info = {}
    x, y = sp.sketch_saddle_point_precondition(A, b, c, mu, 1e-9, 1000, 1.5, info=info)
    xo, yo, ito, convo = orc.saddle_point(A, b, c, mu, 1e-9, 1000, 1.5)
    nrm = np.linalg.norm(xo)
    # the stopping test sits on its threshold after ~75 steps of a cond ~ 10 system: within 5 % of the oracle's count
    assert info["converged"] and convo and abs(info["iterations"] - ito) <= max(2, ito // 20)
    assert np.linalg.norm(x - xo) <= 1e-8 * nrm
    assert np.linalg.norm(y - (b - A @ x)) <= 1e-10 * np.linalg.norm(b)
    assert np.linalg.norm(y - yo) <= 1e-8 * max(np.linalg.norm(yo), np.linalg.norm(b) * 1e-3)
    if mu == 0.0:
        cz = np.zeros((n, 1)) if c is None else c
        xt = np.linalg.solve(A.T @ A, A.T @ b - cz)
        assert np.linalg.norm(x - xt) <= 1e-8 * np.linalg.norm(xt)


def test_saddle_point_reference_errors(rb):
    """src/sketch_and_precondition.rs:320-330: Err for sampling_factor < 1, epsilon <= 0, l = 0; underdetermined systems"""
    from randnla_b200 import sketch_and_precondition as sp
    from randnla_b200.errors import InvalidParameters, NotOverdetermined
    A = random_matrix(100, 10, seed=1); b = random_matrix(100, 1, seed=2); c = random_matrix(10, 1, seed=3)
    for eps, l, sf in ((1e-4, 1000, 0.5), (-1e-4, 1000, 1.5), (0.0, 1000, 1.5), (1e-4, 0, 1.5)):
        with pytest.raises(InvalidParameters):
            sp.sketch_saddle_point_precondition(A, b, c, 1.0, eps, l, sf)
    with pytest.raises(NotOverdetermined):
        sp.sketch_saddle_point_precondition(A.T.copy(), random_matrix(10, 1, seed=4), random_matrix(100, 1, seed=5), 1.0, 1e-4, 10, 1.5)


# ---------------------------------------------------------------------------------------------------------- lsqr
def _lsq_problem(m, n, cond, seed, consistent=False):
    rng = np.random.default_rng(seed)
    U, _ = np.linalg.qr(rng.standard_normal((m, n)))
    V, _ = np.linalg.qr(rng.standard_normal((n, n)))
    A = np.asfortranarray((U * np.logspace(0, -np.log10(cond), n)) @ V.T)
    x = rng.uniform(-100, 100, n)
    b = A @ x + (0.0 if consistent else 1e-2) * rng.standard_normal(m)
    return A, b


@pytest.mark.parametrize("m,n,cond,damp,calc_var,use_x0,consistent", [
    (200, 100, 10.0, 0.0, False, False, False),       # test_overdetermined_system's shape (src/solvers.rs:413-468)
    (3001, 257, 1e3, 0.0, True, False, False),        # odd sizes: scalar tails of both streaming kernels
    (3000, 40, 1e3, 0.5, True, True, False),
    (120, 120, 50.0, 0.0, False, True, True),
    (20000, 300, 1e6, 1e-3, False, False, False),
])
def test_lsqr_matches_oracle(rb, orc, m, n, cond, damp, calc_var, use_x0, consistent):
    """src/solvers.rs:115-278.  (1) six iterations with the stopping tests off: every return value to 1e-10 (summation orders
    differ; beyond ~8 iterations the unorthogonalised recurrences amplify rounding, see tests/test_oracle_next_rows.py);
    (2) to convergence: same stopping reason, iteration count within a few, residual norm to 1e-6 / 1e-3."""
    from randnla_b200 import solvers
    A, b = _lsq_problem(m, n, cond, seed=m + n, consistent=consistent)
    x0 = np.random.default_rng(9).standard_normal(n) if use_x0 else None
    got = solvers.lsqr(A, b, damp, 0.0, 0.0, 0.0, 6, calc_var, x0)
    ref = orc.lsqr(A, b, damp, 0.0, 0.0, 0.0, 6, calc_var, x0)
    assert got[1] == ref[1] == 7 and got[2] == ref[2] == 6
    assert np.abs(got[0] - ref[0]).max() <= 1e-10 * np.abs(ref[0]).max()
    for i in (3, 4, 5, 6, 8):
        assert abs(got[i] - ref[i]) <= 1e-10 * abs(ref[i]), i
    assert len(got[7]) == 6 and np.abs(got[7] - ref[7]).max() <= 1e-10 * ref[7].max()
    if calc_var:
        assert np.abs(got[9] - ref[9]).max() <= 1e-10 * ref[9].max()
    else:
        assert not got[9].any()
    got = solvers.lsqr(A, b, damp, 1e-8, 1e-8, 1e8, None, calc_var, x0)
    ref = orc.lsqr(A, b, damp, 1e-8, 1e-8, 1e8, None, calc_var, x0)
    assert got[1] == ref[1] and abs(got[2] - ref[2]) <= max(2, ref[2] // 20)
    assert len(got[7]) == got[2]
    if cond <= 50.0:
        assert np.abs(got[0] - ref[0]).max() <= 1e-6 * np.abs(ref[0]).max()
    # the ill-conditioned cases run into iter_lim in their slowly converging tail, where the iterates of two implementations
    # have long decorrelated (the oracle itself moves by 2e-4 between two hosts): the residual norms agree loosely, and the
    # running estimate r1norm is the true residual of the x that was returned
    assert abs(got[3] - ref[3]) <= (1e-6 if cond <= 50.0 else 2e-2) * abs(ref[3]) + 1e-7 * np.linalg.norm(b)
    true_r = np.linalg.norm(b - A @ got[0][:, 0])
    assert abs(got[3] - true_r) <= (1e-8 if cond <= 1e3 else 2e-3) * true_r + 1e-7 * np.linalg.norm(b)


def test_lsqr_reference_cases(rb):
    """test_simple_system (src/solvers.rs:391-410) and test_overdetermined_system (:413-468, against LAPACK instead of
    nalgebra's SVD solve); early return on b = 0 (:181-183); iter_lim (istop = 7); scipy's lsqr, which the reference
    translates, gives the same answer"""
    from scipy.sparse.linalg import lsqr as sp_lsqr
    from randnla_b200 import solvers
    A = np.asfortranarray(np.array([[1.0, 0.0], [1.0, 1.0], [0.0, 1.0]]))
    x, istop, itn, r1, r2, an, ac, hist, xn, var = solvers.lsqr(A, np.zeros(3), 0.0, 1e-8, 1e-8, 1e8, None, False, None)
    assert not x.any() and istop == 0 and itn == 0 and list(hist) == [0.0] and r1 == 0.0 and an == 0.0
    x, istop, itn, *_ = solvers.lsqr(A, np.array([1.0, 0.0, -1.0]), 0.0, 1e-8, 1e-8, 1e8, None, False, None)
    assert abs(x[0, 0] - 1.0) < 1e-12 and abs(x[1, 0] + 1.0) < 1e-12 and istop in (1, 2) and itn <= 2
    A, b = _lsq_problem(200, 100, 10.0, seed=2)
    x, istop, itn, r1, *_ = solvers.lsqr(A, b, 0.0, 1e-12, 1e-12, 1e8, None, False, None)
    xs = np.linalg.lstsq(A, b, rcond=None)[0]
    assert np.abs(x[:, 0] - xs).max() <= 1e-8 * np.abs(xs).max()
    assert abs(r1 - np.linalg.norm(A @ xs - b)) <= 1e-8 * r1
    ref = sp_lsqr(A, b, atol=1e-12, btol=1e-12, conlim=1e8, iter_lim=200)
    assert istop == ref[1] and abs(itn - ref[2]) <= 2 and np.abs(x[:, 0] - ref[0]).max() <= 1e-8 * np.abs(xs).max()
    x, istop, itn, *_, hist, xn, var = solvers.lsqr(A, b, 0.0, 1e-14, 1e-14, 1e8, 7, False, None)
    assert istop == 7 and itn == 7 and len(hist) == 7
    with pytest.raises(ValueError):
        solvers.lsqr(A, b[:-1])


def test_lsqr_underdetermined_damped_and_zero_rhs(rb, orc):
    """a wide system (64 x 90) with damping, against the oracle over six iterations and against the regularised normal equations at
    convergence; b = 0 with a non-zero x0 (the early return of :181-183 does not apply: u = -A x0)"""
    from randnla_b200 import solvers
    A = random_matrix(64, 90, seed=12)
    b = np.random.default_rng(13).standard_normal(64)
    got = solvers.lsqr(A, b, 0.1, 0.0, 0.0, 0.0, 6, True, None)
    ref = orc.lsqr(A, b, 0.1, 0.0, 0.0, 0.0, 6, True, None)
    assert got[1] == ref[1] == 7 and np.abs(got[0] - ref[0]).max() <= 1e-11 * np.abs(ref[0]).max()
    for i in (3, 4, 5, 6, 8):
        assert abs(got[i] - ref[i]) <= 1e-11 * abs(ref[i])
    assert np.abs(got[9] - ref[9]).max() <= 1e-11 * ref[9].max()
    x, istop, *_ = solvers.lsqr(A, b, 0.1, 1e-13, 1e-13, 1e8, None, False, None)
    xs = np.linalg.solve(A.T @ A + 0.01 * np.eye(90), A.T @ b)
    assert istop in (1, 2) and np.abs(x[:, 0] - xs).max() <= 1e-8 * np.abs(xs).max()
    x0 = np.random.default_rng(14).standard_normal(90)
    got = solvers.lsqr(A, np.zeros(64), 0.0, 0.0, 0.0, 0.0, 4, False, x0)
    ref = orc.lsqr(A, np.zeros(64), 0.0, 0.0, 0.0, 0.0, 4, False, x0)
    assert got[2] == ref[2] and np.abs(got[0] - ref[0]).max() <= 1e-11 * np.abs(ref[0]).max()


def test_lsqr_at_scale_on_device_buffers(rb):
    """rnla_lsqr_dev on a 1 000 000 x 500 system (4 GB) resident in HBM: planted solution recovered, the normal-equations
    residual is at rounding level, and the running estimates (r1norm, xnorm, anorm <= ||A||_F) agree with the true values"""
    import ctypes as C
    import torch
    from randnla_b200 import runtime as rt, _lib
    lib = _lib.load()
    m, n = 1_000_000, 500
    dA = rt.empty_colmajor(m, n); pA, lda = rt.dev_ptr_ld(dA)
    _lib.check(lib.rnla_sketch_fill_dev(0, 0, 31, 4, m, n, 0, pA, lda)); rt.synchronize()
    torch.manual_seed(5)
    xt = torch.rand(n, 1, dtype=torch.float64, device="cuda") * 200 - 100
    noise = torch.randn(m, 1, dtype=torch.float64, device="cuda") * 1e-2
    db = rt.empty_colmajor(m, 1); db.copy_(dA @ xt + noise)
    dx = rt.empty_colmajor(n, 1); dvar = torch.zeros(n, dtype=torch.float64, device="cuda")
    res = _lib.LsqrResult(); hist = np.zeros(2 * n)
    before = rt.kernel_launches()
    _lib.check(lib.rnla_lsqr_dev(pA, lda, m, n, C.c_void_p(db.data_ptr()), 0.0, 1e-13, 1e-13, 1e8, -1, 1, None,
                                 C.c_void_p(dx.data_ptr()), C.byref(res), C.c_void_p(hist.ctypes.data), hist.size,
                                 C.c_void_p(dvar.data_ptr())))
    rt.synchronize()
    assert rt.kernel_launches() > before
    assert res.istop in (1, 2) and 0 < res.itn < 60 and res.n_arnorms == res.itn
    r = db - dA @ dx
    assert float(torch.linalg.vector_norm(dA.T @ r)) <= 1e-9 * float(torch.linalg.vector_norm(dA)) * float(torch.linalg.vector_norm(r))
    assert float(torch.linalg.vector_norm(dx - xt) / torch.linalg.vector_norm(xt)) < 1e-6        # noise 1e-2 / sqrt(m) per entry
    assert abs(res.r1norm - float(torch.linalg.vector_norm(r))) <= 1e-8 * res.r1norm
    assert abs(res.xnorm - float(torch.linalg.vector_norm(dx))) <= 1e-8 * res.xnorm
    assert res.anorm <= 1.001 * float(torch.linalg.vector_norm(dA)) * np.sqrt(res.itn)
    assert float(dvar.min()) > 0.0
    del dA, db, r
    torch.cuda.empty_cache()


# ----------------------------------------------------------------------------------------------------------- lupp
@pytest.mark.parametrize("n", [1, 2, 3, 7, 64, 257, 500, 1500])
def test_lupp_is_bit_identical_to_the_oracle(rb, orc, n):
    """src/pivot_decompositions.rs:21-86: same pivots, and -- the elimination being kept operation for operation -- the same
    bits in L and U; test_lupp's own assertions (:351-369) on top"""
    from randnla_b200 import pivot_decompositions as pd
    A = random_matrix(n, n, seed=n + 1)
    l, u, p = pd.lupp(A)
    Lo, Uo, po = orc.lupp(A)
    assert p == po.tolist()
    assert np.array_equal(l, Lo) and np.array_equal(u, Uo)
    assert np.abs(l @ u - A[p, :]).max() <= 1e-10 * n * max(1.0, np.abs(A).max())
    assert not np.triu(l, 1).any() and not np.tril(u, -1).any()


def test_lupp_errors_ties_and_structured_inputs(rb, orc):
    from randnla_b200 import pivot_decompositions as pd
    from randnla_b200.errors import NotSquare, SingularMatrix
    with pytest.raises(NotSquare, match="Matrix must be square, found matrix with 3 rows and 4 columns"):
        pd.lupp(np.zeros((3, 4), order="F"))
    with pytest.raises(SingularMatrix, match="Matrix must be nonsingular for an LU decomposition"):
        pd.lupp(np.asfortranarray(np.array([[0.0, 1.0, 2.0], [0.0, 3.0, 4.0], [0.0, 5.0, 6.0]])))
    l, u, p = pd.lupp(np.asfortranarray(np.array([[1.0, 2.0], [2.0, 4.0]])))          # the last diagonal entry is never examined (:32)
    assert u[1, 1] == 0.0 and p == [1, 0]
    # ties: rows of equal magnitude, the first wins; a matrix that needs no pivoting; one that pivots at every step
    for A in (np.array([[2.0, 1.0, 0.0], [-2.0, 0.0, 1.0], [2.0, 3.0, 5.0]]), np.eye(9) * 3 + np.triu(np.ones((9, 9))),
              np.flipud(np.eye(33)) + 1e-3 * random_matrix(33, 33, seed=2), rank_k_matrix(40, 40, 12, seed=3) + 1e-9 * np.eye(40)):
        A = np.asfortranarray(A)
        l, u, p = pd.lupp(A)
        Lo, Uo, po = orc.lupp(A)
        assert p == po.tolist() and np.array_equal(l, Lo) and np.array_equal(u, Uo)


# ------------------------------------------------------------------------------------------------------ src/cg.rs
def _spd(n, cond, seed):
    rng = np.random.default_rng(seed)
    V, _ = np.linalg.qr(rng.standard_normal((n, n)))
    return np.asfortranarray((V * np.logspace(0, -np.log10(cond), n)) @ V.T), rng.standard_normal(n)


def test_cg_reference_cases(rb, orc):
    """src/cg.rs:130-197: the three cgls cases and test_conjugate_gradient, same assertions"""
    from randnla_b200 import cg
    a = np.asfortranarray(np.array([[4.0, 1.0, 2.0], [1.0, 3.0, 0.0], [2.0, 0.0, 1.0]]))
    b = np.array([4.0, 2.0, 2.0])
    assert np.linalg.norm(cg.cgls(a, b, 3.0, 100, None)) < 3.0
    assert np.linalg.norm(cg.cgls(a, b, 3.0, 100, np.ones(3))) < 3.0
    info = {}
    assert np.linalg.norm(cg.cgls(a, b, 1e-20, 1, None, info=info)) > 1e-20 and not info["converged"] and info["iterations"] == 1
    a2 = np.asfortranarray(np.array([[4.0, 1.0, 2.0], [1.0, 3.0, 1.0], [2.0, 1.0, 3.0]]))
    b2 = np.array([1.0, 2.0, 3.0])
    x = cg.conjugate_grad(a2, b2, np.ones(3))
    assert cg.verify_solution(a2, b2, x) < 1e-10
    from randnla_b200.errors import NotPositiveSemiDefinite
    with pytest.raises(NotPositiveSemiDefinite, match="Matrix is not positive semi-definite"):
        cg.conjugate_grad(np.asfortranarray(np.diag([1.0, -2.0, 3.0])), np.ones(3))


@pytest.mark.parametrize("m,n,cond", [(200, 100, 10.0), (3001, 257, 30.0), (20000, 300, 100.0)])
def test_cgls_matches_oracle(rb, orc, m, n, cond):
    """plain cgls (src/cg.rs:18-61) on a dense system, with and without an initial guess: same iteration count (within 5 %: the
    stopping test sits on its threshold after hundreds of steps), same solution to 1e-6; five steps agree to rounding"""
    from randnla_b200 import cg
    A, b = _lsq_problem(m, n, cond, seed=m + n)
    for x0 in (None, np.random.default_rng(3).standard_normal(n)):
        info = {}
        x = cg.cgls(A, b, 1e-9, 4 * n, x0, info=info)
        xo, ito, convo = orc.cgls(A, b, 1e-9, 4 * n, x0)
        assert info["converged"] == convo and abs(info["iterations"] - ito) <= max(1, ito // 20)
        assert np.abs(x - xo).max() <= 1e-6 * np.abs(xo).max()
        assert abs(cg.verify_solution(A, b, x) - orc.verify_solution(A, b, xo)) <= 1e-9 * np.linalg.norm(b)
    # a fixed number of steps, no convergence: iterates agree to rounding
    info = {}
    x = cg.cgls(A, b, 1e-300, 5, None, info=info)
    xo, ito, convo = orc.cgls(A, b, 1e-300, 5)
    assert info["iterations"] == ito == 5 and not info["converged"] and np.abs(x - xo).max() <= 1e-12 * np.abs(xo).max()


@pytest.mark.parametrize("n,cond", [(50, 10.0), (300, 1e3), (600, 50.0)])
def test_conjugate_grad_matches_oracle(rb, orc, n, cond):
    """conjugate_grad (src/cg.rs:77-112), default start (ones) and a given start; n = 600 is beyond the size the PSD check
    is run for on the device (the oracle always runs it: an O(n^3) Jacobi eigen-decomposition, 5 s at n = 600)"""
    from randnla_b200 import cg
    A, b = _spd(n, cond, seed=n)
    for x0 in (None, np.random.default_rng(4).standard_normal(n)):
        info = {}
        x = cg.conjugate_grad(A, b, x0, info=info)
        xo, ito, convo = orc.conjugate_grad(A, b, x0)
        assert info["converged"] == convo and abs(info["iterations"] - ito) <= max(1, ito // 20)
        # both stop at ||r|| < 1e-5, i.e. within cond * 1e-5 of the solution (lambda_max = 1); long runs decorrelate in rounding
        assert np.linalg.norm(x - xo) <= 2.5e-5 * cond
        if info["iterations"] == ito and ito <= 40:
            assert np.abs(x - xo).max() <= 1e-9 * max(1.0, np.abs(xo).max())
        assert cg.verify_solution(A, b, x) < 1e-5


# ------------------------------------------------------------------------ both products of an iteration from one pass over A
def _normal_pass(A, x, cq, y, cy, store):
    """rnla_normal_pass_dev on host operands: returns t (n), u . u, u (or None)"""
    import ctypes as C
    import torch
    from randnla_b200 import runtime as rt, _lib
    lib = _lib.load()
    m, n = A.shape
    dA = rt.to_device_colmajor(A)
    pA, lda = rt.dev_ptr_ld(dA)
    if not lib.rnla_normal_pass_supported(pA, lda, m, n):
        return None
    dx = torch.from_numpy(np.ascontiguousarray(x)).cuda()
    dy = torch.from_numpy(np.ascontiguousarray(y)).cuda() if y is not None else None
    dt = torch.empty(n + 1, dtype=torch.float64, device="cuda")
    du = torch.empty(m, dtype=torch.float64, device="cuda") if store else None
    torch.cuda.synchronize()
    P = lambda t: C.c_void_p(t.data_ptr()) if t is not None else None
    _lib.check(lib.rnla_normal_pass_dev(pA, lda, m, n, P(dx), cq, P(dy), cy, P(du), P(dt)))
    rt.synchronize()
    t = dt.cpu().numpy()
    return t[:n], t[n], (du.cpu().numpy() if store else None)


@pytest.mark.parametrize("m,n", [(2, 1), (32, 8), (64, 9), (1000, 100), (4100, 255), (4096, 256), (5000, 257), (7778, 500), (3002, 1000),
                                 (9000, 2000), (2500, 2048), (100000, 640),
                                 (1, 1), (31, 7), (33, 9), (4099, 255), (4097, 256), (7777, 500), (3001, 1000), (9001, 2000), (2501, 2048), (60001, 640)])
def test_normal_pass_matches_numpy(rb, m, n):
    """csrc/normal_pass.cu: u = cq A x + cy y, t = A^T u, u . u from one pass over A (clusters of 1, 2, 4, 8 CTAs; ragged last slab
    and ragged last column block; odd leading dimensions: even and odd columns through a tensor map each), against numpy;
    bit-reproducible from call to call."""
    rng = np.random.default_rng(m + n)
    A = np.asfortranarray(rng.standard_normal((m, n)))
    x = rng.standard_normal(n); y = rng.standard_normal(m)
    for cq, yy, cy, store in [(1.0, None, 0.0, False), (-1.0, y, 1.0, True), (1.0, y, -0.37, True)]:
        got = _normal_pass(A, x, cq, yy, cy, store)
        assert got is not None, "n <= 2048 must be taken by the one-pass kernel"
        t, uu, u = got
        ur = cq * (A @ x) + (cy * yy if yy is not None else 0.0)
        assert np.abs(t - A.T @ ur).max() <= 1e-13 * (np.abs(A).T @ np.abs(ur)).max()
        assert abs(uu - ur @ ur) <= 1e-13 * (ur @ ur)
        if store:
            assert np.abs(u - ur).max() <= 1e-13 * ((np.abs(A) @ np.abs(x)).max() + np.abs(y).max())
        t2, uu2, _ = _normal_pass(A, x, cq, yy, cy, store)
        assert np.array_equal(t, t2) and uu == uu2


def test_normal_pass_declines_more_than_2048_columns(rb):
    """n > 8 x 256 (portable cluster size): the solvers keep the two streaming kernels"""
    rng = np.random.default_rng(0)
    assert _normal_pass(np.asfortranarray(rng.standard_normal((64, 2050))), np.ones(2050), 1.0, None, 0.0, False) is None


@pytest.mark.parametrize("m,n,ld", [(5000, 300, 5002), (5000, 301, 5003), (4999, 1000, 5000), (777, 20, 779)])
def test_normal_pass_on_a_view_that_starts_8_bytes_off(rb, m, n, ld):
    """a row block of a larger matrix: the base is only 8-byte aligned (tensor maps want 16) and the leading dimension is that of the
    parent, even or odd -- addressed through views whose base sits one element earlier"""
    import ctypes as C
    import torch
    from randnla_b200 import runtime as rt, _lib
    lib = _lib.load()
    rng = np.random.default_rng(m + n)
    parent = np.asfortranarray(rng.standard_normal((ld, n)))
    dP = rt.to_device_colmajor(parent)
    pP, ldp = rt.dev_ptr_ld(dP)
    assert ldp == ld
    base = dP.data_ptr() + 8                                            # row 1 of the parent
    A = parent[1:1 + m]
    x = rng.standard_normal(n); y = rng.standard_normal(m)
    dx, dy = torch.from_numpy(x).cuda(), torch.from_numpy(y).cuda()
    dt = torch.empty(n + 1, dtype=torch.float64, device="cuda"); du = torch.empty(m, dtype=torch.float64, device="cuda")
    torch.cuda.synchronize()
    assert lib.rnla_normal_pass_supported(C.c_void_p(base), ld, m, n)
    _lib.check(lib.rnla_normal_pass_dev(C.c_void_p(base), ld, m, n, C.c_void_p(dx.data_ptr()), -1.0, C.c_void_p(dy.data_ptr()), 1.0,
                                        C.c_void_p(du.data_ptr()), C.c_void_p(dt.data_ptr())))
    rt.synchronize()
    ur = y - A @ x
    t = dt.cpu().numpy()
    assert np.abs(du.cpu().numpy() - ur).max() <= 1e-13 * ((np.abs(A) @ np.abs(x)).max() + np.abs(y).max())
    assert np.abs(t[:n] - A.T @ ur).max() <= 1e-13 * (np.abs(A).T @ np.abs(ur)).max()
    assert abs(t[n] - ur @ ur) <= 1e-13 * (ur @ ur)


@pytest.mark.parametrize("m,n,cond", [(6000, 40, 1e2), (9000, 300, 1e3), (20000, 1200, 1e2)])
def test_one_pass_cgls_equals_the_two_pass_iteration(rb, orc, m, n, cond, monkeypatch):
    """blendenpik with the CGLS iteration on one pass over A (s_new = s - alpha a^T (a p), csrc/solve.cu) against the reference's
    recurrence on the two streaming kernels (RNLA_ONEPASS=0): same iteration count, same solution to rounding; and against the
    oracle (src/sketch_and_precondition.rs:26-59 with src/cg.rs:18-61)."""
    from randnla_b200 import sketch_and_precondition as sp
    rng = np.random.default_rng(n)
    U, _ = np.linalg.qr(rng.standard_normal((m, n))); V, _ = np.linalg.qr(rng.standard_normal((n, n)))
    A = np.asfortranarray((U * np.logspace(0, -np.log10(cond), n)) @ V.T)
    b = A @ rng.uniform(-100, 100, (n, 1)) + 1e-2 * rng.standard_normal((m, 1))
    res = {}
    for mode in ("0", "1"):
        monkeypatch.setenv("RNLA_ONEPASS", mode)
        info = {}
        res[mode] = (sp.blendenpik_overdetermined(A, b, 1e-9, 200, 4.0, kind=2, zeta=8, info=info), info)
    (x2, i2), (x1, i1) = res["0"], res["1"]
    assert i1["converged"] and i2["converged"] and abs(i1["iterations"] - i2["iterations"]) <= 1
    assert np.linalg.norm(x1 - x2) <= 1e-10 * np.linalg.norm(x2)
    xo, ito, convo = orc.blendenpik(A, b, 1e-9, 200, 4.0, kind=2, zeta=8)
    assert convo and abs(i1["iterations"] - ito) <= 2 and np.linalg.norm(x1 - xo) <= 1e-8 * np.linalg.norm(xo)


@pytest.mark.parametrize("m,n", [(4000, 120), (30000, 700)])
def test_one_pass_lsqr_equals_the_two_pass_iteration(rb, orc, m, n, monkeypatch):
    """lsqr (src/solvers.rs:115-278) with u~ = a v - alfa u and a^T u~ from one pass over a, against the two streaming kernels and the
    oracle: six iterations with the stopping tests off agree to rounding."""
    from randnla_b200 import solvers
    A, b = _lsq_problem(m, n, 50.0, seed=m)
    outs = {}
    for mode in ("0", "1"):
        monkeypatch.setenv("RNLA_ONEPASS", mode)
        outs[mode] = solvers.lsqr(A, b, 0.0, 0.0, 0.0, 0.0, 6, True, None)
    o = orc.lsqr(A, b, 0.0, 0.0, 0.0, 0.0, 6, True, None)
    for other in (outs["0"], o):
        g = outs["1"]
        assert np.abs(g[0] - other[0]).max() <= 1e-11 * np.abs(other[0]).max()
        assert g[1] == other[1] and g[2] == other[2]
        for k in (3, 4, 5, 6, 8):
            assert abs(g[k] - other[k]) <= 1e-10 * abs(other[k])
        assert np.abs(np.asarray(g[7]) - np.asarray(other[7])).max() <= 1e-10 * np.abs(np.asarray(other[7])).max()
        assert np.abs(g[9] - other[9]).max() <= 1e-10 * np.abs(other[9]).max()


# -------------------------------------------------------------------------------------------- committed fixtures
def test_next_rows_golden_vectors(rb):
    """tests/golden/next_rows_golden.npz: oracle outputs committed with their generating script; the CUDA path reproduces
    them without the oracle in the loop"""
    import os
    from randnla_b200 import pivot_decompositions as pd, cqrrpt, sketch_and_solve as ss, id as rid, sketch_and_precondition as sp
    from randnla_b200 import runtime as rt
    from randnla_b200.sketch import MatrixAttribute as MA
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "next_rows_golden.npz"))
    F = np.asfortranarray
    q, r, p = pd.qrcp(F(g["qrcp_A"]))
    assert p == g["qrcp_p"].tolist() and np.abs(r - g["qrcp_R"]).max() < 1e-13 and np.abs(q - g["qrcp_Q"]).max() < 1e-13
    q, r, j = cqrrpt.sap_chol_qrcp(F(g["cq_A"]), int(g["cq_d"]))
    assert j == g["cq_J"].tolist() and np.abs(r - g["cq_R"]).max() < 1e-11 and np.abs(q - g["cq_Q"]).max() < 1e-12
    assert np.abs(ss.sketched_least_squares_qr(F(g["sas_A"]), F(g["sas_b"])) - g["sas_x_qr"]).max() < 1e-11
    assert np.abs(ss.sketched_least_squares_svd(F(g["sas_A"]), F(g["sas_b"])) - g["sas_x_svd"]).max() < 1e-11
    M = F(g["id_A"]); k = int(g["id_k"])
    x, j = rid.osid_qrcp(M, k, MA.Column)
    assert j == g["id_col_J"].tolist() and np.abs(x - g["id_col_X"]).max() < 1e-10
    x, i = rid.osid_qrcp(M, k, MA.Row)
    assert i == g["id_row_I"].tolist() and np.abs(x - g["id_row_X"]).max() < 1e-10
    x, j = rid.osid_randomised(M, k, MA.Column)
    assert j == g["id_rand_J"].tolist() and np.abs(x - g["id_rand_X"]).max() < 1e-8
    for name, fn in (("det", rid.cur), ("rand", rid.cur_randomised)):
        j, u, i = fn(M, k)
        assert j == g[f"cur_{name}_J"].tolist() and i == g[f"cur_{name}_I"].tolist()
        assert np.abs(u - g[f"cur_{name}_U"]).max() < 1e-7 * max(1.0, np.abs(g[f"cur_{name}_U"]).max())
    with rt.options(mode=rt.MODE_LITERAL):
        for name, fn in (("det", rid.two_sided_id), ("rand", rid.two_sided_id_randomised)):
            z, i, j, x = fn(M, k)
            assert i == g[f"tsid_{name}_I"].tolist() and j == g[f"tsid_{name}_J"].tolist()
            assert np.abs(z - g[f"tsid_{name}_Z"]).max() < 1e-7 * max(1.0, np.abs(g[f"tsid_{name}_Z"]).max())
    for mu, key in ((0.0, "sp_x_mu0"), (2.0, "sp_x_mu2")):
        x, y = sp.sketch_saddle_point_precondition(F(g["sp_A"]), F(g["sp_b"]), F(g["sp_c"]), mu, 1e-12, 200, 2.0)
        assert np.abs(x - g[key]).max() < 1e-10


def test_solvers_golden_vectors(rb):
    """tests/golden/solvers_golden.npz: scipy's lsqr (six iterations, stopping tests off), LAPACK solutions and pivots, the
    oracle's bit-exact LU factors -- reproduced by the CUDA path without the oracle in the loop"""
    import os
    from randnla_b200 import solvers, cg, pivot_decompositions as pd
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "solvers_golden.npz"))
    A, b = np.asfortranarray(g["lsqr_A"]), g["lsqr_b"]
    for tag, damp, x0 in (("plain", 0.0, None), ("damped_x0", 0.3, g["lsqr_x0"])):
        x, istop, itn, r1, r2, an, ac, hist, xn, var = solvers.lsqr(A, b, damp, 0.0, 0.0, 0.0, 6, True, x0)
        sc = g[f"lsqr_{tag}_scalars"]
        assert (istop, itn) == (int(sc[0]), int(sc[1]))
        assert np.abs(x[:, 0] - g[f"lsqr_{tag}_x"]).max() <= 1e-11 * np.abs(g[f"lsqr_{tag}_x"]).max()
        assert np.allclose([r1, r2, an, ac, xn], sc[2:], rtol=1e-11, atol=0)
        assert np.allclose(var, g[f"lsqr_{tag}_var"], rtol=1e-10, atol=0)
    x, *_ = solvers.lsqr(A, b, 0.0, 1e-14, 1e-14, 1e8, None, False, None)
    assert np.abs(x[:, 0] - g["lsqr_lstsq_x"]).max() <= 1e-9 * np.abs(g["lsqr_lstsq_x"]).max()
    info = {}
    xs = cg.cgls(A, b, 1e-11, 500, None, info=info)
    assert abs(info["iterations"] - int(g["cgls_iterations"][0])) <= 1 and info["converged"] and np.abs(xs[:, 0] - g["lsqr_lstsq_x"]).max() <= 1e-9
    info = {}
    xo = cg.conjugate_grad(np.asfortranarray(g["cg_A"]), g["cg_b"], None, info=info)
    assert abs(info["iterations"] - int(g["cg_iterations"][0])) <= 1 and info["converged"] and np.abs(xo - g["cg_x_lapack"]).max() < 1e-6
    l, u, p = pd.lupp(np.asfortranarray(g["lupp_A"]))
    assert np.array_equal(l, g["lupp_L"]) and np.array_equal(u, g["lupp_U"]) and p == g["lupp_p"].tolist()
