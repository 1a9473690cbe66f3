"""Pins the oracle's restatement of the rows after the hot path (oracle/oracle_next.c) before the CUDA
path is compared against it (-m "not gpu").  What the reference's own tests hold for these functions are PROPERTIES
(src/pivot_decompositions.rs:320-395, src/cqrrpt.rs:73-133, src/id.rs:329-546 -- orthonormal q, triangular r, exact
reconstruction at 1e-4 / 1e-6, panics on bad k and d; the ID, CUR, sketch-and-solve and saddle-point tests only print), so
they are asserted here on seeded restatements of the same fixtures, and everything LAPACK can decide independently (pivot
order, |diag R|, least-squares solutions, pseudo-inverses) is cross-checked against numpy / scipy."""
import os

import numpy as np
import pytest
import scipy.linalg as sl

from conftest import rank_k_matrix, random_matrix

GOLD = os.path.join(os.path.dirname(__file__), "golden", "next_rows_golden.npz")


def perm_t(p):
    n = len(p)
    P = np.zeros((n, n)); P[np.arange(n), p] = 1.0
    return P


@pytest.mark.parametrize("m,n", [(137, 23), (40, 40), (20, 55), (499, 29)])
def test_qrcp_reference_properties_and_lapack_pivots(orc, m, n):
    """test_qrcp (:320-348); LAPACK dgeqp3 picks the same pivots on generic data (norm downdating vs exact recomputation
    only differ on near ties) and the same |diag R|"""
    A = random_matrix(m, n, seed=m * n)
    Q, R, p = orc.qrcp(A)
    assert np.abs(Q.T @ Q - np.eye(m)).max() < 1e-13
    assert np.abs(np.tril(R, -1)).max() < 1e-13 * np.abs(A).max() * np.sqrt(m)
    assert np.abs(Q @ R @ perm_t(p) - A).max() < 1e-12
    Qs, Rs, ps = sl.qr(A, pivoting=True)
    assert np.array_equal(p, ps)
    k = min(m, n)
    assert np.abs(np.abs(np.diag(R))[:k] - np.abs(np.diag(Rs))[:k]).max() < 1e-12 * np.abs(Rs).max()


def test_economic_qrcp_and_first_maximum_rule(orc):
    """test_qrcp_economical (:372-386); ties go to the first column (strict `>`, :122-127)"""
    A = random_matrix(64, 30, seed=4)
    Q, R, p = orc.economic_qrcp(A, 11)
    assert Q.shape == (64, 11) and R.shape == (11, 30)
    assert np.abs(Q.T @ Q - np.eye(11)).max() < 1e-13
    assert np.abs(np.tril(R[:, :11], -1)).max() < 1e-13
    assert np.abs(Q @ R[:, :11] - A[:, p[:11]]).max() < 1e-12
    T = np.zeros((6, 4)); T[0, 1] = T[1, 2] = T[2, 3] = 2.0; T[3, 0] = 2.0      # four columns of equal norm
    _, _, p = orc.qrcp(T)
    assert p[0] == 0
    T[:, 0] *= 0.5
    _, _, p = orc.qrcp(T)
    assert p[0] == 1
    Z = np.zeros((5, 3))
    Q, R, p = orc.qrcp(Z)
    assert list(p) == [0, 1, 2] and np.array_equal(Q, np.eye(5)) and not R.any()


@pytest.mark.parametrize("kind", [0, 1, 2])
def test_sap_chol_qrcp_reference_properties(orc, kind):
    """test_cqrrpt (:73-112): unit, orthogonal columns of q (1e-6 there), triangular r, q r p^T = a (1e-4 there);
    its two panics (:114-133) come back as code 1"""
    A = np.asfortranarray(np.random.default_rng(7).uniform(-1, 1, (2100, 10)))
    for d in (16, 48, 120):
        Q, R, J = orc.sap_chol_qrcp(A, d, kind=kind)
        assert Q.shape == (2100, 10) and R.shape == (10, 10) and sorted(J) == list(range(10))
        assert np.abs(Q.T @ Q - np.eye(10)).max() < 1e-13
        assert np.abs(np.tril(R, -1)).max() < 1e-12
        assert np.abs(Q @ R @ perm_t(J) - A).max() < 1e-12
        # the QR factorisation of a[:, J] is unique up to the signs of diag(r) (r = r_pre r_sk inherits the signs of the
        # Householder diagonal of the sketch's qrcp): compare with LAPACK's after normalising both
        Ql, Rl = np.linalg.qr(A[:, J]); s = np.sign(np.diag(Rl)) * np.sign(np.diag(R))
        assert np.abs(R - s[:, None] * Rl).max() < 1e-11 and np.abs(Q - Ql * s).max() < 1e-12
    for bad in (9, 2101):
        with pytest.raises(ValueError) as e:
            orc.sap_chol_qrcp(A, bad)
        assert e.value.args[0] == 1
    with pytest.raises(ValueError):
        orc.sap_chol_qrcp(A.T.copy(), 50)
    B = rank_k_matrix(300, 20, 6, seed=2)
    Q, R, J = orc.sap_chol_qrcp(B, 60)
    assert Q.shape == (300, 6) and R.shape == (6, 20)                        # numerical rank of the sketch (:37-43)
    assert np.abs(Q @ R - B[:, J]).max() < 1e-9 * np.abs(B).max()


def test_sketched_least_squares_restatement(orc):
    """test_least_squares_qr / _svd (:81-158): both variants solve the SKETCHED problem exactly -- x = lstsq(S a, S b) with S
    recovered from the oracle's own sketch of the identity -- and land near the full solution"""
    rng = np.random.default_rng(3)
    A = np.asfortranarray(rng.standard_normal((480, 25))); hyp = rng.uniform(-100, 100, (25, 1))
    b = A @ hyp + 0.01 * rng.standard_normal((480, 1))
    d = 480 // 4
    for kind in (0, 2):
        S = orc.sketch_apply_dense(np.eye(480), d) if kind == 0 else orc.sketch_apply_saso_block(np.eye(480), d)
        xs = np.linalg.lstsq(S @ A, S @ b, rcond=None)[0]
        for which in (0, 1):
            x = orc.sketched_least_squares(which, A, b, kind=kind)
            assert np.linalg.norm(x - xs) < 1e-10 * np.linalg.norm(xs)
            assert np.linalg.norm(x - hyp) < 1e-3 * np.linalg.norm(hyp)
    with pytest.raises(ValueError) as e:
        orc.sketched_least_squares(0, A[:60].copy(), b[:60])
    assert e.value.args[0] == 2
    # zero pivot / zero singular value: the unknown stays 0 (src/solvers.rs:29, :62)
    A0 = A[:, :6].copy(); A0[:, 5] = 0.0
    for which in (0, 1):
        x = orc.sketched_least_squares(which, A0, b)
        assert x[5, 0] == 0.0 and np.isfinite(x).all()


@pytest.mark.parametrize("m,n,k", [(104, 107, 60), (70, 200, 20)])
def test_id_restatement(orc, m, n, k):
    """test_one_sided_id / test_two_sided_id / test_cur (:463-546) on rank_k_matrix: every decomposition of an exactly rank-k
    matrix is exact; X restricted to the chosen columns is the identity (:297-301); indices are distinct"""
    A = rank_k_matrix(m, n, k, seed=k)
    nrm = np.linalg.norm(A)
    X, J = orc.osid_qrcp(A, k, orc.COLUMN)
    assert np.array_equal(X[:, J], np.eye(k)) and len(set(J)) == k
    assert np.linalg.norm(A[:, J] @ X - A) < 1e-10 * nrm
    _, _, ps = sl.qr(A, pivoting=True)
    assert list(J) == list(ps[:k])
    X, I = orc.osid_qrcp(A, k, orc.ROW)
    assert X.shape == (m, k) and np.linalg.norm(X @ A[I, :] - A) < 1e-10 * nrm
    X, J = orc.osid_randomised(A, k, orc.COLUMN)
    assert np.linalg.norm(A[:, J] @ X - A) < 1e-8 * nrm
    for rnd in (False, True):
        for mode in (0, 1):
            Z, I, J, X = orc.two_sided_id(A, k, rnd, orc.make_opts(mode=mode))
            assert np.linalg.norm(Z @ A[np.ix_(I, J)] @ X - A) < 1e-7 * nrm
        J, U, I = orc.cur(A, k, rnd)
        assert U.shape == (k, k) and np.linalg.norm(A[:, J] @ U @ A[I, :] - A) < 1e-7 * nrm
    for bad in (0, min(m, n) + 1):                                           # the reference's panics (:329-461)
        for call in (lambda: orc.osid_qrcp(A, bad, orc.COLUMN), lambda: orc.osid_randomised(A, bad, orc.COLUMN),
                     lambda: orc.cur(A, bad), lambda: orc.two_sided_id(A, bad)):
            with pytest.raises(ValueError) as e:
                call()
            assert e.value.args[0] == 1
    with pytest.raises(ValueError) as e:
        orc.osid_randomised(A, k, orc.ROW)                                  # a * tsog1(a, k, 2, 1)^T conforms only if n == k
    assert e.value.args[0] == 2


def test_cur_core_is_the_pseudo_inverse_formula(orc):
    """src/id.rs:49-52: u = x pinv(a[i, :]) -- checked against numpy's pinv with the oracle's own x, i"""
    A = rank_k_matrix(60, 45, 9, seed=1)
    J, U, I = orc.cur(A, 9)
    X, J2 = orc.osid_qrcp(A, 9, orc.COLUMN)
    assert list(J) == list(J2)
    assert np.abs(U - X @ np.linalg.pinv(A[I, :])).max() < 1e-9 * np.abs(U).max()


def test_saddle_point_restatement(orc):
    """test_saddle_point (:278-337).  With mu = 0 the result solves a^T a x = a^T b - c.  With mu > 0 the reference
    preconditions with (sigma^2 + mu)^-1/2 but iterates on a M alone, so its x is still the mu = 0 solution of the modified
    right-hand side -- restated as written, and pinned as such."""
    rng = np.random.default_rng(5)
    A = np.asfortranarray(rng.standard_normal((300, 12))); b = rng.standard_normal((300, 1)); c = rng.uniform(-10, 10, (12, 1))
    x, y, it, conv = orc.saddle_point(A, b, None, 0.0, 1e-12, 300, 1.5)
    assert conv and np.linalg.norm(x - np.linalg.lstsq(A, b, rcond=None)[0]) < 1e-10 * np.linalg.norm(x)
    assert np.linalg.norm(y - (b - A @ x)) < 1e-12 * np.linalg.norm(b)
    x, y, it, conv = orc.saddle_point(A, b, c, 0.0, 1e-12, 300, 1.5)
    assert conv and np.linalg.norm(x - np.linalg.solve(A.T @ A, A.T @ b - c)) < 1e-10 * np.linalg.norm(x)
    x2, _, _, conv = orc.saddle_point(A, b, None, 2.5, 1e-12, 300, 1.5)
    assert conv and np.linalg.norm(x2 - np.linalg.lstsq(A, b, rcond=None)[0]) < 1e-9 * np.linalg.norm(x2)
    for bad, code in (((1e-4, 1000, 0.5), 1), ((-1e-4, 1000, 1.5), 1), ((0.0, 1000, 1.5), 1), ((1e-4, 0, 1.5), 1)):
        with pytest.raises(ValueError) as e:
            orc.saddle_point(A, b, c, 1.0, *bad)
        assert e.value.args[0] == code
    with pytest.raises(ValueError) as e:
        orc.saddle_point(A[:8].copy(), b[:8], c, 1.0, 1e-4, 10, 1.5)
    assert e.value.args[0] == 4


def test_oracle_reproduces_next_rows_golden(orc):
    """tests/golden/next_rows_golden.npz (tests/golden/make_golden_next_rows.py) is what the GPU tests are also held to"""
    g = np.load(GOLD)
    F = np.asfortranarray
    Q, R, p = orc.qrcp(F(g["qrcp_A"]))
    assert np.array_equal(p, g["qrcp_p"]) and np.abs(R - g["qrcp_R"]).max() < 1e-14 and np.abs(Q - g["qrcp_Q"]).max() < 1e-14
    Q, R, J = orc.sap_chol_qrcp(F(g["cq_A"]), int(g["cq_d"]))
    assert np.array_equal(J, g["cq_J"]) and np.abs(R - g["cq_R"]).max() < 1e-13 and np.abs(Q - g["cq_Q"]).max() < 1e-13
    assert np.abs(orc.sketched_least_squares(0, F(g["sas_A"]), F(g["sas_b"])) - g["sas_x_qr"]).max() < 1e-13
    assert np.abs(orc.sketched_least_squares(1, F(g["sas_A"]), F(g["sas_b"])) - g["sas_x_svd"]).max() < 1e-13
    M = F(g["id_A"]); k = int(g["id_k"])
    X, J = orc.osid_qrcp(M, k, orc.COLUMN)
    assert np.array_equal(J, g["id_col_J"]) and np.abs(X - g["id_col_X"]).max() < 1e-12
    J, U, I = orc.cur(M, k, True)
    assert np.array_equal(J, g["cur_rand_J"]) and np.array_equal(I, g["cur_rand_I"]) and np.abs(U - g["cur_rand_U"]).max() < 1e-10
    x = orc.saddle_point(F(g["sp_A"]), F(g["sp_b"]), F(g["sp_c"]), 2.0, 1e-12, 200, 2.0)[0]
    assert np.abs(x - g["sp_x_mu2"]).max() < 1e-12


def test_next_rows_validation_precedes_device_access():
    """the reference's asserts / Errs of these rows are raised by the C ABI before any device work (no GPU here)"""
    import randnla_b200 as rb
    from randnla_b200.errors import InvalidParameters, NotOverdetermined
    from randnla_b200.sketch import MatrixAttribute as MA
    A = random_matrix(100, 10, seed=1); b = random_matrix(100, 1, seed=2); c = random_matrix(10, 1, seed=3)
    with pytest.raises(InvalidParameters) as e:
        rb.cqrrpt.sap_chol_qrcp(A, 5)
    assert str(e.value) == "d must satisfy n ≤ d ≪ m"             # src/cqrrpt.rs:29
    for k, msg in ((0, "k must be positive)"), (11, "k must be <= min(l,w)")):   # src/id.rs:278-279 (typo included)
        for call in (lambda: rb.id.osid_qrcp(A, k, MA.Column), lambda: rb.id.osid_randomised(A, k, MA.Row),
                     lambda: rb.id.cur(A, k), lambda: rb.id.cur_randomised(A, k), lambda: rb.id.two_sided_id(A, k),
                     lambda: rb.id.two_sided_id_randomised(A, k)):
            with pytest.raises(InvalidParameters) as e:
                call()
            assert str(e.value) == msg
    with pytest.raises(InvalidParameters) as e:
        rb.sketch_and_precondition.sketch_saddle_point_precondition(A, b, c, 1.0, 1e-4, 1000, 0.5)
    assert str(e.value) == "Sampling factor must be greater than 1, current input is 0.5"
    with pytest.raises(NotOverdetermined):
        rb.sketch_and_precondition.sketch_saddle_point_precondition(A.T.copy(), c, b, 1.0, 1e-4, 10, 1.5)


# ---- lsqr (src/solvers.rs:115-278): the reference says it is a translation of scipy 1.14.1's lsqr, so scipy's own lsqr is the
# ---- known-answer generator: same iterates, same stopping rule, same estimates on the same inputs
def _lsq_problem(m, n, cond, seed, consistent=False):
    rng = np.random.default_rng(seed)
    U, _ = np.linalg.qr(rng.standard_normal((m, n)))
    V, _ = np.linalg.qr(rng.standard_normal((n, n)))
    A = np.asfortranarray((U * np.logspace(0, -np.log10(cond), n)) @ V.T)
    x = rng.uniform(-100, 100, n)
    b = A @ x + (0.0 if consistent else 1e-2) * rng.standard_normal(m)
    return A, b


@pytest.mark.parametrize("m,n,cond,damp,calc_var,use_x0,consistent", [
    (200, 100, 10.0, 0.0, False, False, False),     # test_overdetermined_system's shape (:413-468)
    (300, 40, 1e3, 0.0, True, False, False),
    (300, 40, 1e3, 0.5, True, True, False),
    (120, 120, 50.0, 0.0, False, True, True),
    (500, 30, 1e6, 1e-3, False, False, False),
    (64, 90, 20.0, 0.1, True, False, False),         # underdetermined with damping
])
def test_lsqr_matches_scipy(orc, m, n, cond, damp, calc_var, use_x0, consistent):
    from scipy.sparse.linalg import lsqr as sp_lsqr
    A, b = _lsq_problem(max(m, n), min(m, n), cond, seed=m + n, consistent=consistent)
    if m < n:
        A = np.asfortranarray(A.T); b = np.random.default_rng(5).standard_normal(m)
    x0 = np.random.default_rng(9).standard_normal(A.shape[1]) if use_x0 else None
    # (1) a fixed number of iterations with the stopping tests switched off: every return value agrees to rounding
    # (short: without reorthogonalisation the Golub-Kahan recurrences amplify rounding differences once Ritz values converge --
    # measured here: 1e-15 up to 8 iterations, 1e-8 at 12, 1e-5 at 16 on the cond 1e3 case)
    K = 6
    got = orc.lsqr(A, b, damp=damp, atol=0.0, btol=0.0, conlim=0.0, iter_lim=K, calc_var=calc_var, x0=x0)
    ref = sp_lsqr(A, b, damp=damp, atol=0.0, btol=0.0, conlim=0.0, iter_lim=K, calc_var=calc_var, x0=x0)
    x, istop, itn, r1, r2, an, ac, hist, xn, var = got
    assert istop == ref[1] == 7 and itn == ref[2] == K
    assert np.abs(x[:, 0] - ref[0]).max() <= 1e-10 * np.abs(ref[0]).max()
    for g, r in zip((r1, r2, an, ac, xn), (ref[3], ref[4], ref[5], ref[6], ref[8])):
        assert abs(g - r) <= 1e-10 * abs(r)
    assert len(hist) == itn and np.all(hist >= 0)
    if calc_var:
        assert np.abs(var - ref[9]).max() <= 1e-10 * np.abs(ref[9]).max()
    else:
        assert not var.any()
    # (2) run to convergence: same stopping reason, the iteration count may differ slightly where a test sits on its threshold
    got = orc.lsqr(A, b, damp=damp, atol=1e-8, btol=1e-8, conlim=1e8, iter_lim=None, calc_var=calc_var, x0=x0)
    ref = sp_lsqr(A, b, damp=damp, atol=1e-8, btol=1e-8, conlim=1e8, iter_lim=2 * A.shape[1], calc_var=calc_var, x0=x0)
    assert got[1] == ref[1] and abs(got[2] - ref[2]) <= max(2, ref[2] // 20)
    if cond <= 50.0:                              # x itself is only determined to about cond * atol
        assert np.abs(got[0][:, 0] - ref[0]).max() <= 1e-6 * np.abs(ref[0]).max()
    assert abs(got[3] - ref[3]) <= (1e-6 if cond <= 50.0 else 1e-3) * abs(ref[3]) + 1e-7 * np.linalg.norm(b)


def test_lsqr_reference_cases(orc):
    """test_simple_system (:391-410): b = 0 returns x = 0 at once with the single-entry history [0.0] (:181-183); the 3 x 2
    system with b = (1, 0, -1) has the solution (1, -1) (1e-2 there).  iter_lim is honoured (istop = 7)."""
    A = np.asfortranarray(np.array([[1.0, 0.0], [1.0, 1.0], [0.0, 1.0]]))
    x, istop, itn, r1, r2, an, ac, hist, xn, var = orc.lsqr(A, np.zeros(3), atol=1e-8, btol=1e-8)
    assert not x.any() and istop == 0 and itn == 0 and list(hist) == [0.0] and r1 == 0.0 and an == 0.0
    x, istop, itn, *_ = orc.lsqr(A, np.array([1.0, 0.0, -1.0]), atol=1e-8, btol=1e-8)
    assert np.abs(x[:, 0] - [1.0, -1.0]).max() < 1e-12 and istop in (1, 2) and itn <= 2
    A, b = _lsq_problem(200, 100, 1e4, seed=1)
    x, istop, itn, *_, hist, xn, var = orc.lsqr(A, b, atol=1e-14, btol=1e-14, iter_lim=7)
    assert istop == 7 and itn == 7 and len(hist) == 7
    # the least-squares solution against LAPACK (test_overdetermined_system compares with nalgebra's SVD solve)
    A, b = _lsq_problem(200, 100, 10.0, seed=2)
    x, *_ = orc.lsqr(A, b, atol=1e-12, btol=1e-12)
    xs = np.linalg.lstsq(A, b, rcond=None)[0]
    assert np.abs(x[:, 0] - xs).max() <= 1e-8 * np.abs(xs).max()


# ---- src/cg.rs: conjugate_grad (:77-112) and the reference's own cases (:130-197)
def _spd(n, cond, seed):
    rng = np.random.default_rng(seed)
    V, _ = np.linalg.qr(rng.standard_normal((n, n)))
    return np.asfortranarray((V * np.logspace(0, -np.log10(cond), n)) @ V.T), rng.standard_normal(n)


def test_conjugate_grad_reference_case_and_lapack(orc):
    """test_conjugate_gradient (:187-196): residual below 1e-10 on the 3 x 3 system; LAPACK's solution on larger SPD systems
    (the stopping rule r.r < 1e-10 bounds the residual by 1e-5); NotPositiveSemiDefinite for an indefinite matrix (:80-86);
    the iterates are those of textbook CG (scipy's cg with the same start takes the same steps)."""
    a = np.asfortranarray(np.array([[4.0, 1.0, 2.0], [1.0, 3.0, 1.0], [2.0, 1.0, 3.0]]))
    b = np.array([1.0, 2.0, 3.0])
    x, it, conv = orc.conjugate_grad(a, b, np.ones(3))
    assert conv and it <= 2 and orc.verify_solution(a, b, x) < 1e-10
    for n, cond in ((50, 10.0), (200, 1e3)):
        A, rhs = _spd(n, cond, seed=n)
        x, it, conv = orc.conjugate_grad(A, rhs)
        assert conv and it < 2 * n
        assert np.linalg.norm(A @ x - rhs) < 1e-5
        xs = np.linalg.solve(A, rhs)
        assert np.linalg.norm(x - xs) <= cond * 1e-5
    with pytest.raises(ValueError, match="9"):
        orc.conjugate_grad(np.asfortranarray(np.diag([1.0, -2.0, 3.0])), np.ones(3))
    # a fixed number of steps against scipy's cg from the same starting vector
    from scipy.sparse.linalg import cg as sp_cg
    A, rhs = _spd(60, 30.0, seed=7)
    iterates = []
    sp_cg(A, rhs, x0=np.ones(60), rtol=0.0, atol=0.0, maxiter=5, callback=lambda xk: iterates.append(xk.copy()))
    # the oracle stops on r.r < 1e-10 only, so replay 5 steps by hand with its own update rule
    x = np.ones(60); r = A @ x - rhs; p = -r; rk = r @ r
    for _ in range(5):
        ap = A @ p; alpha = rk / (p @ ap); x = x + alpha * p; r = r + alpha * ap; rk1 = r @ r; p = (rk1 / rk) * p - r; rk = rk1
    assert np.abs(x - iterates[4]).max() <= 1e-12 * np.abs(x).max()


def test_cgls_reference_cases(orc):
    """test_cgls_converges / _with_initial_guess / _does_not_converge (:130-184)"""
    a = np.asfortranarray(np.array([[4.0, 1.0, 2.0], [1.0, 3.0, 0.0], [2.0, 0.0, 1.0]]))
    b = np.array([4.0, 2.0, 2.0])
    x, it, conv = orc.cgls(a, b, 3.0, 100)
    assert conv and np.linalg.norm(x) < 3.0
    x, it, conv = orc.cgls(a, b, 3.0, 100, np.ones(3))
    assert np.linalg.norm(x) < 3.0
    x, it, conv = orc.cgls(a, b, 1e-20, 1)
    assert not conv and it == 1 and np.linalg.norm(x) > 1e-20
    x, it, conv = orc.cgls(a, b, 1e-12, 100)
    assert conv and np.abs(x[:, 0] - np.linalg.solve(a, b)).max() < 1e-10


# ---- lupp (src/pivot_decompositions.rs:21-86)
@pytest.mark.parametrize("n", [1, 2, 7, 64, 500])
def test_lupp_reference_properties_and_lapack(orc, n):
    """test_lupp (:351-369; n = 500 there): triangular factors, P L U = A at 1e-4 (1e-10 here), and agreement with LAPACK's
    dgetrf -- the same pivot rule (first maximum of |.|) gives the same permutation, factors equal to rounding"""
    A = random_matrix(n, n, seed=n)
    L, U, p = orc.lupp(A)
    assert np.array_equal(np.triu(L, 1), np.zeros((n, n))) and np.array_equal(np.diag(L), np.ones(n))
    assert np.array_equal(np.tril(U, -1), np.zeros((n, n)))
    assert np.abs((L @ U) - A[p, :]).max() <= 1e-10 * max(1.0, np.abs(A).max()) * n
    P, Ls, Us = sl.lu(A)
    assert np.array_equal(P.T @ A, A[p, :])
    assert np.abs(L - Ls).max() <= 1e-9 * n and np.abs(U - Us).max() <= 1e-9 * n * np.abs(Us).max()
    assert np.abs(L).max() <= 1.0


def test_lupp_errors_and_ties(orc):
    """NotSquare (:23-27), SingularMatrix on a zero pivot column (:44-48) -- but not for a zero LAST diagonal entry, which the
    loop never examines (:32); ties go to the first row (strict `>`)"""
    with pytest.raises(ValueError, match="5"):
        orc.lupp(np.zeros((3, 4)))
    with pytest.raises(ValueError, match="6"):
        orc.lupp(np.asfortranarray(np.array([[0.0, 1.0, 2.0], [0.0, 3.0, 4.0], [0.0, 5.0, 6.0]])))
    L, U, p = orc.lupp(np.asfortranarray(np.array([[1.0, 2.0], [2.0, 4.0]])))         # singular, but only U[1, 1] shows it
    assert U[1, 1] == 0.0 and list(p) == [1, 0]
    L, U, p = orc.lupp(np.asfortranarray(np.array([[2.0, 1.0, 0.0], [-2.0, 0.0, 1.0], [2.0, 3.0, 5.0]])))
    assert p[0] == 0


def test_solvers_golden_vectors(orc):
    """tests/golden/solvers_golden.npz (scipy lsqr / LAPACK / oracle outputs with their generating script): the oracle
    reproduces them; the same file is the target of the CUDA path in tests/test_gpu_next_rows.py"""
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "solvers_golden.npz"))
    A, b = np.asfortranarray(g["lsqr_A"]), g["lsqr_b"]
    for tag, damp, x0 in (("plain", 0.0, None), ("damped_x0", 0.3, g["lsqr_x0"])):
        x, istop, itn, r1, r2, an, ac, hist, xn, var = orc.lsqr(A, b, damp, 0.0, 0.0, 0.0, 6, True, x0)
        sc = g[f"lsqr_{tag}_scalars"]
        assert (istop, itn) == (int(sc[0]), int(sc[1]))
        assert np.abs(x[:, 0] - g[f"lsqr_{tag}_x"]).max() <= 1e-11 * np.abs(g[f"lsqr_{tag}_x"]).max()
        assert np.allclose([r1, r2, an, ac, xn], sc[2:], rtol=1e-11, atol=0)
        assert np.allclose(var, g[f"lsqr_{tag}_var"], rtol=1e-10, atol=0)
    x, *_ = orc.lsqr(A, b, 0.0, 1e-14, 1e-14, 1e8, None, False, None)
    assert np.abs(x[:, 0] - g["lsqr_lstsq_x"]).max() <= 1e-9 * np.abs(g["lsqr_lstsq_x"]).max()
    xs, it, conv = orc.cgls(A, b, 1e-11, 500)
    assert [it, int(conv)] == g["cgls_iterations"].tolist() and np.abs(xs[:, 0] - g["lsqr_lstsq_x"]).max() <= 1e-9
    xo, ito, convo = orc.conjugate_grad(np.asfortranarray(g["cg_A"]), g["cg_b"])
    assert [ito, int(convo)] == g["cg_iterations"].tolist() and np.abs(xo - g["cg_x_lapack"]).max() < 1e-6
    L, U, p = orc.lupp(np.asfortranarray(g["lupp_A"]))
    assert np.array_equal(L, g["lupp_L"]) and np.array_equal(U, g["lupp_U"]) and np.array_equal(p, g["lupp_p"])
    assert np.abs(L - g["lupp_L_lapack"]).max() < 1e-12 and np.abs(U - g["lupp_U_lapack"]).max() < 1e-12
