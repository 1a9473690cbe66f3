"""Pins the CPU oracle (oracle/) before anything is compared against it (-m "not gpu").

What the reference's own tests hold for this path:
  bit-exact  Philox / ThreeFry known-answer vectors, the seeding KAT
  exact      Orth(0)=I, Orth(I)=I, Stabilizer(0)=I, rand_svd(0), rand_evd1(0), rand_evd2(0) is Err
  numeric    rand_evd2 on a 5x5 PSD matrix vs the deterministic eigen-decomposition (1e-6)
  property   orthonormality, ordering, shapes
Everything nalgebra-specific that is not in the tree is cross-checked against numpy/LAPACK here.
"""
import json
import os

import numpy as np
import pytest

from conftest import rank_k_matrix, random_matrix, random_hermitian, random_psd, lowrank_plus_noise, subspace_angle

GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def kats():
    with open(os.path.join(GOLD, "reference_kats.json")) as f:
        return json.load(f)


# ---------------------------------------------------------------- L0 known-answer tests
def test_philox4x32_kat(orc, kats):
    """rust-random123/src/philox.rs:330-345 `exact_values_philox_4x32`"""
    k = kats["philox4x32_10"]
    key = np.array([[int(x, 16) for x in k["key"]]], dtype=np.uint32)
    ctr = np.zeros((10, 4), dtype=np.uint32); ctr[:, 0] = np.arange(10)
    out = orc.philox4x32_10(ctr, key)
    exp = np.array([[int(x, 16) for x in row] for row in k["out"]], dtype=np.uint32)
    assert (out == exp).all()


def test_threefry2x64_kat(orc, kats):
    """rust-random123/src/threefry.rs:113-127 `exact_values`"""
    k = kats["threefry2x64_20"]
    key = np.array([[int(x, 16) for x in k["key"]]], dtype=np.uint64)
    ctr = np.zeros((10, 2), dtype=np.uint64); ctr[:, 0] = np.arange(10)
    out = orc.threefry2x64_20(ctr, key)
    exp = np.array([[int(x, 16) for x in row] for row in k["out"]], dtype=np.uint64)
    assert (out == exp).all()


def test_seed_from_u64_kat(orc, kats):
    """rust-random123/src/threefry.rs:138-142 (disabled `seedable` test): seed 42 -> first u64"""
    k = kats["seed_from_u64"]
    assert orc.threefry_rng_u64(k["seed"], 0) == int(k["first_u64"])
    # the seed every sketching_operator call uses (src/sketch.rs:112); SURVEY.md Appendix B.1
    key = orc.seed_from_u64(0)
    assert [int(x) for x in key] == [0x45cdb581f973f2ec, 0xad6cad067346f087]
    assert orc.threefry_rng_u64(0, 0) == 142907558101433000
    assert orc.threefry_rng_u64(0, 1) == 18212790059499954013


def test_philox_counter_carry_and_pure_python(orc):
    """independent pure-Python Philox4x32-10 (rust-random123/src/philox.rs:149-154,173-176,211-223)"""
    def ref(ctr, key):
        c = list(ctr); k = list(key)
        for r in range(10):
            if r:
                k = [(k[0] + 0x9E3779B9) & 0xffffffff, (k[1] + 0xBB67AE85) & 0xffffffff]
            p0 = 0xD2511F53 * c[0]; p1 = 0xCD9E8D57 * c[2]
            c = [(p1 >> 32) ^ c[1] ^ k[0], p1 & 0xffffffff, (p0 >> 32) ^ c[3] ^ k[1], p0 & 0xffffffff]
        return c
    rng = np.random.default_rng(5)
    ctr = rng.integers(0, 2**32, size=(64, 4), dtype=np.uint64).astype(np.uint32)
    key = rng.integers(0, 2**32, size=(64, 2), dtype=np.uint64).astype(np.uint32)
    out = orc.philox4x32_10(ctr, key)
    for i in range(64):
        assert [int(x) for x in out[i]] == ref([int(x) for x in ctr[i]], [int(x) for x in key[i]])


# ---------------------------------------------------------------- L1 sketch operators
def test_reference_uniform_and_rademacher_stream(orc):
    """src/sketch.rs:118-126 over the seed-0 ThreeFry stream, rand 0.8.5 transforms, column-major fill order"""
    u = [orc.threefry_rng_u64(0, t) for t in range(12)]
    U = orc.sketching_operator_ref(1, 3, 4)
    R = orc.sketching_operator_ref(2, 3, 4)
    for t in range(12):
        v12 = np.frombuffer(np.uint64((u[t] >> 12) | 0x3FF0000000000000).tobytes(), dtype=np.float64)[0]
        assert U[t % 3, t // 3] == (v12 - 1.0) * 2.0 + (-1.0)
        assert R[t % 3, t // 3] == (1.0 if u[t] < 2**63 else -1.0)
    big = orc.sketching_operator_ref(1, 400, 50)
    assert -1.0 <= big.min() and big.max() < 1.0 and abs(big.mean()) < 0.02 and abs(big.std() - 1 / np.sqrt(3)) < 0.01
    with pytest.raises(NotImplementedError):
        orc.sketching_operator_ref(0, 3, 4)     # Gaussian: ziggurat tables are not in the tree -> parity unpinned


def test_omega_map_is_pure_function_of_indices(orc):
    full = orc.omega_fill(0, 37, 9, seed=11, stream=1)
    part = orc.omega_fill(0, 20, 9, seed=11, stream=1, row_off=13)
    assert (part == full[13:33]).all()
    assert not (orc.omega_fill(0, 37, 9, seed=12, stream=1) == full).any()
    assert not (orc.omega_fill(0, 37, 9, seed=11, stream=2) == full).any()


def test_omega_gaussian_matches_inverse_cdf(orc):
    """T_gauss is the normal quantile of ((k & 0x7fffffff) + 1/2) / 2^31 read from the tail, to FP32 accuracy"""
    from scipy.special import ndtri
    ks = np.concatenate([np.arange(0, 2**31, 2**31 // 2000, dtype=np.uint64), [1, 5, 2**31 - 1, 2**31 - 300]]).astype(np.uint64)
    for k in ks:
        for sign_bit in (0, 1):
            kk = int(k) | (sign_bit << 31)
            v = (int(k) + 0.5) / 2**31
            exact = -ndtri(v / 2.0) * (-1.0 if sign_bit else 1.0)
            got = orc.gauss_from_u32(kk)
            tol = 4e-4 * abs(exact) if v < 2**-23 else max(3e-6 * abs(exact), 1.5e-7)
            assert abs(got - exact) <= tol, (kk, got, exact)


def test_omega_moments(orc):
    g = orc.omega_fill(0, 4000, 64, seed=1)
    assert abs(g.mean()) < 0.01 and abs(g.std() - 1.0) < 0.01 and abs((g**4).mean() - 3.0) < 0.1
    u = orc.omega_fill(1, 4000, 64, seed=1)
    assert -1 < u.min() and u.max() < 1 and abs(u.mean()) < 0.01 and abs(u.var() - 1 / 3) < 0.01
    r = orc.omega_fill(2, 4000, 64, seed=1)
    assert set(np.unique(r)) == {-1.0, 1.0} and abs(r.mean()) < 0.01


# ---------------------------------------------------------------- L2 nalgebra conventions vs LAPACK
@pytest.mark.parametrize("shape", [(50, 8), (8, 8), (9, 20), (300, 60)])
def test_qr_convention(orc, shape):
    """nalgebra qr(): X = Q R, Q^T Q = I, R upper with non-negative diagonal -> equals LAPACK's Q up to column signs"""
    X = random_matrix(*shape, seed=3)
    Q, R = orc.qr(X)
    p = min(shape)
    assert np.abs(Q @ R - X).max() < 1e-13 * np.abs(X).max() * shape[0]
    assert np.abs(Q.T @ Q - np.eye(p)).max() < 1e-14 * shape[0]
    assert np.abs(np.tril(R[:, :p], -1)).max() == 0 and (np.diag(R) >= 0).all()
    Qn, Rn = np.linalg.qr(X)
    sg = np.sign(np.diag(Rn)); sg[sg == 0] = 1
    assert np.abs(Q - Qn[:, :p] * sg).max() < 1e-12


def test_orth_exact_cases(orc):
    """src/lora_helpers.rs:324-340 test_orth_zero_matrix / test_orth_identity_matrix"""
    assert (orc.Orth(np.zeros((5, 5))) == np.eye(5)).all()
    assert np.abs(orc.Orth(np.eye(5)) - np.eye(5)).max() < 1e-15
    Q = orc.Orth(np.zeros((7, 3)))
    assert (Q == np.eye(7, 3)).all()


def _fullpiv_l_numpy(X):
    """independent numpy restatement of nalgebra FullPivLU::new + .l() (first max in column-major order)"""
    W = np.array(X, dtype=np.float64, order="F")
    rows, cols = W.shape
    mn = min(rows, cols)
    for i in range(mn):
        sub = np.abs(W[i:, i:])
        flat = np.argmax(sub.flatten(order="F"))
        cp, rp = i + flat // (rows - i), i + flat % (rows - i)
        d = W[rp, cp]
        if d == 0:
            break
        W[:, [i, cp]] = W[:, [cp, i]]
        W[[i, rp], :] = W[[rp, i], :]
        inv = 1.0 / d
        W[i + 1:, i] = W[i + 1:, i] * inv
        for c in range(i + 1, cols):
            W[i + 1:, c] = (-W[i, c]) * W[i + 1:, i] + W[i + 1:, c]
    L = np.tril(W[:, :mn], -1)
    L[np.arange(mn), np.arange(mn)] = 1.0
    return L


def test_haar_sample_restatement(orc):
    """orc_haar_sample follows src/sketch.rs:45-85: Q of the Householder QR of a Gaussian matrix, columns multiplied by
    signum(R_ii), transposed for Row.  Pinned on the reference's own tests (`test_row_attribute` / `test_column_attribute`,
    :140-214: shapes, Q Q^T = I or Q^T Q = I to 1e-6; the dimension errors of :49-63) and against LAPACK: the same Gaussian matrix
    through numpy's QR, sign-fixed the same way, gives the same Q to rounding."""
    from oracle.oracle import OracleError
    for rows, cols, attr in [(3, 6, 0), (6, 3, 1), (4, 4, 0), (4, 4, 1), (50, 200, 0), (300, 40, 1), (1, 5, 0), (5, 1, 1)]:
        Q = orc.haar_sample(rows, cols, attr, seed=7)
        assert Q.shape == (rows, cols)
        G = orc.omega_fill(0, max(rows, cols), min(rows, cols), seed=7, stream=0)      # m x n with m >= n, column-major fill (:68-72)
        Qn, Rn = np.linalg.qr(G)
        Qn = Qn * np.sign(np.diag(Rn))                  # the factorisation with R_ii > 0: nalgebra's Q (R_ii >= 0), whose signum fix is +1
        want = Qn.T if attr == 0 else Qn
        assert np.abs(Q - want).max() < 1e-12
        eye = np.eye(min(rows, cols))
        assert np.abs((Q @ Q.T if attr == 0 else Q.T @ Q) - eye).max() < 1e-13
    for rows, cols, attr in [(5, 3, 0), (3, 5, 1)]:
        with pytest.raises(OracleError):
            orc.haar_sample(rows, cols, attr)


@pytest.mark.parametrize("shape", [(30, 6), (6, 6), (5, 9), (64, 17)])
def test_stabilizer_matches_numpy_restatement(orc, shape):
    X = random_matrix(*shape, seed=7)
    L = orc.Stabilizer(X)
    assert (L == _fullpiv_l_numpy(X)).all()
    # it really is the L factor of a full-pivot LU: |L_ij| <= 1, unit diagonal
    assert np.abs(L).max() <= 1.0 and (np.diag(L) == 1).all()


def test_stabilizer_exact_cases(orc):
    """src/lora_helpers.rs:358-365 test_stabilizer_zero_matrix"""
    assert (orc.Stabilizer(np.zeros((5, 5))) == np.eye(5)).all()
    assert (orc.Stabilizer(np.zeros((8, 3))) == np.eye(8, 3)).all()
    X = random_matrix(20, 5, seed=1)
    assert orc.Stabilizer(X).shape == (20, 5)          # :342-356 shape tests
    assert orc.Stabilizer(X.T.copy()).shape == (5, 5)


def test_svd_and_eigen_vs_lapack(orc):
    for shape in [(40, 12), (12, 40), (9, 9)]:
        M = random_matrix(*shape, seed=2)
        U, s, Vt = orc.svd(M)
        assert np.abs(s - np.linalg.svd(M, compute_uv=False)).max() < 1e-13 * s[0]
        assert np.abs((U * s) @ Vt - M).max() < 1e-13 * s[0]
        assert (np.diff(s) <= 0).all()
    H = random_hermitian(30, seed=4)
    lam, W = orc.symmetric_eigen(H)
    assert np.abs(lam - np.linalg.eigvalsh(H)).max() < 1e-13 * np.abs(lam).max()
    assert np.abs(W @ np.diag(lam) @ W.T - H).max() < 1e-12
    U, s, Vt = orc.svd(np.zeros((4, 6)))
    assert (U == np.eye(4)).all() and (s == 0).all() and (Vt == np.eye(4, 6)).all()


def test_gemm(orc):
    for (m, K, N) in [(257, 129, 13), (64, 512, 6), (1000, 77, 110), (5, 3, 2)]:
        A, B, Q = random_matrix(m, K, 1), random_matrix(K, N, 2), random_matrix(m, N, 3)
        assert np.abs(orc.gemm_nn(A, B) - A @ B).max() < 1e-12 * K
        assert np.abs(orc.gemm_tn(A, Q) - A.T @ Q).max() < 1e-12 * m


# ---------------------------------------------------------------- L3 / L4: the reference's own tests, run on the oracle
def test_tsog1_shapes(orc):
    """src/lora_helpers.rs:160-232: output has k columns and n rows for passes in {3,4}, stab in {1,2,3}"""
    A = random_matrix(20, 12, seed=9)
    for mode in (0, 1):
        for q in (3, 4):
            for pps in (1, 2, 3):
                S = orc.tsog1(A, 5, q, pps, orc.make_opts(mode=mode))
                assert S.shape == (12, 5) and np.isfinite(S).all()


def test_literal_tsog1_quirks(orc):
    """even pass counts ignore Omega; odd >= 3 all give the q=3 result; q=1 returns zeros"""
    A = random_matrix(30, 18, seed=10)
    o1 = orc.make_opts(mode=1, seed=1); o2 = orc.make_opts(mode=1, seed=2)
    assert (orc.tsog1(A, 6, 2, 1, o1) == orc.tsog1(A, 6, 2, 1, o2)).all()
    assert (orc.tsog1(A, 6, 4, 1, o1) == orc.tsog1(A, 6, 2, 1, o1)).all()
    assert (orc.tsog1(A, 6, 5, 1, o1) == orc.tsog1(A, 6, 3, 1, o1)).all()
    assert not (orc.tsog1(A, 6, 3, 1, o1) == orc.tsog1(A, 6, 3, 1, o2)).all()
    assert (orc.tsog1(A, 6, 1, 1, o1) == 0).all()
    # even q, stabilising every pass: S = Lfactor(first k rows of A, transposed)   (lora_helpers.rs:89-100 with S1 = 0)
    assert (orc.tsog1(A, 6, 2, 1, o1) == orc.Stabilizer(A[:6, :].T.copy())).all()


def test_rf1_qb1(orc):
    """src/lora_helpers.rs:236-304"""
    A = random_matrix(20, 10, seed=11)
    for mode in (0, 1):
        o = orc.make_opts(mode=mode)
        Q = orc.RF1(A, 5, o)
        assert Q.shape == (20, 5) and np.abs(Q.T @ Q - np.eye(5)).max() < 1e-13
        assert np.linalg.norm(A - Q @ Q.T @ A) < np.linalg.norm(A)
        Q, B = orc.QB1(A, 5, 0.01, o)
        assert B.shape == (5, 10) and np.abs(B - Q.T @ A).max() < 1e-13
        assert np.linalg.norm(A - Q @ B) / np.linalg.norm(A) <= 1.0


def test_rand_svd_reference_cases(orc):
    """src/lora_drivers.rs:236-475"""
    for mode in (0, 1):
        o = orc.make_opts(mode=mode)
        for (m, n, k) in [(20, 10, 5), (10, 20, 5), (10, 10, 5)]:
            U, S, Vt = orc.rand_svd(random_matrix(m, n, seed=m), k, 0.1, 5, o)
            assert U.shape == (m, k) and S.shape == (k, k) and Vt.shape == (k, n)
            s = np.diag(S)
            assert (np.diff(s) <= 1e-14).all() and (s >= 0).all()
            assert np.abs(U.T @ U - np.eye(k)).max() < 1e-12 and np.abs(Vt @ Vt.T - np.eye(k)).max() < 1e-12
        U, S, Vt = orc.rand_svd(np.eye(5), 3, 0.01, 2, o)                    # :359-373 identity
        assert U.shape == (5, 3) and np.abs(np.diag(S) - 1).max() < 1e-12
        U, S, Vt = orc.rand_svd(np.zeros((10, 10)), 5, 0.1, 5, o)            # :341-357 zero matrix
        assert np.abs(U - np.eye(10, 5)).max() < 1e-6 and np.abs(S).max() < 1e-6 and np.abs(Vt - np.eye(5, 10)).max() < 1e-6
        with pytest.raises(orc.OracleError) as e:
            orc.rand_svd(random_matrix(5, 5), 0, 0.1, 5, o)                  # :329-339 k = 0
        assert e.value.code == 1
        for bad in [dict(k=2, eps=0.0, s=2), dict(k=2, eps=0.1, s=0)]:
            with pytest.raises(orc.OracleError):
                orc.rand_svd(random_matrix(5, 5), bad["k"], bad["eps"], bad["s"], o)


def test_c1_both_modes_reproduce_the_exact_spectrum(orc):
    """on an exact rank-50 matrix with l = 60 both modes give the true singular values,
    so BASELINE config 1 is pinned against a plain deterministic SVD regardless of Omega."""
    A = rank_k_matrix(2000, 1000, 50, seed=1)
    sv = np.linalg.svd(A, compute_uv=False)[:50]
    for mode in (0, 1):
        U, S, Vt = orc.rand_svd(A, 50, 1e-6, 10, orc.make_opts(mode=mode))
        assert (np.abs(np.diag(S) - sv) / sv).max() < 1e-10
        assert np.linalg.norm(U @ S @ Vt - A) / np.linalg.norm(A) < 1e-12


def test_intended_mode_is_omega_insensitive_with_a_gap(orc):
    """SURVEY.md §8c fact (3)"""
    A, sig = lowrank_plus_noise(600, 300, seed=3, k=20)
    s1 = np.diag(orc.rand_svd(A, 20, 1e-6, 5, orc.make_opts(seed=1))[1])
    s2 = np.diag(orc.rand_svd(A, 20, 1e-6, 5, orc.make_opts(seed=2))[1])
    sv = np.linalg.svd(A, compute_uv=False)[:20]
    assert (np.abs(s1 - s2) / s1).max() < 1e-8
    assert (np.abs(s1 - sv) / sv).max() < 1e-6


def test_rand_evd1_reference_cases(orc):
    """src/lora_drivers.rs:490-673"""
    for mode in (0, 1):
        o = orc.make_opts(mode=mode)
        H = random_hermitian(10, seed=1)
        V, lam = orc.rand_evd1(H, 5, 0.1, 5, o)
        assert V.shape == (10, 5) and len(lam) == 5
        assert np.abs(V.T @ V - np.eye(5)).max() < 1e-6                      # :630-641
        assert (np.diff(np.abs(lam)) <= 1e-12).all()                         # :643-657
        w = np.linalg.eigvalsh(H); w = w[np.argsort(-np.abs(w))][:5]
        assert np.abs(lam - w).max() < 1e-10                                 # l = 10 = n: exact
        with pytest.raises(orc.OracleError) as e:
            orc.rand_evd1(random_matrix(6, 6, seed=2), 3, 0.1, 2, o)         # :503-515 non-symmetric
        assert e.value.code == 8
        V, lam = orc.rand_evd1(np.zeros((10, 10)), 5, 0.1, 5, o)             # :531-548 zero matrix
        assert np.abs(lam).max() == 0 and np.abs(V - np.eye(10, 5)).max() < 1e-6


def test_rand_evd2_reference_cases(orc):
    """src/lora_drivers.rs:688-878"""
    for mode in (0, 1):
        o = orc.make_opts(mode=mode)
        A = random_psd(5, seed=6)
        V, lam = orc.rand_evd2(A, 3, 2, o)                                    # :747-775, the strongest pin on the path
        w, W = np.linalg.eigh(A); idx = np.argsort(-w)[:3]
        assert np.abs(lam - w[idx]).max() < 1e-6
        assert np.abs(V.T @ V - np.eye(3)).max() < 1e-6
        assert np.abs(V @ V.T @ A - W[:, idx] @ W[:, idx].T @ A).max() < 1e-6
        with pytest.raises(orc.OracleError) as e:
            orc.rand_evd2(np.zeros((5, 5)), 3, 2, o)                          # :724-732 zero matrix -> Cholesky fails
        assert e.value.code == 7
        N = -random_psd(5, seed=7)
        with pytest.raises(orc.OracleError) as e:
            orc.rand_evd2(N, 3, 2, o)                                         # :705-722 not PSD
        assert e.value.code == 9
        V, lam = orc.rand_evd2(random_psd(8, seed=8), 3, 0, o)                # :841 s = 0 accepted
        assert V.shape[0] == 8 and len(lam) <= 3
        with pytest.raises(orc.OracleError) as e:
            orc.rand_evd2(A, 0, 2, o)
        assert e.value.code == 1


def test_sketch_step(orc):
    """src/sketch_and_precondition.rs:49,105,172 dimension rules; S A against a materialised S"""
    assert orc.sketch_dim(100, 10, 2.0) == 20 and orc.sketch_dim(15, 10, 2.0) == 15 and orc.sketch_dim(100, 10, 2.55) == 25
    assert orc.sketch_dim(100, 10, 1.0, saddle=True) == 10 and orc.sketch_dim(5, 10, 1.0, saddle=True) == 5
    A = random_matrix(300, 12, seed=5)
    St = orc.omega_fill(0, 300, 40, seed=9, stream=3)
    assert np.abs(orc.sketch_apply_dense(A, 40, seed=9) - St.T @ A).max() < 1e-11
    Ask = orc.sketch_apply_saso(A, 64, zeta=8, seed=9)
    # S has exactly zeta entries +-1/sqrt(zeta) per column: recover S from the identity and compare
    S = orc.sketch_apply_saso(np.eye(300), 64, zeta=8, seed=9)
    assert np.abs(S @ A - Ask).max() < 1e-12
    assert np.allclose((S**2).sum(axis=0), 1.0) or ((S != 0).sum(axis=0) <= 8).all()
    assert abs(np.linalg.norm(Ask) / np.linalg.norm(A) - 1.0) < 0.2


@pytest.mark.parametrize("zeta,width", [(8, 4), (8, 8), (8, 2), (8, 1), (4, 4), (2, 1), (1, 1)])
def test_block_sparse_sign_operator_definition(orc, zeta, width):
    """the block sparse-sign operator this build defines (the reference has no sparse sketch):
    structure of S recovered from S I, balanced dealing per chunk, shard consistency"""
    m, d = 5000, 240
    g, nbs = zeta // width, d // zeta
    S = orc.sketch_apply_saso_block(np.eye(m), d, zeta=zeta, seed=3, width=width)
    assert ((S != 0).sum(axis=0) == zeta).all()
    assert np.allclose(np.abs(S[S != 0]), 1 / np.sqrt(zeta), rtol=0, atol=1e-16)
    assert 0.47 < (S > 0).sum() / (S != 0).sum() < 0.53
    for j in range(0, m, 37):
        rows = np.nonzero(S[:, j])[0]
        for t in range(g):                                   # one aligned block of `width` rows in every stripe
            blk = rows[t * width:(t + 1) * width]
            assert blk[0] % width == 0 and (np.diff(blk) == 1).all()
            assert t * nbs * width <= blk[0] < (t + 1) * nbs * width
    # dealing: inside one chunk of 2048 rows every block of a stripe receives floor or ceil of 2048 / nbs rows
    first = np.array([np.nonzero(S[:, j])[0][0] // width for j in range(2048)])
    counts = np.bincount(first, minlength=nbs)[:nbs]
    assert counts.min() >= 2048 // nbs and counts.max() <= -(-2048 // nbs)
    A = random_matrix(m, 6, seed=8)
    full = orc.sketch_apply_saso_block(A, d, zeta=zeta, seed=3, width=width)
    assert np.abs(full - S @ A).max() < 1e-12
    parts = orc.sketch_apply_saso_block(A[:2300], d, zeta, 3, 0, width) + orc.sketch_apply_saso_block(A[2300:], d, zeta, 3, 2300, width)
    assert np.abs(full - parts).max() < 1e-12
    # different seeds give different operators
    assert (orc.sketch_apply_saso_block(np.eye(300), d, zeta, 4, 0, width) != S[:, :300]).any()


def test_block_sparse_sign_embedding_quality(orc):
    """subspace-embedding quality at d = 4n next to the textbook SASO: width <= 4 holds on a coherent basis,
    width 8 (one block per column) does not -- the reason the default width is min(zeta, 4)"""
    rng = np.random.default_rng(1)
    Qi, _ = np.linalg.qr(rng.standard_normal((20000, 100)))
    Qc, _ = np.linalg.qr(np.vstack([100 * np.eye(100), 1e-2 * rng.standard_normal((19900, 100))]))
    sv = lambda M: np.linalg.svd(M, compute_uv=False)
    for Q in (Qi, Qc):
        t = sv(orc.sketch_apply_saso(Q, 400, zeta=8, seed=0))
        assert 0.35 < t.min() and t.max() < 1.7
        for w in (1, 2, 4):
            b = sv(orc.sketch_apply_saso_block(Q, 400, zeta=8, seed=0, width=w))
            assert 0.25 < b.min() and b.max() < 1.8
    assert sv(orc.sketch_apply_saso_block(Qc, 400, zeta=8, seed=0, width=8)).min() < 0.1


def test_cgls_and_blendenpik_restatement(orc):
    """src/cg.rs:18-61 and src/sketch_and_precondition.rs:26-59 restated: CGLS reproduces a numpy transcription of the
    reference loop iterate for iterate; blendenpik converges to the least-squares solution in a few dozen iterations for
    every sketch operator, and reports the reference's validation errors"""
    rng = np.random.default_rng(3)
    a = rng.standard_normal((200, 12)); b = rng.standard_normal((200, 1))
    x, it, conv = orc.cgls(a, b, 1e-12, 100)
    xr = np.zeros((12, 1)); r = b - a @ xr; s = a.T @ r; p_ = s.copy(); ns = (s.T @ s).item(); itr = 0; cv = False
    for i in range(100):
        ap = a @ p_; alpha = ns / (ap.T @ ap).item(); xr += alpha * p_; r -= alpha * ap
        sn = a.T @ r; nn = (sn.T @ sn).item()
        if np.sqrt(nn) < 1e-12:
            cv = True; itr = i + 1; break
        p_ = sn + (nn / ns) * p_; ns = nn
    assert conv == cv and abs(it - itr) <= 1 and np.abs(x - xr).max() < 1e-12
    assert np.abs(x - np.linalg.lstsq(a, b, rcond=None)[0]).max() < 1e-10
    m, n = 3000, 40
    A = rng.standard_normal((m, n)) * np.logspace(0, -4, n)
    xt = rng.uniform(-100, 100, (n, 1)); bb = A @ xt + 1e-3 * rng.standard_normal((m, 1))
    xl = np.linalg.lstsq(A, bb, rcond=None)[0]
    for kind in (0, 1, 2):
        xs, its, cvs = orc.blendenpik(A, bb, 1e-10, 200, 4.0, kind=kind)
        assert cvs and its < 60 and np.linalg.norm(xs - xl) <= 1e-8 * np.linalg.norm(xl)
    for kind in (0, 2):                                      # src/sketch_and_precondition.rs:82-119
        xs, its, cvs = orc.lsrn(A, bb, 1e-10, 300, 4.0, kind=kind)
        assert cvs and its < 80 and np.linalg.norm(xs - xl) <= 1e-8 * np.linalg.norm(xl)
    for bad in ((1e-6, 10, 0.5), (0.0, 10, 2.0), (1e-6, 0, 2.0)):
        with pytest.raises(ValueError) as e:
            orc.blendenpik(A, bb, bad[0], bad[1], bad[2])
        assert e.value.args[0] == 1
    with pytest.raises(ValueError) as e:
        orc.blendenpik(A[:20].copy(), bb[:20], 1e-6, 10, 2.0)
    assert e.value.args[0] == 4


def test_oracle_reproduces_committed_golden_vectors(orc):
    """tests/golden/path_golden.npz (written by tests/golden/make_golden.py) is what the GPU tests are held to; the oracle
    must keep reproducing it, so a change on either side is caught"""
    import os
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "path_golden.npz"))
    A = np.asfortranarray(g["svd_A"])
    for mode in (0, 1):
        S = orc.rand_svd(A, int(g["svd_k"]), 1e-6, int(g["svd_s"]), orc.make_opts(mode=mode))[1]
        assert np.abs(np.diag(S) - g[f"svd_sigma_mode{mode}"]).max() <= 1e-14
    assert orc.omega_fill(0, 16, 5, seed=int(g["omega_seed"]), stream=1).tobytes() == np.asfortranarray(g["omega_gauss"]).tobytes()
    assert np.abs(orc.rand_evd1(np.asfortranarray(g["evd1_A"]), 6, 0.1, 6, orc.make_opts(mode=0))[1] - g["evd1_lambda"]).max() <= 1e-13
    assert np.abs(orc.rand_evd2(np.asfortranarray(g["evd2_A"]), 6, 4, orc.make_opts(mode=0))[1] - g["evd2_lambda"]).max() <= 1e-13
    assert orc.Stabilizer(np.asfortranarray(g["stab_X"])).tobytes() == np.asfortranarray(g["stab_L"]).tobytes()
    assert orc.sketch_apply_saso_block(np.eye(2100), 48, zeta=8, seed=11, width=4)[:, ::7].tobytes(order="F") == np.asfortranarray(g["sbs_S_z8w4"]).tobytes(order="F")
    assert orc.sketch_apply_saso_block(np.eye(2100), 48, zeta=4, seed=11, width=1)[:, ::7].tobytes(order="F") == np.asfortranarray(g["sbs_S_z4w1"]).tobytes(order="F")
    assert np.abs(orc.sketch_apply_saso_block(np.asfortranarray(g["sbs_T"]), 48, zeta=8, seed=11) - g["sbs_SA"]).max() <= 1e-14
    x = orc.blendenpik(np.asfortranarray(g["lsq_A"]), np.asfortranarray(g["lsq_b"]), 1e-12, 100, 4.0, kind=2, zeta=8)[0]
    assert np.linalg.norm(x - g["lsq_x_block"]) <= 1e-12 * np.linalg.norm(x)
