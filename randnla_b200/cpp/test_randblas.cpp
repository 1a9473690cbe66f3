// C++ host-mirror test: the reference's own unit tests for this path, re-expressed against randblas.hpp.
// Built and run by tests/test_cpp_mirror.py on the GPU box.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include "randblas.hpp"

using namespace randblas;
static int failures = 0;
#define CHECK(cond) do { if (!(cond)) { std::printf("FAIL %s:%d  %s\n", __FILE__, __LINE__, #cond); ++failures; } } while (0)

static bool approx_identity(const DMatrix& m, double tol) {
    for (size_t j = 0; j < m.ncols(); ++j) for (size_t i = 0; i < m.nrows(); ++i)
        if (std::fabs(m(i, j) - (i == j ? 1.0 : 0.0)) > tol) return false;
    return true;
}

int main() {
    using errors::RandNLAError;
    // src/lora_drivers.rs:341-357 test_rand_svd_zero_matrix
    {
        auto [U, S, Vt] = lora_drivers::rand_svd(DMatrix::zeros(10, 10), 5, 0.1, 5);
        CHECK(U.nrows() == 10 && U.ncols() == 5 && S.nrows() == 5 && Vt.nrows() == 5 && Vt.ncols() == 10);
        CHECK(approx_identity(U, 1e-6) && approx_identity(Vt, 1e-6) && S.norm() < 1e-6);
    }
    // src/lora_drivers.rs:329-339 k = 0 -> InvalidParameters
    try { lora_drivers::rand_svd(DMatrix::identity(5, 5), 0, 0.1, 5); CHECK(false); }
    catch (const RandNLAError& e) { CHECK(e.kind == RandNLAError::InvalidParameters); CHECK(std::string(e.what()) == "Rank k must be positive, current input is 0"); }
    // rank-3 matrix: singular values recovered, U orthonormal
    {
        DMatrix A(40, 20);
        for (size_t i = 0; i < 40; ++i) for (size_t j = 0; j < 20; ++j)
            A(i, j) = 3.0 * std::sin(0.1 * i) * std::cos(0.2 * j) + 2.0 * std::cos(0.3 * i) * std::sin(0.1 * j + 1) + ((i * 7 + j * 3) % 5 == 0 ? 1.0 : 0.0) * 0;
        auto [U, S, Vt] = lora_drivers::rand_svd(A, 4, 1e-6, 4);
        DMatrix R = U * S * Vt;
        double err = 0; for (size_t i = 0; i < 40; ++i) for (size_t j = 0; j < 20; ++j) err += (R(i, j) - A(i, j)) * (R(i, j) - A(i, j));
        CHECK(std::sqrt(err) / A.norm() < 1e-10);
        CHECK(approx_identity(U.transpose() * U, 1e-10));
        CHECK(S(0, 0) >= S(1, 1) && S(1, 1) >= S(2, 2));
    }
    // src/lora_helpers.rs:324-340, 358-365
    CHECK(approx_identity(lora_helpers::Orth(DMatrix::zeros(5, 5)), 0.0));
    CHECK(approx_identity(lora_helpers::Orth(DMatrix::identity(5, 5)), 1e-15));
    CHECK(approx_identity(lora_helpers::Stabilizer(DMatrix::zeros(5, 5)), 0.0));
    // src/sketch.rs:216-248
    {
        auto M = sketch::sketching_operator(sketch::DistributionType::Gaussian, 6, 4);
        CHECK(M.nrows() == 6 && M.ncols() == 4);
        try { sketch::sketching_operator(sketch::DistributionType::Uniform, 0, 4); CHECK(false); }
        catch (const RandNLAError& e) { CHECK(e.kind == RandNLAError::InvalidDimensions); }
    }
    // src/lora_drivers.rs:503-515 non-symmetric -> NotHermitian; :724-732 zero matrix -> MatrixDecompositionError
    {
        DMatrix N(4, 4); N(0, 1) = 1.0;
        try { lora_drivers::rand_evd1(N, 2, 0.1, 2); CHECK(false); } catch (const RandNLAError& e) { CHECK(e.kind == RandNLAError::NotHermitian); }
        try { lora_drivers::rand_evd2(DMatrix::zeros(5, 5), 3, 2); CHECK(false); } catch (const RandNLAError& e) { CHECK(e.kind == RandNLAError::MatrixDecompositionError); }
    }
    // sketch step validation, src/sketch_and_precondition.rs:29-48
    try { sketch_and_precondition::sketch_only(DMatrix(3, 4), 0.1, 10, 2.0); CHECK(false); }
    catch (const RandNLAError& e) { CHECK(e.kind == RandNLAError::NotOverdetermined); }
    {
        auto [a_sk, b_sk] = sketch_and_precondition::blendenpik_sketch(DMatrix(200, 5, 1.0), DMatrix(200, 1, 1.0), 1e-6, 10, 4.0);
        CHECK(a_sk.nrows() == 20 && a_sk.ncols() == 5 && b_sk.nrows() == 20);
    }
    // src/sketch_and_precondition.rs:229-290 test_blendenpik_overdetermined: the solve itself and its errors
    {
        const size_t m = 400, n = 6;
        DMatrix A = DMatrix::from_fn(m, n, [](size_t i, size_t j) { return std::sin(0.37 * (double)(i + 1) * (double)(j + 1)) + (i % (j + 2) == 0 ? 1.0 : 0.0); });
        DMatrix xt = DMatrix::from_fn(n, 1, [](size_t i, size_t) { return (double)i - 2.5; });
        DMatrix b = A * xt;
        auto x = sketch_and_precondition::blendenpik_overdetermined(A, b, 1e-12, 100, 4.0);
        double err = 0.0;
        for (size_t i = 0; i < n; ++i) err = std::fmax(err, std::fabs(x(i, 0) - xt(i, 0)));
        CHECK(err < 1e-8);
        auto x2 = sketch_and_precondition::blendenpik_overdetermined(A, b, 1e-12, 100, 4.0, RNLA_SKETCH_SASO_BLOCK, 8);
        err = 0.0;
        for (size_t i = 0; i < n; ++i) err = std::fmax(err, std::fabs(x2(i, 0) - xt(i, 0)));
        CHECK(err < 1e-8);
        try { sketch_and_precondition::blendenpik_overdetermined(A, b, 1e-6, 10, 0.5); CHECK(false); }
        catch (const RandNLAError& e) { CHECK(e.kind == RandNLAError::InvalidParameters); }
    }
    // src/id.rs:463-546 on a rank-4 matrix: one-sided / two-sided ID and CUR reproduce it; :329-461 bad k throws
    {
        const size_t m = 60, n = 45, k = 4;
        DMatrix L = DMatrix::from_fn(m, k, [](size_t i, size_t j) { return std::sin(0.7 * (double)(i + 1) * (double)(j + 1)); });
        DMatrix Rr = DMatrix::from_fn(k, n, [](size_t i, size_t j) { return std::cos(0.3 * (double)(i + 2) * (double)(j + 1)); });
        DMatrix A = L * Rr;
        auto sel_cols = [&](const std::vector<size_t>& J) { return DMatrix::from_fn(m, J.size(), [&](size_t i, size_t c) { return A(i, J[c]); }); };
        auto sel_rows = [&](const std::vector<size_t>& I) { return DMatrix::from_fn(I.size(), n, [&](size_t r, size_t j) { return A(I[r], j); }); };
        auto rel = [&](const DMatrix& B) { double e = 0; for (size_t i = 0; i < m; ++i) for (size_t j = 0; j < n; ++j) e += (B(i, j) - A(i, j)) * (B(i, j) - A(i, j)); return std::sqrt(e) / A.norm(); };
        auto [x, j] = id::osid_qrcp(A, k, sketch::MatrixAttribute::Column);
        CHECK(rel(sel_cols(j) * x) < 1e-10);
        auto [xr, jr] = id::osid_randomised(A, k, sketch::MatrixAttribute::Column);
        CHECK(rel(sel_cols(jr) * xr) < 1e-9);
        for (int rnd = 0; rnd < 2; ++rnd) {
            auto [cj, u, ci] = rnd ? id::cur_randomised(A, k) : id::cur(A, k);
            CHECK(rel(sel_cols(cj) * u * sel_rows(ci)) < 1e-8);
            auto [z, ti, tj, tx] = rnd ? id::two_sided_id_randomised(A, k) : id::two_sided_id(A, k);
            DMatrix core = DMatrix::from_fn(k, k, [&](size_t r, size_t c) { return A(ti[r], tj[c]); });
            CHECK(rel(z * core * tx) < 1e-8);
        }
        try { id::osid_qrcp(A, 0, sketch::MatrixAttribute::Column); CHECK(false); }
        catch (const RandNLAError& e) { CHECK(e.kind == RandNLAError::InvalidParameters && std::string(e.what()) == "k must be positive)"); }
        // src/pivot_decompositions.rs:320-395 and src/cqrrpt.rs:73-133
        auto [q, r, p] = pivot_decompositions::qrcp(L);
        DMatrix qr = q * r;
        double e = 0; for (size_t i = 0; i < m; ++i) for (size_t c = 0; c < k; ++c) e = std::fmax(e, std::fabs(qr(i, c) - L(i, p[c])));
        CHECK(e < 1e-12 && approx_identity(q.transpose() * q, 1e-12));
        auto [qc, rc, jc] = cqrrpt::sap_chol_qrcp(L, 12);
        DMatrix qrc = qc * rc;
        e = 0; for (size_t i = 0; i < m; ++i) for (size_t c = 0; c < k; ++c) e = std::fmax(e, std::fabs(qrc(i, c) - L(i, jc[c])));
        CHECK(e < 1e-12 && approx_identity(qc.transpose() * qc, 1e-12));
        try { cqrrpt::sap_chol_qrcp(L, 2); CHECK(false); } catch (const RandNLAError& e2) { CHECK(e2.kind == RandNLAError::InvalidParameters); }
        // src/sketch_and_solve.rs:81-158 and src/sketch_and_precondition.rs:278-337 on a consistent system
        DMatrix xt = DMatrix::from_fn(k, 1, [](size_t i, size_t) { return 1.5 - (double)i; });
        DMatrix bb = L * xt;
        for (int w = 0; w < 2; ++w) {
            DMatrix xs = w ? sketch_and_solve::sketched_least_squares_svd(L, bb) : sketch_and_solve::sketched_least_squares_qr(L, bb);
            double d = 0; for (size_t i = 0; i < k; ++i) d = std::fmax(d, std::fabs(xs(i, 0) - xt(i, 0)));
            CHECK(d < 1e-9);
        }
        auto [sx, sy] = sketch_and_precondition::sketch_saddle_point_precondition(L, bb, DMatrix(), 0.0, 1e-12, 100, 2.0);
        double d = 0; for (size_t i = 0; i < k; ++i) d = std::fmax(d, std::fabs(sx(i, 0) - xt(i, 0)));
        CHECK(d < 1e-8 && sy.norm() < 1e-8 * bb.norm());
        // src/pivot_decompositions.rs:351-369 (test_lupp) on a 60 x 60 Gaussian matrix
        {
            DMatrix g = sketch::sketching_operator(sketch::DistributionType::Gaussian, 60, 60);
            auto [ll, uu, pp] = pivot_decompositions::lupp(g);
            DMatrix lu = ll * uu;
            double e2 = 0, lo = 0, up = 0;
            for (size_t i = 0; i < 60; ++i) for (size_t c = 0; c < 60; ++c) {
                e2 = std::fmax(e2, std::fabs(lu(i, c) - g(pp[i], c)));
                if (i < c) up = std::fmax(up, std::fabs(ll(i, c)));
                if (i > c) lo = std::fmax(lo, std::fabs(uu(i, c)));
            }
            CHECK(e2 < 1e-12 && up == 0.0 && lo == 0.0);
            try { pivot_decompositions::lupp(DMatrix(3, 4)); CHECK(false); } catch (const RandNLAError& e4) { CHECK(e4.kind == RandNLAError::NotSquare); }
            try { pivot_decompositions::lupp(DMatrix(3, 3)); CHECK(false); } catch (const RandNLAError& e4) { CHECK(e4.kind == RandNLAError::SingularMatrix); }
        }
        // src/cg.rs:130-197
        DMatrix ca = DMatrix::from_fn(3, 3, [](size_t i, size_t j) { const double v[9] = {4, 1, 2, 1, 3, 0, 2, 0, 1}; return v[i * 3 + j]; });
        DMatrix cb = DMatrix::from_fn(3, 1, [](size_t i, size_t) { return i == 0 ? 4.0 : 2.0; });
        DMatrix ones3(3, 1, 1.0);
        CHECK(cg::cgls(ca, cb, 3.0, 100).norm() < 3.0);
        CHECK(cg::cgls(ca, cb, 3.0, 100, &ones3).norm() < 3.0);
        CHECK(cg::cgls(ca, cb, 1e-20, 1).norm() > 1e-20);
        DMatrix sa = DMatrix::from_fn(3, 3, [](size_t i, size_t j) { const double v[9] = {4, 1, 2, 1, 3, 1, 2, 1, 3}; return v[i * 3 + j]; });
        DMatrix sb = DMatrix::from_fn(3, 1, [](size_t i, size_t) { return 1.0 + (double)i; });
        CHECK(cg::verify_solution(sa, sb, cg::conjugate_grad(sa, sb, &ones3)) < 1e-10);
        try { DMatrix ind = DMatrix::identity(3, 3); ind(1, 1) = -2.0; cg::conjugate_grad(ind, sb); CHECK(false); }
        catch (const RandNLAError& e3) { CHECK(e3.kind == RandNLAError::NotPositiveSemiDefinite); }
        // src/solvers.rs:391-410 (test_simple_system) and a consistent tall system through lsqr
        DMatrix a3 = DMatrix::from_fn(3, 2, [](size_t i, size_t j) { return (i == j || i == j + 1) ? 1.0 : 0.0; });
        DMatrix b0(3, 1), b1 = DMatrix::from_fn(3, 1, [](size_t i, size_t) { return i == 0 ? 1.0 : (i == 2 ? -1.0 : 0.0); });
        auto z0 = solvers::lsqr(a3, b0, 0.0, 1e-8, 1e-8, 1e8, -1, false, nullptr);
        CHECK(z0.istop == 0 && z0.itn == 0 && z0.x.norm() == 0.0 && z0.arnorms.size() == 1 && z0.arnorms[0] == 0.0);
        auto z1 = solvers::lsqr(a3, b1, 0.0, 1e-8, 1e-8, 1e8, -1, false, nullptr);
        CHECK(std::fabs(z1.x(0, 0) - 1.0) < 1e-2 && std::fabs(z1.x(1, 0) + 1.0) < 1e-2);
        auto z2 = solvers::lsqr(L, bb, 0.0, 1e-13, 1e-13, 1e8, -1, true, nullptr);
        d = 0; for (size_t i = 0; i < k; ++i) d = std::fmax(d, std::fabs(z2.x(i, 0) - xt(i, 0)));
        CHECK(d < 1e-8 && z2.istop >= 1 && z2.istop <= 2 && z2.arnorms.size() == z2.itn && z2.var(0, 0) > 0.0);
    }
    std::printf(failures ? "CPP MIRROR: %d FAILURES\n" : "CPP MIRROR: ALL OK\n", failures);
    return failures ? 1 : 0;
}
