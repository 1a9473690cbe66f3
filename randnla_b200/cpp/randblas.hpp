// randblas.hpp -- C++ host-side mirror of the reference's Rust API over the C ABI (include/rnla.h).
//
// The reference is a Rust crate (`randblas`); no Rust toolchain exists in this image, so the host side above the
// C ABI is provided in C++ with the reference's module / function names, argument meaning and error behaviour:
//   randblas::sketch::{DistributionType, MatrixAttribute, sketching_operator, haar_sample}   <- src/sketch.rs
//   randblas::lora_helpers::{QB1, RF1, tsog1, Orth, Stabilizer}                               <- src/lora_helpers.rs
//   randblas::lora_drivers::{rand_svd, rand_evd1, rand_evd2}                                  <- src/lora_drivers.rs
//   randblas::sketch_and_precondition::{blendenpik_sketch, lsrn_sketch, saddle_point_sketch} <- src/sketch_and_precondition.rs:26-52,82-107,150-176
//   randblas::sketch_and_precondition::{blendenpik_overdetermined, lsrn_overdetermined, sketch_saddle_point_precondition}
//   randblas::pivot_decompositions::{lupp, qrcp, economic_qrcp}                                    <- src/pivot_decompositions.rs
//   randblas::cqrrpt::sap_chol_qrcp                                                          <- src/cqrrpt.rs
//   randblas::sketch_and_solve::{sketched_least_squares_qr, sketched_least_squares_svd}      <- src/sketch_and_solve.rs
//   randblas::cg::{cgls, conjugate_grad, verify_solution}                                    <- src/cg.rs
//   randblas::solvers::lsqr                                                                 <- src/solvers.rs:115-278
//   randblas::id::{osid_qrcp, osid_randomised, two_sided_id(_randomised), cur(_randomised)}  <- src/id.rs
//   randblas::errors::RandNLAError                                                           <- src/errors.rs
// `Result<T, RandNLAError>` becomes "returns T or throws RandNLAError".  Header-only; link with -lrnla.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <stdexcept>
#include <string>
#include <tuple>
#include <utility>
#include <vector>

#include "rnla.h"

namespace randblas {

// nalgebra::DMatrix<f64>: heap vector, column-major, lda = nrows
class DMatrix {
public:
    DMatrix() : r_(0), c_(0) {}
    DMatrix(size_t rows, size_t cols, double fill = 0.0) : r_(rows), c_(cols), d_(rows * cols, fill) {}
    static DMatrix zeros(size_t rows, size_t cols) { return DMatrix(rows, cols, 0.0); }
    static DMatrix identity(size_t rows, size_t cols) {
        DMatrix m(rows, cols);
        for (size_t i = 0; i < rows && i < cols; ++i) m(i, i) = 1.0;
        return m;
    }
    template <class F> static DMatrix from_fn(size_t rows, size_t cols, F f) {
        DMatrix m(rows, cols);
        for (size_t j = 0; j < cols; ++j) for (size_t i = 0; i < rows; ++i) m(i, j) = f(i, j);
        return m;
    }
    size_t nrows() const { return r_; }
    size_t ncols() const { return c_; }
    double& operator()(size_t i, size_t j) { return d_[i + j * r_]; }
    double operator()(size_t i, size_t j) const { return d_[i + j * r_]; }
    const double* as_ptr() const { return d_.data(); }
    double* as_mut_ptr() { return d_.data(); }
    DMatrix transpose() const { return from_fn(c_, r_, [&](size_t i, size_t j) { return (*this)(j, i); }); }
    DMatrix columns(size_t first, size_t n) const { return from_fn(r_, n, [&](size_t i, size_t j) { return (*this)(i, first + j); }); }
    DMatrix operator*(const DMatrix& b) const {      // small host products for tests; the hot products live on the GPU
        DMatrix o(r_, b.c_);
        for (size_t j = 0; j < b.c_; ++j) for (size_t k = 0; k < c_; ++k) { const double v = b(k, j); for (size_t i = 0; i < r_; ++i) o(i, j) += (*this)(i, k) * v; }
        return o;
    }
    double norm() const { double s = 0; for (double v : d_) s += v * v; return std::sqrt(s); }
private:
    size_t r_, c_;
    std::vector<double> d_;
};

namespace errors {
// reference src/errors.rs:3-14; Display :16-31
struct RandNLAError : public std::exception {
    enum Kind { InvalidParameters = 1, InvalidDimensions, NegativeDimensions, NotOverdetermined, NotSquare, SingularMatrix,
                MatrixDecompositionError, NotHermitian, NotPositiveSemiDefinite, ComputationError };
    Kind kind;
    std::string msg, shown;
    RandNLAError(Kind k, std::string m) : kind(k), msg(std::move(m)) {
        switch (k) {
            case MatrixDecompositionError: shown = "Matrix decomposition error: " + msg; break;
            case NotHermitian: shown = "Not a Hermitian matrix: " + msg; break;
            case NotPositiveSemiDefinite: shown = "Not a positive semi-definite matrix: " + msg; break;
            case ComputationError: shown = "Computation error: " + msg; break;
            default: shown = msg;
        }
    }
    const char* what() const noexcept override { return shown.c_str(); }
};
inline void check(rnla_status s) {
    if (s != RNLA_OK) throw RandNLAError(static_cast<RandNLAError::Kind>(s), rnla_last_error_message());
}
}  // namespace errors

namespace sketch {
enum class DistributionType { Gaussian = 0, Uniform = 1, Rademacher = 2 };   // src/sketch.rs:9-13
enum class MatrixAttribute { Row = 0, Column = 1 };                          // src/sketch.rs:18-21

// src/sketch.rs:102-130
inline DMatrix sketching_operator(DistributionType dist, size_t rows, size_t cols) {
    DMatrix m(rows, cols);
    errors::check(rnla_sketching_operator(static_cast<int32_t>(dist), (int64_t)rows, (int64_t)cols, m.as_mut_ptr()));
    return m;
}
// src/sketch.rs:45-85
inline DMatrix haar_sample(size_t rows, size_t cols, MatrixAttribute attr) {
    DMatrix m(rows, cols);
    errors::check(rnla_haar_sample((int64_t)rows, (int64_t)cols, static_cast<int32_t>(attr), m.as_mut_ptr()));
    return m;
}
}  // namespace sketch

namespace lora_helpers {
// src/lora_helpers.rs:131-133
inline DMatrix Orth(const DMatrix& X) {
    const size_t p = std::min(X.nrows(), X.ncols());
    DMatrix Q(X.nrows(), p);
    int64_t qc = 0;
    errors::check(rnla_orth(X.as_ptr(), (int64_t)X.nrows(), (int64_t)X.ncols(), Q.as_mut_ptr(), nullptr, &qc));
    return Q;
}
// src/lora_helpers.rs:144-146
inline DMatrix Stabilizer(const DMatrix& X) {
    const size_t p = std::min(X.nrows(), X.ncols());
    DMatrix L(X.nrows(), p);
    int64_t lc = 0;
    errors::check(rnla_stabilizer(X.as_ptr(), (int64_t)X.nrows(), (int64_t)X.ncols(), L.as_mut_ptr(), &lc));
    return L;
}
// src/lora_helpers.rs:58-105 (infallible signature in the reference: failures abort, like its unwrap())
inline DMatrix tsog1(const DMatrix& A, size_t k, int32_t num_passes, int32_t passes_per_stab) {
    DMatrix S(A.ncols(), k);
    errors::check(rnla_tsog1(A.as_ptr(), (int64_t)A.nrows(), (int64_t)A.ncols(), (int64_t)k, num_passes, passes_per_stab, S.as_mut_ptr()));
    return S;
}
// src/lora_helpers.rs:37-44
inline DMatrix RF1(const DMatrix& A, size_t k) {
    const size_t l = std::min(k, std::min(A.nrows(), A.ncols()));
    DMatrix Q(A.nrows(), l);
    int64_t qc = 0;
    errors::check(rnla_rf1(A.as_ptr(), (int64_t)A.nrows(), (int64_t)A.ncols(), (int64_t)k, Q.as_mut_ptr(), &qc));
    return Q;
}
// src/lora_helpers.rs:17-23
inline std::pair<DMatrix, DMatrix> QB1(const DMatrix& A, size_t k, double epsilon) {
    const size_t l = std::min(k, std::min(A.nrows(), A.ncols()));
    DMatrix Q(A.nrows(), l), B(l, A.ncols());
    int64_t qc = 0;
    errors::check(rnla_qb1(A.as_ptr(), (int64_t)A.nrows(), (int64_t)A.ncols(), (int64_t)k, epsilon, Q.as_mut_ptr(), B.as_mut_ptr(), &qc));
    return {std::move(Q), std::move(B)};
}
}  // namespace lora_helpers

namespace lora_drivers {
// src/lora_drivers.rs:30-69 -> (U m x r, S r x r dense diagonal, V^T r x n)
inline std::tuple<DMatrix, DMatrix, DMatrix> rand_svd(const DMatrix& A, size_t k, double epsilon, size_t s) {
    const size_t m = A.nrows(), n = A.ncols();
    const size_t cap = std::max<size_t>(std::min(std::max<size_t>(k, 1), std::min(m, n)), 1);
    DMatrix U(m, cap), S(cap, cap), Vt(cap, n);
    std::puts("Running RSVD");                                                   // :47
    int64_t r = 0;
    errors::check(rnla_rand_svd(A.as_ptr(), (int64_t)m, (int64_t)n, (int64_t)k, epsilon, (int64_t)s, U.as_mut_ptr(), S.as_mut_ptr(), Vt.as_mut_ptr(), &r));
    return {std::move(U), std::move(S), std::move(Vt)};
}
// src/lora_drivers.rs:87-151
inline std::pair<DMatrix, std::vector<double>> rand_evd1(const DMatrix& A, size_t k, double epsilon, size_t s) {
    const size_t n = A.nrows();
    if (A.ncols() != n) throw errors::RandNLAError(errors::RandNLAError::NotHermitian, "Input matrix is not Hermitian");
    const size_t cap = std::max<size_t>(std::min(std::max<size_t>(k, 1), n), 1);
    DMatrix V(n, cap);
    std::vector<double> lam(cap);
    std::puts("Running REVD1");                                                  // :112
    int64_t r = 0;
    errors::check(rnla_rand_evd1(A.as_ptr(), (int64_t)n, (int64_t)k, epsilon, (int64_t)s, V.as_mut_ptr(), lam.data(), &r));
    lam.resize((size_t)r);
    return {V.columns(0, (size_t)r), std::move(lam)};
}
// src/lora_drivers.rs:167-224
inline std::pair<DMatrix, std::vector<double>> rand_evd2(const DMatrix& A, size_t k, size_t s) {
    const size_t n = A.nrows();
    if (A.ncols() != n) throw errors::RandNLAError(errors::RandNLAError::NotSquare, "rand_evd2 needs a square matrix");
    const size_t cap = std::max<size_t>(std::min(std::max<size_t>(k, 1), n), 1);
    DMatrix V(n, cap);
    std::vector<double> lam(cap);
    std::puts("Running REVD2");                                                  // :175
    int64_t r = 0;
    errors::check(rnla_rand_evd2(A.as_ptr(), (int64_t)n, (int64_t)k, (int64_t)s, V.as_mut_ptr(), lam.data(), &r));
    lam.resize((size_t)r);
    return {V.columns(0, (size_t)r), std::move(lam)};
}
}  // namespace lora_drivers

namespace sketch_and_precondition {
using errors::RandNLAError;
inline void validate(const DMatrix& a, double epsilon, size_t l, double sampling_factor) {
    char b[160];
    if (a.nrows() < a.ncols()) {                                                 // src/sketch_and_precondition.rs:29-33
        std::snprintf(b, sizeof b, "Need more columns than rows, found %zu rows and %zu columns", a.nrows(), a.ncols());
        throw RandNLAError(RandNLAError::NotOverdetermined, b);
    }
    if (sampling_factor < 1.0) {                                                 // :34-38
        std::snprintf(b, sizeof b, "Sampling factor must be greater than 1, current input is %g", sampling_factor);
        throw RandNLAError(RandNLAError::InvalidParameters, b);
    }
    if (epsilon <= 0.0) {                                                        // :39-43
        std::snprintf(b, sizeof b, "Epsilon must be positive, current input is %g", epsilon);
        throw RandNLAError(RandNLAError::InvalidParameters, b);
    }
    if (l == 0) throw RandNLAError(RandNLAError::InvalidParameters, "Number of iterations must be positive, current input is 0");   // :44-48
}
// sketch step of blendenpik_overdetermined (:49-52): (a_sk, b_sk) = (S a, S b), dense Gaussian S (d x m)
inline std::pair<DMatrix, DMatrix> blendenpik_sketch(const DMatrix& a, const DMatrix& b, double epsilon, size_t l, double sampling_factor,
                                                     rnla_sketch_kind kind = RNLA_SKETCH_DENSE, int zeta = 8) {
    validate(a, epsilon, l, sampling_factor);
    const int64_t d = rnla_sketch_dim((int64_t)a.nrows(), (int64_t)a.ncols(), sampling_factor, 0);
    DMatrix a_sk((size_t)d, a.ncols()), b_sk((size_t)d, b.ncols());
    rnla_options o; rnla_get_options(&o);
    errors::check(rnla_sketch_apply(kind, RNLA_GAUSSIAN, o.seed, d, zeta, a.as_ptr(), (int64_t)a.nrows(), (int64_t)a.ncols(),
                                    b.as_ptr(), (int64_t)b.ncols(), a_sk.as_mut_ptr(), b_sk.as_mut_ptr()));
    return {std::move(a_sk), std::move(b_sk)};
}
// blendenpik_overdetermined end to end (:26-59): returns x (n x 1)
inline DMatrix blendenpik_overdetermined(const DMatrix& a, const DMatrix& b, double epsilon, size_t l, double sampling_factor,
                                         rnla_sketch_kind kind = RNLA_SKETCH_DENSE, int zeta = 8, int width = 0) {
    validate(a, epsilon, l, sampling_factor);
    DMatrix x(a.ncols(), 1);
    int64_t iters = 0; int32_t conv = 0;
    errors::check(rnla_blendenpik_overdetermined(a.as_ptr(), (int64_t)a.nrows(), (int64_t)a.ncols(), b.as_ptr(), epsilon, (int64_t)l,
                                                 sampling_factor, kind, kind == RNLA_SKETCH_SASO_BLOCK ? width : RNLA_GAUSSIAN, zeta,
                                                 x.as_mut_ptr(), &iters, &conv));
    if (conv) std::printf("CGLS converged after %lld iterations\n", (long long)iters);             // src/cg.rs:46
    else std::printf("CGLS failed to converged after %zu iterations\n", l);                        // src/cg.rs:58
    return x;
}
// lsrn_overdetermined end to end (:82-119), n <= 1024: returns x (n x 1)
inline DMatrix lsrn_overdetermined(const DMatrix& a, const DMatrix& b, double epsilon, size_t l, double sampling_factor,
                                   rnla_sketch_kind kind = RNLA_SKETCH_DENSE, int zeta = 8, int width = 0) {
    validate(a, epsilon, l, sampling_factor);
    DMatrix x(a.ncols(), 1);
    int64_t iters = 0; int32_t conv = 0;
    errors::check(rnla_lsrn_overdetermined(a.as_ptr(), (int64_t)a.nrows(), (int64_t)a.ncols(), b.as_ptr(), epsilon, (int64_t)l,
                                           sampling_factor, kind, kind == RNLA_SKETCH_SASO_BLOCK ? width : RNLA_GAUSSIAN, zeta,
                                           x.as_mut_ptr(), &iters, &conv));
    return x;
}
// sketch step of lsrn_overdetermined (:105-107) and sketch_saddle_point_precondition (:172-176, saddle = true)
inline DMatrix sketch_only(const DMatrix& a, double epsilon, size_t l, double sampling_factor, bool saddle = false,
                           rnla_sketch_kind kind = RNLA_SKETCH_DENSE, int zeta = 8) {
    validate(a, epsilon, l, sampling_factor);
    const int64_t d = rnla_sketch_dim((int64_t)a.nrows(), (int64_t)a.ncols(), sampling_factor, saddle ? 1 : 0);
    DMatrix a_sk((size_t)d, a.ncols());
    rnla_options o; rnla_get_options(&o);
    errors::check(rnla_sketch_apply(kind, RNLA_GAUSSIAN, o.seed, d, zeta, a.as_ptr(), (int64_t)a.nrows(), (int64_t)a.ncols(),
                                    nullptr, 0, a_sk.as_mut_ptr(), nullptr));
    return a_sk;
}
}  // namespace sketch_and_precondition

// sketch_saddle_point_precondition end to end (src/sketch_and_precondition.rs:150-216), n <= 1024: returns (x, y).
// `c` may be empty (0 x 0), as `c.is_empty()` at :195
namespace sketch_and_precondition {
inline std::pair<DMatrix, DMatrix> sketch_saddle_point_precondition(const DMatrix& a, const DMatrix& b, const DMatrix& c, double mu,
                                                                    double epsilon, size_t l, double sampling_factor) {
    validate(a, epsilon, l, sampling_factor);
    DMatrix x(a.ncols(), 1), y(a.nrows(), 1);
    int64_t iters = 0; int32_t conv = 0;
    errors::check(rnla_sketch_saddle_point_precondition(a.as_ptr(), (int64_t)a.nrows(), (int64_t)a.ncols(), b.as_ptr(),
                                                        c.nrows() * c.ncols() ? c.as_ptr() : nullptr, mu, epsilon, (int64_t)l,
                                                        sampling_factor, x.as_mut_ptr(), y.as_mut_ptr(), &iters, &conv));
    return {std::move(x), std::move(y)};
}
}  // namespace sketch_and_precondition

// ---- rows after the hot path --------------------------------------------------------------
inline std::vector<size_t> to_usize(const std::vector<int64_t>& v, size_t k) { return std::vector<size_t>(v.begin(), v.begin() + (long)k); }

namespace pivot_decompositions {
// src/pivot_decompositions.rs:105-180: (q m x m, r m x n, p)
inline std::tuple<DMatrix, DMatrix, std::vector<size_t>> qrcp(const DMatrix& a) {
    const size_t m = a.nrows(), n = a.ncols();
    DMatrix q(m, m), r(m, n);
    std::vector<int64_t> p(std::max<size_t>(n, 1));
    errors::check(rnla_qrcp(a.as_ptr(), (int64_t)m, (int64_t)n, (int64_t)std::min(m, n), (int64_t)m, q.as_mut_ptr(), r.as_mut_ptr(), p.data()));
    return {std::move(q), std::move(r), to_usize(p, n)};
}
// :196-269: (q_eco m x k, r_eco k x n, p); the reference's asserts ("k must be positive", "k must be <= min(m,n)") throw
inline std::tuple<DMatrix, DMatrix, std::vector<size_t>> economic_qrcp(const DMatrix& a, size_t k) {
    const size_t m = a.nrows(), n = a.ncols();
    DMatrix q(m, std::max<size_t>(k, 1)), r(m, n);
    std::vector<int64_t> p(std::max<size_t>(n, 1));
    errors::check(rnla_qrcp(a.as_ptr(), (int64_t)m, (int64_t)n, (int64_t)k, (int64_t)k, q.as_mut_ptr(), r.as_mut_ptr(), p.data()));
    return {std::move(q), DMatrix::from_fn(k, n, [&](size_t i, size_t j) { return r(i, j); }), to_usize(p, n)};
}
// src/pivot_decompositions.rs:21-86; throws NotSquare / SingularMatrix as the reference returns them
inline std::tuple<DMatrix, DMatrix, std::vector<size_t>> lupp(const DMatrix& matrix) {
    const size_t n = std::max<size_t>(matrix.nrows(), 1);
    DMatrix l(n, n), u(n, n);
    std::vector<int64_t> p(n);
    errors::check(rnla_lupp(matrix.as_ptr(), (int64_t)matrix.nrows(), (int64_t)matrix.ncols(), l.as_mut_ptr(), u.as_mut_ptr(), p.data()));
    return {std::move(l), std::move(u), to_usize(p, n)};
}
}  // namespace pivot_decompositions

namespace cqrrpt {
// src/cqrrpt.rs:27-58: (q m x k, r k x n, j)
inline std::tuple<DMatrix, DMatrix, std::vector<size_t>> sap_chol_qrcp(const DMatrix& a, size_t d) {
    const size_t m = a.nrows(), n = a.ncols();
    DMatrix q(m, n);
    std::vector<double> r(std::max<size_t>(n * n, 1));
    std::vector<int64_t> j(std::max<size_t>(n, 1));
    int64_t k = 0;
    errors::check(rnla_sap_chol_qrcp(a.as_ptr(), (int64_t)m, (int64_t)n, (int64_t)d, RNLA_SKETCH_DENSE, RNLA_GAUSSIAN, 0, q.as_mut_ptr(),
                                     r.data(), j.data(), &k));
    return {q.columns(0, (size_t)k), DMatrix::from_fn((size_t)k, n, [&](size_t i, size_t c) { return r[i + c * (size_t)k]; }), to_usize(j, n)};
}
}  // namespace cqrrpt

namespace sketch_and_solve {
// src/sketch_and_solve.rs:24-33
inline DMatrix sketched_least_squares_qr(const DMatrix& a, const DMatrix& b) {
    DMatrix x(a.ncols(), 1);
    errors::check(rnla_sketched_least_squares_qr(a.as_ptr(), (int64_t)a.nrows(), (int64_t)a.ncols(), b.as_ptr(), RNLA_SKETCH_DENSE, RNLA_GAUSSIAN, 0, x.as_mut_ptr()));
    return x;
}
// :54-66
inline DMatrix sketched_least_squares_svd(const DMatrix& a, const DMatrix& b) {
    DMatrix x(a.ncols(), 1);
    errors::check(rnla_sketched_least_squares_svd(a.as_ptr(), (int64_t)a.nrows(), (int64_t)a.ncols(), b.as_ptr(), RNLA_SKETCH_DENSE, RNLA_GAUSSIAN, 0, x.as_mut_ptr()));
    return x;
}
}  // namespace sketch_and_solve

namespace cg {
// src/cg.rs:18-61; x == nullptr is the reference's `None` (zeros).  Prints what the reference prints.
inline DMatrix cgls(const DMatrix& a, const DMatrix& b, double tolerance, size_t num_iterations, const DMatrix* x = nullptr) {
    if (b.nrows() != a.nrows() || (x && x->nrows() != a.ncols())) throw std::invalid_argument("cgls: shapes do not conform");
    DMatrix out(a.ncols(), 1);
    int64_t iters = 0; int32_t conv = 0;
    errors::check(rnla_cgls(a.as_ptr(), (int64_t)a.nrows(), (int64_t)a.ncols(), b.as_ptr(), tolerance, (int64_t)num_iterations,
                            x ? x->as_ptr() : nullptr, out.as_mut_ptr(), &iters, &conv));
    if (conv) std::printf("CGLS converged after %lld iterations\n", (long long)iters);               // :46
    else std::printf("CGLS failed to converged after %zu iterations\n", num_iterations);             // :58
    return out;
}
// :77-112; x == nullptr: the vector of ones (:88).  Throws RandNLAError::NotPositiveSemiDefinite (:80-86).
inline DMatrix conjugate_grad(const DMatrix& a, const DMatrix& b, const DMatrix* x = nullptr) {
    const size_t n = a.nrows();
    if (a.ncols() != n || b.nrows() != n || (x && x->nrows() != n)) throw std::invalid_argument("conjugate_grad: shapes do not conform");
    DMatrix out(n, 1);
    int64_t iters = 0; int32_t conv = 0;
    errors::check(rnla_conjugate_grad(a.as_ptr(), (int64_t)n, b.as_ptr(), x ? x->as_ptr() : nullptr, out.as_mut_ptr(), &iters, &conv));
    if (conv) std::printf("Converged after %lld iterations\n", (long long)iters);                    // :101
    return out;
}
// :115-117
inline double verify_solution(const DMatrix& a, const DMatrix& b, const DMatrix& x) {
    double r = 0.0;
    errors::check(rnla_verify_solution(a.as_ptr(), (int64_t)a.nrows(), (int64_t)a.ncols(), b.as_ptr(), x.as_ptr(), &r));
    return r;
}
}  // namespace cg

namespace solvers {
// src/solvers.rs:115-278: the reference's 10-tuple, in its order
struct LsqrOutput {
    DMatrix x;                     // solution vector (n x 1)
    size_t istop, itn;             // reason for termination, iterations performed
    double r1norm, r2norm, anorm, acond;
    std::vector<double> arnorms;   // history of the ||A^T r|| estimates
    double xnorm;
    DMatrix var;                   // variance estimate (n x 1), zeros unless calc_var
};
// iter_lim < 0 is the reference's `None` (2 n); x0 == nullptr its `None`
inline LsqrOutput lsqr(const DMatrix& a, const DMatrix& b, double damp, double atol, double btol, double conlim, long iter_lim,
                       bool calc_var, const DMatrix* x0) {
    const size_t n = a.ncols();
    if (b.nrows() != a.nrows() || (x0 && x0->nrows() != n)) throw std::invalid_argument("lsqr: shapes do not conform");   // nalgebra panics
    LsqrOutput o{DMatrix(n, 1), 0, 0, 0, 0, 0, 0, {}, 0, DMatrix(n, 1)};
    o.arnorms.assign((size_t)std::max<long>(iter_lim < 0 ? 2 * (long)n : iter_lim, 1), 0.0);
    rnla_lsqr_result r{};
    errors::check(rnla_lsqr(a.as_ptr(), (int64_t)a.nrows(), (int64_t)n, b.as_ptr(), damp, atol, btol, conlim, (int64_t)iter_lim,
                            calc_var ? 1 : 0, x0 ? x0->as_ptr() : nullptr, o.x.as_mut_ptr(), &r, o.arnorms.data(),
                            (int64_t)o.arnorms.size(), o.var.as_mut_ptr()));
    o.istop = (size_t)r.istop; o.itn = (size_t)r.itn; o.r1norm = r.r1norm; o.r2norm = r.r2norm; o.anorm = r.anorm; o.acond = r.acond;
    o.xnorm = r.xnorm;
    o.arnorms.resize((size_t)std::min<int64_t>(r.n_arnorms, (int64_t)o.arnorms.size()));
    return o;
}
}  // namespace solvers

namespace id {
using sketch::MatrixAttribute;
// src/id.rs:272-318
inline std::pair<DMatrix, std::vector<size_t>> osid_qrcp(const DMatrix& y, size_t k, MatrixAttribute attr) {
    const bool col = attr == MatrixAttribute::Column;
    DMatrix x(col ? std::max<size_t>(k, 1) : y.nrows(), col ? y.ncols() : std::max<size_t>(k, 1));
    std::vector<int64_t> j(std::max<size_t>(k, 1));
    errors::check(rnla_osid_qrcp(y.as_ptr(), (int64_t)y.nrows(), (int64_t)y.ncols(), (int64_t)k, (int32_t)attr, x.as_mut_ptr(), j.data()));
    return {std::move(x), to_usize(j, k)};
}
// :217-249
inline std::pair<DMatrix, std::vector<size_t>> osid_randomised(const DMatrix& a, size_t k, MatrixAttribute attr) {
    const bool col = attr == MatrixAttribute::Column;
    DMatrix x(col ? std::max<size_t>(k, 1) : a.nrows(), col ? a.ncols() : std::max<size_t>(k, 1));
    std::vector<int64_t> j(std::max<size_t>(k, 1));
    errors::check(rnla_osid_randomised(a.as_ptr(), (int64_t)a.nrows(), (int64_t)a.ncols(), (int64_t)k, (int32_t)attr, x.as_mut_ptr(), j.data()));
    return {std::move(x), to_usize(j, k)};
}
inline std::tuple<DMatrix, std::vector<size_t>, std::vector<size_t>, DMatrix> two_sided_impl(const DMatrix& a, size_t k, int randomised) {
    DMatrix z(a.nrows(), std::max<size_t>(k, 1)), x(std::max<size_t>(k, 1), a.ncols());
    std::vector<int64_t> i(std::max<size_t>(k, 1)), j(std::max<size_t>(k, 1));
    errors::check(rnla_two_sided_id(a.as_ptr(), (int64_t)a.nrows(), (int64_t)a.ncols(), (int64_t)k, randomised, z.as_mut_ptr(), i.data(), j.data(), x.as_mut_ptr()));
    return {std::move(z), to_usize(i, k), to_usize(j, k), std::move(x)};
}
// :118-129 and :94-101
inline auto two_sided_id(const DMatrix& a, size_t k) { return two_sided_impl(a, k, 0); }
inline auto two_sided_id_randomised(const DMatrix& a, size_t k) { return two_sided_impl(a, k, 1); }
inline std::tuple<std::vector<size_t>, DMatrix, std::vector<size_t>> cur_impl(const DMatrix& a, size_t k, int randomised) {
    DMatrix u(std::max<size_t>(k, 1), std::max<size_t>(k, 1));
    std::vector<int64_t> i(std::max<size_t>(k, 1)), j(std::max<size_t>(k, 1));
    errors::check(rnla_cur(a.as_ptr(), (int64_t)a.nrows(), (int64_t)a.ncols(), (int64_t)k, randomised, j.data(), u.as_mut_ptr(), i.data()));
    return {to_usize(j, k), std::move(u), to_usize(i, k)};
}
// :34-71 and :154-193
inline auto cur(const DMatrix& a, size_t k) { return cur_impl(a, k, 0); }
inline auto cur_randomised(const DMatrix& a, size_t k) { return cur_impl(a, k, 1); }
}  // namespace id

}  // namespace randblas
