// The callers of the sketch that SURVEY.md section 8(f) ranks after the hot path, end to end on the device, built from the
// kernels of the path (streaming GEMMs, sketch operators, CholeskyQR2 panels, Jacobi core, CGLS) plus the column-pivoted
// QR of pivot.cu:
//   sap_chol_qrcp                       reference src/cqrrpt.rs:27-58
//   sketched_least_squares_qr / _svd    reference src/sketch_and_solve.rs:24-65
//   osid_qrcp, osid_randomised, two_sided_id(_randomised), cur(_randomised)      reference src/id.rs:34-318
//   sketch_saddle_point_precondition    reference src/sketch_and_precondition.rs:150-216
// Single GPU (the matrices these drivers factor are sketches or selected rows/columns).
#include "drivers.cuh"
#include "gemm.cuh"
#include "panel.cuh"
#include <algorithm>
#include <cmath>
#include <vector>

namespace rnla {

namespace {

rnla_status single_rank(const char* who) {
    if (ctx().nranks > 1) return fail(RNLA_ERR_INVALID_PARAMETERS, std::string(who) + ": single-GPU driver (destroy the communicator first)");
    return RNLA_OK;
}

// host copy of diag(R)[0..k)
rnla_status read_diag(const double* R, int64_t ld, int64_t k, std::vector<double>& out) {
    out.assign((size_t)std::max<int64_t>(k, 0), 0.0);
    if (k <= 0) return RNLA_OK;
    RNLA_CUDA(cudaMemcpy2DAsync(out.data(), 8, R, (size_t)(ld + 1) * 8, 8, (size_t)k, cudaMemcpyDeviceToHost, ctx().stream));
    RNLA_CUDA(cudaStreamSynchronize(ctx().stream));
    return RNLA_OK;
}

// thin SVD of X (rows x n, rows >= n, n <= 1024): X is overwritten by the orthonormal factor Q of X = Q R, and
// R = Ur diag(sig) Vr^T (one-sided Jacobi on R^T).  So U = Q Ur, V = Vr.
rnla_status thin_svd(double* X, int64_t ldx, int64_t rows, int n, double* Ur, double* sig, double* Vr, const char* who) {
    Ctx& c = ctx();
    if (n > 1024) return fail(RNLA_ERR_INVALID_DIMENSIONS, std::string(who) + " (device): n <= 1024 (size of the on-device SVD core)");
    DevBuf R, work, info;
    RNLA_CUDA(R.alloc((size_t)n * n * 8)); RNLA_CUDA(work.alloc(jacobi_svd_work_doubles(n) * 8)); RNLA_CUDA(info.alloc(8));
    int64_t def = 0;
    RNLA_TRY(dev_qr_blocked(X, ldx, rows, n, R.d(), &def));
    RNLA_CUDA(jacobi_svd(R.d(), n, n, Ur, n, sig, Vr, n, work.d(), info.as<int>(), c.stream, 1));
    int h[2];
    RNLA_CUDA(cudaMemcpyAsync(h, info.p, 8, cudaMemcpyDeviceToHost, c.stream));
    RNLA_CUDA(cudaStreamSynchronize(c.stream));
    g_last_jacobi_sweeps = h[0];
    if (h[1]) return fail(RNLA_ERR_MATRIX_DECOMPOSITION, std::string(who) + ": SVD of the sketch did not converge");
    return RNLA_OK;
}

// P (rows x p) = pinv(T)^T for a tall T (rows x p, rows >= p, p <= 1024; destroyed): T = (Q Ur) S Vr^T, pinv(T)^T = Q Ur S^+ Vr^T,
// S^+ the reciprocals of the singular values > 0 (nalgebra `pseudo_inverse(0.0)`, reference src/id.rs:52,69)
rnla_status pinv_transposed_tall(double* T, int64_t ldt, int64_t rows, int p, double* P, int64_t ldp) {
    Ctx& c = ctx();
    DevBuf Ur, sig, Vr, Vrt, M2, sinv;
    RNLA_CUDA(Ur.alloc((size_t)p * p * 8)); RNLA_CUDA(sig.alloc((size_t)p * 8)); RNLA_CUDA(Vr.alloc((size_t)p * p * 8));
    RNLA_CUDA(Vrt.alloc((size_t)p * p * 8)); RNLA_CUDA(M2.alloc((size_t)p * p * 8)); RNLA_CUDA(sinv.alloc((size_t)p * 8));
    RNLA_TRY(thin_svd(T, ldt, rows, p, Ur.d(), sig.d(), Vr.d(), "pseudo_inverse"));
    // sinv = 1 / sigma where sigma > 0 else 0: diag_solve on a vector of ones
    std::vector<double> ones((size_t)p, 1.0);
    RNLA_CUDA(cudaMemcpyAsync(sinv.p, ones.data(), (size_t)p * 8, cudaMemcpyHostToDevice, c.stream));
    RNLA_TRY(dev_diag_solve(sig.d(), p, sinv.d()));
    RNLA_CUDA(scale_columns(Ur.d(), p, p, p, sinv.d(), c.stream));
    RNLA_CUDA(transpose_matrix(Vr.d(), p, Vrt.d(), p, p, p, c.stream));
    RNLA_TRY(dev_gemm_nn(Ur.d(), p, p, p, Vrt.d(), p, p, M2.d(), p));
    RNLA_TRY(dev_gemm_nn(T, ldt, rows, p, M2.d(), p, p, P, ldp));
    RNLA_CUDA(cudaStreamSynchronize(c.stream));      // `ones` leaves scope
    return RNLA_OK;
}

}  // namespace

// ---------------------------------------------------------------------------------------------- src/cqrrpt.rs:27-58
// Q: m x n buffer (ldq), first k columns valid.  R: n x n buffer (ldr), k x n valid.  dJ: n int64 on the device.
rnla_status dev_sap_chol_qrcp(const double* A, int64_t lda, int64_t m, int64_t n, int64_t d, int kind, int dist_or_width, int zeta,
                              uint64_t seed, double* Q, int64_t ldq, double* R, int64_t ldr, int64_t* dJ, int64_t* k_out) {
    Ctx& c = ctx();
    phases_reset();
    RNLA_TRY(single_rank("sap_chol_qrcp"));
    if (!(n <= d && d <= m) || n <= 0) return fail(RNLA_ERR_INVALID_PARAMETERS, "d must satisfy n \xe2\x89\xa4 d \xe2\x89\xaa m");   // :29
    if (n > 16384) return fail(RNLA_ERR_INVALID_DIMENSIONS, "sap_chol_qrcp (device): n <= 16384");
    DevBuf Ask, Rinv, W, Rpre;
    RNLA_CUDA(Ask.alloc((size_t)d * n * 8));
    RNLA_TRY(dev_sketch_apply(kind, dist_or_width, seed, d, zeta, A, lda, m, n, 0, Ask.d(), d));        // :31-33
    RNLA_TRY(dev_qrcp(Ask.d(), d, d, n, n, dJ, nullptr, 0, 0));                                          // :35
    std::vector<double> diag;
    RNLA_TRY(read_diag(Ask.d(), d, n, diag));
    int64_t k = 0;
    for (double v : diag) if (std::fabs(v) > 1e-10) ++k;                                                 // :37-43
    *k_out = k;
    if (k == 0) return RNLA_OK;
    for (int64_t i = 0; i < k; ++i)
        if (diag[(size_t)i] == 0.0) return fail(RNLA_ERR_SINGULAR_MATRIX, "sap_chol_qrcp: leading block of the sketch's R factor is singular");
    const int kk = (int)k;
    RNLA_CUDA(Rinv.alloc((size_t)k * k * 8)); RNLA_CUDA(W.alloc((size_t)n * k * 8)); RNLA_CUDA(Rpre.alloc((size_t)k * k * 8));
    {
        PhaseScope ph("cqrrpt:precondition");
        RNLA_TRY(dev_tri_inv_blocked(Ask.d(), d, kk, Rinv.d(), k));                                      // :47
        // A[:, J[:k]] Rinv = A W with W(J[i], :) = Rinv(i, :): no permuted copy of the tall matrix          :46, :48
        RNLA_CUDA(axpby_matrix(0.0, nullptr, 0, 0.0, nullptr, 0, W.d(), n, n, k, c.stream));
        RNLA_TRY(dev_scatter_rows(Rinv.d(), k, k, dJ, W.d(), n));
        RNLA_TRY(dev_gemm_nn(A, lda, m, n, W.d(), n, k, Q, ldq));
    }
    {
        // the reference takes one Cholesky of A_pre^T A_pre (:50-53); the blocked CholeskyQR2 panels produce the same
        // factors (QR with a positive diagonal is unique) with the orthogonality of two passes
        PhaseScope ph("cqrrpt:cholqr(A_pre)");
        int64_t def = 0;
        RNLA_TRY(dev_qr_blocked(Q, ldq, m, kk, Rpre.d(), &def));
        if (def) return fail(RNLA_ERR_MATRIX_DECOMPOSITION, "Cholesky decomposition failed");            // :51
    }
    RNLA_TRY(dev_gemm_nn(Rpre.d(), k, k, k, Ask.d(), d, n, R, ldr));                                     // :55
    RNLA_CUDA(cudaStreamSynchronize(c.stream));
    return RNLA_OK;
}

// --------------------------------------------------------------------------------------- src/sketch_and_solve.rs:24-65
rnla_status dev_sketched_least_squares(int which, const double* A, int64_t lda, int64_t m, int64_t n, const double* b, int kind,
                                       int dist_or_width, int zeta, uint64_t seed, double* x) {
    Ctx& c = ctx();
    phases_reset();
    RNLA_TRY(single_rank("sketched_least_squares"));
    const int64_t d = m / 4;                                                                             // :26, :56
    if (n <= 0 || d < n) return fail(RNLA_ERR_INVALID_DIMENSIONS, "sketched_least_squares: the sketch has rows/4 rows, fewer than the columns of a");
    if (n > 16384) return fail(RNLA_ERR_INVALID_DIMENSIONS, "sketched_least_squares (device): n <= 16384");
    const int nn = (int)n;
    DevBuf Ask, bsk, z;
    RNLA_CUDA(Ask.alloc((size_t)d * n * 8)); RNLA_CUDA(bsk.alloc((size_t)d * 8)); RNLA_CUDA(z.alloc((size_t)n * 8));
    RNLA_TRY(dev_sketch_apply(kind, dist_or_width, seed, d, zeta, A, lda, m, n, 0, Ask.d(), d));         // :27 / :57
    RNLA_TRY(dev_sketch_apply(kind, dist_or_width, seed, d, zeta, b, m, m, 1, 0, bsk.d(), d));           // :28 / :58
    if (which == 0) {
        PhaseScope ph("solve:qr");
        DevBuf R;
        RNLA_CUDA(R.alloc((size_t)n * n * 8));
        int64_t def = 0;
        RNLA_TRY(dev_qr_blocked(Ask.d(), d, d, nn, R.d(), &def));                                        // :29
        RNLA_TRY(dev_gemm_tn(Ask.d(), d, d, n, bsk.d(), d, 1, z.d(), n, false));                         // :30
        RNLA_TRY(dev_backsolve_upper(R.d(), n, nn, z.d(), x));                                           // :31
    } else {
        PhaseScope ph("solve:svd");
        DevBuf Ur, sig, Vr, z2;
        RNLA_CUDA(Ur.alloc((size_t)n * n * 8)); RNLA_CUDA(sig.alloc((size_t)n * 8)); RNLA_CUDA(Vr.alloc((size_t)n * n * 8));
        RNLA_CUDA(z2.alloc((size_t)n * 8));
        RNLA_TRY(thin_svd(Ask.d(), d, d, nn, Ur.d(), sig.d(), Vr.d(), "sketched_least_squares_svd"));    // :59-62
        RNLA_TRY(dev_gemm_tn(Ask.d(), d, d, n, bsk.d(), d, 1, z.d(), n, false));
        RNLA_TRY(dev_small_gemv(Ur.d(), n, nn, 1, z.d(), z2.d()));                                       // :63  u^T b_sk
        RNLA_TRY(dev_diag_solve(sig.d(), nn, z2.d()));                                                   // :64
        RNLA_TRY(dev_small_gemv(Vr.d(), n, nn, 0, z2.d(), x));                                           // :65  v x
    }
    RNLA_CUDA(cudaStreamSynchronize(c.stream));
    return RNLA_OK;
}

// ------------------------------------------------------------------------------------------------ src/id.rs:272-318
// attr RNLA_COLUMN: Y (l x w) ~ Y[:, J] X with X k x w;  RNLA_ROW: Y ~ X Y[J, :] with X l x k.  dJ: k int64 on the device.
rnla_status dev_osid_qrcp(const double* Y, int64_t ldy, int64_t l, int64_t w, int64_t k, int attr, double* X, int64_t ldx, int64_t* dJ) {
    Ctx& c = ctx();
    if (k <= 0) return fail(RNLA_ERR_INVALID_PARAMETERS, "k must be positive)");                         // :278
    if (k > std::min(l, w)) return fail(RNLA_ERR_INVALID_PARAMETERS, "k must be <= min(l,w)");           // :279
    if (k > 16384) return fail(RNLA_ERR_INVALID_DIMENSIONS, "osid_qrcp (device): k <= 16384");
    if (attr == RNLA_ROW) {                                                                              // :309-314
        DevBuf Yt, Xt;
        RNLA_CUDA(Yt.alloc((size_t)l * w * 8)); RNLA_CUDA(Xt.alloc((size_t)k * l * 8));
        RNLA_CUDA(transpose_matrix(Y, ldy, Yt.d(), w, l, w, c.stream));
        RNLA_TRY(dev_osid_qrcp(Yt.d(), w, w, l, k, RNLA_COLUMN, Xt.d(), k, dJ));
        RNLA_CUDA(transpose_matrix(Xt.d(), k, X, ldx, k, l, c.stream));
        return RNLA_OK;
    }
    DevBuf R, perm, Rinv, T;
    RNLA_CUDA(R.alloc((size_t)l * w * 8)); RNLA_CUDA(perm.alloc((size_t)w * 8));
    RNLA_CUDA(copy_matrix(Y, ldy, R.d(), l, l, w, c.stream));
    RNLA_TRY(dev_qrcp(R.d(), l, l, w, k, perm.as<int64_t>(), nullptr, 0, 0));                            // :283
    std::vector<double> diag;
    RNLA_TRY(read_diag(R.d(), l, k, diag));
    for (double v : diag)
        if (v == 0.0) return fail(RNLA_ERR_SINGULAR_MATRIX, "osid_qrcp: R1 is singular (rank of the matrix is below k)");   // unwrap :290
    PhaseScope ph("id:interpolation");
    RNLA_CUDA(Rinv.alloc((size_t)k * k * 8)); RNLA_CUDA(T.alloc((size_t)k * std::max<int64_t>(w - k, 1) * 8));
    RNLA_TRY(dev_tri_inv_blocked(R.d(), l, (int)k, Rinv.d(), k));
    if (w > k) RNLA_TRY(dev_gemm_nn(Rinv.d(), k, k, k, R.d() + k * l, l, w - k, T.d(), k));              // T = R1^-1 R2   :285-290
    RNLA_TRY(dev_build_interp(T.d(), k, k, w, perm.as<int64_t>(), X, ldx));                              // :293-308
    RNLA_CUDA(cudaMemcpyAsync(dJ, perm.p, (size_t)k * 8, cudaMemcpyDeviceToDevice, c.stream));
    return RNLA_OK;
}

// ------------------------------------------------------------------------------------------------ src/id.rs:217-249
rnla_status dev_osid_randomised(const double* A, int64_t lda, int64_t m, int64_t n, int64_t k, int attr, const rnla_options& o,
                                double* X, int64_t ldx, int64_t* dJ) {
    Ctx& c = ctx();
    if (k <= 0) return fail(RNLA_ERR_INVALID_PARAMETERS, "k must be positive)");                         // :223
    if (k > std::min(m, n)) return fail(RNLA_ERR_INVALID_PARAMETERS, "k must be <= min(l,w)");           // :224
    if (attr == RNLA_COLUMN) {
        DevBuf Y;
        RNLA_CUDA(Y.alloc((size_t)k * n * 8));
        RNLA_TRY(dev_sketch_apply(RNLA_SKETCH_DENSE, RNLA_GAUSSIAN, o.seed, k, 0, A, lda, m, n, 0, Y.d(), k));   // :238-243
        return dev_osid_qrcp(Y.d(), k, k, n, k, RNLA_COLUMN, X, ldx, dJ);                                // :246
    }
    // Row: `a * s_matrix.transpose()` with s_matrix = tsog1(a, k, 2, 1) (n x k) conforms only when n == k (:230-233)
    if (n != k) return fail(RNLA_ERR_INVALID_DIMENSIONS, "osid_randomised(Row): a * tsog1(a, k, 2, 1)^T needs a.ncols() == k");
    DevBuf S, St, Y;
    RNLA_CUDA(S.alloc((size_t)n * k * 8)); RNLA_CUDA(St.alloc((size_t)n * k * 8)); RNLA_CUDA(Y.alloc((size_t)m * k * 8));
    ShardInfo sh{m, 0, m};
    RNLA_TRY(dev_tsog1(A, lda, sh, n, (int)k, 2, 1, o, S.d()));                                          // :230
    RNLA_CUDA(transpose_matrix(S.d(), n, St.d(), k, n, k, c.stream));
    RNLA_TRY(dev_gemm_nn(A, lda, m, n, St.d(), k, n, Y.d(), m));                                         // :233
    return dev_osid_qrcp(Y.d(), m, m, k, k, RNLA_ROW, X, ldx, dJ);                                       // :236
}

// -------------------------------------------------------------------------------- src/id.rs:118-129 and :94-101
rnla_status dev_two_sided_id(int randomised, const double* A, int64_t lda, int64_t m, int64_t n, int64_t k, const rnla_options& o,
                             double* Z, int64_t ldz, int64_t* dI, int64_t* dJ, double* X, int64_t ldx) {
    phases_reset();
    RNLA_TRY(single_rank("two_sided_id"));
    if (randomised) RNLA_TRY(dev_osid_randomised(A, lda, m, n, k, RNLA_COLUMN, o, X, ldx, dJ));
    else RNLA_TRY(dev_osid_qrcp(A, lda, m, n, k, RNLA_COLUMN, X, ldx, dJ));
    DevBuf Ac;
    RNLA_CUDA(Ac.alloc((size_t)m * k * 8));
    RNLA_TRY(dev_gather_columns(A, lda, m, dJ, k, Ac.d(), m));
    if (randomised) RNLA_TRY(dev_osid_randomised(Ac.d(), m, m, k, k, RNLA_ROW, o, Z, ldz, dI));
    else RNLA_TRY(dev_osid_qrcp(Ac.d(), m, m, k, k, RNLA_ROW, Z, ldz, dI));
    RNLA_CUDA(cudaStreamSynchronize(ctx().stream));
    return RNLA_OK;
}

// --------------------------------------------------------------------------------- src/id.rs:34-71 and :154-193
rnla_status dev_cur(int randomised, const double* A, int64_t lda, int64_t m, int64_t n, int64_t k, const rnla_options& o,
                    int64_t* dJ, double* U, int64_t ldu, int64_t* dI) {
    Ctx& c = ctx();
    phases_reset();
    RNLA_TRY(single_rank("cur"));
    if (k <= 0) return fail(RNLA_ERR_INVALID_PARAMETERS, "k must be positive)");
    if (k > std::min(m, n)) return fail(RNLA_ERR_INVALID_PARAMETERS, "k must be <= min(l,w)");
    if (k > 1024) return fail(RNLA_ERR_INVALID_DIMENSIONS, "cur (device): k <= 1024 (size of the on-device SVD core)");
    if (m >= n) {
        DevBuf X, Ac, Act, perm, Ar, Art, P;
        RNLA_CUDA(X.alloc((size_t)k * n * 8)); RNLA_CUDA(Ac.alloc((size_t)m * k * 8)); RNLA_CUDA(Act.alloc((size_t)m * k * 8));
        RNLA_CUDA(perm.alloc((size_t)m * 8)); RNLA_CUDA(Ar.alloc((size_t)k * n * 8)); RNLA_CUDA(Art.alloc((size_t)k * n * 8));
        RNLA_CUDA(P.alloc((size_t)k * n * 8));
        if (randomised) RNLA_TRY(dev_osid_randomised(A, lda, m, n, k, RNLA_COLUMN, o, X.d(), k, dJ));    // :162
        else RNLA_TRY(dev_osid_qrcp(A, lda, m, n, k, RNLA_COLUMN, X.d(), k, dJ));                        // :42
        RNLA_TRY(dev_gather_columns(A, lda, m, dJ, k, Ac.d(), m));                                       // :44
        RNLA_CUDA(transpose_matrix(Ac.d(), m, Act.d(), k, m, k, c.stream));
        RNLA_TRY(dev_qrcp(Act.d(), k, k, m, k, perm.as<int64_t>(), nullptr, 0, 0));                      // :46
        RNLA_CUDA(cudaMemcpyAsync(dI, perm.p, (size_t)k * 8, cudaMemcpyDeviceToDevice, c.stream));       // :48
        RNLA_TRY(dev_gather_rows(A, lda, n, dI, k, Ar.d(), k));                                          // :51
        RNLA_CUDA(transpose_matrix(Ar.d(), k, Art.d(), n, k, n, c.stream));
        RNLA_TRY(pinv_transposed_tall(Art.d(), n, n, (int)k, P.d(), n));                                 // pinv(A[I, :]) = P (n x k)
        RNLA_TRY(dev_gemm_nn(X.d(), k, k, n, P.d(), n, k, U, ldu));                                      // :52
    } else {
        DevBuf Z, Ar, perm, Ac, P, Ut;
        RNLA_CUDA(Z.alloc((size_t)k * m * 8)); RNLA_CUDA(Ar.alloc((size_t)k * n * 8)); RNLA_CUDA(perm.alloc((size_t)n * 8));
        RNLA_CUDA(Ac.alloc((size_t)m * k * 8)); RNLA_CUDA(P.alloc((size_t)m * k * 8)); RNLA_CUDA(Ut.alloc((size_t)k * k * 8));
        if (randomised) {
            // osid_randomised(a^T, k, Column) (:177): Y = S a^T = (a S^T)^T with S^T (n x k) the same dense operator, so the
            // transposed copy of a (:176) is never formed
            DevBuf St, Yt, Y;
            RNLA_CUDA(St.alloc((size_t)n * k * 8)); RNLA_CUDA(Yt.alloc((size_t)m * k * 8)); RNLA_CUDA(Y.alloc((size_t)m * k * 8));
            RNLA_CUDA(fill_philox(RNLA_GAUSSIAN, o.seed, 3 /* STREAM_SKETCH_DENSE */, n, k, 0, St.d(), n, c.stream));
            RNLA_TRY(dev_gemm_nn(A, lda, m, n, St.d(), n, k, Yt.d(), m));
            RNLA_CUDA(transpose_matrix(Yt.d(), m, Y.d(), k, m, k, c.stream));
            RNLA_TRY(dev_osid_qrcp(Y.d(), k, k, m, k, RNLA_COLUMN, Z.d(), k, dI));
        } else {
            DevBuf At;
            RNLA_CUDA(At.alloc((size_t)m * n * 8));
            RNLA_CUDA(transpose_matrix(A, lda, At.d(), n, m, n, c.stream));                              // :56
            RNLA_TRY(dev_osid_qrcp(At.d(), n, n, m, k, RNLA_COLUMN, Z.d(), k, dI));                      // :57
        }
        RNLA_TRY(dev_gather_rows(A, lda, n, dI, k, Ar.d(), k));                                          // :59
        RNLA_TRY(dev_qrcp(Ar.d(), k, k, n, k, perm.as<int64_t>(), nullptr, 0, 0));                       // :62
        RNLA_CUDA(cudaMemcpyAsync(dJ, perm.p, (size_t)k * 8, cudaMemcpyDeviceToDevice, c.stream));       // :65
        RNLA_TRY(dev_gather_columns(A, lda, m, dJ, k, Ac.d(), m));                                       // :68
        RNLA_TRY(pinv_transposed_tall(Ac.d(), m, m, (int)k, P.d(), m));                                  // pinv(A[:, J]) = P^T   :69
        RNLA_TRY(dev_gemm_nn(Z.d(), k, k, m, P.d(), m, k, Ut.d(), k));                                   // (pinv z^T)^T = z P
        RNLA_CUDA(transpose_matrix(Ut.d(), k, U, ldu, k, k, c.stream));                                  // :70
    }
    RNLA_CUDA(cudaStreamSynchronize(c.stream));
    return RNLA_OK;
}

// ------------------------------------------------------------------------- src/sketch_and_precondition.rs:150-216
// Dense operator only (the reference's own): S^T (m x d) is materialised once, because the driver applies S (to A and to
// b_mod) and S^T (to a d-vector, :199-203).  The preconditioned matrix a M (:192) is never formed: CGLS runs in operator form.
// cvec may be NULL (`c.is_empty()`).  x: n, y: m.
rnla_status dev_saddle_point(const double* A, int64_t lda, int64_t m, int64_t n, const double* b, const double* cvec, double mu,
                             double epsilon, int64_t maxit, double sampling_factor, int dist, uint64_t seed, double* x, double* y,
                             int64_t* iters_out, int32_t* converged_out) {
    Ctx& c = ctx();
    phases_reset();
    RNLA_TRY(single_rank("sketch_saddle_point_precondition"));
    if (n <= 0 || n > 1024) return fail(RNLA_ERR_INVALID_DIMENSIONS, "sketch_saddle_point_precondition (device): 1 <= n <= 1024 (size of the on-device SVD core)");
    int64_t d = (int64_t)std::floor(sampling_factor * (double)n);                                        // :172
    d = std::min(std::max<int64_t>(d, 1), m);
    const int nn = (int)n;
    DevBuf St, Ask, Ur, sig, Vr, w, M, bmod, sb, z1, z, t1, t2, tm;
    RNLA_CUDA(St.alloc((size_t)m * d * 8)); RNLA_CUDA(Ask.alloc((size_t)d * n * 8));
    RNLA_CUDA(Ur.alloc((size_t)n * n * 8)); RNLA_CUDA(sig.alloc((size_t)n * 8)); RNLA_CUDA(Vr.alloc((size_t)n * n * 8));
    RNLA_CUDA(w.alloc((size_t)n * 8)); RNLA_CUDA(M.alloc((size_t)n * n * 8)); RNLA_CUDA(bmod.alloc((size_t)m * 8));
    RNLA_CUDA(sb.alloc((size_t)d * 8)); RNLA_CUDA(z1.alloc((size_t)n * 8)); RNLA_CUDA(z.alloc((size_t)n * 8));
    {
        PhaseScope ph("sketch:dense");
        RNLA_CUDA(fill_philox(dist, seed, 3 /* STREAM_SKETCH_DENSE */, m, d, 0, St.d(), m, c.stream));   // :173
        RNLA_TRY(dev_gemm_tn(St.d(), m, m, d, A, lda, n, Ask.d(), d, false));                            // :176
    }
    {
        PhaseScope ph("precond:svd(A_sk)");
        RNLA_TRY(thin_svd(Ask.d(), d, d, nn, Ur.d(), sig.d(), Vr.d(), "sketch_saddle_point_precondition"));   // :179-182
        if (!(mu > 0.0)) {                                                                               // :188
            std::vector<double> hs((size_t)n);
            RNLA_CUDA(cudaMemcpyAsync(hs.data(), sig.p, (size_t)n * 8, cudaMemcpyDeviceToHost, c.stream));
            RNLA_CUDA(cudaStreamSynchronize(c.stream));
            for (double s : hs)
                if (!(s > 1e-10)) return fail(RNLA_ERR_INVALID_DIMENSIONS,
                    "sketch_saddle_point_precondition: with mu = 0 the sketch must have full column rank (the reference's shapes do not conform otherwise)");
        }
        RNLA_TRY(dev_saddle_weights(sig.d(), nn, mu, w.d()));
        RNLA_CUDA(copy_matrix(Vr.d(), n, M.d(), n, n, n, c.stream));
        RNLA_CUDA(scale_columns(M.d(), n, n, n, w.d(), c.stream));                                       // M = V diag(w)   :185-190
    }
    RNLA_CUDA(cudaMemcpyAsync(bmod.p, b, (size_t)m * 8, cudaMemcpyDeviceToDevice, c.stream));            // :194
    if (cvec) {                                                                                          // :195-206
        PhaseScope ph("saddle:b_mod");
        RNLA_CUDA(t1.alloc((size_t)n * 8)); RNLA_CUDA(t2.alloc((size_t)d * 8)); RNLA_CUDA(tm.alloc((size_t)m * 8));
        RNLA_TRY(dev_small_gemv(Vr.d(), n, nn, 1, cvec, z1.d()));                                        // vt c            :196
        RNLA_TRY(dev_mul_vec(w.d(), nn, z1.d()));
        RNLA_TRY(dev_small_gemv(Ur.d(), n, nn, 0, z1.d(), t1.d()));                                      // u = Q Ur
        RNLA_TRY(dev_gemv_n(Ask.d(), d, d, n, t1.d(), t2.d()));
        RNLA_TRY(dev_gemv_n(St.d(), m, m, d, t2.d(), tm.d()));                                           // s^T (..)        :199-203
        RNLA_TRY(dev_axpby_vec(-1.0, tm.d(), 1.0, bmod.d(), m));                                         // :205
    }
    {
        PhaseScope ph("saddle:z0");
        RNLA_TRY(dev_gemv_t(St.d(), m, m, d, bmod.d(), sb.d()));                                         // s b_mod
        RNLA_TRY(dev_gemv_t(Ask.d(), d, d, n, sb.d(), z1.d()));
        RNLA_TRY(dev_small_gemv(Ur.d(), n, nn, 1, z1.d(), z.d()));                                       // z0 = u^T (s b_mod)   :208
    }
    St.release();
    int64_t it = 0; int32_t conv = 0;
    RNLA_TRY(dev_cgls_operator(A, lda, m, n, bmod.d(), M.d(), z.d(), epsilon, maxit, &it, &conv));        // :211
    RNLA_TRY(dev_small_gemv(M.d(), n, nn, 0, z.d(), x));                                                 // x = m z   :213
    RNLA_TRY(dev_gemv_n(A, lda, m, n, x, y));
    RNLA_TRY(dev_axpby_vec(1.0, b, -1.0, y, m));                                                         // y = b - a x   :214
    RNLA_CUDA(cudaStreamSynchronize(c.stream));
    if (iters_out) *iters_out = it;
    if (converged_out) *converged_out = conv;
    return RNLA_OK;
}

}  // namespace rnla
