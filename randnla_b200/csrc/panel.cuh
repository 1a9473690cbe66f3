// Small dense core and panel utilities: everything that is O(l^2)..O(l^3)
// or a single sweep over an (rows x l) panel.  Column-major, explicit leading dimensions.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace rnla {

// K0: out(r, c) = omega(row_off + r, c), r < rows, c < cols   (rng.cuh map)
cudaError_t fill_philox(int dist, uint64_t seed, uint32_t stream, int64_t rows, int64_t cols,
                        int64_t row_off, double* out, int64_t ld, cudaStream_t st);
// reference-compatible sequential ThreeFry2x64 stream (src/sketch.rs:112-127): sample t -> entry
// (t % rows, t / rows), one u64 per sample; Uniform / Rademacher only.
cudaError_t fill_threefry(int dist, uint64_t key0, uint64_t key1, int64_t rows, int64_t cols,
                          double* out, int64_t ld, cudaStream_t st);
cudaError_t philox_blocks(int64_t n, const uint32_t* ctr, const uint32_t* key, uint32_t* out, cudaStream_t st);
cudaError_t threefry_blocks(int64_t n, const uint64_t* ctr, const uint64_t* key, uint64_t* out, cudaStream_t st);

// In-place upper Cholesky G = R^T R of a p x p Gram matrix with rank-deficiency detection.
// Column j is declared deficient when its pivot d_j <= tol2 * G_jj (or G_jj == 0); then R_jj = 1,
// R_j,j+1.. = 0 and flags[j] = 1.  info[0] = number of deficient columns, info[1] = 1 if non-finite,
// info[2] = number of exactly-zero columns (flag 2).  info must hold 3 ints.
cudaError_t chol_upper(double* G, int64_t ld, int p, double tol2, int* flags, int* info, cudaStream_t st);
// Rinv = R^-1 (upper triangular, p x p)
cudaError_t tri_inv_upper(const double* R, int64_t ldr, int p, double* Rinv, int64_t ldi, cudaStream_t st);
// C = op(A) * B for small p x p matrices; zero_diag_flags: rows j of B's *left factor* with flags[j] set get
// their diagonal entry of A treated as 0 (see orth: R'' = R with R_jj := 0 for replaced columns)
cudaError_t small_gemm(const double* A, int64_t lda, const double* B, int64_t ldb, double* C, int64_t ldc,
                       int M, int N, int K, cudaStream_t st);
cudaError_t zero_flagged_diag(double* R, int64_t ld, int p, const int* flags, cudaStream_t st);
// X(:, j) = e_{p_j} for flagged columns; p_j = global row index target[j]; rows are [row_off, row_off+rows)
cudaError_t replace_columns(double* X, int64_t ld, int64_t rows, int64_t row_off, int p, const int* flags,
                            const int64_t* target, cudaStream_t st);
// device-driven replace_columns + zero_flagged_diag (no host round trip): see panel.cu.  info = chol_upper's (3 ints),
// state[0] = attempt counter (zero it before the first pass), hist receives 3 ints per pass.
cudaError_t orth_fixup(double* X, int64_t ld, int64_t rows, int64_t row_off, int64_t rows_global, int p, const int* flags,
                       const int* info, int* state, int* hist, int pass, double* G, int64_t ldg, cudaStream_t st);
cudaError_t set_identity(double* X, int64_t ld, int64_t rows, int64_t cols, cudaStream_t st);
cudaError_t copy_matrix(const double* src, int64_t lds, double* dst, int64_t ldd, int64_t rows, int64_t cols, cudaStream_t st);
cudaError_t transpose_matrix(const double* src, int64_t lds, double* dst, int64_t ldd, int64_t rows, int64_t cols, cudaStream_t st);
// dst = a * x + b * y (elementwise over rows x cols)
cudaError_t axpby_matrix(double a, const double* x, int64_t ldx, double b, const double* y, int64_t ldy,
                         double* dst, int64_t ldd, int64_t rows, int64_t cols, cudaStream_t st);
// out[0] += sum of squares (out must be zeroed by the caller); deterministic two-stage reduction
cudaError_t sumsq(const double* X, int64_t ld, int64_t rows, int64_t cols, double* out, double* scratch, int nscratch, cudaStream_t st);
// exact symmetry test: flag[0] = 1 if any A(i,j) != A(j,i)   (reference lora_drivers.rs:106)
cudaError_t check_symmetric(const double* A, int64_t lda, int64_t n, int* flag, cudaStream_t st);
// flag[0] = 1 if a diagonal entry A(i, col_off + i), i < rows, is negative or NaN (necessary condition of PSD-ness, O(n))
cudaError_t check_negative_diag(const double* A, int64_t lda, int64_t rows, int64_t col_off, int* flag, cudaStream_t st);
// scale column j of X by s[j]
cudaError_t scale_columns(double* X, int64_t ld, int64_t rows, int64_t cols, const double* s, cudaStream_t st);

// One-sided Jacobi SVD of M (p x p): M = U diag(sigma) V^T, sigma sorted descending; zero singular values get
// an orthonormal completion of U built from unit vectors (so SVD(0) = I * 0 * I, as the reference's tests expect).
// work: jacobi_svd_work_doubles(p) doubles.  info[0] = sweeps used, info[1] = 1 if not converged.
// Blocked for shared memory (panel.cu); synchronises the stream between sweeps when p*p*16 bytes exceed one SM's shared memory.
size_t jacobi_svd_work_doubles(int p);
// transpose != 0: the sweeps run on M^T (the outputs are still the factors of M).  For an upper-triangular M = R from
// a QR factorisation this is the Drmac-Veselic preconditioning: Jacobi on the lower-triangular R^T converges in 6-9
// sweeps where R itself can take 30+ on graded spectra.
cudaError_t jacobi_svd(const double* M, int64_t ldm, int p, double* U, int64_t ldu, double* sigma,
                       double* V, int64_t ldv, double* work, int* info, cudaStream_t st, int transpose = 0);
// Two-sided Jacobi eigen-decomposition of symmetric C (p x p): C = W diag(lambda) W^T.
// order: 0 = descending by value, 1 = descending by |value| (reference lora_drivers.rs:134-138)
cudaError_t jacobi_eigh(const double* C, int64_t ldc, int p, double* W, int64_t ldw, double* lambda,
                        int order, double* work, int* info, cudaStream_t st);

// literal `Stabilizer`: unit-lower-trapezoidal L of a full-pivot LU with the permutations dropped
// (reference lora_helpers.rs:144-146, nalgebra 0.33 FullPivLU).  X (rows x cols) is destroyed; L is rows x min(rows, cols).
cudaError_t fullpiv_lu_L(double* X, int64_t ldx, int64_t rows, int64_t cols, double* L, int64_t ldl,
                         double* scratch, cudaStream_t st);

// sparse-sign sketch (SASO, K6): A_sk(d x n) += S A_local, S has zeta nonzeros (+-1/sqrt(zeta)) per column (= per row of A)
cudaError_t saso_apply(uint64_t seed, int64_t d, int zeta, const double* A, int64_t lda, int64_t m, int64_t n,
                       int64_t row_off, double* Ask, int64_t ldk, int sms, cudaStream_t st);

// synthetic low-rank-plus-noise: A(i,j) += eta * gauss(row_off + i, j) / sqrt(m_global)
cudaError_t add_noise(double* A, int64_t lda, int64_t rows, int64_t cols, int64_t row_off, double scale,
                      uint64_t seed, uint32_t stream, cudaStream_t st);

}  // namespace rnla
