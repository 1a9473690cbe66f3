/* rnla.h -- C ABI of the B200-native sketch-and-factor hot path of randnla (crate `randblas`).
 *
 * This is the drop-in boundary: every entry point below is what a Rust FFI binding
 * for the corresponding reference function would call.  Plain pointers and sizes only; all matrices
 * are column-major f64 with lda = nrows unless an explicit leading dimension is given (nalgebra
 * `DMatrix<f64>` layout).  The caller owns every buffer: the library never returns memory the caller
 * must free, and never frees caller memory.
 *
 * Two flavours per operation:
 *   rnla_xxx      host buffers in/out (what the reference-signature Rust function binds);
 *                 host<->device copies happen inside the call
 *   rnla_xxx_dev  device buffers in/out on the library stream (needed for the 32-320 GB configs
 *                 where a host DMatrix cannot be the carrier, and for benchmarks)
 *
 * Every function returns an rnla_status; codes map 1:1 onto the reference's `RandNLAError`
 * variants (reference src/errors.rs:3-14).  rnla_last_error_message() returns the text the
 * reference would have put in the variant's String (thread-local, valid until the next call).
 *
 * There is NO CPU fallback: if no CUDA device is usable every compute entry point fails with
 * RNLA_ERR_COMPUTATION.
 */
#ifndef RNLA_H
#define RNLA_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum rnla_status {
    RNLA_OK = 0,
    RNLA_ERR_INVALID_PARAMETERS = 1,   /* RandNLAError::InvalidParameters        errors.rs:4  */
    RNLA_ERR_INVALID_DIMENSIONS = 2,   /* RandNLAError::InvalidDimensions        errors.rs:5  */
    RNLA_ERR_NEGATIVE_DIMENSIONS = 3,  /* RandNLAError::NegativeDimensions       errors.rs:6  */
    RNLA_ERR_NOT_OVERDETERMINED = 4,   /* RandNLAError::NotOverdetermined        errors.rs:7  */
    RNLA_ERR_NOT_SQUARE = 5,           /* RandNLAError::NotSquare                errors.rs:8  */
    RNLA_ERR_SINGULAR_MATRIX = 6,      /* RandNLAError::SingularMatrix           errors.rs:9  */
    RNLA_ERR_MATRIX_DECOMPOSITION = 7, /* RandNLAError::MatrixDecompositionError errors.rs:10 */
    RNLA_ERR_NOT_HERMITIAN = 8,        /* RandNLAError::NotHermitian             errors.rs:11 */
    RNLA_ERR_NOT_PSD = 9,              /* RandNLAError::NotPositiveSemiDefinite  errors.rs:12 */
    RNLA_ERR_COMPUTATION = 10          /* RandNLAError::ComputationError         errors.rs:13 (CUDA/NCCL failures land here) */
} rnla_status;

/* reference src/sketch.rs:9-13 */
typedef enum rnla_dist { RNLA_GAUSSIAN = 0, RNLA_UNIFORM = 1, RNLA_RADEMACHER = 2 } rnla_dist;
/* reference src/sketch.rs:18-21 */
typedef enum rnla_attr { RNLA_ROW = 0, RNLA_COLUMN = 1 } rnla_attr;

/* Which `tsog1` the drivers run:
 *   INTENDED  monograph TSOG1: S = Omega; (S = A S; stab; S = A^T S; stab)*, QR-based stabiliser.
 *   LITERAL   statement-for-statement src/lora_helpers.rs:58-105 including the zero-S1 defect and the
 *             permutation-dropping full-pivot-LU `Stabilizer` (:144-146): Omega is never used for even
 *             num_passes, exactly as in the reference. */
typedef enum rnla_mode { RNLA_MODE_INTENDED = 0, RNLA_MODE_LITERAL = 1 } rnla_mode;

/* Generator behind rnla_sketching_operator:
 *   PHILOX    this build's counter-based map (Philox4x32-10, entry = pure function of seed,row,col)
 *   THREEFRY  the reference's ThreeFry2x64-20 sequential stream seeded with seed_from_u64(seed)
 *             (src/sketch.rs:112-127); available for UNIFORM and RADEMACHER, whose rand 0.8.5
 *             transforms are closed-form.  GAUSSIAN needs rand_distr's ziggurat tables, which are
 *             not in the reference tree -> RNLA_ERR_INVALID_PARAMETERS. */
typedef enum rnla_generator { RNLA_GEN_PHILOX = 0, RNLA_GEN_THREEFRY = 1 } rnla_generator;

typedef struct rnla_options {
    int32_t mode;             /* rnla_mode; default INTENDED */
    int32_t dist;             /* rnla_dist of the range-finder sketch; reference hard-codes GAUSSIAN (lora_helpers.rs:71,74) */
    uint64_t seed;            /* reference hard-codes 0 (sketch.rs:112) */
    int32_t num_passes;       /* <=0: reference constants (2 for rand_svd/rand_evd1 lora_helpers.rs:40, 3 for rand_evd2 lora_drivers.rs:186) */
    int32_t passes_per_stab;  /* <=0: 1 (same lines) */
    int32_t fused_sketch;     /* 1: Omega is generated inside the A*Omega kernel and never materialised; 0: materialise (K0) then multiply;
                               * 2 (default): fuse iff the n x l operand would not stay L2-resident (> 48 MiB) */
    int32_t range_passes_int8; /* which passes over A run on the INT8 tensor cores (tcgen05 kind::i8) from a balanced-digit fixed-point
                               * split of A with EXACT int32 accumulation (csrc/i8gemm.cu):
                               * -1 (default, "auto"): level 3 where the shape is supported, else 0.
                               *  0: every pass in FP64 (DMMA).
                               *  3: every pass FP64-grade on the integer pipe: all four products on a 55-bit split (7 digit planes of widths
                               *     7, 8, ..., 8 bits; 28 digit pairs): the error of a product is below 2^-53 of (row max of A) x (column max of
                               *     the thin operand) x (contraction length)^(1/2), the normwise error model of an FP64 GEMM, independent of
                               *     the spectrum of A; singular values agree with the FP64 kernels' like two FP64 GEMMs with different
                               *     summation orders do.
                               *  1: A Omega, A^T Y on 31-bit operands (4 planes, 10 pairs), A S with all 16 pairs; Q^T A stays FP64.
                               *  2: as 1, Q^T A on the 55-bit split.  Levels 1 and 2 are SPECTRUM-CONDITIONAL: they reproduce the FP64 result
                               *     to 1e-10 only when the part of A outside the captured range is small against sigma_k and
                               *     sigma_1 / sigma_k <~ 1e4 (DESIGN.md 5c has the measured sweep); opt-in, never chosen by auto.
                               * Applies to rand_svd / rand_evd1 (dev_qb1) and rand_evd2 in the intended mode, l <= 256 (thin operands wider than
                               * 128 columns go through the MMA kernels in 128-column tiles), m n >= 2^22; other shapes keep FP64.  A matrix
                               * with Inf / NaN entries, or with a non-zero row below 2^-959, keeps the FP64 kernels too.
                               * Also RNLA_RANGE_INT8=0|1|2|3|auto in the environment. */
} rnla_options;

/* ---- library / context ------------------------------------------------------------------------- */
int32_t rnla_version(void);
const char* rnla_last_error_message(void);
/* bind the calling process to a CUDA device (default 0) and create the context (stream, workspace pool) */
rnla_status rnla_init(int32_t device);
void rnla_shutdown(void);
/* library stream as a cudaStream_t (so callers can order their own work / events against it) */
void* rnla_stream(void);
/* run on a caller-owned cudaStream_t instead (e.g. torch's current stream); NULL restores the library stream */
rnla_status rnla_set_stream(void* cuda_stream);
rnla_status rnla_synchronize(void);
/* the int8 passes (rnla_options.range_passes_int8) keep their digit-plane workspace (about 4 to 11 bytes per element of the
 * largest A seen) across calls; this frees it (rnla_shutdown does too) */
rnla_status rnla_release_workspace(void);
void rnla_default_options(rnla_options* opt);
/* process-wide defaults used by the reference-signature entry points */
rnla_status rnla_set_options(const rnla_options* opt);
void rnla_get_options(rnla_options* opt);
/* kernels launched by this library since load (bench.py `gpu_launches`) */
uint64_t rnla_kernel_launches(void);
/* on != 0: the integer tensor-core kernels of the passes also record kernel-level entries, named "k:...", nested inside the
 * driver-level phases (bench.py's roofline block reads per-launch durations from them); off by default */
rnla_status rnla_set_kernel_timing(int32_t on);
/* per-phase device timings of the last driver call: names[i] / ms[i]; returns the number of phases */
int32_t rnla_get_timings(const char** names, double* ms, int32_t cap);

/* ---- multi-GPU: one process per GPU, A row-sharded ------------------------------ */
/* 128-byte NCCL unique id; rank 0 creates it, the host side broadcasts it (torch.distributed / MPI / files) */
rnla_status rnla_comm_unique_id(uint8_t id[128]);
rnla_status rnla_comm_init(int32_t nranks, int32_t rank, const uint8_t id[128]);
rnla_status rnla_comm_destroy(void);
int32_t rnla_comm_size(void);
int32_t rnla_comm_rank(void);

/* ---- L0: counter-based RNG (reference rust-random123/src/philox.rs:211-223, threefry.rs:69-93) -- */
/* nblocks independent Philox4x32-10 blocks evaluated ON THE DEVICE: ctr[4*i..], key[2*i..] -> out[4*i..] (host buffers) */
rnla_status rnla_philox4x32_10(int64_t nblocks, const uint32_t* ctr, const uint32_t* key, uint32_t* out);
/* nblocks ThreeFry2x64-20 blocks on the device: ctr[2*i..], key[2*i..] -> out[2*i..] */
rnla_status rnla_threefry2x64_20(int64_t nblocks, const uint64_t* ctr, const uint64_t* key, uint64_t* out);

/* ---- L1: sketch operators (reference src/sketch.rs) -------------------------------------------- */
/* sketching_operator(dist, rows, cols)  src/sketch.rs:102-130.  Errors: InvalidDimensions if rows==0||cols==0 (:107-111). */
rnla_status rnla_sketching_operator(int32_t dist, int64_t rows, int64_t cols, double* out);
/* extended: explicit generator / seed / stream / global row offset (row-sharded operators) / leading dimension */
rnla_status rnla_sketch_fill(int32_t generator, int32_t dist, uint64_t seed, uint32_t stream,
                             int64_t rows, int64_t cols, int64_t row_offset, double* out, int64_t ld);
rnla_status rnla_sketch_fill_dev(int32_t generator, int32_t dist, uint64_t seed, uint32_t stream,
                                 int64_t rows, int64_t cols, int64_t row_offset, double* d_out, int64_t ld);
/* haar_sample(rows, cols, attr)  src/sketch.rs:45-85 */
rnla_status rnla_haar_sample(int64_t rows, int64_t cols, int32_t attr, double* out);

/* ---- L3: range-finder helpers (reference src/lora_helpers.rs) ----------------------------------- */
/* Orth(X) :131-133 -> thin Q (rows x min(rows,cols)), R diag >= 0 convention.  R (qcols x cols) optional (NULL ok). */
rnla_status rnla_orth(const double* X, int64_t rows, int64_t cols, double* Q, double* R, int64_t* qcols);
/* Stabilizer(X) :144-146 -> unit-lower-trapezoidal L of full-pivot LU, permutations dropped (rows x min(rows,cols)) */
rnla_status rnla_stabilizer(const double* X, int64_t rows, int64_t cols, double* L, int64_t* lcols);
/* tsog1(A, k, num_passes, passes_per_stab) :58-105 -> S (n x k).  mode from the current options. */
rnla_status rnla_tsog1(const double* A, int64_t m, int64_t n, int64_t k, int32_t num_passes,
                       int32_t passes_per_stab, double* S);
/* RF1(A, k) :37-44 -> Q (m x qcols) */
rnla_status rnla_rf1(const double* A, int64_t m, int64_t n, int64_t k, double* Q, int64_t* qcols);
/* QB1(A, k, epsilon) :17-23 -> Q (m x qcols), B (qcols x n, ld = qcols) */
rnla_status rnla_qb1(const double* A, int64_t m, int64_t n, int64_t k, double epsilon,
                     double* Q, double* B, int64_t* qcols);

/* ---- L4: drivers (reference src/lora_drivers.rs) ------------------------------------------------ */
/* rand_svd(A, k, epsilon, s) :30-69.  U: m x k buffer, S: k x k buffer (dense diagonal matrix, as the
 * reference returns), Vt: k x n buffer; *r = min(k, Q.ncols()) columns/rows are valid, packed with
 * leading dimensions m, r, r.  Errors: InvalidParameters for k==0, epsilon<=0, s==0 (:31-45). */
rnla_status rnla_rand_svd(const double* A, int64_t m, int64_t n, int64_t k, double epsilon, int64_t s,
                          double* U, double* S, double* Vt, int64_t* r);
/* rand_evd1(A, k, epsilon, s) :87-151.  V: n x k, lambda: k.  NotHermitian if A != A^T exactly (:106). */
rnla_status rnla_rand_evd1(const double* A, int64_t n, int64_t k, double epsilon, int64_t s,
                           double* V, double* lambda, int64_t* r);
/* rand_evd2(A, k, s) :167-224 (Nystrom).  NotPositiveSemiDefinite / MatrixDecompositionError as the reference. */
rnla_status rnla_rand_evd2(const double* A, int64_t n, int64_t k, int64_t s,
                           double* V, double* lambda, int64_t* r);

/* device-resident drivers: A is the LOCAL row shard (m_local x n, lda) when a communicator is active,
 * U is the matching local row shard; n-side outputs are replicated on every rank.  Unlike the host-buffer entry points
 * these return as soon as the work is enqueued on the library stream (rnla_stream / rnla_set_stream): order your reads
 * against that stream or call rnla_synchronize(). */
rnla_status rnla_rand_svd_dev(const double* dA, int64_t lda, int64_t m_local, int64_t n, int64_t k, int64_t s,
                              const rnla_options* opt, double* dU, int64_t ldu, double* dSigma /* k */,
                              double* dVt, int64_t ldvt, int64_t* r);
rnla_status rnla_rand_evd1_dev(const double* dA, int64_t lda, int64_t n, int64_t k, int64_t s,
                               const rnla_options* opt, double* dV, int64_t ldv, double* dLambda, int64_t* r);
rnla_status rnla_rand_evd2_dev(const double* dA, int64_t lda, int64_t n, int64_t k, int64_t s,
                               const rnla_options* opt, double* dV, int64_t ldv, double* dLambda, int64_t* r);

/* ---- sketch step of sketch_and_precondition (reference src/sketch_and_precondition.rs:49-52,105-107,172-176) */
/* DENSE: i.i.d. `dist` entries (what the reference draws).  SASO: textbook sparse-sign operator, zeta independent row
 * indices per column (shared-memory scatter; 1 <= zeta <= 8, d <= 25600).  SASO_BLOCK: sparse-sign operator whose zeta
 * zeta = g*w non-zeros per column come as g groups of w consecutive rows, one group per stripe of d/g rows (OSNAP block
 * construction), dealt chunk-wise by keyed bijections (register accumulators, no atomics, fixed summation order, A
 * streamed once; zeta in {1,2,4,8}, zeta <= d <= 16384).  For this kind the `dist` argument carries the block width w
 * (1, 2 or 4, w | zeta; 0 = default min(zeta, 4); w = 1 is statistically the textbook SASO).  DESIGN.md section 5. */
typedef enum rnla_sketch_kind { RNLA_SKETCH_DENSE = 0, RNLA_SKETCH_SASO = 1, RNLA_SKETCH_SASO_BLOCK = 2 } rnla_sketch_kind;
/* d = sketch dimension rule of the reference: blendenpik/lsrn (:49,:105) rule 0, saddle point (:172) rule 1 */
int64_t rnla_sketch_dim(int64_t m, int64_t n, double sampling_factor, int32_t rule);
/* A_sk (d x n) = S A, b_sk (d x nrhs) = S b with S (d x m): dense i.i.d. `dist`, or sparse-sign with zeta nonzeros per column.
 * b may be NULL (nrhs = 0).  Validation of m>=n, sampling_factor, epsilon, l stays with the caller-facing wrappers. */
rnla_status rnla_sketch_apply(int32_t kind, int32_t dist, uint64_t seed, int64_t d, int32_t zeta,
                              const double* A, int64_t m, int64_t n, const double* b, int64_t nrhs,
                              double* A_sk, double* b_sk);
rnla_status rnla_sketch_apply_dev(int32_t kind, int32_t dist, uint64_t seed, int64_t d, int32_t zeta,
                                  const double* dA, int64_t lda, int64_t m_local, int64_t n, int64_t row_offset,
                                  double* dA_sk, int64_t ld_sk);

/* ---- blendenpik_overdetermined end to end (reference src/sketch_and_precondition.rs:26-59, src/cg.rs:18-61): sketch, QR of
 * the sketch, z0 = Q^T b_sk, R^-1, CGLS on A R^-1 in OPERATOR form (the reference forms the dense product), x = R^-1 z.
 * Validation and messages follow :29-48.  kind/dist/zeta as in rnla_sketch_apply (the reference's own choice is DENSE,
 * GAUSSIAN).  x: n doubles.  iterations / converged (may be NULL): CGLS iterations used and whether ||s|| < epsilon fired
 * (the reference only prints that).  A numerically singular R is RNLA_ERR_SINGULAR_MATRIX (the reference unwraps :55). */
rnla_status rnla_blendenpik_overdetermined(const double* A, int64_t m, int64_t n, const double* b, double epsilon, int64_t l,
                                           double sampling_factor, int32_t kind, int32_t dist, int32_t zeta, double* x,
                                           int64_t* iterations, int32_t* converged);
/* device buffers; A and b are the caller's row shard when a communicator is active, x is replicated */
rnla_status rnla_blendenpik_overdetermined_dev(const double* dA, int64_t lda, int64_t m_local, int64_t n, const double* db,
                                               double epsilon, int64_t l, double sampling_factor, int32_t kind, int32_t dist,
                                               int32_t zeta, double* dx, int64_t* iterations, int32_t* converged);

/* lsrn_overdetermined end to end (reference src/sketch_and_precondition.rs:82-119): sketch, SVD of the sketch, N = V Sigma^-1
 * (0 where sigma == 0, :112), CGLS on A N in operator form from y = 0, x = N y.  Same arguments as blendenpik.  The on-device
 * SVD core limits n to 1024 (RNLA_ERR_INVALID_DIMENSIONS beyond). */
rnla_status rnla_lsrn_overdetermined(const double* A, int64_t m, int64_t n, const double* b, double epsilon, int64_t l,
                                     double sampling_factor, int32_t kind, int32_t dist, int32_t zeta, double* x,
                                     int64_t* iterations, int32_t* converged);
rnla_status rnla_lsrn_overdetermined_dev(const double* dA, int64_t lda, int64_t m_local, int64_t n, const double* db, double epsilon,
                                         int64_t l, double sampling_factor, int32_t kind, int32_t dist, int32_t zeta, double* dx,
                                         int64_t* iterations, int32_t* converged);

/* ==== rows after the hot path: the callers and neighbours of the sketch ==================== */
/* index vectors are int64 (the reference's Vec<usize>) */

/* qrcp (reference src/pivot_decompositions.rs:105-180; steps = min(m, n)) and economic_qrcp(a, k) (:196-269; steps = k):
 * Householder QR with column pivoting on exactly recomputed trailing column norms, first maximum wins.  R: m x n, the
 * reference's work matrix `r` after `steps` reflections (`r_eco` = its first k rows); perm: n; Q (optional, NULL / qcols = 0
 * to skip): the first qcols columns of the accumulated reflectors (qcols = k: `q_eco`; qcols = m: the full `q`).
 * Errors (the reference asserts): "k must be positive", "k must be <= min(m,n)" -> INVALID_PARAMETERS. */
rnla_status rnla_qrcp(const double* A, int64_t m, int64_t n, int64_t steps, int64_t qcols, double* Q, double* R, int64_t* perm);
/* in place on the device: dR holds A on entry (m x n, ldr), R on exit */
rnla_status rnla_qrcp_dev(double* dR, int64_t ldr, int64_t m, int64_t n, int64_t steps, int64_t* dperm, double* dQ, int64_t ldq,
                          int64_t qcols);

/* lupp(matrix) (reference src/pivot_decompositions.rs:21-86): LU with partial row pivoting, first maximum wins, the
 * reference's elimination order with separately rounded multiply and subtract -- L, U and perm are bit-identical to that
 * arithmetic.  L: n x n unit lower triangular, U: n x n upper triangular, perm: n (`p[k]` = original index of row k).
 * Errors: rows != cols -> NOT_SQUARE "Matrix must be square, found matrix with {} rows and {} columns" (:23-27); a zero pivot
 * column -> SINGULAR_MATRIX "Matrix must be nonsingular for an LU decomposition" (:44-48; like the reference, the last
 * diagonal entry is not examined). */
rnla_status rnla_lupp(const double* A, int64_t rows, int64_t cols, double* L, double* U, int64_t* perm);
/* device buffers: dW n x n holds the matrix on entry and is destroyed; dL, dU n x n; dperm n */
rnla_status rnla_lupp_dev(double* dW, int64_t ldw, int64_t n, double* dL, int64_t ldl, double* dU, int64_t ldu, int64_t* dperm);

/* sap_chol_qrcp(a, d) (reference src/cqrrpt.rs:27-58; CQRRPT): sketch (d x m operator, kind/dist/zeta as in
 * rnla_sketch_apply; the reference's own is DENSE GAUSSIAN), qrcp of the sketch, numerical rank k = #{|R_ii| > 1e-10},
 * A_pre = A[:, J[:k]] R_k^-1, Cholesky QR of A_pre, R = R_pre R_sk[:k, :].  Q: m x n buffer, first *k columns valid (ld m);
 * R: n x n buffer, holds the k x n factor packed with ld = *k; J: n.  Errors: "d must satisfy n <= d << m" (the reference's
 * assert :29) -> INVALID_PARAMETERS; failed Cholesky (:51) -> MATRIX_DECOMPOSITION. */
rnla_status rnla_sap_chol_qrcp(const double* A, int64_t m, int64_t n, int64_t d, int32_t kind, int32_t dist, int32_t zeta,
                               double* Q, double* R, int64_t* J, int64_t* k);
/* device buffers: dQ m x n (ldq), dR n x n (ldr; k x n valid, NOT repacked), dJ n */
rnla_status rnla_sap_chol_qrcp_dev(const double* dA, int64_t lda, int64_t m, int64_t n, int64_t d, int32_t kind, int32_t dist,
                                   int32_t zeta, double* dQ, int64_t ldq, double* dR, int64_t ldr, int64_t* dJ, int64_t* k);

/* sketched_least_squares_qr / _svd (reference src/sketch_and_solve.rs:24-33, :54-66): sketch with rows/4 rows, QR (or SVD)
 * of the sketch, x = R^-1 Q^T b_sk with `solve_upper_triangular_system`'s rule for zero pivots (src/solvers.rs:22-41), or
 * x = V Sigma^-1 U^T b_sk (`solve_diagonal_system` :57-69).  x: n.  The SVD variant needs n <= 1024 on the device.
 * INVALID_DIMENSIONS when rows/4 < n (the reference then indexes out of bounds). */
rnla_status rnla_sketched_least_squares_qr(const double* A, int64_t m, int64_t n, const double* b, int32_t kind, int32_t dist,
                                           int32_t zeta, double* x);
rnla_status rnla_sketched_least_squares_svd(const double* A, int64_t m, int64_t n, const double* b, int32_t kind, int32_t dist,
                                            int32_t zeta, double* x);
rnla_status rnla_sketched_least_squares_dev(int32_t which /* 0 QR, 1 SVD */, const double* dA, int64_t lda, int64_t m, int64_t n,
                                            const double* db, int32_t kind, int32_t dist, int32_t zeta, double* dx);

/* interpolative decompositions (reference src/id.rs).  attr: RNLA_COLUMN  Y ~ Y[:, J] X, X k x w;  RNLA_ROW  Y ~ X Y[J, :],
 * X l x k.  J: k.  Errors: "k must be positive)", "k must be <= min(l,w)" (asserts :278-279) -> INVALID_PARAMETERS; a singular
 * R1 (the reference unwraps :290) -> SINGULAR_MATRIX. */
rnla_status rnla_osid_qrcp(const double* Y, int64_t l, int64_t w, int64_t k, int32_t attr, double* X, int64_t* J);        /* :272-318 */
/* osid_randomised (:217-249): Column sketches with a k x m Gaussian; Row multiplies by tsog1(a, k, 2, 1)^T, which conforms
 * only when a has k columns (INVALID_DIMENSIONS otherwise), as it is used by two_sided_id_randomised (:99) */
rnla_status rnla_osid_randomised(const double* A, int64_t m, int64_t n, int64_t k, int32_t attr, double* X, int64_t* J);
rnla_status rnla_osid_randomised_dev(const double* dA, int64_t lda, int64_t m, int64_t n, int64_t k, int32_t attr,
                                     const rnla_options* opt, double* dX, int64_t ldx, int64_t* dJ);
/* two_sided_id (:118-129, randomised = 0) / two_sided_id_randomised (:94-101): A ~ Z A[I, J] X; Z m x k, I k, J k, X k x n */
rnla_status rnla_two_sided_id(const double* A, int64_t m, int64_t n, int64_t k, int32_t randomised, double* Z, int64_t* I, int64_t* J,
                              double* X);
/* cur (:34-71, randomised = 0) / cur_randomised (:154-193): A ~ A[:, J] U A[I, :]; J k, U k x k, I k.  k <= 1024 on the device */
rnla_status rnla_cur(const double* A, int64_t m, int64_t n, int64_t k, int32_t randomised, int64_t* J, double* U, int64_t* I);
rnla_status rnla_cur_dev(const double* dA, int64_t lda, int64_t m, int64_t n, int64_t k, int32_t randomised, const rnla_options* opt,
                         int64_t* dJ, double* dU, int64_t ldu, int64_t* dI);

/* sketch_saddle_point_precondition (reference src/sketch_and_precondition.rs:150-216), dense Gaussian operator (the
 * reference's): SVD of the sketch, M = V (Sigma^2 + mu)^-1/2 (mu > 0) or V Sigma^-1, b_mod = b - S^T U diag(.) V^T c,
 * z0 = U^T S b_mod, CGLS on A M in operator form, x = M z, y = b - A x.  c may be NULL (`c.is_empty()` :195).  x: n, y: m.
 * Validation as blendenpik (:152-171).  n <= 1024 on the device; with mu = 0 a rank-deficient sketch is INVALID_DIMENSIONS
 * (the reference's shapes do not conform at :211). */
rnla_status rnla_sketch_saddle_point_precondition(const double* A, int64_t m, int64_t n, const double* b, const double* c, double mu,
                                                  double epsilon, int64_t l, double sampling_factor, double* x, double* y,
                                                  int64_t* iterations, int32_t* converged);
rnla_status rnla_sketch_saddle_point_precondition_dev(const double* dA, int64_t lda, int64_t m, int64_t n, const double* db,
                                                      const double* dc, double mu, double epsilon, int64_t l, double sampling_factor,
                                                      double* dx, double* dy, int64_t* iterations, int32_t* converged);

/* ---- reference src/cg.rs: the iterative solvers the drivers above are built on, under their own names ----
 * cgls(a, b, tolerance, num_iterations, x) (:18-61) on a dense m x n system: x0 may be NULL (the reference's `None`: zeros).
 * x: n.  iterations / converged (may be NULL): what the reference prints (:46, :58).  No validation in the reference. */
rnla_status rnla_cgls(const double* A, int64_t m, int64_t n, const double* b, double tolerance, int64_t num_iterations,
                      const double* x0, double* x, int64_t* iterations, int32_t* converged);
/* device buffers; dA / db the caller's row shard when a communicator is active, dx replicated (in: initial guess, out: solution) */
rnla_status rnla_cgls_dev(const double* dA, int64_t lda, int64_t m_local, int64_t n, const double* db, double tolerance,
                          int64_t num_iterations, double* dx, int64_t* iterations, int32_t* converged);
/* conjugate_grad(a, b, x) (:77-112): a n x n symmetric positive semi-definite, at most 2 n iterations, stops when r.r < 1e-10.
 * x0 may be NULL (the reference's `None`: the vector of ones).  NOT_POSITIVE_SEMI_DEFINITE from the reference's eigenvalue
 * check (:80-86), which is run for n <= 512 only (an O(n^3) decomposition in front of an O(n^2) iteration; same policy as
 * rnla_rand_evd2).  iterations: the loop index the reference prints when it converges (2 n otherwise). */
rnla_status rnla_conjugate_grad(const double* A, int64_t n, const double* b, const double* x0, double* x, int64_t* iterations,
                                int32_t* converged);
rnla_status rnla_conjugate_grad_dev(const double* dA, int64_t lda, int64_t n, const double* db, double* dx, int64_t* iterations,
                                    int32_t* converged);
/* verify_solution(a, b, x) = ||a x - b||_2 (:115-117) */
rnla_status rnla_verify_solution(const double* A, int64_t m, int64_t n, const double* b, const double* x, double* residual_norm);

/* lsqr(a, b, damp, atol, btol, conlim, iter_lim, calc_var, x0) (reference src/solvers.rs:115-278, its translation of scipy
 * 1.14.1 sparse.linalg.lsqr; called by no driver of the reference, SURVEY.md section 8f row 1 "wire lsqr").  The
 * Golub-Kahan bidiagonalisation runs on the device (A streamed twice per iteration), the scalar recurrences on the host.
 * iter_lim < 0: the reference's `None` (2 n).  x0 may be NULL.  x: n.  var: n (zeros unless calc_var; may be NULL when
 * calc_var == 0).  arnorms: the reference's history of ||A^T r|| estimates (its 8th return value), the first
 * min(n_arnorms, arnorms_cap) entries are written; may be NULL.  The other return values arrive in *result in the
 * reference's order.  No validation in the reference (shape mismatches panic inside nalgebra): m >= 1, n >= 1 here. */
typedef struct rnla_lsqr_result {
    int64_t istop;      /* reason for termination, 0..7 as scipy */
    int64_t itn;        /* iterations performed */
    double r1norm;      /* norm(r) */
    double r2norm;      /* sqrt(norm(r)^2 + damp^2 norm(x - x0)^2) */
    double anorm;       /* estimate of the Frobenius norm of Abar */
    double acond;       /* estimate of cond(Abar) */
    double xnorm;       /* norm(x) */
    int64_t n_arnorms;  /* length of the arnorm history */
} rnla_lsqr_result;
rnla_status rnla_lsqr(const double* A, int64_t m, int64_t n, const double* b, double damp, double atol, double btol, double conlim,
                      int64_t iter_lim, int32_t calc_var, const double* x0, double* x, rnla_lsqr_result* result, double* arnorms,
                      int64_t arnorms_cap, double* var);
/* device buffers: dA / db the caller's row shard when a communicator is active; dx0, dx, dvar replicated; arnorms on the HOST */
rnla_status rnla_lsqr_dev(const double* dA, int64_t lda, int64_t m_local, int64_t n, const double* db, double damp, double atol,
                          double btol, double conlim, int64_t iter_lim, int32_t calc_var, const double* dx0, double* dx,
                          rnla_lsqr_result* result, double* arnorms, int64_t arnorms_cap, double* dvar);

/* ---- building blocks on device buffers (tests, benches, host mirrors) ---------------------------- */
/* y (m) = A x (trans = 0) or y (n) = A^T x (trans != 0, all-reduced over the communicator): the two streaming kernels of CGLS */
rnla_status rnla_gemv_dev(const double* dA, int64_t lda, int64_t m, int64_t n, int32_t trans, const double* dx, double* dy);
/* C (m x N) = A (m x K) * B (K x N) */
rnla_status rnla_gemm_nn_dev(const double* dA, int64_t lda, int64_t m, int64_t K,
                             const double* dB, int64_t ldb, int64_t N, double* dC, int64_t ldc);
/* C (m x N) = A (m x K) * Omega(K x N), Omega generated in-kernel (never materialised) */
rnla_status rnla_sketch_gemm_dev(const double* dA, int64_t lda, int64_t m, int64_t K, int32_t dist, uint64_t seed,
                                 uint32_t stream, int64_t N, double* dC, int64_t ldc);
/* Z (n x N) = A (m x n)^T * Q (m x N); all-reduced over the communicator if one is active and allreduce != 0 */
rnla_status rnla_gemm_tn_dev(const double* dA, int64_t lda, int64_t m, int64_t n,
                             const double* dQ, int64_t ldq, int64_t N, double* dZ, int64_t ldz, int32_t allreduce);
/* in-place orthonormalisation of a (row-sharded if `sharded`) panel; R (cols x cols, ld = cols) optional */
rnla_status rnla_orth_dev(double* dX, int64_t ldx, int64_t rows_local, int64_t cols, int32_t sharded,
                          double* dR, int64_t* deficient);
/* small dense core on one GPU: SVD of a (rows x cols) matrix with cols <= 512 columns... see DESIGN.md */
rnla_status rnla_small_svd_dev(const double* dM, int64_t ldm, int64_t p, double* dU, double* dSigma, double* dV);
/* planning decisions of the launches, computed on the host (no GPU needed; tests pin them):
 * rnla_plan_gemm: out = {split-K parts of gemm_nn for (m x n) * (n x N), row chunks of gemm_tn for (m x n)^T (m x N), rows per
 * chunk, tiles}.  rnla_plan_saso_block: the work list of the block sparse-sign kernel for `nchunks` chunks of 2048 rows:
 * shape = {blocks per thread, columns per CTA, parts, column groups}, desc = (column group, first chunk, end chunk, slot) per
 * CTA in launch order; returns the number of CTAs (-1: unsupported shape). */
void rnla_plan_gemm(int64_t m, int64_t n, int64_t N, int32_t sms, int32_t* out);
int32_t rnla_plan_saso_block(int64_t d, int32_t zeta, int32_t width, int64_t n, int64_t nchunks, int32_t sms, int32_t* shape,
                             int32_t* desc, int32_t cap, int32_t* nslots);
/* diagnostics: Jacobi sweeps used by the last small SVD (drivers and rnla_small_svd_dev) */
int32_t rnla_last_jacobi_sweeps(void);
rnla_status rnla_small_eigh_dev(const double* dC, int64_t ldc, int64_t p, double* dW, double* dLambda);

/* the integer tensor-core products behind rnla_options.range_passes_int8 (csrc/i8gemm.cu), for tests and benches: one split of A
 * into `planes` digit planes (4: 31-bit operands -- all_pairs != 0 adds the sweep over digit-pair groups 4..6; 6: 47-bit; 7: 55-bit),
 * then `reps` products.  trans = 0: C (m x N) = A B with B n x N;  trans != 0: C (n x N) = A^T B with B m x N.  N <= 256. */
rnla_status rnla_i8_gemm_dev(int32_t trans, int32_t planes, int32_t all_pairs, const double* dA, int64_t lda, int64_t m, int64_t n,
                             const double* dB, int64_t ldb, int64_t N, double* dC, int64_t ldc, int32_t reps);
/* test hook: drain the int32 accumulators of the integer kernels every `stages` stages of 64 contraction indices instead of at the
 * exactness bound (340 stages for 7 planes); 0 restores the bound.  Results do not change (the drains are exact). */
rnla_status rnla_debug_i8_flush(int32_t stages);
/* round-1 form of the same: reps > 0: 4 planes, ten leading pairs; reps < 0: |reps| products, A B with all 16 pairs of the 31-bit
 * split, A^T B on the 55-bit split */
rnla_status rnla_i8_range_gemm_dev(int32_t trans, const double* dA, int64_t lda, int64_t m, int64_t n, const double* dB, int64_t ldb,
                                   int64_t N, double* dC, int64_t ldc, int32_t reps);

/* synthetic inputs, generated on the device shard by shard: A = U0 diag(sigma) V0^T + eta G */
rnla_status rnla_generate_lowrank_dev(double* dA, int64_t lda, int64_t m_local, int64_t n, int64_t row_offset,
                                      int64_t m_global, int64_t r0, const double* sigma_host, double eta, uint64_t seed);

/* live roof measurement for bench.py (same box, same run): peak DMMA.8x8x4 rate of all SMs in TFLOP/s (register-resident
 * chains, no memory) and a read-only HBM stream over `hbm_bytes` (>= 1 GiB recommended) in GB/s.  Either pointer may be NULL. */
rnla_status rnla_measure_roofs(double* fp64_dmma_tflops, double* hbm_read_gbs, size_t hbm_bytes);
/* int8 tensor-core rate (tcgen05.mma kind::i8, M = N = 128, K = 32, operands resident in shared memory, every SM) in TOP/s:
 * best short launch (`burst`) and a 0.25 s back-to-back run under the box's power cap (`sustained`); either pointer may be NULL */
rnla_status rnla_measure_int8_roof(double* burst_tops, double* sustained_tops);

/* raw device memory helpers for hosts without a CUDA runtime binding (Rust shim, ctypes) */
rnla_status rnla_malloc(void** dptr, size_t bytes);
rnla_status rnla_free(void* dptr);
rnla_status rnla_memcpy_h2d(void* dst, const void* src, size_t bytes);
rnla_status rnla_memcpy_d2h(void* dst, const void* src, size_t bytes);

#ifdef __cplusplus
}
#endif
#endif /* RNLA_H */
