/* oracle.h -- CPU restatement of the reference's sketch-and-factor path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may build,
 * load or call this; nothing under randnla_b200/ does.  See oracle/README.md for what is pinned how.
 *
 * All matrices column-major f64 with lda = nrows (nalgebra DMatrix layout).
 */
#ifndef RNLA_ORACLE_H
#define RNLA_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- L0 RNG: rust-random123/src/philox.rs:149-154,173-176,211-223; threefry.rs:30-93; rand_core seed_from_u64 */
void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);
void orc_threefry2x64_20(const uint64_t ctr[2], const uint64_t key[2], uint64_t out[2]);
void orc_seed_from_u64(uint64_t state, uint64_t key[2]);
/* t-th u64 of ThreeFry2x64Rng::seed_from_u64(seed) (BlockRng64 order: x[0], x[1] of block 0, then block 1 ...) */
uint64_t orc_threefry_rng_u64(uint64_t seed, uint64_t t);

/* ---- L1 sketch operators ----------------------------------------------------------------------- */
/* this build's counter map (restated independently of randnla_b200/csrc/rng.cuh):
 * out(r,c) = T_dist(philox4x32_10((q_lo,q_hi,c,stream),(seed_lo,seed_hi))[R&3]), R = row_off + r, q = R>>2 */
void orc_omega_fill(int dist, uint64_t seed, uint32_t stream, int64_t rows, int64_t cols, int64_t row_off,
                    double* out, int64_t ld);
float orc_gauss_from_u32(uint32_t k);
/* what the reference's sketching_operator returns for Uniform (dist 1) / Rademacher (dist 2):
 * src/sketch.rs:112-127 with rand 0.8.5 Uniform<f64> / Bernoulli over the seed-0 ThreeFry stream.  Returns -1 for Gaussian
 * (needs rand_distr's ziggurat tables, not in the tree). */
int orc_sketching_operator_ref(int dist, uint64_t seed, int64_t rows, int64_t cols, double* out);
/* rand_distr 0.4.3 StandardNormal (ziggurat) on the ThreeFry2x64Rng stream: tables, and `count` samples (returns the stream words consumed) */
void orc_ziggurat_tables(double* x257, double* f257);
int64_t orc_threefry_gaussian(uint64_t seed, int64_t count, double* out);

/* ---- L2 dense kernels with nalgebra 0.33 conventions ------------------------------------------- */
void orc_gemm_nn(const double* A, int64_t lda, int64_t m, int64_t K, const double* B, int64_t ldb, int64_t N, double* C, int64_t ldc);
void orc_gemm_tn(const double* A, int64_t lda, int64_t m, int64_t n, const double* Q, int64_t ldq, int64_t N, double* Z, int64_t ldz);
/* X.qr(): thin Q (rows x p), R (p x cols), p = min(rows, cols), R diag >= 0.  R may be NULL. */
void orc_qr(const double* X, int64_t rows, int64_t cols, double* Q, double* R);
/* src/sketch.rs:45-85; attr 0 = Row, 1 = Column; returns 0 or 2 (InvalidDimensions) */
int orc_haar_sample(int64_t rows, int64_t cols, int attr, uint64_t seed, double* out);
/* X.full_piv_lu().l(): rows x min(rows, cols) */
void orc_stabilizer(const double* X, int64_t rows, int64_t cols, double* L);
/* lower Cholesky; returns 0 on success, -1 if not positive definite */
int orc_cholesky_lower(const double* A, int64_t n, double* L);
/* thin SVD of M (rows x cols, any shape): U rows x p, sigma p (descending), Vt p x cols */
int orc_svd(const double* M, int64_t rows, int64_t cols, double* U, double* sigma, double* Vt);
/* symmetric eigen-decomposition (values ascending) */
int orc_symmetric_eigen(const double* A, int64_t n, double* W, double* lambda);

/* ---- L3/L4: the path.  mode 0 = intended (monograph TSOG1, QR stabiliser), 1 = literal (as written).
 * omega_n: optional n x l operator used where tsog1 draws sketching_operator(Gaussian, n, l) (lora_helpers.rs:71);
 * omega_m: optional m x l operator for the odd branch (:74).  NULL -> generated with orc_omega_fill(dist, seed, stream 1 / 2). */
typedef struct orc_opts {
    int mode; int dist; uint64_t seed; int num_passes; int passes_per_stab;
    const double* omega_n; const double* omega_m;
    int skip_psd_check;   /* rand_evd2 only: skip the O(n^3) eigen-decomposition of src/lora_drivers.rs:178-184 (tests at sizes where
                           * it would take minutes; the inputs of those tests are PSD by construction) */
} orc_opts;
void orc_tsog1(const double* A, int64_t m, int64_t n, int64_t k, int num_passes, int passes_per_stab, const orc_opts* o, double* S);
void orc_rf1(const double* A, int64_t m, int64_t n, int64_t k, const orc_opts* o, double* Q);
void orc_qb1(const double* A, int64_t m, int64_t n, int64_t k, const orc_opts* o, double* Q, double* B);
/* returns 0 or the RandNLAError status code (same numbering as include/rnla.h); *r = min(k, min(k+s, m, n)) */
int orc_rand_svd(const double* A, int64_t m, int64_t n, int64_t k, double epsilon, int64_t s, const orc_opts* o,
                 double* U, double* S, double* Vt, int64_t* r);
int orc_rand_evd1(const double* A, int64_t n, int64_t k, double epsilon, int64_t s, const orc_opts* o, double* V, double* lambda, int64_t* r);
int orc_rand_evd2(const double* A, int64_t n, int64_t k, int64_t s, const orc_opts* o, double* V, double* lambda, int64_t* r);

/* ---- sketch step of sketch_and_precondition (src/sketch_and_precondition.rs:49-52,105-107,172-176) */
int64_t orc_sketch_dim(int64_t m, int64_t n, double sampling_factor, int rule);
/* dense: A_sk = S A with S(i,j) = omega(row j, col i) (stream 3), the definition the GPU path uses */
void orc_sketch_apply_dense(int dist, uint64_t seed, int64_t d, const double* A, int64_t m, int64_t n, double* A_sk);
/* sparse sign: zeta non-zeros per column of S */
void orc_sketch_apply_saso(uint64_t seed, int64_t d, int zeta, const double* A, int64_t m, int64_t n, double* A_sk);
int64_t orc_cgls(const double* a, int64_t m, int64_t n, const double* b, double tolerance, int64_t num_iterations, double* x, int* converged_out);
int orc_blendenpik(const double* A, int64_t m, int64_t n, const double* b, double epsilon, int64_t l, double sampling_factor,
                   int kind, int dist_or_width, int zeta, uint64_t seed, double* x, int64_t* iters_out, int* converged_out);
int orc_lsrn(const double* A, int64_t m, int64_t n, const double* b, double epsilon, int64_t l, double sampling_factor,
             int kind, int dist_or_width, int zeta, uint64_t seed, double* x, int64_t* iters_out, int* converged_out);
void orc_sketch_apply_saso_block(uint64_t seed, int64_t d, int zeta, int w, const double* A, int64_t m, int64_t n,
                                 int64_t row_off, double* A_sk);

/* ---- next rows, oracle_next.c ------------------------------------------------ */
/* src/pivot_decompositions.rs:105-180 (steps = min(m,n)) / :196-269 (steps = k).  R m x n, Q m x m or NULL, perm n */
void orc_qrcp_steps(const double* A, int64_t m, int64_t n, int64_t steps, double* R, double* Q, int64_t* perm);
/* src/cqrrpt.rs:27-58 */
int orc_sap_chol_qrcp(const double* A, int64_t m, int64_t n, int64_t d, int kind, int dist_or_width, int zeta, uint64_t seed,
                      double* Q, double* R, int64_t* J, int64_t* k_out);
/* src/sketch_and_solve.rs:24-33 (which 0, QR) / :54-66 (which 1, SVD) */
int orc_sketched_least_squares(int which, const double* A, int64_t m, int64_t n, const double* b, int kind, int dist_or_width,
                               int zeta, uint64_t seed, double* x);
/* src/id.rs */
int orc_osid_qrcp(const double* Y, int64_t l, int64_t w, int64_t k, int attr, double* X, int64_t* J);
int orc_osid_randomised(const double* A, int64_t m, int64_t n, int64_t k, int attr, const orc_opts* o, double* X, int64_t* J);
int orc_two_sided_id(int randomised, const double* A, int64_t m, int64_t n, int64_t k, const orc_opts* o,
                     double* Z, int64_t* I, int64_t* J, double* X);
int orc_cur(int randomised, const double* A, int64_t m, int64_t n, int64_t k, const orc_opts* o, int64_t* J, double* U, int64_t* I);
/* src/sketch_and_precondition.rs:150-216 */
int orc_saddle_point(const double* A, int64_t m, int64_t n, const double* b, const double* c, double mu, double epsilon, int64_t l,
                     double sampling_factor, int dist, uint64_t seed, double* x, double* y, int64_t* iters_out, int* converged_out);
/* src/solvers.rs:115-278 (translation of scipy 1.14.1 lsqr).  iter_lim < 0: 2 n.  Returns the length of the arnorms history. */
int64_t orc_lsqr(const double* a, int64_t m, int64_t n, const double* b, double damp, double atol, double btol, double conlim,
                 int64_t iter_lim, int calc_var, const double* x0, double* x, int64_t* istop_out, int64_t* itn_out,
                 double* r1norm_out, double* r2norm_out, double* anorm_out, double* acond_out, double* arnorms, double* xnorm_out,
                 double* var);

/* src/cg.rs:77-112; returns 0 or 9 (NotPositiveSemiDefinite) */
int orc_conjugate_grad(const double* a, int64_t n, const double* b, double* x, int64_t* iters_out, int* converged_out);

/* src/pivot_decompositions.rs:21-86; returns 0 or 6 (SingularMatrix) */
int orc_lupp(const double* a, int64_t n, double* L, double* U, int64_t* perm);

void orc_set_threads(int nthreads);
int orc_get_threads(void);

#ifdef __cplusplus
}
#endif
#endif
