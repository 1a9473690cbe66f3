#pragma once
#include "context.cuh"
#include <vector>

namespace rnla {

struct ShardInfo { int64_t rows_local; int64_t row_off; int64_t rows_global; };

rnla_status shard_layout(int64_t rows_local, ShardInfo* out);

rnla_status dev_gemm_nn(const double* A, int64_t lda, int64_t m, int64_t K, const double* B, int64_t ldb, int64_t N,
                        double* C, int64_t ldc);
rnla_status dev_sketch_gemm(const double* A, int64_t lda, int64_t m, int64_t K, int dist, uint64_t seed, uint32_t stream,
                            int64_t N, double* C, int64_t ldc);
rnla_status dev_gemm_tn(const double* A, int64_t lda, int64_t m, int64_t n, const double* Q, int64_t ldq, int64_t N,
                        double* Z, int64_t ldz, bool allreduce);
rnla_status orth_inplace(double* X, int64_t ldx, const ShardInfo& sh, int p, bool sharded, double* R_out, int64_t* deficient_out,
                         int need_clean = 2);

rnla_status dev_tsog1(const double* A, int64_t lda, const ShardInfo& sh, int64_t n, int l, int q, int pps,
                      const rnla_options& o, double* S);
rnla_status dev_rf1(const double* A, int64_t lda, const ShardInfo& sh, int64_t n, int l, int q, int pps,
                    const rnla_options& o, double* Q, int64_t ldq);
rnla_status dev_qb1(const double* A, int64_t lda, const ShardInfo& sh, int64_t n, int l, int q, int pps,
                    const rnla_options& o, double* Q, int64_t ldq, double* Bt);
rnla_status dev_rand_svd(const double* A, int64_t lda, int64_t m_local, int64_t n, int64_t k, int64_t s,
                         const rnla_options& o, double* U, int64_t ldu, double* Sigma, double* Vt, int64_t ldvt, int64_t* r_out);
rnla_status dev_rand_evd1(const double* A, int64_t lda, int64_t m_local, int64_t n, int64_t k, int64_t s,
                          const rnla_options& o, double* V, int64_t ldv, double* Lambda, int64_t* r_out);
rnla_status dev_rand_evd2(const double* A, int64_t lda, int64_t m_local, int64_t n, int64_t k, int64_t s,
                          const rnla_options& o, double* V, int64_t ldv, double* Lambda, int64_t* r_out);

// saso_block.cu: block sparse-sign sketch, A_sk fully overwritten with S A_local
struct SbFrag { int cg, lo, hi, slot; };     // work-list entry: column group, chunk range, partial slot (-1: writes A_sk)
void saso_block_worklist(int ncg, int64_t nchunks, int sms, std::vector<SbFrag>& work, std::vector<int>& fix_cg,
                         std::vector<int>& fix_off, std::vector<int>& fix_cnt, std::vector<int>& slots);
bool saso_block_shape(int64_t d, int zeta, int w, int64_t n, int* bpt_out, int* cb_out, int* parts_out);
rnla_status saso_block_apply(uint64_t seed, int64_t d, int zeta, int w, const double* A, int64_t lda, int64_t m_local,
                             int64_t n, int64_t row_off, double* Ask, int64_t ldk);

extern int g_last_jacobi_sweeps;

// solve.cu: blendenpik_overdetermined end to end on the device
rnla_status dev_sketch_apply(int kind, int dist, uint64_t seed, int64_t d, int zeta, const double* dA, int64_t lda,
                             int64_t m_local, int64_t n, int64_t row_offset, double* dAsk, int64_t ldk);
rnla_status dev_gemv_n(const double* A, int64_t lda, int64_t m, int64_t n, const double* x, double* y);
rnla_status dev_gemv_t(const double* A, int64_t lda, int64_t m, int64_t n, const double* r, double* u);
rnla_status dev_qr_blocked(double* X, int64_t ldx, int64_t rows, int n, double* R, int64_t* deficient);
rnla_status dev_tri_inv_blocked(const double* R, int64_t ldr, int n, double* Rinv, int64_t ldi);
rnla_status dev_blendenpik(const double* A, int64_t lda, int64_t m_local, int64_t n, const double* b, double epsilon,
                           int64_t maxit, double sampling_factor, int kind, int dist_or_width, int zeta, uint64_t seed,
                           double* x, int64_t* iters_out, int32_t* converged_out);

rnla_status dev_lsrn(const double* A, int64_t lda, int64_t m_local, int64_t n, const double* b, double epsilon, int64_t maxit,
                     double sampling_factor, int kind, int dist_or_width, int zeta, uint64_t seed, double* x,
                     int64_t* iters_out, int32_t* converged_out);

// src/cg.rs:77-117 on device buffers
rnla_status dev_conjugate_grad(const double* A, int64_t lda, int64_t n, const double* b, double* x, int64_t* iterations_out,
                               int32_t* converged_out);
rnla_status dev_verify_solution(const double* A, int64_t lda, int64_t m_local, int64_t n, const double* b, const double* x, double* out);
// lsqr of src/solvers.rs:115-278 on device buffers (u row-sharded with A; v, w, x, var replicated); arnorms is a HOST array
rnla_status dev_lsqr(const double* A, int64_t lda, int64_t m_local, int64_t n, const double* b, double damp, double atol, double btol,
                     double conlim, int64_t iter_lim, int calc_var, const double* x0, double* x, rnla_lsqr_result* res,
                     double* arnorms, int64_t arnorms_cap, double* var);

// normal_pass.cu: u = cq (A x) + cy y, t[0..n) = A^T u, t[n] = u . u with A streamed once (clusters hold row slabs in shared memory)
bool normal_pass_supported(const double* A, int64_t lda, int64_t m_local, int64_t n);
struct NormalPassPlan { int cluster, ncb, ne, shift_e, shift_o, pitch, stage_bytes; };
NormalPassPlan normal_pass_plan(uint64_t base_address, int64_t lda, int64_t n);
rnla_status dev_normal_pass(const double* A, int64_t lda, int64_t m_local, int64_t n, const double* x, double cq, const double* y, double cy,
                            double* uout, double* t);

rnla_status dev_small_gemv(const double* M, int64_t ld, int n, int trans, const double* x, double* y);
rnla_status dev_axpby_vec(double a, const double* x, double b, double* y, int64_t n);
rnla_status dev_cgls_operator(const double* A, int64_t lda, int64_t m_local, int64_t n, const double* b, const double* M, double* z,
                              double epsilon, int64_t maxit, int64_t* it_out, int32_t* conv_out);

// pivot.cu: column-pivoted Householder QR (reference src/pivot_decompositions.rs:105-269) and index shuffles
rnla_status dev_qrcp(double* R, int64_t ldr, int64_t m, int64_t n, int64_t steps, int64_t* dperm, double* Q, int64_t ldq, int64_t qcols);
rnla_status dev_lupp(double* W, int64_t ld, int64_t n, double* L, int64_t ldl, double* U, int64_t ldu, int64_t* dperm,
                     int64_t* singular_step);
rnla_status dev_gather_columns(const double* A, int64_t lda, int64_t m, const int64_t* dJ, int64_t k, double* out, int64_t ldo);
rnla_status dev_gather_rows(const double* A, int64_t lda, int64_t n, const int64_t* dI, int64_t k, double* out, int64_t ldo);
rnla_status dev_scatter_rows(const double* M, int64_t ldm, int64_t k, const int64_t* dJ, double* W, int64_t ldw);
rnla_status dev_build_interp(const double* T, int64_t ldt, int64_t k, int64_t w, const int64_t* dperm, double* X, int64_t ldx);
rnla_status dev_backsolve_upper(const double* U, int64_t ldu, int n, double* y, double* x);
rnla_status dev_diag_solve(const double* s, int n, double* z);
rnla_status dev_saddle_weights(const double* s, int n, double mu, double* w);
rnla_status dev_mul_vec(const double* w, int n, double* z);

// next_rows.cu: SURVEY.md section 8f rows 2-4 and the saddle-point driver
rnla_status dev_sap_chol_qrcp(const double* A, int64_t lda, int64_t m, int64_t n, int64_t d, int kind, int dist_or_width, int zeta,
                              uint64_t seed, double* Q, int64_t ldq, double* R, int64_t ldr, int64_t* dJ, int64_t* k_out);
rnla_status dev_sketched_least_squares(int which, const double* A, int64_t lda, int64_t m, int64_t n, const double* b, int kind,
                                       int dist_or_width, int zeta, uint64_t seed, double* x);
rnla_status dev_osid_qrcp(const double* Y, int64_t ldy, int64_t l, int64_t w, int64_t k, int attr, double* X, int64_t ldx, int64_t* dJ);
rnla_status dev_osid_randomised(const double* A, int64_t lda, int64_t m, int64_t n, int64_t k, int attr, const rnla_options& o,
                                double* X, int64_t ldx, int64_t* dJ);
rnla_status dev_two_sided_id(int randomised, const double* A, int64_t lda, int64_t m, int64_t n, int64_t k, const rnla_options& o,
                             double* Z, int64_t ldz, int64_t* dI, int64_t* dJ, double* X, int64_t ldx);
rnla_status dev_cur(int randomised, const double* A, int64_t lda, int64_t m, int64_t n, int64_t k, const rnla_options& o,
                    int64_t* dJ, double* U, int64_t ldu, int64_t* dI);
rnla_status dev_saddle_point(const double* A, int64_t lda, int64_t m, int64_t n, const double* b, const double* cvec, double mu,
                             double epsilon, int64_t maxit, double sampling_factor, int dist, uint64_t seed, double* x, double* y,
                             int64_t* iters_out, int32_t* converged_out);

// i8gemm.cu: the passes over A on the INT8 tensor cores (tcgen05 kind::i8), rnla_options.range_passes_int8
bool i8_supported(int64_t m, int64_t n, int l);
bool i8_active_for(const double* A, int64_t lda, int64_t m, int64_t n, int64_t N);
void i8_deactivate();
// precision of the products that follow: 4 digit planes (31-bit operands; all_pairs adds the sweep over groups 4..6),
// 6 (47-bit) or 7 (55-bit, FP64-grade)
void i8_set_precision(int planes, bool all_pairs);
void i8_release();
void i8_debug_flush(int stages);
void i8_free_workspace();
// usable = false: A holds Inf / NaN or rows too small to scale -> the caller keeps the FP64 kernels
rnla_status i8_prepare(const double* A, int64_t lda, int64_t m, int64_t n, int planes, bool* usable);
rnla_status i8_prepare_begin(const double* A, int64_t lda, int64_t m, int64_t n, int planes);
rnla_status i8_prepare_rows(int64_t r0, int64_t count, bool phases = false);
rnla_status i8_prepare_end(bool* usable);
rnla_status i8_gemm_nn(const double* B, int64_t ldb, int64_t N, double* C, int64_t ldc);
rnla_status i8_gemm_tn(const double* Q, int64_t ldq, int64_t N, double* Z, int64_t ldz);
// C = A * Omega with Omega(k, c) drawn from the Philox entry map inside the operand kernels (never materialised in FP64)
rnla_status i8_gemm_nn_omega(int dist, uint64_t seed, uint32_t stream, int64_t N, double* C, int64_t ldc);
// how a driver spends the integer tensor cores (resolved from rnla_options.range_passes_int8 and the shape)
struct I8Plan {
    int stored = 0;                       // digit planes of the split of A (0: FP64 kernels everywhere)
    int early = 4; bool early_all = false;  // A Omega, A^T Y: products that shape the sketch
    int last = 4;  bool last_all = true;    // the last A S, whose range becomes Q
    int carry = 0;                        // the product that carries the values (Q^T A, Nystrom Y = A S): planes, 0 = FP64
};
I8Plan i8_plan(const rnla_options& o, int64_t m_local, int64_t n, int l);

// ziggurat.cu: the reference's Gaussian entries (rand_distr 0.4.3 StandardNormal on the sequential ThreeFry stream, sketch.rs:112-117)
rnla_status fill_threefry_gaussian(uint64_t key0, uint64_t key1, int64_t rows, int64_t cols, double* out, int64_t ld, int64_t* words_consumed);
void ziggurat_tables_host(double* x257, double* f257);
void threefry_key_from_u64(uint64_t state, uint64_t key[2]);
rnla_status fill_operator(int generator, int dist, uint64_t seed, uint32_t stream, int64_t rows, int64_t cols, int64_t row_off, double* out,
                          int64_t ld);

// literal.cu: bug-compatible pieces of the reference
rnla_status literal_tsog1(const double* A, int64_t lda, const ShardInfo& sh, int64_t n, int l, int q, int pps,
                          const rnla_options& o, double* S);
rnla_status dev_stabilizer(const double* X, int64_t ldx, int64_t rows, int64_t cols, double* L, int64_t ldl);

}  // namespace rnla
