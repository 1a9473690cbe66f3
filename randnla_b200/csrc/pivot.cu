// Column-pivoted Householder QR on the device and the index-shuffling kernels around it:
// reference src/pivot_decompositions.rs:105-180 (`qrcp`) and :196-269 (`economic_qrcp`), src/solvers.rs:22-69.
//
// The reference pivots on EXACTLY recomputed trailing column norms (it re-reads every trailing column after each
// reflection, :169-171), first maximum wins (:122-127).  The same rule is kept here, fused: the kernel that reflects a column
// also accumulates the norm of its rows below the new diagonal, so the recomputation costs no extra pass.  Two launches per
// step: (1) one CTA picks the pivot, swaps the two columns (all rows, as `swap_columns` does), builds the unit reflector
// v = (x + sign(x0) |x| e1) / |..| (:137-146); (2) one warp (short columns) or one CTA (long columns) per trailing column
// applies r_j -= 2 v (v . r_j) (:149-154).  These matrices are sketches (d x n with d = O(n), or k x n): the whole trailing
// matrix is L2-resident or close to it, and the work is latency-bound BLAS-2 by construction of the algorithm.
#include "drivers.cuh"
#include "gemm.cuh"
#include "panel.cuh"
#include <algorithm>
#include <cmath>
#include <vector>

namespace rnla {

namespace {

constexpr int PV_T = 1024;

__device__ __forceinline__ double block_sum(double v, double* red) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const int nw = blockDim.x >> 5;
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    double t = 0.0;
    for (int w = 0; w < nw; ++w) t += red[w];
    return t;
}

// norms[j] = || R[row0.., j] ||, one warp per column
__global__ void __launch_bounds__(256)
col_norms_kernel(const double* __restrict__ R, int64_t ld, int64_t m, int64_t n, int64_t row0, double* __restrict__ norms) {
    const int64_t j = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (j >= n) return;
    const int lane = threadIdx.x & 31;
    const double* c = R + j * ld;
    double s = 0.0;
    for (int64_t i = row0 + lane; i < m; i += 32) s = fma(c[i], c[i], s);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) norms[j] = sqrt(s);
}

// step ks: pivot choice, column swap, reflector.  v (length m - ks) is written to vout; flag[0] = 1 iff |x| != 0.
__global__ void __launch_bounds__(PV_T)
qrcp_pivot_kernel(double* __restrict__ R, int64_t ld, int64_t m, int64_t n, int64_t ks, double* __restrict__ norms,
                  int64_t* __restrict__ perm, double* __restrict__ vout, int* __restrict__ flag) {
    __shared__ double red[32];
    __shared__ double bval[32];
    __shared__ long long bidx[32];
    __shared__ long long s_mi;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // first maximum of norms[ks..n): strict > inside a thread's increasing walk, (value, -index) order across threads
    double best = -1.0; long long bi = n;
    for (int64_t j = ks + tid; j < n; j += PV_T) { const double v = norms[j]; if (v > best) { best = v; bi = j; } }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double ov = __shfl_xor_sync(0xffffffffu, best, o);
        const long long oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
    }
    if (lane == 0) { bval[warp] = best; bidx[warp] = bi; }
    __syncthreads();
    if (tid == 0) {
        double b = bval[0]; long long i = bidx[0];
        for (int w = 1; w < PV_T / 32; ++w) if (bval[w] > b || (bval[w] == b && bidx[w] < i)) { b = bval[w]; i = bidx[w]; }
        if (i >= n) i = ks;           // all NaN: keep the column where it is
        s_mi = i;
        if (i != ks) {
            const int64_t tp = perm[ks]; perm[ks] = perm[i]; perm[i] = tp;
            const double tn = norms[ks]; norms[ks] = norms[i]; norms[i] = tn;
        }
    }
    __syncthreads();
    const int64_t mi = s_mi;
    double* ck = R + ks * ld;
    if (mi != ks) {
        double* cm = R + mi * ld;
        for (int64_t i = tid; i < m; i += PV_T) { const double t = ck[i]; ck[i] = cm[i]; cm[i] = t; }
    }
    __syncthreads();
    const int64_t len = m - ks;
    double s = 0.0;
    for (int64_t i = tid; i < len; i += PV_T) { const double x = ck[ks + i]; s = fma(x, x, s); }
    s = block_sum(s, red);
    const double norm_x = sqrt(s);
    if (norm_x == 0.0) { if (tid == 0) flag[0] = 0; return; }
    const double x0 = ck[ks];
    const double v0 = x0 + (x0 >= 0.0 ? norm_x : -norm_x);
    double s2 = 0.0;
    for (int64_t i = tid; i < len; i += PV_T) { const double x = (i == 0) ? v0 : ck[ks + i]; s2 = fma(x, x, s2); }
    s2 = block_sum(s2, red);
    const double nv = sqrt(s2);
    for (int64_t i = tid; i < len; i += PV_T) { const double x = (i == 0) ? v0 : ck[ks + i]; vout[i] = x / nv; }
    if (tid == 0) flag[0] = 1;
}

// columns j0 + [0, ncols): c[row0 + i] -= 2 v[i] (v . c[row0..]); optionally norms[j] = || c[row0 + 1 ..] ||.
// G threads per column (32: shuffles only; 256: one CTA per column).
template <int G>
__global__ void __launch_bounds__(256)
reflect_kernel(double* __restrict__ R, int64_t ld, int64_t m, int64_t row0, int64_t j0, int64_t ncols,
               const double* __restrict__ v, const int* __restrict__ flag, double* __restrict__ norms, int64_t norm_from) {
    __shared__ double red[8];
    if (flag && flag[0] == 0) return;
    constexpr int CPB = 256 / G;
    const int64_t jj = (int64_t)blockIdx.x * CPB + threadIdx.x / G;
    const int t = threadIdx.x % G;
    const bool live = jj < ncols;
    const int64_t len = m - row0;
    double* c = R + (j0 + (live ? jj : 0)) * ld + row0;
    double dot = 0.0;
    if (live) for (int64_t i = t; i < len; i += G) dot = fma(v[i], c[i], dot);
    if (G == 32) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
    } else {
        dot = block_sum(dot, red);
    }
    double nn = 0.0;
    if (live) {
        for (int64_t i = t; i < len; i += G) {
            // `r[(i, j)] -= 2.0 * v[i - k] * dot_product` (:152): three separately rounded operations
            const double x = __dsub_rn(c[i], __dmul_rn(__dmul_rn(2.0, v[i]), dot));
            c[i] = x;
            if (i >= 1) nn = fma(x, x, nn);
        }
    }
    if (norms) {
        if (G == 32) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) nn += __shfl_xor_sync(0xffffffffu, nn, o);
        } else {
            nn = block_sum(nn, red);
        }
        if (live && t == 0 && j0 + jj >= norm_from) norms[j0 + jj] = sqrt(nn);
    }
}

__global__ void iota_kernel(int64_t* __restrict__ p, int64_t n) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) p[i] = i;
}

// out(:, c) = A(:, J[c])
__global__ void __launch_bounds__(256)
gather_columns_kernel(const double* __restrict__ A, int64_t lda, int64_t m, const int64_t* __restrict__ J, int64_t k,
                      double* __restrict__ out, int64_t ldo) {
    const int64_t c = blockIdx.y;
    const double* src = A + J[c] * lda;
    double* dst = out + c * ldo;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += (int64_t)gridDim.x * blockDim.x) dst[i] = src[i];
}
// out(i, :) = A(I[i], :), out is k x n
__global__ void __launch_bounds__(256)
gather_rows_kernel(const double* __restrict__ A, int64_t lda, int64_t n, const int64_t* __restrict__ I, int64_t k,
                   double* __restrict__ out, int64_t ldo) {
    const int64_t total = k * n;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
        const int64_t j = idx / k, i = idx - j * k;
        out[i + j * ldo] = A[I[i] + j * lda];
    }
}
// W (n x k, zeroed by the caller): W(J[i], :) = M(i, :) for i < k   (A[:, J[:k]] M = A W)
__global__ void __launch_bounds__(256)
scatter_rows_kernel(const double* __restrict__ M, int64_t ldm, int64_t k, const int64_t* __restrict__ J, double* __restrict__ W, int64_t ldw) {
    const int64_t total = k * k;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
        const int64_t c = idx / k, i = idx - c * k;
        W[J[i] + c * ldw] = M[i + c * ldm];
    }
}
// interpolation matrix of the column ID (src/id.rs:293-308): X (k x w), X(:, p[idx]) = e_idx (idx < k), X(:, p[k + c]) = T(:, c)
__global__ void __launch_bounds__(256)
build_interp_kernel(const double* __restrict__ T, int64_t ldt, int64_t k, int64_t w, const int64_t* __restrict__ p,
                    double* __restrict__ X, int64_t ldx) {
    const int64_t total = k * w;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
        const int64_t c = idx / k, i = idx - c * k;
        X[i + p[c] * ldx] = (c < k) ? (i == c ? 1.0 : 0.0) : T[i + (c - k) * ldt];
    }
}

// solve_upper_triangular_system (src/solvers.rs:22-41): x_i = (y_i - sum_{j>i} u_ij x_j) / u_ii, rows with u_ii == 0 keep
// x_i = 0.  One CTA, column sweeps (coalesced): after x_i is known, y[0..i) -= x_i u[0..i, i].  y is destroyed.
__global__ void __launch_bounds__(1024)
backsolve_upper_kernel(const double* __restrict__ U, int64_t ldu, int n, double* __restrict__ y, double* __restrict__ x) {
    __shared__ double xi_s;
    for (int i = n - 1; i >= 0; --i) {
        if (threadIdx.x == 0) {
            const double d = U[i + (int64_t)i * ldu];
            const double xi = (d != 0.0) ? y[i] / d : 0.0;
            x[i] = xi; xi_s = xi;
        }
        __syncthreads();
        const double xi = xi_s;
        if (xi != 0.0) {
            const double* col = U + (int64_t)i * ldu;
            for (int r = threadIdx.x; r < i; r += blockDim.x) y[r] = fma(-xi, col[r], y[r]);
        }
        __syncthreads();
    }
}
// solve_diagonal_system (src/solvers.rs:57-69): z_i /= s_i where s_i != 0, else 0
__global__ void diag_solve_kernel(const double* __restrict__ s, int n, double* __restrict__ z) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) z[i] = (s[i] != 0.0) ? z[i] / s[i] : 0.0;
}
// w_j = 1 / sqrt(s_j^2 + mu) (mu > 0) or 1 / s_j (mu == 0)                 src/sketch_and_precondition.rs:186, :189
__global__ void saddle_weights_kernel(const double* __restrict__ s, int n, double mu, double* __restrict__ w) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) w[i] = (mu > 0.0) ? 1.0 / sqrt(s[i] * s[i] + mu) : 1.0 / s[i];
}
__global__ void mul_vec_kernel(const double* __restrict__ w, int n, double* __restrict__ z) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) z[i] *= w[i];
}

inline unsigned grid_for(int64_t n, int cap = 148 * 16) { return (unsigned)std::max<int64_t>(1, std::min<int64_t>((n + 255) / 256, cap)); }

rnla_status launch_reflect(double* R, int64_t ld, int64_t m, int64_t row0, int64_t j0, int64_t ncols, const double* v,
                           const int* flag, double* norms, int64_t norm_from) {
    Ctx& c = ctx();
    if (ncols <= 0 || m - row0 <= 0) return RNLA_OK;
    if (m - row0 <= 2048)
        reflect_kernel<32><<<(unsigned)((ncols + 7) / 8), 256, 0, c.stream>>>(R, ld, m, row0, j0, ncols, v, flag, norms, norm_from);
    else
        reflect_kernel<256><<<(unsigned)ncols, 256, 0, c.stream>>>(R, ld, m, row0, j0, ncols, v, flag, norms, norm_from);
    ++g_kernel_launches;
    RNLA_CUDA(cudaGetLastError());
    return RNLA_OK;
}

}  // namespace

// Householder QR with column pivoting, `steps` steps, in place on R (m x n, ld): afterwards R holds what the reference's
// work matrix `r` holds (rows 0..steps of it are `r_eco`).  dperm: n int64 on the device.  Q (m x qcols, optional): the first
// qcols columns of H_0 H_1 ... H_{steps-1} (qcols = k: `q_eco`, qcols = m: the reference's full `q`).
rnla_status dev_qrcp(double* R, int64_t ldr, int64_t m, int64_t n, int64_t steps, int64_t* dperm, double* Q, int64_t ldq, int64_t qcols) {
    Ctx& c = ctx();
    if (m <= 0 || n <= 0) return fail(RNLA_ERR_INVALID_DIMENSIONS, "qrcp: empty matrix");
    if (n * 65535LL < 0 || (n + 7) / 8 > 2147483647LL) return fail(RNLA_ERR_INVALID_DIMENSIONS, "qrcp: too many columns");
    steps = std::min(steps, std::min(m, n));
    PhaseScope ph("qrcp");
    DevBuf norms, V, flags;
    const bool keep = Q != nullptr && qcols > 0;
    RNLA_CUDA(norms.alloc((size_t)n * 8));
    RNLA_CUDA(V.alloc((size_t)m * (keep ? std::max<int64_t>(steps, 1) : 1) * 8));
    RNLA_CUDA(flags.alloc((size_t)std::max<int64_t>(steps, 1) * 4));
    iota_kernel<<<grid_for(n), 256, 0, c.stream>>>(dperm, n);
    col_norms_kernel<<<(unsigned)((n + 7) / 8), 256, 0, c.stream>>>(R, ldr, m, n, 0, norms.d());
    g_kernel_launches += 2;
    RNLA_CUDA(cudaGetLastError());
    for (int64_t ks = 0; ks < steps; ++ks) {
        double* v = V.d() + (keep ? ks * m : 0);
        int* flag = flags.as<int>() + ks;
        qrcp_pivot_kernel<<<1, PV_T, 0, c.stream>>>(R, ldr, m, n, ks, norms.d(), dperm, v, flag);
        ++g_kernel_launches;
        RNLA_CUDA(cudaGetLastError());
        RNLA_TRY(launch_reflect(R, ldr, m, ks, ks, n - ks, v, flag, norms.d(), ks + 1));
    }
    if (keep) {
        RNLA_CUDA(set_identity(Q, ldq, m, qcols, c.stream));
        for (int64_t ks = steps - 1; ks >= 0; --ks)        // H_ks leaves e_c, c < ks, alone
            RNLA_TRY(launch_reflect(Q, ldq, m, ks, ks, qcols - ks, V.d() + ks * m, flags.as<int>() + ks, nullptr, 0));
    }
    return RNLA_OK;
}

rnla_status dev_gather_columns(const double* A, int64_t lda, int64_t m, const int64_t* dJ, int64_t k, double* out, int64_t ldo) {
    if (m <= 0 || k <= 0) return RNLA_OK;
    gather_columns_kernel<<<dim3(grid_for(m, 64), (unsigned)k), 256, 0, ctx().stream>>>(A, lda, m, dJ, k, out, ldo);
    ++g_kernel_launches;
    RNLA_CUDA(cudaGetLastError());
    return RNLA_OK;
}
rnla_status dev_gather_rows(const double* A, int64_t lda, int64_t n, const int64_t* dI, int64_t k, double* out, int64_t ldo) {
    if (n <= 0 || k <= 0) return RNLA_OK;
    gather_rows_kernel<<<grid_for(k * n), 256, 0, ctx().stream>>>(A, lda, n, dI, k, out, ldo);
    ++g_kernel_launches;
    RNLA_CUDA(cudaGetLastError());
    return RNLA_OK;
}
rnla_status dev_scatter_rows(const double* M, int64_t ldm, int64_t k, const int64_t* dJ, double* W, int64_t ldw) {
    if (k <= 0) return RNLA_OK;
    scatter_rows_kernel<<<grid_for(k * k), 256, 0, ctx().stream>>>(M, ldm, k, dJ, W, ldw);
    ++g_kernel_launches;
    RNLA_CUDA(cudaGetLastError());
    return RNLA_OK;
}
rnla_status dev_build_interp(const double* T, int64_t ldt, int64_t k, int64_t w, const int64_t* dperm, double* X, int64_t ldx) {
    build_interp_kernel<<<grid_for(k * w), 256, 0, ctx().stream>>>(T, ldt, k, w, dperm, X, ldx);
    ++g_kernel_launches;
    RNLA_CUDA(cudaGetLastError());
    return RNLA_OK;
}
rnla_status dev_backsolve_upper(const double* U, int64_t ldu, int n, double* y, double* x) {
    backsolve_upper_kernel<<<1, 1024, 0, ctx().stream>>>(U, ldu, n, y, x);
    ++g_kernel_launches;
    RNLA_CUDA(cudaGetLastError());
    return RNLA_OK;
}
rnla_status dev_diag_solve(const double* s, int n, double* z) {
    diag_solve_kernel<<<(n + 255) / 256, 256, 0, ctx().stream>>>(s, n, z);
    ++g_kernel_launches;
    RNLA_CUDA(cudaGetLastError());
    return RNLA_OK;
}
rnla_status dev_saddle_weights(const double* s, int n, double mu, double* w) {
    saddle_weights_kernel<<<(n + 255) / 256, 256, 0, ctx().stream>>>(s, n, mu, w);
    ++g_kernel_launches;
    RNLA_CUDA(cudaGetLastError());
    return RNLA_OK;
}
rnla_status dev_mul_vec(const double* w, int n, double* z) {
    mul_vec_kernel<<<(n + 255) / 256, 256, 0, ctx().stream>>>(w, n, z);
    ++g_kernel_launches;
    RNLA_CUDA(cudaGetLastError());
    return RNLA_OK;
}

// ======================================================================================================================
// lupp(matrix) of src/pivot_decompositions.rs:21-86: LU with partial (row) pivoting, first maximum wins (strict `>`, :36-42),
// Gaussian elimination with separately rounded multiply and subtract (:63-70) -- kept operation for operation
// (__ddiv_rn / __dmul_rn / __dsub_rn, no FMA contraction), so L, U and p are BIT-IDENTICAL to the reference's arithmetic.
// One launch per elimination step: every CTA owns LU_CPB trailing columns and redundantly finds the pivot of column k (an
// L2-resident column), takes the swapped values of rows k / pivot_row of its own columns into registers, and applies the
// rank-1 update to them; CTA 0 also records the multipliers (L), the pivot (U_kk), the permutation and swaps the two rows of
// the earlier L columns.  Column k of the work matrix is never written during its own step (the other CTAs are reading it):
// L and U are separate outputs.  flag[0] = 1 + k marks a zero pivot at step k (SingularMatrix, :44-48).
namespace {
constexpr int LU_T = 256;
constexpr int LU_CPB = 4;
__global__ void __launch_bounds__(LU_T)
lupp_step_kernel(double* __restrict__ W, int64_t ld, int64_t n, int64_t k, double* __restrict__ L, int64_t ldl,
                 double* __restrict__ U, int64_t ldu, int64_t* __restrict__ perm, int* __restrict__ flag) {
    __shared__ double bval[LU_T / 32];
    __shared__ long long bidx[LU_T / 32];
    __shared__ long long s_pr;
    __shared__ double s_pv;
    if (flag[0] != 0) return;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const double* colk = W + k * ld;
    double best = -1.0; long long bi = n;
    for (int64_t i = k + tid; i < n; i += LU_T) { const double v = fabs(colk[i]); if (v > best) { best = v; bi = i; } }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double ov = __shfl_xor_sync(0xffffffffu, best, o);
        const long long oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
    }
    if (lane == 0) { bval[warp] = best; bidx[warp] = bi; }
    __syncthreads();
    if (tid == 0) {
        double b = bval[0]; long long i = bidx[0];
        for (int w = 1; w < LU_T / 32; ++w) if (bval[w] > b || (bval[w] == b && bidx[w] < i)) { b = bval[w]; i = bidx[w]; }
        if (i >= n) { i = k; b = fabs(colk[k]); }           // all NaN: `val > pivot_val` never fires, the diagonal stays
        s_pr = i; s_pv = b;
    }
    __syncthreads();
    const int64_t pr = s_pr;
    if (s_pv == 0.0) { if (blockIdx.x == 0 && tid == 0) flag[0] = 1 + (int)k; return; }
    const double pivot = colk[pr];                           // lu[(k, k)] after the swap
    const double okk = colk[k];                              // what row pr holds in column k after the swap
    const int64_t j0 = k + 1 + (int64_t)blockIdx.x * LU_CPB;
    const int nc = (int)min((int64_t)LU_CPB, n - j0);
    double ukj[LU_CPB], okj[LU_CPB];
#pragma unroll
    for (int c = 0; c < LU_CPB; ++c) {
        ukj[c] = c < nc ? W[pr + (j0 + c) * ld] : 0.0;      // row k of column j after the swap
        okj[c] = c < nc ? W[k + (j0 + c) * ld] : 0.0;       // row pr of column j after the swap
    }
    __syncthreads();                                         // every thread holds both rows before anyone overwrites them
    for (int64_t i = k + 1 + tid; i < n; i += LU_T) {
        const double mult = __ddiv_rn(i == pr ? okk : colk[i], pivot);                             // :64
#pragma unroll
        for (int c = 0; c < LU_CPB; ++c)
            if (c < nc) {
                double* w = W + i + (j0 + c) * ld;
                *w = __dsub_rn(i == pr ? okj[c] : *w, __dmul_rn(mult, ukj[c]));                     // :67-69
            }
        if (blockIdx.x == 0) L[i + k * ldl] = mult;                                                // :65
    }
    if (tid < nc) { W[k + (j0 + tid) * ld] = ukj[tid]; U[k + (j0 + tid) * ldu] = ukj[tid]; }
    if (blockIdx.x == 0) {
        if (tid == 0) {
            U[k + k * ldu] = pivot;
            if (pr != k) { const int64_t t = perm[k]; perm[k] = perm[pr]; perm[pr] = t; }          // :57-59
        }
        if (pr != k)                                                                               // :51-56 on the stored multipliers
            for (int64_t j = tid; j < k; j += LU_T) { const double t = L[k + j * ldl]; L[k + j * ldl] = L[pr + j * ldl]; L[pr + j * ldl] = t; }
    }
}
__global__ void lupp_init_kernel(double* __restrict__ L, int64_t ldl, double* __restrict__ U, int64_t ldu, int64_t n,
                                 int64_t* __restrict__ perm, int* __restrict__ flag) {
    const int64_t total = n * n;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
        const int64_t c = idx / n, r = idx - c * n;
        L[r + c * ldl] = r == c ? 1.0 : 0.0;                                                      // :74
        U[r + c * ldu] = 0.0;                                                                     // :75
        if (c == 0) perm[r] = r;                                                                  // :30
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) flag[0] = 0;
}
}  // namespace

// W: n x n, holds the matrix on entry and is destroyed.  L, U: n x n outputs.  dperm: n.  *singular_step = -1, or the step
// at which the pivot was exactly zero (the reference's SingularMatrix error).
rnla_status dev_lupp(double* W, int64_t ld, int64_t n, double* L, int64_t ldl, double* U, int64_t ldu, int64_t* dperm,
                     int64_t* singular_step) {
    Ctx& c = ctx();
    PhaseScope ph("lupp");
    DevBuf flag;
    RNLA_CUDA(flag.alloc(8));
    lupp_init_kernel<<<(unsigned)std::max<int64_t>(1, std::min<int64_t>((n * n + 255) / 256, 148 * 8)), 256, 0, c.stream>>>(
        L, ldl, U, ldu, n, dperm, flag.as<int>());
    ++g_kernel_launches;
    for (int64_t k = 0; k + 1 < n; ++k) {                                                         // :32
        const unsigned grid = (unsigned)((n - k - 1 + LU_CPB - 1) / LU_CPB);
        lupp_step_kernel<<<grid, LU_T, 0, c.stream>>>(W, ld, n, k, L, ldl, U, ldu, dperm, flag.as<int>());
        ++g_kernel_launches;
    }
    RNLA_CUDA(cudaGetLastError());
    // the last diagonal entry is never a pivot (:32 stops at n - 2): it is what the eliminations left
    RNLA_CUDA(cudaMemcpyAsync(U + (n - 1) + (n - 1) * ldu, W + (n - 1) + (n - 1) * ld, 8, cudaMemcpyDeviceToDevice, c.stream));
    int hflag = 0;
    RNLA_CUDA(cudaMemcpyAsync(&hflag, flag.p, 4, cudaMemcpyDeviceToHost, c.stream));
    RNLA_CUDA(cudaStreamSynchronize(c.stream));
    *singular_step = hflag ? hflag - 1 : -1;
    return RNLA_OK;
}

}  // namespace rnla
