// Bug-compatible ("literal") pieces of the reference, on the GPU:
//   * Stabilizer(X) = X.clone().full_piv_lu().l()            reference src/lora_helpers.rs:144-146
//   * tsog1 exactly as written, including the zero-S1 defect  reference src/lora_helpers.rs:58-105
//
// The full-pivot LU follows nalgebra 0.33.0 `FullPivLU::new` + `lu::gauss_step(_swap)` [not vendored in
// the reference tree; restated from the published algorithm, see DESIGN.md "oracle pinning"]:
//   pivot  = first entry of maximal modulus of the trailing block in column-major scan order (`icamax_full`)
//   swap   = whole columns i <-> col_piv, whole rows i <-> row_piv
//   step   = coeffs *= 1/diag ;  trailing(r,c) = (-pivot_row[c]) * coeffs[r] + trailing(r,c)   (separate mul and add)
//   break  = on an exactly zero pivot
// and `.l()` = strictly-lower part of the first min(rows, cols) columns with a unit diagonal.  The row
// permutation is NOT applied to the returned L (that is the reference's behaviour).
// Every floating-point operation is issued as an explicit _rn intrinsic so that the CPU oracle, which does
// the same operations in the same order, agrees bit for bit.
#include "drivers.cuh"
#include "gemm.cuh"
#include "panel.cuh"
#include <algorithm>
#include <vector>

namespace rnla {

namespace {

__device__ __forceinline__ bool better(double v, long long i, double bv, long long bi) {
    return v > bv || (v == bv && i < bi);
}

__global__ void __launch_bounds__(256)
lu_argmax_partial_kernel(const double* __restrict__ X, int64_t ld, int64_t rows, int64_t cols, int64_t i,
                         double* __restrict__ pval, long long* __restrict__ pidx) {
    __shared__ double s_v[8];
    __shared__ long long s_i[8];
    const int64_t sr = rows - i, sc = cols - i, total = sr * sc;
    double bv = -1.0; long long bi = 0x7fffffffffffffffLL;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
        const int64_t c = idx / sr, r = idx - c * sr;
        const double v = fabs(X[(i + r) + (i + c) * ld]);
        if (better(v, idx, bv, bi)) { bv = v; bi = idx; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double ov = __shfl_xor_sync(0xffffffffu, bv, o);
        const long long oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (better(ov, oi, bv, bi)) { bv = ov; bi = oi; }
    }
    if ((threadIdx.x & 31) == 0) { s_v[threadIdx.x >> 5] = bv; s_i[threadIdx.x >> 5] = bi; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; ++w) if (better(s_v[w], s_i[w], bv, bi)) { bv = s_v[w]; bi = s_i[w]; }
        pval[blockIdx.x] = bv; pidx[blockIdx.x] = bi;
    }
}

// scratch layout: rowI[cols] rowP[cols] colI[rows] colC[rows]
__global__ void __launch_bounds__(256)
lu_gather_kernel(const double* __restrict__ X, int64_t ld, int64_t rows, int64_t cols, int64_t i,
                 const double* __restrict__ pval, const long long* __restrict__ pidx, int nb,
                 double* __restrict__ scratch, long long* __restrict__ piv, int* __restrict__ stop) {
    if (*stop) return;
    __shared__ double s_v[8];
    __shared__ long long s_i[8];
    __shared__ long long s_best;
    __shared__ double s_bestv;
    double bv = -1.0; long long bi = 0x7fffffffffffffffLL;
    for (int k = threadIdx.x; k < nb; k += blockDim.x) if (better(pval[k], pidx[k], bv, bi)) { bv = pval[k]; bi = pidx[k]; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double ov = __shfl_xor_sync(0xffffffffu, bv, o);
        const long long oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (better(ov, oi, bv, bi)) { bv = ov; bi = oi; }
    }
    if ((threadIdx.x & 31) == 0) { s_v[threadIdx.x >> 5] = bv; s_i[threadIdx.x >> 5] = bi; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; ++w) if (better(s_v[w], s_i[w], bv, bi)) { bv = s_v[w]; bi = s_i[w]; }
        s_best = bi; s_bestv = bv;
    }
    __syncthreads();
    const int64_t sr = rows - i;
    const int64_t cp = i + s_best / sr, rp = i + s_best % sr;
    if (blockIdx.x == 0 && threadIdx.x == 0) { piv[0] = rp; piv[1] = cp; piv[2] = (s_bestv == 0.0) ? 1 : 0; }
    if (s_bestv == 0.0) return;            // the update kernel raises *stop
    double* rowI = scratch; double* rowP = rowI + cols; double* colI = rowP + cols; double* colC = colI + rows;
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, nt = (int64_t)gridDim.x * blockDim.x;
    for (int64_t c = tid; c < cols; c += nt) { rowI[c] = X[i + c * ld]; rowP[c] = X[rp + c * ld]; }
    for (int64_t r = tid; r < rows; r += nt) { colI[r] = X[r + i * ld]; colC[r] = X[r + cp * ld]; }
}

__global__ void __launch_bounds__(256)
lu_update_kernel(double* __restrict__ X, int64_t ld, int64_t rows, int64_t cols, int64_t i,
                 const double* __restrict__ scratch, const long long* __restrict__ piv, int* __restrict__ stop) {
    if (*stop) return;
    if (piv[2]) {                       // exactly zero pivot: nalgebra breaks out of the loop
        if (blockIdx.x == 0 && threadIdx.x == 0) *stop = 1;
        return;
    }
    const int64_t rp = piv[0], cp = piv[1];
    const double* rowI = scratch; const double* rowP = rowI + cols; const double* colI = rowP + cols; const double* colC = colI + rows;
    const double diag = rowP[cp];
    const double inv = __ddiv_rn(1.0, diag);
    const int64_t sr = rows - i, sc = cols - i, regB = sr * sc, total = regB + 2 * i;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
        if (idx >= regB) {
            // rows i and rp of the already factored columns c < i are exchanged
            const int64_t t = idx - regB, c = t >> 1;
            if (rp != i) { if (t & 1) X[rp + c * ld] = rowI[c]; else X[i + c * ld] = rowP[c]; }
            continue;
        }
        const int64_t c = i + idx / sr, r = i + idx % sr;
        // value of the permuted matrix at (rr, cc)
        auto perm = [&](int64_t rq, int64_t cq) -> double {
            const int64_t rr = (rq == i) ? rp : (rq == rp) ? i : rq;
            const int64_t cc = (cq == i) ? cp : (cq == cp) ? i : cq;
            if (rr == i) return rowI[cc];
            if (rr == rp) return rowP[cc];
            if (cc == i) return colI[rr];
            if (cc == cp) return colC[rr];
            return X[rr + cc * ld];     // rr == rq, cc == cq: the thread's own element
        };
        double out;
        if (r == i) out = perm(i, c);
        else {
            const double lr = __dmul_rn(perm(r, i), inv);
            if (c == i) out = lr;
            else out = __dadd_rn(__dmul_rn(-perm(i, c), lr), perm(r, c));
        }
        X[r + c * ld] = out;
    }
}

__global__ void __launch_bounds__(256)
lu_extract_L_kernel(const double* __restrict__ X, int64_t ldx, int64_t rows, int64_t mn, double* __restrict__ L, int64_t ldl) {
    const int64_t total = rows * mn;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
        const int64_t c = idx / rows, r = idx - c * rows;
        L[r + c * ldl] = (r > c) ? X[r + c * ldx] : (r == c ? 1.0 : 0.0);
    }
}

}  // namespace

cudaError_t fullpiv_lu_L(double* X, int64_t ldx, int64_t rows, int64_t cols, double* L, int64_t ldl,
                         double* scratch, cudaStream_t st) {
    // scratch: 2*cols + 2*rows doubles, then 1024 doubles + 1024 int64 partials, 3 int64 pivot record, 1 int stop
    const int64_t mn = std::min(rows, cols);
    double* sc = scratch;
    double* pval = sc + 2 * cols + 2 * rows;
    long long* pidx = reinterpret_cast<long long*>(pval + 1024);
    long long* piv = pidx + 1024;
    int* stop = reinterpret_cast<int*>(piv + 4);
    cudaError_t e = cudaMemsetAsync(stop, 0, sizeof(int), st);
    if (e != cudaSuccess) return e;
    for (int64_t i = 0; i < mn; ++i) {
        const int64_t total = (rows - i) * (cols - i);
        int nb = (int)std::min<int64_t>(1024, (total + 255) / 256);
        if (nb < 1) nb = 1;
        lu_argmax_partial_kernel<<<nb, 256, 0, st>>>(X, ldx, rows, cols, i, pval, pidx);
        ++g_kernel_launches;
        int gb = (int)std::min<int64_t>(256, (std::max(rows, cols) + 255) / 256);
        lu_gather_kernel<<<gb, 256, 0, st>>>(X, ldx, rows, cols, i, pval, pidx, nb, sc, piv, stop);
        ++g_kernel_launches;
        int ub = (int)std::min<int64_t>(148 * 8, (total + 2 * i + 255) / 256);
        if (ub < 1) ub = 1;
        lu_update_kernel<<<ub, 256, 0, st>>>(X, ldx, rows, cols, i, sc, piv, stop);
        ++g_kernel_launches;
        if (i == 0) {
            // a zero matrix stops at the first pivot: avoid 3*mn empty launches
            int h = 0;
            e = cudaMemcpyAsync(&h, stop, sizeof(int), cudaMemcpyDeviceToHost, st);
            if (e != cudaSuccess) return e;
            e = cudaStreamSynchronize(st);
            if (e != cudaSuccess) return e;
            if (h) break;
        }
    }
    const int64_t total = rows * mn;
    int eb = (int)std::min<int64_t>(148 * 8, (total + 255) / 256);
    if (eb < 1) eb = 1;
    lu_extract_L_kernel<<<eb, 256, 0, st>>>(X, ldx, rows, mn, L, ldl);
    ++g_kernel_launches;
    return cudaGetLastError();
}

size_t fullpiv_lu_scratch_bytes(int64_t rows, int64_t cols) {
    return (size_t)(2 * cols + 2 * rows + 1024) * 8 + 1024 * 8 + 4 * 8 + 16;
}

// Stabilizer on a device matrix (not destroyed)
rnla_status dev_stabilizer(const double* X, int64_t ldx, int64_t rows, int64_t cols, double* L, int64_t ldl) {
    Ctx& c = ctx();
    if (rows <= 0 || cols <= 0) return RNLA_OK;
    DevBuf W, sc;
    RNLA_CUDA(W.alloc((size_t)rows * cols * 8));
    RNLA_CUDA(sc.alloc(fullpiv_lu_scratch_bytes(rows, cols)));
    RNLA_CUDA(copy_matrix(X, ldx, W.d(), rows, rows, cols, c.stream));
    RNLA_CUDA(fullpiv_lu_L(W.d(), rows, rows, cols, L, ldl, sc.d(), c.stream));
    return RNLA_OK;
}

// tsog1 as written (reference src/lora_helpers.rs:58-105).  Single GPU only: the full-pivot LU of a
// row-sharded panel has no distributed counterpart in this build.
rnla_status literal_tsog1(const double* A, int64_t lda, const ShardInfo& sh, int64_t n, int l, int q, int pps,
                          const rnla_options& o, double* S) {
    Ctx& c = ctx();
    if (c.nranks > 1) return fail(RNLA_ERR_INVALID_PARAMETERS, "literal mode is single-GPU only");
    const int64_t m = sh.rows_local;
    const int64_t mm = std::max<int64_t>(m, 1);
    int done = 0;
    DevBuf S1, T, T2;
    RNLA_CUDA(S1.alloc((size_t)n * l * 8)); RNLA_CUDA(T.alloc((size_t)mm * l * 8)); RNLA_CUDA(T2.alloc((size_t)std::max(mm, n) * l * 8));
    bool s1_zero = true;
    RNLA_CUDA(cudaMemsetAsync(S, 0, (size_t)n * l * 8, c.stream));                                   // :67
    if (q % 2 == 0) {
        // :70-72  S = Omega(n x k) -- overwritten by the loop below, so only materialised when the loop does not run
        if (q < 2) RNLA_TRY(fill_operator(o.generator, o.dist, o.seed, 1 /* STREAM_RANGE_N */, n, l, 0, S, n));
    } else {
        // :73-82  S1 = A^T Omega(m x k); its stabilised copy `_S2` is discarded
        PhaseScope ph("tsog1:At_Omega");
        RNLA_TRY(fill_operator(o.generator, o.dist, o.seed, 2 /* STREAM_RANGE_M */, m, l, 0, T.d(), mm));
        RNLA_TRY(dev_gemm_tn(A, lda, m, n, T.d(), mm, l, S1.d(), n, false));
        s1_zero = false;
        done = 1;
    }
    int diff = q - done;                                                                              // :85
    if (diff >= 2) {
        // every loop iteration restarts from S1 (:89) and overwrites S, so only the LAST iteration is observable;
        // `passes_done` keeps counting across iterations and decides where that iteration stabilises (:91, :97)
        done += 2 * (diff / 2 - 1);
        {
            PhaseScope ph("pass:A*S1");
            if (s1_zero) RNLA_CUDA(cudaMemsetAsync(T.d(), 0, (size_t)mm * l * 8, c.stream));        // A * zeros
            else RNLA_TRY(dev_gemm_nn(A, lda, m, n, S1.d(), n, l, T.d(), mm));
        }
        ++done;
        const double* tall = T.d();
        if (done % pps == 0) {                                                                        // :91-94
            PhaseScope ph("stab:L(Y)");
            RNLA_TRY(dev_stabilizer(T.d(), mm, m, l, T2.d(), mm));
            tall = T2.d();
        }
        {
            PhaseScope ph("pass:At*Y");
            RNLA_TRY(dev_gemm_tn(A, lda, m, n, tall, mm, l, S, n, false));                            // :95
        }
        ++done;
        if (done % pps == 0) {                                                                        // :97-100
            PhaseScope ph("stab:L(S)");
            RNLA_TRY(dev_stabilizer(S, n, n, l, T2.d(), n));
            RNLA_CUDA(copy_matrix(T2.d(), n, S, n, n, l, c.stream));
        }
    } else if (q % 2 != 0) {
        // q == 1: S stays zeros (:67, :104)
    }
    return RNLA_OK;
}

}  // namespace rnla
