// Sketch step of sketch_and_precondition (reference src/sketch_and_precondition.rs:49-52, 105-107, 172-176)
// and the on-device synthetic test matrices.
//
// Dense operator: the reference draws S (d x m) and forms S*A.  Here S is defined through its transpose
// S^T (m x d), S^T(j, i) = omega(row j, col i): rows of S^T follow the rows of A, so a row-sharded A
// needs only its own rows of S^T and S*A = (S^T)^T A is the streaming TN kernel (K2) plus one all-reduce.
//
// Sparse-sign operator (SASO, K6; not in the reference): column j of S has `zeta`
// non-zeros +-1/sqrt(zeta) at rows h_1(j)..h_zeta(j) drawn from one Philox block per column.  A is
// column-major, so each CTA keeps the d-vector accumulators of a few output columns in shared memory and
// streams the matching columns of A exactly once with coalesced loads; the scatter is shared-memory atomics.
#include "drivers.cuh"
#include "gemm.cuh"
#include "panel.cuh"
#include "rng.cuh"
#include <algorithm>
#include <cmath>
#include <vector>

namespace rnla {

namespace {

constexpr int SASO_THREADS = 512;
constexpr int SASO_SMEM = 200 * 1024;

// row index and sign of the t-th non-zero of column j
struct SasoHash { int idx[8]; };   // sign in bit 31

__device__ __forceinline__ void saso_hash(uint64_t seed, uint64_t j, int64_t d, int zeta, int* idx, double* sgn) {
    if (d <= 32768) {
        const u32x4 b = philox4x32_10((uint32_t)j, (uint32_t)(j >> 32), 0u, 4u /* STREAM_SASO_ROWS */, (uint32_t)seed, (uint32_t)(seed >> 32));
        const uint32_t w[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
        for (int t = 0; t < 8; ++t) {
            if (t < zeta) {
                const uint32_t f = (w[t >> 1] >> ((t & 1) * 16)) & 0xffffu;
                idx[t] = (int)(((uint64_t)(f >> 1) * (uint64_t)d) >> 15);
                sgn[t] = (f & 1u) ? -1.0 : 1.0;
            }
        }
    } else {
#pragma unroll
        for (int blk = 0; blk < 2; ++blk) {
            if (blk * 4 < zeta) {
                const u32x4 b = philox4x32_10((uint32_t)j, (uint32_t)(j >> 32), (uint32_t)blk, 4u, (uint32_t)seed, (uint32_t)(seed >> 32));
                const uint32_t w[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int t = blk * 4 + e;
                    if (t < zeta) {
                        idx[t] = (int)(((uint64_t)(w[e] >> 1) * (uint64_t)d) >> 31);
                        sgn[t] = (w[e] & 1u) ? -1.0 : 1.0;
                    }
                }
            }
        }
    }
}

__global__ void __launch_bounds__(SASO_THREADS)
saso_kernel(uint64_t seed, int64_t d, int zeta, const double* __restrict__ A, int64_t lda, int64_t m, int64_t n,
            int64_t row_off, double* __restrict__ Ask, int64_t ldk, int CB, int64_t rows_per_chunk, double scale) {
    extern __shared__ double acc[];
    const int64_t c0 = (int64_t)blockIdx.x * CB;
    const int cbv = (int)min((int64_t)CB, n - c0);
    const int64_t j0 = (int64_t)blockIdx.y * rows_per_chunk;
    const int64_t j1 = min(m, j0 + rows_per_chunk);
    for (int64_t i = threadIdx.x; i < (int64_t)cbv * d; i += blockDim.x) acc[i] = 0.0;
    __syncthreads();
    for (int64_t j = j0 + threadIdx.x; j < j1; j += blockDim.x) {
        int idx[8]; double sgn[8];
        saso_hash(seed, (uint64_t)(row_off + j), d, zeta, idx, sgn);
        for (int cc = 0; cc < cbv; ++cc) {
            const double v = A[j + (c0 + cc) * lda];
            double* a = acc + (int64_t)cc * d;
#pragma unroll
            for (int t = 0; t < 8; ++t) if (t < zeta) atomicAdd(a + idx[t], sgn[t] * v);
        }
    }
    __syncthreads();
    for (int64_t i = threadIdx.x; i < (int64_t)cbv * d; i += blockDim.x) {
        const double v = acc[i];
        if (v != 0.0) atomicAdd(Ask + (i % d) + (c0 + i / d) * ldk, v * scale);
    }
}

__global__ void __launch_bounds__(256)
add_noise_kernel(double* __restrict__ A, int64_t lda, int64_t rows, int64_t cols, int64_t row_off, double scale,
                 uint64_t seed, uint32_t stream) {
    const uint64_t q_first = (uint64_t)row_off >> 2;
    const uint64_t q_last = (uint64_t)(row_off + rows - 1) >> 2;
    const int64_t nq = (int64_t)(q_last - q_first + 1);
    const int64_t total = nq * cols;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
        const int64_t c = idx / nq, qi = idx - c * nq;
        const uint64_t q = q_first + (uint64_t)qi;
        const u32x4 b = omega_block(seed, stream, q, (uint32_t)c);
        const uint32_t w[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int64_t r = (int64_t)(4 * q + e) - row_off;
            if (r >= 0 && r < rows) A[r + c * lda] += scale * (double)gauss_from_u32(w[e]);
        }
    }
}

}  // namespace

cudaError_t saso_apply(uint64_t seed, int64_t d, int zeta, const double* A, int64_t lda, int64_t m, int64_t n,
                       int64_t row_off, double* Ask, int64_t ldk, int sms, cudaStream_t st) {
    if (m <= 0 || n <= 0) return cudaSuccess;
    int CB = (int)std::min<int64_t>(8, SASO_SMEM / (d * 8));
    if (CB < 1) return cudaErrorInvalidValue;
    CB = (int)std::min<int64_t>(CB, n);
    const int64_t groups = (n + CB - 1) / CB;
    int64_t chunks = std::max<int64_t>(1, (4LL * sms + groups - 1) / groups);
    const int64_t maxchunks = std::max<int64_t>(1, m / 4096);
    chunks = std::min(chunks, maxchunks);
    chunks = std::min<int64_t>(chunks, 65535);
    const int64_t rpc = (m + chunks - 1) / chunks;
    chunks = (m + rpc - 1) / rpc;
    static bool attr = false;
    if (!attr) {
        cudaError_t e = cudaFuncSetAttribute(saso_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SASO_SMEM);
        if (e != cudaSuccess) return e;
        attr = true;
    }
    dim3 grid((unsigned)groups, (unsigned)chunks);
    saso_kernel<<<grid, SASO_THREADS, (size_t)CB * d * 8, st>>>(seed, d, zeta, A, lda, m, n, row_off, Ask, ldk, CB, rpc,
                                                               1.0 / std::sqrt((double)zeta));
    ++g_kernel_launches;
    return cudaGetLastError();
}

cudaError_t add_noise(double* A, int64_t lda, int64_t rows, int64_t cols, int64_t row_off, double scale,
                      uint64_t seed, uint32_t stream, cudaStream_t st) {
    if (rows <= 0 || cols <= 0) return cudaSuccess;
    const int64_t total = (rows / 4 + 2) * cols;
    int blocks = (int)std::min<int64_t>(148 * 16, (total + 255) / 256);
    add_noise_kernel<<<blocks, 256, 0, st>>>(A, lda, rows, cols, row_off, scale, seed, stream);
    ++g_kernel_launches;
    return cudaGetLastError();
}

rnla_status dev_sketch_apply(int kind, int dist, uint64_t seed, int64_t d, int zeta, const double* dA, int64_t lda,
                             int64_t m_local, int64_t n, int64_t row_offset, double* dAsk, int64_t ldk) {
    Ctx& c = ctx();
    if (d <= 0 || n <= 0) return fail(RNLA_ERR_INVALID_DIMENSIONS, "Rows and columns must be greater than 0");
    if (ldk < d) return fail(RNLA_ERR_INVALID_DIMENSIONS, "leading dimension of the sketch smaller than d");
    if (kind == RNLA_SKETCH_DENSE) {
        if (dist < RNLA_GAUSSIAN || dist > RNLA_RADEMACHER) return fail(RNLA_ERR_INVALID_PARAMETERS, "unknown distribution");
        PhaseScope ph("sketch:dense");
        DevBuf St;
        const int64_t mm = std::max<int64_t>(m_local, 1);
        RNLA_CUDA(St.alloc((size_t)mm * d * 8));
        RNLA_CUDA(fill_philox(dist, seed, 3 /* STREAM_SKETCH_DENSE */, m_local, d, row_offset, St.d(), mm, c.stream));
        return dev_gemm_tn(St.d(), mm, m_local, d, dA, lda, n, dAsk, ldk, true);
    }
    if (kind == RNLA_SKETCH_SASO) {
        if (zeta < 1 || zeta > 8) return fail(RNLA_ERR_INVALID_PARAMETERS, "SASO: 1 <= zeta <= 8");
        if (d * 8 > SASO_SMEM) return fail(RNLA_ERR_INVALID_DIMENSIONS, "SASO: sketch dimension d must be <= 25600");
        PhaseScope ph("sketch:saso");
        const bool ar = c.nranks > 1;
        DevBuf packed;
        double* out = dAsk; int64_t ldo = ldk;
        if (ar && ldk != d) { RNLA_CUDA(packed.alloc((size_t)d * n * 8)); out = packed.d(); ldo = d; }
        RNLA_CUDA(axpby_matrix(0.0, nullptr, 0, 0.0, nullptr, 0, out, ldo, d, n, c.stream));
        RNLA_CUDA(saso_apply(seed, d, zeta, dA, lda, m_local, n, row_offset, out, ldo, c.sms, c.stream));
        if (ar) {
            RNLA_TRY(allreduce_sum_f64(out, (size_t)d * n));
            if (out != dAsk) RNLA_CUDA(copy_matrix(out, ldo, dAsk, ldk, d, n, c.stream));
        }
        return RNLA_OK;
    }
    if (kind == RNLA_SKETCH_SASO_BLOCK) {
        PhaseScope ph("sketch:saso_block");
        const bool ar = c.nranks > 1;
        DevBuf packed;
        double* out = dAsk; int64_t ldo = ldk;
        if (ar && ldk != d) { RNLA_CUDA(packed.alloc((size_t)d * n * 8)); out = packed.d(); ldo = d; }
        RNLA_TRY(saso_block_apply(seed, d, zeta, dist /* block width, 0 = default */, dA, lda, m_local, n, row_offset, out, ldo));
        if (ar) {
            RNLA_TRY(allreduce_sum_f64(out, (size_t)d * n));
            if (out != dAsk) RNLA_CUDA(copy_matrix(out, ldo, dAsk, ldk, d, n, c.stream));
        }
        return RNLA_OK;
    }
    return fail(RNLA_ERR_INVALID_PARAMETERS, "unknown sketch kind");
}

// A = U0 diag(sigma) V0^T + (eta / sqrt(m_global)) G, generated shard by shard (pure function of the global indices)
rnla_status dev_generate_lowrank(double* dA, int64_t lda, int64_t m_local, int64_t n, int64_t row_offset, int64_t m_global,
                                 int64_t r0, const double* sigma_host, double eta, uint64_t seed) {
    Ctx& c = ctx();
    if (r0 <= 0 || r0 > std::min(m_global, n)) return fail(RNLA_ERR_INVALID_DIMENSIONS, "generate_lowrank: 1 <= r0 <= min(m, n)");
    ShardInfo sh;
    RNLA_TRY(shard_layout(m_local, &sh));
    if (sh.rows_global != m_global || sh.row_off != row_offset)
        return fail(RNLA_ERR_INVALID_DIMENSIONS, "generate_lowrank: shard layout does not match the communicator");
    const int64_t mm = std::max<int64_t>(m_local, 1);
    DevBuf U0, V0, V0t, sig;
    RNLA_CUDA(U0.alloc((size_t)mm * r0 * 8)); RNLA_CUDA(V0.alloc((size_t)n * r0 * 8)); RNLA_CUDA(V0t.alloc((size_t)n * r0 * 8));
    RNLA_CUDA(sig.alloc((size_t)r0 * 8));
    RNLA_CUDA(fill_philox(DIST_GAUSSIAN, seed, 16 /* STREAM_SYNTH_U */, m_local, r0, row_offset, U0.d(), mm, c.stream));
    RNLA_CUDA(fill_philox(DIST_GAUSSIAN, seed, 17 /* STREAM_SYNTH_V */, n, r0, 0, V0.d(), n, c.stream));
    RNLA_TRY(orth_inplace(U0.d(), mm, sh, (int)r0, true, nullptr, nullptr));
    ShardInfo nside{n, 0, n};
    RNLA_TRY(orth_inplace(V0.d(), n, nside, (int)r0, false, nullptr, nullptr));
    RNLA_CUDA(cudaMemcpyAsync(sig.p, sigma_host, (size_t)r0 * 8, cudaMemcpyHostToDevice, c.stream));
    RNLA_CUDA(scale_columns(U0.d(), mm, m_local, r0, sig.d(), c.stream));
    RNLA_CUDA(transpose_matrix(V0.d(), n, V0t.d(), r0, n, r0, c.stream));
    RNLA_TRY(dev_gemm_nn(U0.d(), mm, m_local, r0, V0t.d(), r0, n, dA, lda));
    if (eta != 0.0)
        RNLA_CUDA(add_noise(dA, lda, m_local, n, row_offset, eta / std::sqrt((double)m_global), seed, 18 /* STREAM_SYNTH_NOISE */, c.stream));
    RNLA_CUDA(cudaStreamSynchronize(c.stream));
    return RNLA_OK;
}

}  // namespace rnla
