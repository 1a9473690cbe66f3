// Streaming FP64 GEMM kernels of the sketch-and-factor path (SURVEY.md §2.2 K1, K1', K2, K5).
// All matrices are column-major with explicit leading dimensions (nalgebra DMatrix layout).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace rnla {

// C[m x N] = A[m x K] * B[K x N].  If gen != 0, B is never read: B(k, j) = omega(k + k_off, j) is
// evaluated inside the kernel from Philox4x32-10 (rng.cuh), so the tall A is the only HBM stream.
struct GemmNN {
    const double* A; int64_t lda; int64_t m; int64_t K;
    const double* B; int64_t ldb; int64_t N;
    double* C; int64_t ldc;
    int gen; int dist; uint64_t seed; uint32_t stream; uint64_t k_off;
    // filled in by gemm_nn: split of the K loop over blockIdx.z (partials P[z], m x N packed, reduced in fixed order)
    int ksplit = 1; int kt_per = 0; double* P = nullptr; int64_t pstride = 0;
};

// Z[n x N] = A[m x n]^T * Q[m x N]  (K = m is the long, streamed dimension; split over row chunks and
// reduced in a fixed order, so results are deterministic).  workspace holds the per-chunk partials.
struct GemmTN {
    const double* A; int64_t lda; int64_t m; int64_t n;
    const double* Q; int64_t ldq; int64_t N;
    double* Z; int64_t ldz;
    int accumulate;   // Z += result instead of Z = result (used by the row-sharded reduction)
};

cudaError_t gemm_nn(const GemmNN& p, cudaStream_t st);
size_t gemm_tn_workspace_bytes(int64_t m, int64_t n, int64_t N, int sms);
cudaError_t gemm_tn(const GemmTN& p, double* workspace, size_t workspace_bytes, int sms, cudaStream_t st);

// planning decisions (host logic, exported through rnla_plan_* for CPU tests)
int gemm_nn_ksplit(int64_t m, int64_t K, int64_t N, int sms);
void gemm_tn_plan_info(int64_t m, int64_t n, int64_t N, int sms, int* chunks, int64_t* chunk_rows, int* tiles);

// launch counter (bench.py reports gpu_launches from it)
extern unsigned long long g_kernel_launches;

}  // namespace rnla
