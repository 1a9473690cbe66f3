/* oracle_gemm.c -- the reference's dense products (`*` on DMatrix, i.e. nalgebra -> matrixmultiply 0.3.9 gemm;
 * call sites src/lora_helpers.rs:21,41,76,89,95, src/lora_drivers.rs:66,121,148,187,191) restated as a blocked,
 * OpenMP-parallel f64 GEMM.  TEST / CPU-BASELINE INFRASTRUCTURE ONLY (see oracle.h).
 * Built with -O3 -mavx2 -mfma (x86-64-v3: present on the build container and on the GPU box hosts). */
#include "oracle.h"
#include <omp.h>
#include <stdlib.h>
#include <string.h>

typedef double v4 __attribute__((vector_size(32), aligned(8)));
#define AT(M, ld, i, j) (M)[(int64_t)(i) + (int64_t)(j) * (int64_t)(ld)]

static int g_threads = 0;
void orc_set_threads(int n) { g_threads = n; if (n > 0) omp_set_num_threads(n); }
int orc_get_threads(void) { return g_threads > 0 ? g_threads : omp_get_max_threads(); }

static inline v4 ld4(const double* p) { v4 v; memcpy(&v, p, 32); return v; }
static inline void st4(double* p, v4 v) { memcpy(p, &v, 32); }

/* C[8 x nc] += A[8 x kc] * B[kc x nc], nc <= 6 */
static void micro_nn(const double* A, int64_t lda, const double* B, int64_t ldb, double* C, int64_t ldc, int64_t kc, int nc) {
    v4 acc[6][2];
    for (int c = 0; c < 6; ++c) { acc[c][0] = (v4){0, 0, 0, 0}; acc[c][1] = (v4){0, 0, 0, 0}; }
    if (nc == 6) {
        for (int64_t k = 0; k < kc; ++k) {
            const v4 a0 = ld4(A + k * lda), a1 = ld4(A + k * lda + 4);
            for (int c = 0; c < 6; ++c) {
                const double b = B[k + c * ldb];
                const v4 bb = {b, b, b, b};
                acc[c][0] += a0 * bb; acc[c][1] += a1 * bb;
            }
        }
    } else {
        for (int64_t k = 0; k < kc; ++k) {
            const v4 a0 = ld4(A + k * lda), a1 = ld4(A + k * lda + 4);
            for (int c = 0; c < nc; ++c) {
                const double b = B[k + c * ldb];
                const v4 bb = {b, b, b, b};
                acc[c][0] += a0 * bb; acc[c][1] += a1 * bb;
            }
        }
    }
    for (int c = 0; c < nc; ++c) {
        st4(C + c * ldc, ld4(C + c * ldc) + acc[c][0]);
        st4(C + c * ldc + 4, ld4(C + c * ldc + 4) + acc[c][1]);
    }
}

void orc_gemm_nn(const double* A, int64_t lda, int64_t m, int64_t K, const double* B, int64_t ldb, int64_t N, double* C, int64_t ldc) {
    const int64_t MB = 64, KB = 512;
    const int64_t nblk = (m + MB - 1) / MB;
#pragma omp parallel for schedule(dynamic, 4)
    for (int64_t ib = 0; ib < nblk; ++ib) {
        const int64_t i0 = ib * MB, i1 = i0 + MB < m ? i0 + MB : m;
        for (int64_t c = 0; c < N; ++c) for (int64_t i = i0; i < i1; ++i) AT(C, ldc, i, c) = 0.0;
        for (int64_t k0 = 0; k0 < K; k0 += KB) {
            const int64_t kc = k0 + KB < K ? KB : K - k0;
            int64_t i = i0;
            for (; i + 8 <= i1; i += 8)
                for (int64_t c = 0; c < N; c += 6)
                    micro_nn(&AT(A, lda, i, k0), lda, &AT(B, ldb, k0, c), ldb, &AT(C, ldc, i, c), ldc, kc, (int)(N - c < 6 ? N - c : 6));
            for (; i < i1; ++i)
                for (int64_t c = 0; c < N; ++c) {
                    double s = 0.0;
                    for (int64_t k = 0; k < kc; ++k) s += AT(A, lda, i, k0 + k) * AT(B, ldb, k0 + k, c);
                    AT(C, ldc, i, c) += s;
                }
        }
    }
}

static inline double hsum(v4 v) { return (v[0] + v[1]) + (v[2] + v[3]); }

/* Z[4 x 4] += A[rows x 4]^T Q[rows x 4] */
static void micro_tn(const double* A, int64_t lda, const double* Q, int64_t ldq, int64_t rows, double* Z, int64_t ldz, int nj, int ncq) {
    v4 acc[4][4];
    for (int a = 0; a < 4; ++a) for (int b = 0; b < 4; ++b) acc[a][b] = (v4){0, 0, 0, 0};
    int64_t i = 0;
    if (nj == 4 && ncq == 4) {
        for (; i + 4 <= rows; i += 4) {
            const v4 a0 = ld4(A + i), a1 = ld4(A + i + lda), a2 = ld4(A + i + 2 * lda), a3 = ld4(A + i + 3 * lda);
            for (int b = 0; b < 4; ++b) {
                const v4 q = ld4(Q + i + b * ldq);
                acc[0][b] += a0 * q; acc[1][b] += a1 * q; acc[2][b] += a2 * q; acc[3][b] += a3 * q;
            }
        }
    }
    for (int a = 0; a < nj; ++a)
        for (int b = 0; b < ncq; ++b) {
            double s = hsum(acc[a][b]);
            for (int64_t t = i; t < rows; ++t) s += A[t + a * lda] * Q[t + b * ldq];
            Z[a + b * ldz] += s;
        }
}

void orc_gemm_tn(const double* A, int64_t lda, int64_t m, int64_t n, const double* Q, int64_t ldq, int64_t N, double* Z, int64_t ldz) {
    const int64_t RB = 1024;
    const int64_t jgroups = (n + 3) / 4;
#pragma omp parallel
    {
        const int nt = omp_get_num_threads(), tid = omp_get_thread_num();
        const int64_t g0 = jgroups * tid / nt, g1 = jgroups * (tid + 1) / nt;
        for (int64_t g = g0; g < g1; ++g) {
            const int64_t j0 = g * 4, nj = n - j0 < 4 ? n - j0 : 4;
            for (int64_t c = 0; c < N; ++c) for (int64_t j = 0; j < nj; ++j) AT(Z, ldz, j0 + j, c) = 0.0;
        }
        for (int64_t r0 = 0; r0 < m; r0 += RB) {
            const int64_t rows = r0 + RB < m ? RB : m - r0;
            for (int64_t g = g0; g < g1; ++g) {
                const int64_t j0 = g * 4; const int nj = (int)(n - j0 < 4 ? n - j0 : 4);
                for (int64_t c = 0; c < N; c += 4)
                    micro_tn(&AT(A, lda, r0, j0), lda, &AT(Q, ldq, r0, c), ldq, rows, &AT(Z, ldz, j0, c), ldz, nj, (int)(N - c < 4 ? N - c : 4));
            }
        }
    }
}
