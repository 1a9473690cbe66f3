// Small dense core and panel utilities; see panel.cuh.  These are latency-/HBM-bound helpers next to
// the GEMM passes (<3 % of a rand_svd at the headline size), written for clarity.
#include "panel.cuh"
#include "gemm.cuh"
#include "rng.cuh"
#include <algorithm>
#include <cstdlib>
#include <cfloat>
#include <cooperative_groups.h>
#include <cmath>

namespace cg = cooperative_groups;

namespace rnla {

#define LAUNCHED() (++g_kernel_launches, cudaGetLastError())

namespace {

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// ---------------------------------------------------------------- K0 fills
__global__ void __launch_bounds__(256)
fill_philox_kernel(int dist, uint64_t seed, uint32_t stream, int64_t rows, int64_t cols, int64_t row_off,
                   double* __restrict__ out, int64_t ld) {
    const uint64_t q_first = (uint64_t)row_off >> 2;
    const uint64_t q_last = (uint64_t)(row_off + rows - 1) >> 2;
    const int64_t nq = (int64_t)(q_last - q_first + 1);
    const int64_t total = nq * cols;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (int64_t)gridDim.x * blockDim.x) {
        const int64_t c = idx / nq, qi = idx - c * nq;
        const uint64_t q = q_first + (uint64_t)qi;
        const u32x4 b = omega_block(seed, stream, q, (uint32_t)c);
        const uint32_t w[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int64_t r = (int64_t)(4 * q + e) - row_off;
            if (r >= 0 && r < rows) out[r + c * ld] = sample_from_u32(dist, w[e]);
        }
    }
}

__global__ void __launch_bounds__(256)
fill_threefry_kernel(int dist, uint64_t k0, uint64_t k1, int64_t rows, int64_t cols, double* __restrict__ out, int64_t ld) {
    const int64_t total = rows * cols;
    const int64_t nblk = (total + 1) / 2;
    for (int64_t blk = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; blk < nblk;
         blk += (int64_t)gridDim.x * blockDim.x) {
        uint64_t x[2];
        threefry2x64_20((uint64_t)blk, 0ull, k0, k1, x[0], x[1]);
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            const int64_t t = 2 * blk + e;
            if (t < total) {
                const int64_t c = t / rows, r = t - c * rows;
                double v;
                if (dist == DIST_UNIFORM) {
                    // rand 0.8.5 UniformFloat<f64>::sample: [1,2) from the top 52 bits, -1, * scale(2) + low(-1), no fma
                    const double v12 = __longlong_as_double((long long)((x[e] >> 12) | 0x3FF0000000000000ull));
                    v = __dadd_rn(__dmul_rn(__dadd_rn(v12, -1.0), 2.0), -1.0);
                } else {
                    v = (x[e] < 0x8000000000000000ull) ? 1.0 : -1.0;   // Bernoulli(0.5): u64 < 2^63
                }
                out[r + c * ld] = v;
            }
        }
    }
}

__global__ void philox_blocks_kernel(int64_t n, const uint32_t* ctr, const uint32_t* key, uint32_t* out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        const u32x4 r = philox4x32_10(ctr[4 * i], ctr[4 * i + 1], ctr[4 * i + 2], ctr[4 * i + 3], key[2 * i], key[2 * i + 1]);
        out[4 * i] = r.x; out[4 * i + 1] = r.y; out[4 * i + 2] = r.z; out[4 * i + 3] = r.w;
    }
}
__global__ void threefry_blocks_kernel(int64_t n, const uint64_t* ctr, const uint64_t* key, uint64_t* out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) threefry2x64_20(ctr[2 * i], ctr[2 * i + 1], key[2 * i], key[2 * i + 1], out[2 * i], out[2 * i + 1]);
}

// ---------------------------------------------------------------- Cholesky with deficiency detection
__global__ void __launch_bounds__(1024)
chol_upper_kernel(double* __restrict__ G, int64_t ld, int p, double tol2, int* __restrict__ flags, int* __restrict__ info) {
    __shared__ double s_red[32];
    __shared__ double s_piv;
    __shared__ int s_def, s_ndef, s_bad, s_nzero;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
    if (tid == 0) { s_ndef = 0; s_bad = 0; s_nzero = 0; }
    __syncthreads();
    for (int j = 0; j < p; ++j) {
        // pivot d = G_jj - sum_{k<j} R_kj^2
        double part = 0.0;
        for (int k = tid; k < j; k += blockDim.x) { const double v = G[k + j * ld]; part += v * v; }
        part = warp_sum(part);
        if (lane == 0) s_red[warp] = part;
        __syncthreads();
        if (warp == 0) {
            double v = lane < nwarps ? s_red[lane] : 0.0;
            v = warp_sum(v);
            if (lane == 0) {
                const double gjj = G[j + j * ld];
                const double d = gjj - v;
                // 0 = regular column, 1 = numerically dependent on the previous ones (keep its residual,
                // it gets a fresh chance in the next pass), 2 = exactly zero column (caller replaces it)
                int def = 0;
                if (!(gjj > 0.0)) def = 2;
                else if (!(d > tol2 * gjj)) def = 1;
                if (!isfinite(gjj) || !isfinite(v)) { s_bad = 1; def = 2; }
                const double piv = def ? 1.0 : sqrt(d);
                G[j + j * ld] = piv;
                flags[j] = def;
                s_ndef += (def != 0); s_nzero += (def == 2);
                s_piv = piv; s_def = def;
            }
        }
        __syncthreads();
        const double piv = s_piv; const int def = s_def;
        // row j of R, one warp per column i
        for (int i = j + 1 + warp; i < p; i += nwarps) {
            double dot = 0.0;
            if (!def) for (int k = lane; k < j; k += 32) dot += G[k + j * ld] * G[k + (int64_t)i * ld];
            dot = warp_sum(dot);
            if (lane == 0) G[j + (int64_t)i * ld] = def ? 0.0 : (G[j + (int64_t)i * ld] - dot) / piv;
        }
        __syncthreads();
    }
    for (int idx = tid; idx < p * p; idx += blockDim.x) {
        const int c = idx / p, r = idx - c * p;
        if (r > c) G[r + (int64_t)c * ld] = 0.0;
    }
    if (tid == 0) { info[0] = s_ndef; info[1] = s_bad; info[2] = s_nzero; }
}

__global__ void __launch_bounds__(1024)
tri_inv_upper_kernel(const double* __restrict__ R, int64_t ldr, int p, double* __restrict__ X, int64_t ldi) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    for (int j = warp; j < p; j += nwarps) {
        double* x = X + (int64_t)j * ldi;
        for (int i = lane; i < p; i += 32) if (i > j) x[i] = 0.0;
        __syncwarp();
        for (int i = j; i >= 0; --i) {
            double s = 0.0;
            for (int k = i + 1 + lane; k <= j; k += 32) s += R[i + (int64_t)k * ldr] * x[k];
            s = warp_sum(s);
            if (lane == 0) x[i] = ((i == j ? 1.0 : 0.0) - s) / R[i + (int64_t)i * ldr];
            __syncwarp();
        }
    }
}

// Shared-memory variants of the two kernels above.  Each of the p dependent steps of the global-memory versions costs an
// L2 round trip, a block-wide reduction and an IEEE division (p = 110: 280 us per factorisation, 165 us per inverse).
//   chol_upper_smem_kernel: right-looking Cholesky on the packed upper triangle (element (k, j), k <= j, at
//     j(j+1)/2 + k; 177 KB at p = 210): per step one scalar pivot, a row scaling and a rank-1 update of the trailing
//     triangle spread over the whole CTA -- no reductions.  Same deficiency rule and flags as chol_upper_kernel.
//   tri_inv_upper_smem_kernel: R row-packed in shared memory, reciprocal diagonal precomputed, one 16-lane group per
//     column of the inverse (64 columns in flight), private p-vector per group.
__device__ __forceinline__ int pk(int k, int j) { return j * (j + 1) / 2 + k; }

__global__ void __launch_bounds__(1024)
chol_upper_smem_kernel(double* __restrict__ G, int64_t ld, int p, double tol2, int* __restrict__ flags, int* __restrict__ info) {
    extern __shared__ double gs[];
    double* diag0 = gs + (size_t)p * (p + 1) / 2;       // original diagonal
    double* rowj = diag0 + p;                           // scaled row j, contiguous
    __shared__ double s_inv;
    __shared__ int s_def, s_ndef, s_bad, s_nzero;
    const int tid = threadIdx.x, tx = tid & 31, ty = tid >> 5;
    for (int idx = tid; idx < p * p; idx += blockDim.x) {
        const int c = idx / p, r = idx - c * p;
        if (r <= c) { const double v = G[r + (int64_t)c * ld]; gs[pk(r, c)] = v; if (r == c) diag0[c] = v; }
    }
    if (tid == 0) { s_ndef = 0; s_bad = 0; s_nzero = 0; }
    __syncthreads();
    for (int j = 0; j < p; ++j) {
        if (tid == 0) {
            const double gjj = diag0[j];
            const double d = gs[pk(j, j)];              // G_jj - sum_{k<j} R_kj^2 after the previous rank-1 updates
            // 0 = regular column, 1 = numerically dependent on the previous ones (keeps its residual, re-examined in the
            // next pass), 2 = exactly zero column (the caller replaces it)
            int def = 0;
            if (!(gjj > 0.0)) def = 2;
            else if (!(d > tol2 * gjj)) def = 1;
            if (!isfinite(gjj) || !isfinite(d)) { s_bad = 1; def = 2; }
            const double piv = def ? 1.0 : sqrt(d);
            gs[pk(j, j)] = piv;
            flags[j] = def;
            s_ndef += (def != 0); s_nzero += (def == 2);
            s_inv = def ? 0.0 : 1.0 / piv; s_def = def;
        }
        __syncthreads();
        const double inv = s_inv; const int def = s_def;
        for (int i = j + 1 + tid; i < p; i += blockDim.x) {
            const double v = def ? 0.0 : gs[pk(j, i)] * inv;
            gs[pk(j, i)] = v; rowj[i] = v;
        }
        __syncthreads();
        if (!def) {
            for (int i = j + 1 + ty; i < p; i += 32) {
                const double rji = rowj[i];
                double* gi = gs + pk(0, i);
                for (int k = j + 1 + tx; k <= i; k += 32) gi[k] = fma(-rowj[k], rji, gi[k]);
            }
        }
        __syncthreads();
    }
    for (int idx = tid; idx < p * p; idx += blockDim.x) {
        const int c = idx / p, r = idx - c * p;
        G[r + (int64_t)c * ld] = r <= c ? gs[pk(r, c)] : 0.0;
    }
    if (tid == 0) { info[0] = s_ndef; info[1] = s_bad; info[2] = s_nzero; }
}

__global__ void __launch_bounds__(1024)
tri_inv_upper_smem_kernel(const double* __restrict__ R, int64_t ldr, int p, double* __restrict__ X, int64_t ldi) {
    extern __shared__ double gs[];
    const int tid = threadIdx.x, l16 = tid & 15, grp = tid >> 4, ngrp = blockDim.x >> 4;
    double* rs = gs;                                    // row-packed R: (i, k), k >= i, at i p - i(i-1)/2 + k - i
    double* rinv = gs + (size_t)p * (p + 1) / 2;        // 1 / R_ii
    double* xs = rinv + p + (size_t)grp * p;            // this group's column of the inverse
    for (int idx = tid; idx < p * p; idx += blockDim.x) {
        const int c = idx / p, r = idx - c * p;
        if (r <= c) {
            const double v = R[r + (int64_t)c * ldr];
            rs[r * p - r * (r - 1) / 2 + c - r] = v;
            if (r == c) rinv[r] = 1.0 / v;
        }
    }
    __syncthreads();
    const unsigned gmask = 0xffffu << (16 * ((tid >> 4) & 1));     // the two groups of a warp run loops of different length
    for (int j = grp; j < p; j += ngrp) {
        for (int i = j; i >= 0; --i) {
            const double* ri = rs + (i * p - i * (i - 1) / 2 - i);  // ri[k] = R(i, k)
            double s = 0.0;
            for (int k = i + 1 + l16; k <= j; k += 16) s = fma(ri[k], xs[k], s);
#pragma unroll
            for (int o = 8; o > 0; o >>= 1) s += __shfl_xor_sync(gmask, s, o);
            if (l16 == 0) xs[i] = ((i == j ? 1.0 : 0.0) - s) * rinv[i];
            __syncwarp(gmask);
        }
        double* x = X + (int64_t)j * ldi;
        for (int i = l16; i < p; i += 16) x[i] = i <= j ? xs[i] : 0.0;
        __syncwarp(gmask);
    }
}

// Blocked variants (round 2): the kernels above pay three CTA-wide barriers, a square root and a division per COLUMN (p = 128: 140 us
// per factorisation, 112 us per inverse -- 80 such launches are 10 ms of blendenpik's preconditioner).  Here the p dependent steps
// become p / 16 (p / 32) block steps.
//   chol_upper_blocked_kernel: right-looking over column blocks of 16 on the packed upper triangle: warp 0 factors the 16 x 16
//     diagonal block (same pivot / deficiency rule per column, in the same order), one thread per trailing column solves the 16-row
//     panel by forward substitution in registers, and the whole CTA applies the rank-16 update.  A deficient column has pivot 1, a
//     zero row and contributes nothing, exactly as in the unblocked kernels.
//   tri_inv_upper_blocked_kernel (p <= 144): blocks of 32; every diagonal block is inverted by one warp (lane = column of the
//     inverse, private back substitution on its own column in shared memory, reciprocal diagonal precomputed), then block back substitution
//     X_IJ = - X_II sum_{I < K <= J} R_IK X_KJ for I = J - 1, J - 2, ... with the whole CTA on every block (lane = row: the loads of R
//     and X_II are consecutive, those of X_KJ and of the intermediate product broadcasts).
constexpr int CHB = 16;
__global__ void __launch_bounds__(512)
chol_upper_blocked_kernel(double* __restrict__ G, int64_t ld, int p, double tol2, int* __restrict__ flags, int* __restrict__ info) {
    extern __shared__ double gs[];
    double* diag0 = gs + (size_t)p * (p + 1) / 2;       // original diagonal
    __shared__ double b_inv[CHB];
    __shared__ int s_ndef, s_bad, s_nzero;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int idx = tid; idx < p * p; idx += blockDim.x) {
        const int c = idx / p, r = idx - c * p;
        if (r <= c) { const double v = G[r + (int64_t)c * ld]; gs[pk(r, c)] = v; if (r == c) diag0[c] = v; }
    }
    if (tid == 0) { s_ndef = 0; s_bad = 0; s_nzero = 0; }
    __syncthreads();
    for (int j0 = 0; j0 < p; j0 += CHB) {
        const int nb = min(CHB, p - j0), j1 = j0 + nb;
        if (warp == 0) {
            for (int jj = 0; jj < nb; ++jj) {
                const int j = j0 + jj;
                const double gjj = diag0[j];
                const double d = gs[pk(j, j)];          // G_jj - sum_{k<j} R_kj^2 after the updates so far
                int def = 0;
                if (!(gjj > 0.0)) def = 2;
                else if (!(d > tol2 * gjj)) def = 1;
                const bool bad = !isfinite(gjj) || !isfinite(d);
                if (bad) def = 2;
                const double piv = def ? 1.0 : sqrt(d), inv = def ? 0.0 : 1.0 / piv;
                __syncwarp();                           // every lane has read d
                if (lane == 0) {
                    gs[pk(j, j)] = piv; flags[j] = def; b_inv[jj] = inv;
                    if (bad) s_bad = 1;
                    s_ndef += (def != 0); s_nzero += (def == 2);
                }
                const int i = j + 1 + lane;             // scale row j inside the block
                if (i < j1) gs[pk(j, i)] = def ? 0.0 : gs[pk(j, i)] * inv;
                __syncwarp();
                if (!def) {
                    // rank-1 update of the rest of the block, its 16 x 16 index square spread over the lanes
#pragma unroll
                    for (int e = lane; e < CHB * CHB; e += 32) {
                        const int k = j0 + (e & (CHB - 1)), c = j0 + (e >> 4);
                        if (k > j && k <= c && c < j1) gs[pk(k, c)] = fma(-gs[pk(j, k)], gs[pk(j, c)], gs[pk(k, c)]);
                    }
                }
                __syncwarp();
            }
        }
        __syncthreads();
        // the 16-row panel right of the block: R[j0.., i] = D^-T G[j0.., i], one thread per column i
        for (int i = j1 + tid; i < p; i += blockDim.x) {
            double* gi = gs + pk(0, i);
            double g[CHB];
#pragma unroll
            for (int kk = 0; kk < CHB; ++kk) g[kk] = kk < nb ? gi[j0 + kk] : 0.0;
#pragma unroll
            for (int kk = 0; kk < CHB; ++kk) {
                if (kk < nb) {
                    const double* rk = gs + pk(0, j0 + kk);
                    double sum = g[kk];
#pragma unroll
                    for (int l = 0; l < kk; ++l) sum = fma(-rk[j0 + l], g[l], sum);
                    g[kk] = sum * b_inv[kk];            // a deficient column: inverse pivot 0, zero row
                    gi[j0 + kk] = g[kk];
                }
            }
        }
        __syncthreads();
        // rank-nb update of the trailing triangle
        for (int i = j1 + warp; i < p; i += (int)(blockDim.x >> 5)) {
            double ri[CHB];
            double* gi = gs + pk(0, i);
#pragma unroll
            for (int kk = 0; kk < CHB; ++kk) ri[kk] = kk < nb ? gi[j0 + kk] : 0.0;
            for (int k = j1 + lane; k <= i; k += 32) {
                const double* gk = gs + pk(0, k);
                double sum = gi[k];
#pragma unroll
                for (int kk = 0; kk < CHB; ++kk) sum = fma(-gk[min(j0 + kk, j1 - 1)], ri[kk], sum);
                gi[k] = sum;
            }
        }
        __syncthreads();
    }
    for (int idx = tid; idx < p * p; idx += blockDim.x) {
        const int c = idx / p, r = idx - c * p;
        G[r + (int64_t)c * ld] = r <= c ? gs[pk(r, c)] : 0.0;
    }
    if (tid == 0) { info[0] = s_ndef; info[1] = s_bad; info[2] = s_nzero; }
}

constexpr int TIB = 32;
__global__ void __launch_bounds__(1024)
tri_inv_upper_blocked_kernel(const double* __restrict__ R, int64_t ldr, int p, double* __restrict__ X, int64_t ldi) {
    extern __shared__ double gs[];
    const size_t tri = (size_t)p * (p + 1) / 2;
    double* rs = gs;                                    // R, packed by columns: (k, j), k <= j, at j (j + 1) / 2 + k
    double* xs = gs + tri;                              // the inverse, same packing
    double* dinv = xs + tri;                            // 1 / R_ii
    double* tb = dinv + p;                              // 32 x 32 intermediate block, [c][r]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nbk = (p + TIB - 1) / TIB;
    for (int idx = tid; idx < p * p; idx += blockDim.x) {
        const int c = idx / p, r = idx - c * p;
        if (r <= c) { const double v = R[r + (int64_t)c * ldr]; rs[pk(r, c)] = v; if (r == c) dinv[r] = 1.0 / v; }
    }
    __syncthreads();
    // diagonal blocks: warp b inverts block b, lane = column of the inverse
    if (warp < nbk) {
        const int c0 = warp * TIB, j = c0 + lane;
        if (j < p) {
            double* xc = xs + pk(0, j);                 // this lane's column of the inverse, rows c0 .. j
            for (int ii = lane; ii >= 0; --ii) {
                double sum = (ii == lane) ? 1.0 : 0.0;
                for (int kk = ii + 1; kk <= lane; ++kk) sum = fma(-rs[pk(c0 + ii, c0 + kk)], xc[c0 + kk], sum);
                xc[c0 + ii] = sum * dinv[c0 + ii];
            }
        }
    }
    __syncthreads();
    // block back substitution, distance s from the diagonal: every thread one element (r = lane, c = warp) of every block (I, I + s)
    for (int sdist = 1; sdist < nbk; ++sdist) {
        for (int J = sdist; J < nbk; ++J) {
            const int I = J - sdist;
            const int row = I * TIB + lane, col = J * TIB + warp;
            double t = 0.0;
            if (col < p) {
                const double* xc = xs + pk(0, col);
                for (int k = (I + 1) * TIB; k <= min(col, (J + 1) * TIB - 1); ++k) t = fma(rs[pk(row, k)], xc[k], t);     // row < k: inside the triangle
            }
            tb[warp * TIB + lane] = t;
            __syncthreads();
            if (col < p) {
                double v = 0.0;
                for (int kk = lane; kk < TIB; ++kk) v = fma(xs[pk(row, I * TIB + kk)], tb[warp * TIB + kk], v);
                xs[pk(row, col)] = -v;
            }
            __syncthreads();
        }
    }
    for (int idx = tid; idx < p * p; idx += blockDim.x) {
        const int c = idx / p, r = idx - c * p;
        X[r + (int64_t)c * ldi] = r <= c ? xs[pk(r, c)] : 0.0;
    }
}

// The same inverse IN PLACE on one packed triangle, for 144 < p <= 210 (two triangles no longer fit in shared memory): diagonal blocks
// one after the other through a 32 x 32 scratch block, then block columns from the LAST to the first and, inside a block column, block
// rows upwards: X_IJ = - X_II sum_{I < K <= J} R_IK X_KJ overwrites R_IJ when nothing needs it any more (the blocks R_IK, K < J, belong
// to block columns that come later; X_KJ, K > I, were finished just before).
__global__ void __launch_bounds__(1024)
tri_inv_upper_inplace_kernel(const double* __restrict__ R, int64_t ldr, int p, double* __restrict__ X, int64_t ldi) {
    extern __shared__ double gs[];
    double* rs = gs;                                    // R on entry, its inverse on exit; packed by columns
    double* dinv = gs + (size_t)p * (p + 1) / 2;
    double* tb = dinv + p;                              // 32 x 32 scratch, [c][r]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nbk = (p + TIB - 1) / TIB;
    for (int idx = tid; idx < p * p; idx += blockDim.x) {
        const int c = idx / p, r = idx - c * p;
        if (r <= c) { const double v = R[r + (int64_t)c * ldr]; rs[pk(r, c)] = v; if (r == c) dinv[r] = 1.0 / v; }
    }
    __syncthreads();
    for (int b = 0; b < nbk; ++b) {
        const int c0 = b * TIB;
        if (warp == 0) {
            const int j = c0 + lane;
            double* xc = tb + lane * TIB;               // this lane's column of the block's inverse
            if (j < p) {
                for (int ii = lane; ii >= 0; --ii) {
                    double sum = (ii == lane) ? 1.0 : 0.0;
                    for (int kk = ii + 1; kk <= lane; ++kk) sum = fma(-rs[pk(c0 + ii, c0 + kk)], xc[kk], sum);
                    xc[ii] = sum * dinv[c0 + ii];
                }
            }
            __syncwarp();
            if (j < p) for (int ii = 0; ii <= lane; ++ii) rs[pk(c0 + ii, j)] = xc[ii];
        }
        __syncthreads();
    }
    for (int J = nbk - 1; J >= 1; --J) {
        for (int I = J - 1; I >= 0; --I) {
            const int row = I * TIB + lane, col = J * TIB + warp;
            double t = 0.0;
            if (col < p) {
                const double* xc = rs + pk(0, col);
                for (int k = (I + 1) * TIB; k <= min(col, (J + 1) * TIB - 1); ++k) t = fma(rs[pk(row, k)], xc[k], t);
            }
            tb[warp * TIB + lane] = t;
            __syncthreads();
            if (col < p) {
                double v = 0.0;
                for (int kk = lane; kk < TIB; ++kk) v = fma(rs[pk(row, I * TIB + kk)], tb[warp * TIB + kk], v);
                rs[pk(row, col)] = -v;
            }
            __syncthreads();
        }
    }
    for (int idx = tid; idx < p * p; idx += blockDim.x) {
        const int c = idx / p, r = idx - c * p;
        X[r + (int64_t)c * ldi] = r <= c ? rs[pk(r, c)] : 0.0;
    }
}

__global__ void __launch_bounds__(256)
small_gemm_kernel(const double* __restrict__ A, int64_t lda, const double* __restrict__ B, int64_t ldb,
                  double* __restrict__ C, int64_t ldc, int M, int N, int K) {
    const int64_t total = (int64_t)M * N;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(idx / M), r = (int)(idx - (int64_t)c * M);
        double s = 0.0;
        for (int k = 0; k < K; ++k) s += A[r + (int64_t)k * lda] * B[k + (int64_t)c * ldb];
        C[r + (int64_t)c * ldc] = s;
    }
}

__global__ void zero_flagged_diag_kernel(double* R, int64_t ld, int p, const int* flags) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < p && flags[j] == 2) R[j + (int64_t)j * ld] = 0.0;
}

__global__ void __launch_bounds__(256)
replace_columns_kernel(double* __restrict__ X, int64_t ld, int64_t rows, int64_t row_off, int p,
                       const int* __restrict__ flags, const int64_t* __restrict__ target) {
    const int j = blockIdx.y;
    if (j >= p || flags[j] != 2) return;
    const int64_t tgt = target[j] - row_off;
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < rows; r += (int64_t)gridDim.x * blockDim.x)
        X[r + (int64_t)j * ld] = (r == tgt) ? 1.0 : 0.0;
}

// Device-driven version of (replace_columns + zero_flagged_diag) for orth_inplace's passes that run without a host round
// trip: exactly-zero columns (flag 2) become unit vectors e_t, t = j on the first attempt and a hashed global row later
// (same rule as the host-side code in drivers.cu), and their R_jj is set to 0.  state[0] = attempt counter.
__global__ void __launch_bounds__(256)
orth_fixup_kernel(double* __restrict__ X, int64_t ld, int64_t rows, int64_t row_off, int64_t rows_global, int p,
                  const int* __restrict__ flags, const int* __restrict__ info, const int* __restrict__ state,
                  double* __restrict__ G, int64_t ldg) {
    if (info[2] == 0) return;
    const int attempt = state[0];
    for (int j = 0; j < p; ++j) {
        if (flags[j] != 2) continue;
        const int64_t tgt_g = attempt == 0 ? (int64_t)j
                                           : (int64_t)(((uint64_t)j * 7919u + (uint64_t)attempt * 104729u + 13u) % (uint64_t)rows_global);
        const int64_t tgt = tgt_g - row_off;
        for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < rows; r += (int64_t)gridDim.x * blockDim.x)
            X[r + (int64_t)j * ld] = (r == tgt) ? 1.0 : 0.0;
        if (blockIdx.x == 0 && threadIdx.x == 0) G[j + (int64_t)j * ldg] = 0.0;
    }
}
// hist[3*pass .. 3*pass+2] = info (deficient, non-finite, exactly-zero); attempt counter advances when a column was replaced
__global__ void orth_advance_kernel(const int* __restrict__ info, int* __restrict__ state, int* __restrict__ hist, int pass) {
    hist[3 * pass] = info[0]; hist[3 * pass + 1] = info[1]; hist[3 * pass + 2] = info[2];
    if (info[2] > 0) state[0] += 1;
}

__global__ void __launch_bounds__(256)
set_identity_kernel(double* X, int64_t ld, int64_t rows, int64_t cols) {
    const int64_t total = rows * cols;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
        const int64_t c = idx / rows, r = idx - c * rows;
        X[r + c * ld] = (r == c) ? 1.0 : 0.0;
    }
}

__global__ void __launch_bounds__(256)
axpby_kernel(double a, const double* __restrict__ x, int64_t ldx, double b, const double* __restrict__ y, int64_t ldy,
             double* __restrict__ dst, int64_t ldd, int64_t rows, int64_t cols) {
    const int64_t total = rows * cols;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
        const int64_t c = idx / rows, r = idx - c * rows;
        double v = 0.0;
        if (x) v = a * x[r + c * ldx];
        if (y) v += b * y[r + c * ldy];
        dst[r + c * ldd] = v;
    }
}

__global__ void __launch_bounds__(256)
transpose_kernel(const double* __restrict__ src, int64_t lds, double* __restrict__ dst, int64_t ldd, int64_t rows, int64_t cols) {
    __shared__ double tile[32][33];
    const int64_t r0 = (int64_t)blockIdx.x * 32, c0 = (int64_t)blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
    for (int k = ty; k < 32; k += 8) {
        const int64_t r = r0 + tx, c = c0 + k;
        tile[k][tx] = (r < rows && c < cols) ? src[r + c * lds] : 0.0;
    }
    __syncthreads();
    for (int k = ty; k < 32; k += 8) {
        const int64_t c = c0 + tx, r = r0 + k;       // dst(c, r) = src(r, c)
        if (r < rows && c < cols) dst[c + r * ldd] = tile[tx][k];
    }
}

__global__ void __launch_bounds__(256)
sumsq_partial_kernel(const double* __restrict__ X, int64_t ld, int64_t rows, int64_t cols, double* __restrict__ scratch) {
    __shared__ double s_red[8];
    const int64_t total = rows * cols;
    double s = 0.0;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
        const int64_t c = idx / rows, r = idx - c * rows;
        const double v = X[r + c * ld];
        s += v * v;
    }
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < 8; ++w) t += s_red[w];
        scratch[blockIdx.x] = t;
    }
}
__global__ void sumsq_final_kernel(const double* scratch, int n, double* out) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        double t = 0.0;
        for (int i = 0; i < n; ++i) t += scratch[i];
        out[0] += t;
    }
}

__global__ void __launch_bounds__(256)
check_symmetric_kernel(const double* __restrict__ A, int64_t lda, int64_t n, int* flag) {
    const int64_t total = n * n;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
        const int64_t c = idx / n, r = idx - c * n;
        if (r > c && A[r + c * lda] != A[c + r * lda]) *flag = 1;
    }
}

__global__ void __launch_bounds__(256)
scale_columns_kernel(double* X, int64_t ld, int64_t rows, int64_t cols, const double* s) {
    const int64_t total = rows * cols;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
        const int64_t c = idx / rows, r = idx - c * rows;
        X[r + c * ld] *= s[c];
    }
}

// ---------------------------------------------------------------- Jacobi kernels (single CTA)
__device__ __forceinline__ void rr_pair(int round, int idx, int P, int& a, int& b) {
    // round-robin tournament on P (even) players
    if (idx == 0) { a = P - 1; b = round; }
    else { a = (round + idx) % (P - 1); b = (round - idx + (P - 1)) % (P - 1); }
    if (a > b) { const int t = a; a = b; b = t; }
}

constexpr int JACOBI_MAX_SWEEPS = 40;

// Reciprocal square root and square root for the rotation parameters: the MUFU.RSQ64H seed
// (~20 bits, full double exponent range) plus two Newton steps, a few ulp -- the IEEE division and square root of the
// math library are ~10x longer dependent chains, and the rotation parameters sit on the critical path of every round.
__device__ __forceinline__ double fast_rsqrt(double x) {     // x > 0, normal
    double r;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    const double h = 0.5 * x;
    r = r * fma(-h * r, r, 1.5);
    r = r * fma(-h * r, r, 1.5);
    return r;
}
__device__ __forceinline__ double fast_sqrt(double x) {      // x >= 0
    if (!(x > DBL_MIN)) return sqrt(x);
    const double r = fast_rsqrt(x);
    double s = x * r;
    s = fma(fma(-s, s, x), 0.5 * r, s);                      // one correction step on the root itself
    return s;
}

// One-sided (Hestenes) Jacobi, blocked for shared memory.  W (p x p, ld p) starts as M and V as I, both in global
// `work`; the columns are cut into blocks of b columns and one CTA orthogonalises the union of two blocks (2b columns of
// W and of V, 32 b p bytes) entirely in shared memory with a cyclic round-robin sweep, one warp per column pair.  Block
// pairs of one round-robin round are disjoint and run on different CTAs; rounds are separate launches.  When the whole
// matrix fits (2b >= p, e.g. p = 110 at the headline size) a single CTA runs every sweep without leaving shared memory.
// flags[0] = some rotation happened in this sweep; single mode also sets flags[1] = sweeps, flags[2] = converged.
__global__ void __launch_bounds__(256)
jacobi_init_kernel(const double* __restrict__ M, int64_t ldm, int p, int transpose, double* __restrict__ W,
                   double* __restrict__ V, int* flags) {
    if (blockIdx.x == 0)
        for (int i = threadIdx.x; i < 4 + JACOBI_MAX_SWEEPS; i += blockDim.x) flags[i] = 0;
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < p * p; idx += gridDim.x * blockDim.x) {
        const int c = idx / p, r = idx - c * p;
        W[idx] = transpose ? M[c + (int64_t)r * ldm] : M[r + (int64_t)c * ldm];
        V[idx] = (r == c) ? 1.0 : 0.0;
    }
}

// One block pair: columns [i0, i1) and [j0, j1) of W and V go to shared memory, up to max_inner cyclic sweeps over their
// union, and back.  Returns (to every thread) whether the first sweep rotated anything; *sweeps_out / *conv_out as named.
template <int LG>   // lanes per column pair: 16 (up to 64 pairs per round in one pass) or 32 (up to 32 pairs)
__device__ int jacobi_pair_block(double* __restrict__ W, double* __restrict__ V, int p, int i0, int i1, int j0, int j1,
                                 int max_inner, int cross_only, double* jsm, int* s_rot, int* sweeps_out, int* conv_out) {
    const int tid = threadIdx.x;
    const int nI = i1 - i0, nc = nI + (j1 - j0);
    double* sW = jsm;
    double* sV = jsm + (size_t)nc * p;
    for (int idx = tid; idx < nc * p; idx += blockDim.x) {
        const int c = idx / p, r = idx - c * p;
        const int gc = c < nI ? i0 + c : j0 + (c - nI);
        sW[idx] = W[(size_t)gc * p + r];
        sV[idx] = V[(size_t)gc * p + r];
    }
    __syncthreads();
    const int NC = (nc & 1) ? nc + 1 : nc;
    // cross_only: only pairs (a in I, b in J), a bipartite cyclic schedule of max(nI, nJ) rounds -- the pairs inside a block
    // are visited once per sweep elsewhere (outer round 0), not again in every block pair that contains the block
    const int nJ = nc - nI, mx = nI > nJ ? nI : nJ;
    const int nrounds = cross_only ? mx : NC - 1, npairs = cross_only ? mx : NC / 2;
    const double tol = sqrt((double)(p > 1 ? p : 1)) * DBL_EPSILON;
    int sweeps = 0, conv = NC < 2 ? 1 : 0, first = 0;
    // one LG-lane group per column pair; a round is a single pass of the groups when NC / 2 <= 1024 / LG
    const int l16 = tid & (LG - 1), grp = tid / LG, ngrp = blockDim.x / LG;
    for (int sw = 0; sw < max_inner && NC >= 2; ++sw) {
        if (tid == 0) *s_rot = 0;
        __syncthreads();
        for (int rd = 0; rd < nrounds; ++rd) {
            for (int pi0 = 0; pi0 < npairs; pi0 += ngrp) {          // uniform trip count: the shuffles below need whole warps
                const int pi = pi0 + grp;
                int ca = 0, cb = 0;
                bool valid = pi < npairs;
                if (valid) {
                    if (cross_only) { int jj = pi + rd; if (jj >= mx) jj -= mx; ca = pi; cb = nI + jj; valid = pi < nI && jj < nJ; }
                    else { rr_pair(rd, pi, NC, ca, cb); valid = cb < nc; }
                }
                if (!valid) { ca = 0; cb = 0; }
                double* wa = sW + (size_t)ca * p; double* wb = sW + (size_t)cb * p;
                double al = 0.0, be = 0.0, ga = 0.0;
                if (valid) {
#pragma unroll 4
                    for (int k = l16; k < p; k += LG) { const double x = wa[k], y = wb[k]; al = fma(x, x, al); be = fma(y, y, be); ga = fma(x, y, ga); }
                }
#pragma unroll
                for (int o = LG / 2; o > 0; o >>= 1) {
                    al += __shfl_xor_sync(0xffffffffu, al, o);
                    be += __shfl_xor_sync(0xffffffffu, be, o);
                    ga += __shfl_xor_sync(0xffffffffu, ga, o);
                }
                // Rotation parameters, computed unconditionally so that this chain overlaps the threshold chain.  With
                // a = be - al, b = 2 ga (both scaled by one power of two against under/overflow of the squares):
                // tan(2 theta) = b / a, |theta| <= pi/4;  r = 1/hypot(a, b);  cos(2 theta) = |a| r;
                // u = (1 + cos 2theta) / 2 = cos^2(theta);  cs = u rsqrt(u);  sn = sign(ab) |b| r rsqrt(u) / 2.
                // Two reciprocal square roots (MUFU seed + two Newton steps each) instead of two divisions and three
                // square roots: the sequence of dependent FP64 operations is the latency of a round.
                const double aga = fabs(ga);
                const double thr = tol * (fast_sqrt(al) * fast_sqrt(be));
                const double a0 = be - al, b0 = 2.0 * ga;
                const int ex = (__double2hiint(fmax(fabs(a0), fabs(b0))) >> 20) & 0x7ff;
                const double sc = __hiloint2double((2046 - ex) << 20, 0);
                const double a1 = a0 * sc, b1 = b0 * sc;
                const double r = fast_rsqrt(fma(a1, a1, b1 * b1));
                const double u = fma(0.5 * fabs(a1), r, 0.5);
                const double ru = fast_rsqrt(u);
                const double cs = u * ru;
                const double sn = copysign(0.5 * fabs(b1) * r * ru, a1 * b1);
                if (valid && aga > DBL_MIN && aga > thr) {
                    double* va = sV + (size_t)ca * p; double* vb = sV + (size_t)cb * p;
#pragma unroll 4
                    for (int k = l16; k < p; k += LG) {
                        const double x = wa[k], y = wb[k];
                        wa[k] = cs * x - sn * y; wb[k] = sn * x + cs * y;
                        const double vx = va[k], vy = vb[k];
                        va[k] = cs * vx - sn * vy; vb[k] = sn * vx + cs * vy;
                    }
                    if (l16 == 0) *s_rot = 1;
                }
            }
            __syncthreads();
        }
        const int rot = *s_rot;
        if (sw == 0) first = rot;
        __syncthreads();
        sweeps = sw + 1;
        if (!rot) { conv = 1; break; }
    }
    for (int idx = tid; idx < nc * p; idx += blockDim.x) {
        const int c = idx / p, r = idx - c * p;
        const int gc = c < nI ? i0 + c : j0 + (c - nI);
        W[(size_t)gc * p + r] = sW[idx];
        V[(size_t)gc * p + r] = sV[idx];
    }
    *sweeps_out = sweeps; *conv_out = conv;
    return first;
}

// whole matrix in one CTA (p <= 72): flags[1] = sweeps, flags[2] = converged
__global__ void __launch_bounds__(1024)
jacobi_single_kernel(double* __restrict__ W, double* __restrict__ V, int p, int max_sweeps, int* __restrict__ flags) {
    extern __shared__ double jsm[];
    __shared__ int s_rot;
    const int b = (p + 1) / 2;
    int sweeps, conv;
    if (p <= 64) jacobi_pair_block<32>(W, V, p, 0, b, b, p, max_sweeps, 0, jsm, &s_rot, &sweeps, &conv);
    else jacobi_pair_block<16>(W, V, p, 0, b, b, p, max_sweeps, 0, jsm, &s_rot, &sweeps, &conv);
    if (threadIdx.x == 0) { flags[1] = sweeps; flags[2] = conv; }
}

// Cooperative launch, NB/2 CTAs (all co-resident): every sweep is NB-1 rounds of disjoint block pairs with a grid barrier
// between rounds; the sweep loop and its convergence test run on the device (flags[4 + sweep] collects "some rotation").
__global__ void __launch_bounds__(1024)
jacobi_coop_kernel(double* __restrict__ W, double* __restrict__ V, int p, int b, int nblocks, int NB, int max_sweeps,
                   int* __restrict__ flags) {
    extern __shared__ double jsm[];
    __shared__ int s_rot;
    cg::grid_group grid = cg::this_grid();
    int sweeps = 0, conv = 0;
    for (int sw = 0; sw < max_sweeps; ++sw) {
        for (int round = 0; round < NB - 1; ++round) {
            // outer round 0 pairs every block once: full tournament over the union (covers the pairs inside both blocks);
            // the other rounds only take the cross pairs.  A block paired with the padding block is swept alone in round 0.
            int I, J; rr_pair(round, blockIdx.x, NB, I, J);
            const bool pad = J >= nblocks;
            if (!pad || round == 0) {
                int isw, icv;
                const int i0 = I * b, i1 = min(p, I * b + b), j0 = pad ? i1 : J * b, j1 = pad ? i1 : min(p, J * b + b);
                const int rot = b <= 32 ? jacobi_pair_block<32>(W, V, p, i0, i1, j0, j1, 1, round != 0, jsm, &s_rot, &isw, &icv)
                                        : jacobi_pair_block<16>(W, V, p, i0, i1, j0, j1, 1, round != 0, jsm, &s_rot, &isw, &icv);
                if (rot && threadIdx.x == 0) atomicOr(&flags[4 + sw], 1);
            }
            __threadfence();
            grid.sync();
        }
        sweeps = sw + 1;
        if (*(volatile int*)&flags[4 + sw] == 0) { conv = 1; break; }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) { flags[1] = sweeps; flags[2] = conv; }
}

// W = C + 1.01 ||C||_F I, V = I, flags cleared (single CTA; p <= 1024)
__global__ void __launch_bounds__(1024)
eigh_shift_kernel(const double* __restrict__ C, int64_t ldc, int p, double* __restrict__ W, double* __restrict__ V, int* flags) {
    __shared__ double red[32];
    __shared__ double s_shift;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    double s = 0.0;
    for (int idx = tid; idx < p * p; idx += blockDim.x) { const double v = C[(idx % p) + (int64_t)(idx / p) * ldc]; s = fma(v, v, s); }
    s = warp_sum(s);
    if (lane == 0) red[warp] = s;
    __syncthreads();
    if (warp == 0) {
        double t = lane < (blockDim.x >> 5) ? red[lane] : 0.0;
        t = warp_sum(t);
        if (lane == 0) s_shift = 1.01 * sqrt(t);
    }
    __syncthreads();
    const double shift = s_shift;
    for (int idx = tid; idx < p * p; idx += blockDim.x) {
        const int c = idx / p, r = idx - c * p;
        // symmetrised read: the two triangles of C may differ in the last bit when C came out of a GEMM
        const double v = 0.5 * (C[r + (int64_t)c * ldc] + C[c + (int64_t)r * ldc]);
        W[idx] = v + (r == c ? shift : 0.0);
        V[idx] = (r == c) ? 1.0 : 0.0;
    }
    for (int i = tid; i < 4 + JACOBI_MAX_SWEEPS; i += blockDim.x) flags[i] = 0;
}

// lambda_j = v_j^T C v_j, ordering (0: descending by value, 1: descending by |value|), W_out = sorted eigenvectors
__global__ void __launch_bounds__(1024)
eigh_finish_kernel(const double* __restrict__ C, int64_t ldc, int p, const double* __restrict__ V, double* __restrict__ Wout,
                   int64_t ldw, double* __restrict__ lambda, int order, double* __restrict__ tmp /* p */, const int* __restrict__ flags,
                   int* info) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
    for (int j = warp; j < p; j += nwarps) {
        const double* v = V + (size_t)j * p;
        double num = 0.0, den = 0.0;
        for (int k = lane; k < p; k += 32) {
            double t = 0.0;                                     // (C v)_k with the symmetrised C
            for (int i = 0; i < p; ++i) t = fma(0.5 * (C[k + (int64_t)i * ldc] + C[i + (int64_t)k * ldc]), v[i], t);
            num = fma(v[k], t, num); den = fma(v[k], v[k], den);
        }
        num = warp_sum(num); den = warp_sum(den);
        if (lane == 0) tmp[j] = den > 0.0 ? num / den : 0.0;
    }
    __syncthreads();
    for (int j = warp; j < p; j += nwarps) {
        const double lj = tmp[j];
        const double kj = order ? fabs(lj) : lj;
        int rank = 0;
        for (int i = lane; i < p; i += 32) { const double li = tmp[i]; const double ki = order ? fabs(li) : li; rank += (ki > kj) || (ki == kj && i < j); }
        rank = (int)(warp_sum((double)rank) + 0.5);
        for (int k = lane; k < p; k += 32) Wout[k + (int64_t)rank * ldw] = V[(size_t)j * p + k];
        if (lane == 0) lambda[rank] = lj;
    }
    if (tid == 0) { info[0] = flags[1]; info[1] = flags[2] ? 0 : 1; }
}

// sigma, ordering, U = W / sigma, orthonormal completion; W, V as left by the sweeps.  flags -> info
__global__ void __launch_bounds__(1024)
jacobi_svd_finish_kernel(int p, double* __restrict__ U, int64_t ldu, double* __restrict__ sigma, double* __restrict__ Vout,
                         int64_t ldv, double* __restrict__ work, const int* __restrict__ flags, int* info) {
    double* W = work;                    // p x p, ld p
    double* V = work + (size_t)p * p;    // p x p, ld p
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
    const int s_sweeps = flags[1], s_conv = flags[2];
    // column norms -> sigma (unsorted, stored temporarily in sigma[]), then rank by value (stable, descending)
    for (int j = warp; j < p; j += nwarps) {
        double s = 0.0;
        for (int k = lane; k < p; k += 32) { const double x = W[(size_t)j * p + k]; s += x * x; }
        s = warp_sum(s);
        if (lane == 0) sigma[j] = sqrt(s);
    }
    __syncthreads();
    // rank: one thread per column; ranks kept in registers; needs sigma unsorted snapshot -> copy into U's first column? use V diag trick: keep in work tail
    double* sig_tmp = work + 2 * (size_t)p * p;   // p doubles (caller provides 2*p*p + p)
    for (int j = tid; j < p; j += blockDim.x) sig_tmp[j] = sigma[j];
    __syncthreads();
    for (int j = warp; j < p; j += nwarps) {
        const double sj = sig_tmp[j];
        int rank = 0;
        for (int i = lane; i < p; i += 32) { const double si = sig_tmp[i]; rank += (si > sj) || (si == sj && i < j); }
        rank = (int)(warp_sum((double)rank) + 0.5);
        const double inv = sj > 0.0 ? 1.0 / sj : 0.0;
        for (int k = lane; k < p; k += 32) {
            U[k + (int64_t)rank * ldu] = W[(size_t)j * p + k] * inv;
            Vout[k + (int64_t)rank * ldv] = V[(size_t)j * p + k];
        }
        if (lane == 0) sigma[rank] = sj;
    }
    __syncthreads();
    // orthonormal completion of U for exactly-zero singular values: candidates e_0, e_1, ... (two Gram-Schmidt passes)
    // executed by warp 0 only (rare path)
    if (warp == 0) {
        int cand = 0;
        for (int j = 0; j < p; ++j) {
            if (sigma[j] > 0.0) continue;
            double* uj = U + (int64_t)j * ldu;
            for (; cand < p; ++cand) {
                for (int k = lane; k < p; k += 32) uj[k] = (k == cand) ? 1.0 : 0.0;
                __syncwarp();
                for (int pass = 0; pass < 2; ++pass) {
                    for (int i = 0; i < p; ++i) {
                        if (i == j) continue;
                        if (i > j && !(sigma[i] > 0.0)) continue;      // not yet built
                        const double* ui = U + (int64_t)i * ldu;
                        double d = 0.0;
                        for (int k = lane; k < p; k += 32) d += ui[k] * uj[k];
                        d = warp_sum(d);
                        for (int k = lane; k < p; k += 32) uj[k] -= d * ui[k];
                        __syncwarp();
                    }
                }
                double nn = 0.0;
                for (int k = lane; k < p; k += 32) nn += uj[k] * uj[k];
                nn = warp_sum(nn);
                if (nn > 0.25) {
                    const double inv = 1.0 / sqrt(nn);
                    for (int k = lane; k < p; k += 32) uj[k] *= inv;
                    __syncwarp();
                    ++cand;
                    break;
                }
            }
        }
    }
    if (tid == 0) { info[0] = s_sweeps; info[1] = s_conv ? 0 : 1; }
}

inline int grid_for(int64_t total, int threads, int cap = 148 * 16) {
    int64_t b = (total + threads - 1) / threads;
    if (b < 1) b = 1;
    if (b > cap) b = cap;
    return (int)b;
}

}  // namespace

cudaError_t fill_philox(int dist, uint64_t seed, uint32_t stream, int64_t rows, int64_t cols, int64_t row_off,
                        double* out, int64_t ld, cudaStream_t st) {
    if (rows <= 0 || cols <= 0) return cudaSuccess;
    fill_philox_kernel<<<grid_for((rows / 4 + 2) * cols, 256), 256, 0, st>>>(dist, seed, stream, rows, cols, row_off, out, ld);
    return LAUNCHED();
}
cudaError_t fill_threefry(int dist, uint64_t key0, uint64_t key1, int64_t rows, int64_t cols, double* out, int64_t ld, cudaStream_t st) {
    if (rows <= 0 || cols <= 0) return cudaSuccess;
    fill_threefry_kernel<<<grid_for((rows * cols + 1) / 2, 256), 256, 0, st>>>(dist, key0, key1, rows, cols, out, ld);
    return LAUNCHED();
}
cudaError_t philox_blocks(int64_t n, const uint32_t* ctr, const uint32_t* key, uint32_t* out, cudaStream_t st) {
    if (n <= 0) return cudaSuccess;
    philox_blocks_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(n, ctr, key, out);
    return LAUNCHED();
}
cudaError_t threefry_blocks(int64_t n, const uint64_t* ctr, const uint64_t* key, uint64_t* out, cudaStream_t st) {
    if (n <= 0) return cudaSuccess;
    threefry_blocks_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(n, ctr, key, out);
    return LAUNCHED();
}
constexpr size_t SMALL_SMEM = 224 * 1024;
cudaError_t chol_upper(double* G, int64_t ld, int p, double tol2, int* flags, int* info, cudaStream_t st) {
    const size_t need = ((size_t)p * (p + 1) / 2 + 2 * (size_t)p) * 8;
    if (need <= SMALL_SMEM) {
        static bool attr = false;
        if (!attr) {
            cudaError_t e = cudaFuncSetAttribute(chol_upper_smem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMALL_SMEM);
            if (e == cudaSuccess) e = cudaFuncSetAttribute(chol_upper_blocked_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMALL_SMEM);
            if (e != cudaSuccess) return e;
            attr = true;
        }
        const char* oe = getenv("RNLA_SMALL_OLD");              // the unblocked kernels, for comparison
        const bool old_kernels = oe && oe[0] == '1';
        if (old_kernels) chol_upper_smem_kernel<<<1, 1024, need, st>>>(G, ld, p, tol2, flags, info);
        else chol_upper_blocked_kernel<<<1, 512, need, st>>>(G, ld, p, tol2, flags, info);
    } else {
        chol_upper_kernel<<<1, 1024, 0, st>>>(G, ld, p, tol2, flags, info);
    }
    return LAUNCHED();
}
cudaError_t tri_inv_upper(const double* R, int64_t ldr, int p, double* Rinv, int64_t ldi, cudaStream_t st) {
    const char* oe = getenv("RNLA_SMALL_OLD");
    const bool old_kernels = oe && oe[0] == '1';
    const size_t need_blocked = ((size_t)p * (p + 1) + (size_t)p + TIB * TIB) * 8;
    if (!old_kernels && p >= 1 && need_blocked <= SMALL_SMEM) {
        static bool battr = false;
        if (!battr) {
            cudaError_t e = cudaFuncSetAttribute(tri_inv_upper_blocked_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMALL_SMEM);
            if (e != cudaSuccess) return e;
            battr = true;
        }
        tri_inv_upper_blocked_kernel<<<1, 1024, need_blocked, st>>>(R, ldr, p, Rinv, ldi);
        return LAUNCHED();
    }
    const size_t need_inplace = ((size_t)p * (p + 1) / 2 + (size_t)p + TIB * TIB) * 8;
    if (!old_kernels && p >= 1 && need_inplace <= SMALL_SMEM) {
        static bool iattr = false;
        if (!iattr) {
            cudaError_t e = cudaFuncSetAttribute(tri_inv_upper_inplace_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMALL_SMEM);
            if (e != cudaSuccess) return e;
            iattr = true;
        }
        tri_inv_upper_inplace_kernel<<<1, 1024, need_inplace, st>>>(R, ldr, p, Rinv, ldi);
        return LAUNCHED();
    }
    const size_t base = ((size_t)p * (p + 1) / 2 + (size_t)p) * 8;
    int ngrp = base + 2 * (size_t)p * 8 <= SMALL_SMEM ? (int)((SMALL_SMEM - base) / ((size_t)p * 8)) : 0;
    ngrp = std::min(64, ngrp) & ~1;                       // whole warps
    if (ngrp >= 2) {
        static bool attr = false;
        if (!attr) {
            cudaError_t e = cudaFuncSetAttribute(tri_inv_upper_smem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMALL_SMEM);
            if (e != cudaSuccess) return e;
            attr = true;
        }
        tri_inv_upper_smem_kernel<<<1, 16 * ngrp, base + (size_t)ngrp * p * 8, st>>>(R, ldr, p, Rinv, ldi);
    } else {
        tri_inv_upper_kernel<<<1, 1024, 0, st>>>(R, ldr, p, Rinv, ldi);
    }
    return LAUNCHED();
}
cudaError_t small_gemm(const double* A, int64_t lda, const double* B, int64_t ldb, double* C, int64_t ldc,
                       int M, int N, int K, cudaStream_t st) {
    if (M <= 0 || N <= 0) return cudaSuccess;
    small_gemm_kernel<<<grid_for((int64_t)M * N, 256), 256, 0, st>>>(A, lda, B, ldb, C, ldc, M, N, K);
    return LAUNCHED();
}
cudaError_t zero_flagged_diag(double* R, int64_t ld, int p, const int* flags, cudaStream_t st) {
    zero_flagged_diag_kernel<<<(p + 255) / 256, 256, 0, st>>>(R, ld, p, flags);
    return LAUNCHED();
}
cudaError_t replace_columns(double* X, int64_t ld, int64_t rows, int64_t row_off, int p, const int* flags,
                            const int64_t* target, cudaStream_t st) {
    if (rows <= 0 || p <= 0) return cudaSuccess;
    dim3 grid((unsigned)grid_for(rows, 256, 64), (unsigned)p);
    replace_columns_kernel<<<grid, 256, 0, st>>>(X, ld, rows, row_off, p, flags, target);
    return LAUNCHED();
}
cudaError_t orth_fixup(double* X, int64_t ld, int64_t rows, int64_t row_off, int64_t rows_global, int p, const int* flags,
                       const int* info, int* state, int* hist, int pass, double* G, int64_t ldg, cudaStream_t st) {
    orth_fixup_kernel<<<grid_for(std::max<int64_t>(rows, 1), 256, 64), 256, 0, st>>>(X, ld, rows, row_off, rows_global, p, flags, info, state, G, ldg);
    ++g_kernel_launches;
    orth_advance_kernel<<<1, 1, 0, st>>>(info, state, hist, pass);
    return LAUNCHED();
}
cudaError_t set_identity(double* X, int64_t ld, int64_t rows, int64_t cols, cudaStream_t st) {
    if (rows <= 0 || cols <= 0) return cudaSuccess;
    set_identity_kernel<<<grid_for(rows * cols, 256), 256, 0, st>>>(X, ld, rows, cols);
    return LAUNCHED();
}
cudaError_t copy_matrix(const double* src, int64_t lds, double* dst, int64_t ldd, int64_t rows, int64_t cols, cudaStream_t st) {
    if (rows <= 0 || cols <= 0) return cudaSuccess;
    return cudaMemcpy2DAsync(dst, (size_t)ldd * 8, src, (size_t)lds * 8, (size_t)rows * 8, (size_t)cols, cudaMemcpyDeviceToDevice, st);
}
cudaError_t transpose_matrix(const double* src, int64_t lds, double* dst, int64_t ldd, int64_t rows, int64_t cols, cudaStream_t st) {
    if (rows <= 0 || cols <= 0) return cudaSuccess;
    dim3 grid((unsigned)((rows + 31) / 32), (unsigned)((cols + 31) / 32));
    transpose_kernel<<<grid, 256, 0, st>>>(src, lds, dst, ldd, rows, cols);
    return LAUNCHED();
}
cudaError_t axpby_matrix(double a, const double* x, int64_t ldx, double b, const double* y, int64_t ldy,
                         double* dst, int64_t ldd, int64_t rows, int64_t cols, cudaStream_t st) {
    if (rows <= 0 || cols <= 0) return cudaSuccess;
    axpby_kernel<<<grid_for(rows * cols, 256), 256, 0, st>>>(a, x, ldx, b, y, ldy, dst, ldd, rows, cols);
    return LAUNCHED();
}
cudaError_t sumsq(const double* X, int64_t ld, int64_t rows, int64_t cols, double* out, double* scratch, int nscratch, cudaStream_t st) {
    if (rows <= 0 || cols <= 0) return cudaSuccess;
    int blocks = grid_for(rows * cols, 256, nscratch);
    sumsq_partial_kernel<<<blocks, 256, 0, st>>>(X, ld, rows, cols, scratch);
    ++g_kernel_launches;
    sumsq_final_kernel<<<1, 32, 0, st>>>(scratch, blocks, out);
    return LAUNCHED();
}
// flag[0] = 1 if any diagonal entry A(i, col_off + i) of the local row block is negative (or NaN): a negative diagonal entry
// proves a negative eigenvalue (e_i^T A e_i < 0)
__global__ void negative_diag_kernel(const double* __restrict__ A, int64_t lda, int64_t rows, int64_t col_off, int* flag) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < rows && !(A[i + (col_off + i) * lda] >= 0.0)) atomicExch(flag, 1);
}
cudaError_t check_negative_diag(const double* A, int64_t lda, int64_t rows, int64_t col_off, int* flag, cudaStream_t st) {
    if (rows <= 0) return cudaSuccess;
    negative_diag_kernel<<<(unsigned)((rows + 255) / 256), 256, 0, st>>>(A, lda, rows, col_off, flag);
    ++g_kernel_launches;
    return cudaGetLastError();
}
cudaError_t check_symmetric(const double* A, int64_t lda, int64_t n, int* flag, cudaStream_t st) {
    if (n <= 0) return cudaSuccess;
    check_symmetric_kernel<<<grid_for(n * n, 256), 256, 0, st>>>(A, lda, n, flag);
    return LAUNCHED();
}
cudaError_t scale_columns(double* X, int64_t ld, int64_t rows, int64_t cols, const double* s, cudaStream_t st) {
    if (rows <= 0 || cols <= 0) return cudaSuccess;
    scale_columns_kernel<<<grid_for(rows * cols, 256), 256, 0, st>>>(X, ld, rows, cols, s);
    return LAUNCHED();
}
size_t jacobi_svd_work_doubles(int p) { return 2 * (size_t)p * p + (size_t)p + 8 + (4 + JACOBI_MAX_SWEEPS + 1) / 2; }

// the sweeps on W (p x p, ld p) with V accumulating the rotations: one CTA for small p, a cooperative grid otherwise
static cudaError_t jacobi_run_sweeps(double* W, double* Vw, int p, int* flags, cudaStream_t st) {
    constexpr int JSMEM = 224 * 1024;
    static bool attr = false;
    if (!attr) {
        cudaError_t e = cudaFuncSetAttribute(jacobi_single_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, JSMEM);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(jacobi_coop_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, JSMEM);
        if (e != cudaSuccess) return e;
        attr = true;
    }
    if (p <= 72) {
        // small enough that one SM's shared-memory bandwidth is not the limit: every sweep inside one CTA
        jacobi_single_kernel<<<1, 1024, 16 * (size_t)p * p + 16, st>>>(W, Vw, p, JACOBI_MAX_SWEEPS, flags);
        ++g_kernel_launches;
    } else {
        // A block pair must fit one SM's shared memory, but a sweep is bound by that SM's shared-memory bandwidth (80 bytes
        // move per element of a column pair) and by the latency of a round, so ~14-20 narrow blocks spread every round over
        // 7-10 SMs (p = 110 in one CTA: 3.2 ms; 7 CTAs of 2 x 8 columns: < 1 ms)
        int b = (int)(JSMEM / (32 * (size_t)p));
        if (b > 64) b = 64;
        if (b < 1) return cudaErrorInvalidValue;
        b = std::min(b, std::max(8, (p + 19) / 20));
        int nblocks = (p + b - 1) / b;
        int NB = (nblocks & 1) ? nblocks + 1 : nblocks;
        int max_sweeps = JACOBI_MAX_SWEEPS;
        void* args[] = {&W, &Vw, &p, &b, &nblocks, &NB, &max_sweeps, &flags};
        cudaError_t e = cudaLaunchCooperativeKernel((const void*)jacobi_coop_kernel, dim3(NB / 2), dim3(1024), args,
                                                    32 * (size_t)b * p + 16, st);
        if (e != cudaSuccess) return e;
        ++g_kernel_launches;
    }
    return cudaSuccess;
}

cudaError_t jacobi_svd(const double* M, int64_t ldm, int p, double* U, int64_t ldu, double* sigma,
                       double* V, int64_t ldv, double* work, int* info, cudaStream_t st, int transpose) {
    double* W = work;
    double* Vw = work + (size_t)p * p;
    int* flags = reinterpret_cast<int*>(work + 2 * (size_t)p * p + (size_t)p);     // 4 + JACOBI_MAX_SWEEPS ints
    jacobi_init_kernel<<<grid_for((int64_t)p * p, 256), 256, 0, st>>>(M, ldm, p, transpose, W, Vw, flags);
    ++g_kernel_launches;
    { cudaError_t e = jacobi_run_sweeps(W, Vw, p, flags, st); if (e != cudaSuccess) return e; }
    // the sweeps factor W0 = X diag(sigma) Y^T with X from the rotated columns and Y the accumulated rotations;
    // W0 = M^T swaps the roles of the two sides
    if (transpose) jacobi_svd_finish_kernel<<<1, 1024, 0, st>>>(p, V, ldv, sigma, U, ldu, work, flags, info);
    else jacobi_svd_finish_kernel<<<1, 1024, 0, st>>>(p, U, ldu, sigma, V, ldv, work, flags, info);
    return LAUNCHED();
}
cudaError_t jacobi_eigh(const double* C, int64_t ldc, int p, double* W, int64_t ldw, double* lambda,
                        int order, double* work, int* info, cudaStream_t st) {
    // Symmetric eigen-decomposition through the one-sided sweeps above: M = C + shift I with shift = 1.01 ||C||_F is
    // positive definite with condition number <= 201, its one-sided Jacobi factors M V = V diag(mu) give the eigenvectors of
    // C, and the eigenvalues are the Rayleigh quotients v^T C v (absolute accuracy eps ||C||, as for any Jacobi method on an
    // indefinite matrix).  The old two-sided kernel ran in one CTA on global memory: 11 ms at p = 110.
    double* Wk = work;
    double* Vw = work + (size_t)p * p;
    int* flags = reinterpret_cast<int*>(work + 2 * (size_t)p * p + (size_t)p);
    eigh_shift_kernel<<<1, 1024, 0, st>>>(C, ldc, p, Wk, Vw, flags);
    ++g_kernel_launches;
    { cudaError_t e = jacobi_run_sweeps(Wk, Vw, p, flags, st); if (e != cudaSuccess) return e; }
    eigh_finish_kernel<<<1, 1024, 0, st>>>(C, ldc, p, Vw, W, ldw, lambda, order, work + 2 * (size_t)p * p, flags, info);
    return LAUNCHED();
}

}  // namespace rnla
