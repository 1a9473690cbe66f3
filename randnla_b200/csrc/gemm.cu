// K1 / K1' / K5  (gemm_nn: tall A streamed once, optional in-kernel Philox right operand) and
// K2 (gemm_tn: A^T Q / Gram, split over the long row dimension with a fixed-order reduction).
//
// Replaces the nalgebra products at reference src/lora_helpers.rs:21,41,76,89,95 and
// src/lora_drivers.rs:66,121,148,187,191.
//
// Shape of both kernels (one CTA per SM, 384 threads):
//   warps 0-3  producers: TMA-engine bulk copies (cp.async.bulk, one per tile column, into a padded
//              bank-conflict-free smem layout), Philox/Gaussian generation of the Omega tile,
//              manual zero-filling loads for ragged edges / unaligned inputs
//   warps 4-11 consumers: 4 (M) x 2 (N) warp grid over a 128 x (16*NT) tile, DMMA.8x8x4 from smem
//   mbarrier full/empty ring between them.
#include "gemm.cuh"
#include <algorithm>
#include "ptx.cuh"
#include "rng.cuh"

namespace rnla {

unsigned long long g_kernel_launches = 0;

namespace {

constexpr int BM = 128;
constexpr int NCONS = 8;
constexpr int NPROD = 4;
constexpr int NPROD_THREADS = NPROD * 32;
constexpr int NTHREADS = (NCONS + NPROD) * 32;
constexpr int SMEM_BUDGET = 226 * 1024;
// 384 threads x 168 regs at launch; producers give registers back, consumers take them:
// 128 x 96 + 256 x 200 = 63488 <= 64512 = 384 x 168
constexpr int PROD_REGS = 96;
constexpr int CONS_REGS = 200;
// TN producers only issue copies: 128 x 40 + 256 x 224 = 62464
constexpr int TN_PROD_REGS = 40;
constexpr int TN_CONS_REGS = 224;

constexpr int cmin(int a, int b) { return a < b ? a : b; }

// ------------------------------------------------------------------------------------------
//  NN
// ------------------------------------------------------------------------------------------
constexpr int NN_BK = 32;
constexpr int GEN_ILP = 2;
constexpr int NN_BMP = BM + 4;      // A tile  [k][BMP]   (row index contiguous, as in global memory)
constexpr int NN_BKP = NN_BK + 4;   // B tile  [j][BKP]   (k contiguous, as in global memory)

template <int NT>
struct NNCfg {
    static constexpr int BN = 16 * NT;
    static constexpr int A_TILE = NN_BK * NN_BMP;
    static constexpr int B_TILE = BN * NN_BKP;
    static constexpr int STAGE = A_TILE + B_TILE;                       // doubles
    static constexpr int STAGES = cmin(6, SMEM_BUDGET / (STAGE * 8));
    static constexpr int SMEM = STAGES * STAGE * 8 + 2 * STAGES * 8;
};

template <int NT, int GEN>
__global__ void __launch_bounds__(NTHREADS, 1)
gemm_nn_kernel(const GemmNN p, const int a_bulk, const int b_bulk) {
    using Cfg = NNCfg<NT>;
    constexpr int BN = Cfg::BN, BK = NN_BK, BMP = NN_BMP, BKP = NN_BKP, STAGES = Cfg::STAGES;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double* tiles = reinterpret_cast<double*>(smem_raw);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + (size_t)STAGES * Cfg::STAGE * 8);
    uint64_t* empty = full + STAGES;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t i0 = (int64_t)blockIdx.x * BM;
    const int64_t j0 = (int64_t)blockIdx.y * BN;
    const int rows_valid = (int)min((int64_t)BM, p.m - i0);
    const int cols_valid = (int)min((int64_t)BN, p.N - j0);
    const int KT_all = (int)((p.K + BK - 1) / BK);
    // split-K: this CTA takes the BK-steps [kt_lo, kt_lo + KT) and writes a partial tile (gemm_nn reduces them in order)
    const int kt_lo = p.ksplit > 1 ? (int)blockIdx.z * p.kt_per : 0;
    const int KT = p.ksplit > 1 ? max(0, min(KT_all - kt_lo, p.kt_per)) : KT_all;

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], NPROD); mbar_init(&empty[s], NCONS); }
        mbar_fence_init();
    }
    __syncthreads();

    if (warp < NPROD) {
        // ================= producers =================
        reg_dec<PROD_REGS>();
        const int tid = threadIdx.x;   // 0..127
        const double* Ablk = p.A + i0;
        for (int kt = 0; kt < KT; ++kt) {
            const int s = kt % STAGES;
            const uint32_t ph = (uint32_t)(kt / STAGES) & 1u;
            mbar_wait(&empty[s], ph ^ 1u);
            double* As = tiles + (size_t)s * Cfg::STAGE;
            double* Bs = As + Cfg::A_TILE;
            const int64_t k0 = (int64_t)(kt_lo + kt) * BK;
            const int kv = (int)min((int64_t)BK, p.K - k0);
            const bool fullk = (kv == BK);
            uint32_t tx = 0;
            // ---- A tile ----
            if (a_bulk && fullk) {
                if (warp == 0) {
                    static_assert(BK == 32, "one A column per lane of warp 0");
                    const uint32_t bytes = (uint32_t)(rows_valid & ~1) * 8u;
                    const double* src = Ablk + (k0 + lane) * p.lda;
                    if (bytes) bulk_g2s(As + lane * BMP, src, bytes, &full[s]);
                    if (rows_valid & 1) As[lane * BMP + rows_valid - 1] = src[rows_valid - 1];
                    tx += bytes * BK;
                }
            } else {
                for (int idx = tid; idx < BK * BM; idx += NPROD_THREADS) {
                    const int kk = idx / BM, r = idx - kk * BM;
                    double v = 0.0;
                    if (kk < kv && r < rows_valid) v = ldg_stream(Ablk + r + (k0 + kk) * p.lda);
                    As[kk * BMP + r] = v;
                }
            }
            // ---- B tile ----
            if (GEN) {
                // (BK/4) * BN Philox blocks per stage = exactly NT per producer thread; groups of GEN_ILP
                // independent blocks are evaluated together so the 10 dependent rounds of one block overlap
                // with those of the others (a single warp per scheduler has no other latency hiding)
                const uint64_t q0 = (p.k_off + (uint64_t)k0) >> 2;
                constexpr int QPS = BK / 4;                      // row-quads per stage
                constexpr int PER_THREAD = QPS * BN / NPROD_THREADS;
                static_assert(QPS * BN % NPROD_THREADS == 0 && PER_THREAD == NT, "generation tiling");
#pragma unroll
                for (int base = 0; base < PER_THREAD; base += GEN_ILP) {
                    u32x4 blk[GEN_ILP];
#pragma unroll
                    for (int u = 0; u < GEN_ILP; ++u) {
                        if (base + u < PER_THREAD) {
                            const int idx = tid + (base + u) * NPROD_THREADS;
                            blk[u] = omega_block(p.seed, p.stream, q0 + (idx % QPS), (uint32_t)(j0 + idx / QPS));
                        }
                    }
#pragma unroll
                    for (int u = 0; u < GEN_ILP; ++u) {
                        if (base + u < PER_THREAD) {
                            const int idx = tid + (base + u) * NPROD_THREADS;
                            const int j = idx / QPS, qd = idx % QPS;
                            double2 lo, hi;
                            if (GEN == 1) {
                                float z[4];
                                gauss_block4(blk[u], z);       // all lanes take part (warp vote inside)
                                lo.x = bits_d(f32_bits_to_f64_bits(f_bits(z[0]))); lo.y = bits_d(f32_bits_to_f64_bits(f_bits(z[1])));
                                hi.x = bits_d(f32_bits_to_f64_bits(f_bits(z[2]))); hi.y = bits_d(f32_bits_to_f64_bits(f_bits(z[3])));
                            } else {
                                lo.x = sample_from_u32(GEN - 1, blk[u].x); lo.y = sample_from_u32(GEN - 1, blk[u].y);
                                hi.x = sample_from_u32(GEN - 1, blk[u].z); hi.y = sample_from_u32(GEN - 1, blk[u].w);
                            }
                            if (j < cols_valid) {
                                double2* dst = reinterpret_cast<double2*>(Bs + j * BKP + 4 * qd);
                                dst[0] = lo; dst[1] = hi;
                            }
                        }
                    }
                }
            } else if (b_bulk && fullk) {
                static_assert(BN <= NPROD_THREADS, "one B column per producer thread");
                if (tid < cols_valid) bulk_g2s(Bs + tid * BKP, p.B + k0 + (j0 + tid) * p.ldb, BK * 8, &full[s]);
                const int wcnt = max(0, min(32, cols_valid - warp * 32));
                tx += (uint32_t)wcnt * BK * 8u;
            } else {
                for (int idx = tid; idx < BK * BN; idx += NPROD_THREADS) {
                    const int j = idx / BK, kk = idx - j * BK;
                    double v = 0.0;
                    if (kk < kv && j < cols_valid) v = p.B[k0 + kk + (j0 + j) * p.ldb];
                    Bs[j * BKP + kk] = v;
                }
            }
            __syncwarp();
            if (lane == 0) {
                if (tx) mbar_arrive_expect_tx(&full[s], tx);
                else mbar_arrive(&full[s]);
            }
        }
    } else {
        // ================= consumers =================
        reg_inc<CONS_REGS>();
        const int cw = warp - NPROD;
        const int wm = cw & 3, wn = cw >> 2;
        const int g = lane >> 2, t = lane & 3;
        double acc[4][NT][2];
#pragma unroll
        for (int mi = 0; mi < 4; ++mi)
#pragma unroll
            for (int ni = 0; ni < NT; ++ni) { acc[mi][ni][0] = 0.0; acc[mi][ni][1] = 0.0; }

        const int a_off = t * BMP + wm * 32 + g;
        const int b_off = (wn * 8 * NT + g) * BKP + t;
        for (int kt = 0; kt < KT; ++kt) {
            const int s = kt % STAGES;
            const uint32_t ph = (uint32_t)(kt / STAGES) & 1u;
            mbar_wait(&full[s], ph);
            const double* As = tiles + (size_t)s * Cfg::STAGE + a_off;
            const double* Bs = tiles + (size_t)s * Cfg::STAGE + Cfg::A_TILE + b_off;
#pragma unroll
            for (int ks = 0; ks < BK / 4; ++ks) {
                double a[4], b[NT];
#pragma unroll
                for (int mi = 0; mi < 4; ++mi) a[mi] = As[ks * 4 * BMP + mi * 8];
#pragma unroll
                for (int ni = 0; ni < NT; ++ni) b[ni] = Bs[ni * 8 * BKP + ks * 4];
#pragma unroll
                for (int mi = 0; mi < 4; ++mi)
#pragma unroll
                    for (int ni = 0; ni < NT; ++ni) dmma884(acc[mi][ni][0], acc[mi][ni][1], a[mi], b[ni]);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[s]);
        }
        // ---- epilogue ----
        double* Cout = p.ksplit > 1 ? p.P + (int64_t)blockIdx.z * p.pstride : p.C;
        const int64_t ldo = p.ksplit > 1 ? p.m : p.ldc;
#pragma unroll
        for (int mi = 0; mi < 4; ++mi) {
            const int64_t row = i0 + wm * 32 + mi * 8 + g;
            if (row < p.m) {
#pragma unroll
                for (int ni = 0; ni < NT; ++ni) {
                    const int64_t col = j0 + wn * 8 * NT + ni * 8 + 2 * t;
                    if (col < p.N) Cout[row + col * ldo] = acc[mi][ni][0];
                    if (col + 1 < p.N) Cout[row + (col + 1) * ldo] = acc[mi][ni][1];
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------
//  TN
// ------------------------------------------------------------------------------------------
template <int NT, int BK>
struct TNCfg {
    static constexpr int BN = 16 * NT;
    static constexpr int BKP = BK + 4;
    static constexpr int A_TILE = BM * BKP;      // [j][BKP]
    static constexpr int B_TILE = BN * BKP;      // [c][BKP]
    static constexpr int STAGE = A_TILE + B_TILE;
    static constexpr int STAGES = cmin(6, SMEM_BUDGET / (STAGE * 8));
    static constexpr int SMEM = STAGES * STAGE * 8 + 2 * STAGES * 8;
};

template <int NT, int BK>
__global__ void __launch_bounds__(NTHREADS, 1)
gemm_tn_kernel(const GemmTN p, double* __restrict__ P, const int64_t ldp, const int64_t pstride,
               const int chunks, const int64_t chunk_rows, const int njb, const int a_bulk, const int q_bulk) {
    using Cfg = TNCfg<NT, BK>;
    constexpr int BN = Cfg::BN, BKP = Cfg::BKP, STAGES = Cfg::STAGES;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double* tiles = reinterpret_cast<double*>(smem_raw);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + (size_t)STAGES * Cfg::STAGE * 8);
    uint64_t* empty = full + STAGES;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int jb = blockIdx.x % njb, chunk = blockIdx.x / njb;
    const int64_t j0 = (int64_t)jb * BM;              // first column of A (= output row)
    const int64_t c0 = (int64_t)blockIdx.y * BN;      // first column of Q (= output column)
    const int jv = (int)min((int64_t)BM, p.n - j0);
    const int cv = (int)min((int64_t)BN, p.N - c0);
    const int64_t r0 = (int64_t)chunk * chunk_rows;
    const int64_t r1 = min(p.m, r0 + chunk_rows);
    const int KT = (int)((r1 - r0 + BK - 1) / BK);

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], NPROD); mbar_init(&empty[s], NCONS); }
        mbar_fence_init();
    }
    __syncthreads();

    if (warp < NPROD) {
        reg_dec<TN_PROD_REGS>();
        const int tid = threadIdx.x;
        for (int kt = 0; kt < KT; ++kt) {
            const int s = kt % STAGES;
            const uint32_t ph = (uint32_t)(kt / STAGES) & 1u;
            mbar_wait(&empty[s], ph ^ 1u);
            double* As = tiles + (size_t)s * Cfg::STAGE;
            double* Qs = As + Cfg::A_TILE;
            const int64_t r = r0 + (int64_t)kt * BK;
            const int kv = (int)min((int64_t)BK, r1 - r);
            const bool fullk = (kv == BK);
            uint32_t tx = 0;
            if (a_bulk && fullk) {
                int cnt = 0;
                for (int j = tid; j < jv; j += NPROD_THREADS, ++cnt)
                    bulk_g2s(As + j * BKP, p.A + r + (j0 + j) * p.lda, BK * 8, &full[s]);
                // every lane must know the warp's byte count: recompute it arithmetically
                const int wfirst = warp * 32;
                int wcnt = 0;
                for (int base = wfirst; base < jv; base += NPROD_THREADS) wcnt += min(32, jv - base);
                tx += (uint32_t)wcnt * BK * 8u;
            } else {
                for (int idx = tid; idx < BM * BK; idx += NPROD_THREADS) {
                    const int j = idx / BK, kk = idx - j * BK;
                    double v = 0.0;
                    if (kk < kv && j < jv) v = ldg_stream(p.A + r + kk + (j0 + j) * p.lda);
                    As[j * BKP + kk] = v;
                }
            }
            if (q_bulk && fullk) {
                for (int c = tid; c < cv; c += NPROD_THREADS)
                    bulk_g2s(Qs + c * BKP, p.Q + r + (c0 + c) * p.ldq, BK * 8, &full[s]);
                const int wfirst = warp * 32;
                int wcnt = 0;
                for (int base = wfirst; base < cv; base += NPROD_THREADS) wcnt += min(32, cv - base);
                tx += (uint32_t)wcnt * BK * 8u;
            } else {
                for (int idx = tid; idx < BN * BK; idx += NPROD_THREADS) {
                    const int c = idx / BK, kk = idx - c * BK;
                    double v = 0.0;
                    if (kk < kv && c < cv) v = p.Q[r + kk + (c0 + c) * p.ldq];
                    Qs[c * BKP + kk] = v;
                }
            }
            __syncwarp();
            if (lane == 0) {
                if (tx) mbar_arrive_expect_tx(&full[s], tx);
                else mbar_arrive(&full[s]);
            }
        }
    } else {
        reg_inc<TN_CONS_REGS>();
        const int cw = warp - NPROD;
        const int wm = cw & 3, wn = cw >> 2;
        const int g = lane >> 2, t = lane & 3;
        double acc[4][NT][2];
#pragma unroll
        for (int mi = 0; mi < 4; ++mi)
#pragma unroll
            for (int ni = 0; ni < NT; ++ni) { acc[mi][ni][0] = 0.0; acc[mi][ni][1] = 0.0; }

        const int a_off = (wm * 32 + g) * BKP + t;
        const int b_off = (wn * 8 * NT + g) * BKP + t;
        for (int kt = 0; kt < KT; ++kt) {
            const int s = kt % STAGES;
            const uint32_t ph = (uint32_t)(kt / STAGES) & 1u;
            mbar_wait(&full[s], ph);
            const double* As = tiles + (size_t)s * Cfg::STAGE + a_off;
            const double* Qs = tiles + (size_t)s * Cfg::STAGE + Cfg::A_TILE + b_off;
#pragma unroll
            for (int ks = 0; ks < BK / 4; ++ks) {
                double a[4], b[NT];
#pragma unroll
                for (int mi = 0; mi < 4; ++mi) a[mi] = As[mi * 8 * BKP + ks * 4];
#pragma unroll
                for (int ni = 0; ni < NT; ++ni) b[ni] = Qs[ni * 8 * BKP + ks * 4];
#pragma unroll
                for (int mi = 0; mi < 4; ++mi)
#pragma unroll
                    for (int ni = 0; ni < NT; ++ni) dmma884(acc[mi][ni][0], acc[mi][ni][1], a[mi], b[ni]);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[s]);
        }
        double* out; int64_t ld;
        if (chunks > 1) { out = P + (int64_t)chunk * pstride; ld = ldp; }
        else { out = p.Z; ld = p.ldz; }
        const bool accum = (chunks == 1) && p.accumulate;
#pragma unroll
        for (int mi = 0; mi < 4; ++mi) {
            const int64_t row = j0 + wm * 32 + mi * 8 + g;
            if (row < p.n) {
#pragma unroll
                for (int ni = 0; ni < NT; ++ni) {
                    const int64_t col = c0 + wn * 8 * NT + ni * 8 + 2 * t;
                    if (col < p.N) {
                        double* d = out + row + col * ld;
                        *d = accum ? *d + acc[mi][ni][0] : acc[mi][ni][0];
                    }
                    if (col + 1 < p.N) {
                        double* d = out + row + (col + 1) * ld;
                        *d = accum ? *d + acc[mi][ni][1] : acc[mi][ni][1];
                    }
                }
            }
        }
    }
}

// fixed-order reduction of the per-chunk partials
__global__ void __launch_bounds__(256)
tn_reduce_kernel(const double* __restrict__ P, int64_t ldp, int64_t pstride, int chunks,
                 double* __restrict__ Z, int64_t ldz, int64_t n, int64_t N, int accumulate) {
    const int64_t total = n * N;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (int64_t)gridDim.x * blockDim.x) {
        const int64_t j = idx / n, i = idx - j * n;
        const double* src = P + i + j * ldp;
        double s = 0.0;
        for (int c = 0; c < chunks; ++c) s += src[(int64_t)c * pstride];
        double* d = Z + i + j * ldz;
        *d = accumulate ? *d + s : s;
    }
}

// same reduction for small outputs with many chunks (Gram matrices of tall panels: 110 x 110 outputs, ~370 chunks):
// 8 threads share one output, thread g sums chunks g, g+8, ... and the 8 partial sums are combined in a fixed tree,
// so the result is still independent of scheduling.  32 consecutive outputs per block keep the loads coalesced.
__global__ void __launch_bounds__(256)
tn_reduce_grouped_kernel(const double* __restrict__ P, int64_t ldp, int64_t pstride, int chunks,
                         double* __restrict__ Z, int64_t ldz, int64_t n, int64_t N, int accumulate) {
    __shared__ double sh[8][32];
    const int o = threadIdx.x & 31, g = threadIdx.x >> 5;
    const int64_t total = n * N;
    const int64_t idx = (int64_t)blockIdx.x * 32 + o;
    double s = 0.0;
    if (idx < total) {
        const int64_t j = idx / n, i = idx - j * n;
        const double* src = P + i + j * ldp;
        for (int c = g; c < chunks; c += 8) s += src[(int64_t)c * pstride];
    }
    sh[g][o] = s;
    __syncthreads();
    if (g == 0 && idx < total) {
        const int64_t j = idx / n, i = idx - j * n;
        const double t = ((sh[0][o] + sh[1][o]) + (sh[2][o] + sh[3][o])) + ((sh[4][o] + sh[5][o]) + (sh[6][o] + sh[7][o]));
        double* d = Z + i + j * ldz;
        *d = accumulate ? *d + t : t;
    }
}

// ------------------------------------------------------------------------------------------
//  NN, thin right operand (N <= 24): the HBM-bound regime (l = k + p below the ~21-column ridge).
//  The 128 x 32 A tiles of the kernel above arrive as 1 KB pieces, one per column, 1.6 MB apart: measured 2.98 TB/s at
//  l = 16 (36 % of DRAM peak, consumers starved at the full barrier).  DRAM wants long bursts, so this variant gives a CTA
//  TBM = 1024 (N <= 16) or 512 (N <= 32) consecutive rows and stages A as column segments of 8 / 4 KB; a warp then owns
//  128 / 64 rows, i.e. 16 / 8 DMMA row tiles x N/8 column tiles of accumulators.  A stage is ONE DMMA k-step (4 columns,
//  32 KB) so that 6+ stages are in flight and a freed stage is refilled at once: the producer does nothing but issue
//  copies, the thin B operand (K x N, L2-resident) is fetched by the consumers one stage ahead into registers.
// ------------------------------------------------------------------------------------------
// N <= 24: with four 8-column tiles the prefetched operand no longer fits the registers (measured at l = 32: 14.5 / 28 ms
// per pass against 7.6 / 7.0 ms at l = 24); wider operands take the general kernels
constexpr int THIN_MAX_N = 24;
constexpr int TBK = 4;                  // one DMMA k-step per stage: 32 KB stages, 6 of them in flight
constexpr int THIN_CONS = 8;
constexpr int THIN_THREADS = (THIN_CONS + 1) * 32;

template <int MI /* 8-row tiles per warp */, int NI /* 8-column tiles */>
struct ThinCfg {
    static constexpr int TBM = THIN_CONS * 8 * MI;
    static constexpr int TBMP = TBM + 4;
    static constexpr int STAGE = TBK * TBMP;             // doubles: the A tile only, B comes straight from L2
    static constexpr int STAGES = cmin(8, SMEM_BUDGET / (STAGE * 8));
    static constexpr int SMEM = STAGES * STAGE * 8 + 2 * STAGES * 8;
};

template <int MI, int NI>
__global__ void __launch_bounds__(THIN_THREADS, 1)
gemm_nn_thin_kernel(const GemmNN p) {
    using Cfg = ThinCfg<MI, NI>;
    constexpr int TBM = Cfg::TBM, TBMP = Cfg::TBMP, STAGES = Cfg::STAGES;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double* tiles = reinterpret_cast<double*>(smem_raw);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + (size_t)STAGES * Cfg::STAGE * 8);
    uint64_t* empty = full + STAGES;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t i0 = (int64_t)blockIdx.x * TBM;
    const int rows_valid = (int)min((int64_t)TBM, p.m - i0);
    const int KT_all = (int)((p.K + TBK - 1) / TBK);
    const int kt_lo = p.ksplit > 1 ? (int)blockIdx.z * p.kt_per : 0;
    const int KT = p.ksplit > 1 ? max(0, min(KT_all - kt_lo, p.kt_per)) : KT_all;
    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], THIN_CONS); }
        mbar_fence_init();
    }
    __syncthreads();
    if (warp == THIN_CONS) {
        // ================= producer warp: nothing but the copies, issued the moment a stage is free =================
        const double* Ablk = p.A + i0;
        const uint32_t bytes = (uint32_t)(rows_valid & ~1) * 8u;
        for (int kt = 0; kt < KT; ++kt) {
            const int s = kt % STAGES;
            mbar_wait(&empty[s], ((uint32_t)(kt / STAGES) & 1u) ^ 1u);
            double* As = tiles + (size_t)s * Cfg::STAGE;
            const int64_t k0 = (int64_t)(kt_lo + kt) * TBK;
            const int kv = (int)min((int64_t)TBK, p.K - k0);
            if (kv == TBK && !(rows_valid & 1)) {
                if (lane == 0) {
                    mbar_arrive_expect_tx(&full[s], bytes * TBK);
#pragma unroll
                    for (int kk = 0; kk < TBK; ++kk) bulk_g2s(As + kk * TBMP, Ablk + (k0 + kk) * p.lda, bytes, &full[s]);
                }
            } else {
                // last K step / odd last row: columns past K are zero (their B rows are zero too, but 0 * garbage is NaN)
                for (int kk = kv; kk < TBK; ++kk)
                    for (int r = lane; r < TBM; r += 32) As[kk * TBMP + r] = 0.0;
                if (rows_valid & 1)
                    for (int kk = lane; kk < kv; kk += 32) As[kk * TBMP + rows_valid - 1] = Ablk[rows_valid - 1 + (k0 + kk) * p.lda];
                __syncwarp();
                if (lane == 0) {
                    mbar_arrive_expect_tx(&full[s], bytes * (uint32_t)kv);
                    if (bytes)
                        for (int kk = 0; kk < kv; ++kk) bulk_g2s(As + kk * TBMP, Ablk + (k0 + kk) * p.lda, bytes, &full[s]);
                }
            }
        }
    } else {
        // ================= consumers: warp w owns rows [w * 8 MI, (w + 1) * 8 MI) of the tile =================
        const int g = lane >> 2, t = lane & 3;
        double acc[MI][NI][2];
#pragma unroll
        for (int mi = 0; mi < MI; ++mi)
#pragma unroll
            for (int ni = 0; ni < NI; ++ni) { acc[mi][ni][0] = 0.0; acc[mi][ni][1] = 0.0; }
        const int a_off = t * TBMP + warp * 8 * MI + g;
        // B fragment b[ni] = B(k0 + t, 8 ni + g): K x N is L2-resident, fetched one stage ahead into registers
        auto load_b = [&](int kt, double* b) {
            const int64_t k = (int64_t)(kt_lo + kt) * TBK + t;
#pragma unroll
            for (int ni = 0; ni < NI; ++ni) {
                const int col = ni * 8 + g;
                b[ni] = (kt < KT && k < p.K && col < p.N) ? __ldg(p.B + k + (int64_t)col * p.ldb) : 0.0;
            }
        };
        // three stages of lookahead: a 16-32 KB stage is 700-1400 cycles of work, an L2 round trip under load is more
        double b1[NI], b2[NI], b3[NI];
        load_b(0, b1); load_b(1, b2); load_b(2, b3);
        for (int kt = 0; kt < KT; ++kt) {
            const int s = kt % STAGES;
            double b[NI];
#pragma unroll
            for (int ni = 0; ni < NI; ++ni) { b[ni] = b1[ni]; b1[ni] = b2[ni]; b2[ni] = b3[ni]; }
            load_b(kt + 3, b3);
            mbar_wait(&full[s], (uint32_t)(kt / STAGES) & 1u);
            const double* As = tiles + (size_t)s * Cfg::STAGE + a_off;
#pragma unroll
            for (int mi = 0; mi < MI; ++mi) {
                const double a = As[mi * 8];
#pragma unroll
                for (int ni = 0; ni < NI; ++ni) dmma884(acc[mi][ni][0], acc[mi][ni][1], a, b[ni]);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[s]);
        }
        double* Cout = p.ksplit > 1 ? p.P + (int64_t)blockIdx.z * p.pstride : p.C;
        const int64_t ldo = p.ksplit > 1 ? p.m : p.ldc;
#pragma unroll
        for (int mi = 0; mi < MI; ++mi) {
            const int64_t row = i0 + warp * 8 * MI + mi * 8 + g;
            if (row < p.m) {
#pragma unroll
                for (int ni = 0; ni < NI; ++ni) {
                    const int64_t col = ni * 8 + 2 * t;
                    if (col < p.N) Cout[row + col * ldo] = acc[mi][ni][0];
                    if (col + 1 < p.N) Cout[row + (col + 1) * ldo] = acc[mi][ni][1];
                }
            }
        }
    }
}

// C = sum_z P[z] for the split-K partials of gemm_nn, fixed order
__global__ void __launch_bounds__(256)
nn_reduce_kernel(const double* __restrict__ P, int64_t pstride, int ksplit, int64_t m, int64_t N, double* __restrict__ C, int64_t ldc) {
    const int64_t total = m * N;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
        double s = 0.0;
        for (int z = 0; z < ksplit; ++z) s += P[(int64_t)z * pstride + idx];
        C[(idx % m) + (idx / m) * ldc] = s;
    }
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

template <int NT, int GEN>
cudaError_t launch_nn_g(const GemmNN& p, int nblkN, cudaStream_t st) {
    using Cfg = NNCfg<NT>;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(gemm_nn_kernel<NT, GEN>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM);
        if (e != cudaSuccess) return e;
        attr_set = true;
    }
    const int a_bulk = aligned16(p.A) && (p.lda % 2 == 0);
    const int b_bulk = !p.gen && aligned16(p.B) && (p.ldb % 2 == 0);
    dim3 grid((unsigned)((p.m + BM - 1) / BM), (unsigned)nblkN, (unsigned)p.ksplit);
    gemm_nn_kernel<NT, GEN><<<grid, NTHREADS, Cfg::SMEM, st>>>(p, a_bulk, b_bulk);
    ++g_kernel_launches;
    return cudaGetLastError();
}
template <int NT>
cudaError_t launch_nn(const GemmNN& p, int nblkN, cudaStream_t st) {
    if (!p.gen) return launch_nn_g<NT, 0>(p, nblkN, st);
    switch (p.dist) {
        case DIST_GAUSSIAN: return launch_nn_g<NT, 1>(p, nblkN, st);
        case DIST_UNIFORM: return launch_nn_g<NT, 2>(p, nblkN, st);
        default: return launch_nn_g<NT, 3>(p, nblkN, st);
    }
}

constexpr int TN_BK = 32;

struct TNPlan { int nblkN, NT, njb, chunks; int64_t chunk_rows, ldp, pstride; };

TNPlan plan_tn(int64_t m, int64_t n, int64_t N, int sms) {
    TNPlan pl;
    pl.nblkN = (int)((N + 127) / 128);
    const int64_t per = (N + pl.nblkN - 1) / pl.nblkN;
    pl.NT = (int)((per + 15) / 16);
    pl.njb = (int)((n + BM - 1) / BM);
    const int64_t tiles = (int64_t)pl.njb * pl.nblkN;
    int64_t want = (8LL * sms + tiles - 1) / tiles;
    const int64_t maxc = m / (16 * TN_BK) > 0 ? m / (16 * TN_BK) : 1;
    if (want > maxc) want = maxc;
    if (want < 1) want = 1;
    // one CTA per SM: among the chunk counts around `want`, take the one whose tiles * chunks fills whole waves best
    // (157 column blocks x 8 chunks = 8.49 waves at the headline size wastes 5.7 %; x 16 = 16.97 waves wastes 0.2 %)
    {
        double best = -1.0; int64_t bestc = want;
        const int64_t lo = want / 2 > 1 ? want / 2 : 1, hi = want * 2 + 1 < maxc ? want * 2 + 1 : maxc;
        for (int64_t cnd = lo; cnd <= hi; ++cnd) {
            const int64_t units = tiles * cnd, waves = (units + sms - 1) / sms;
            // the partials cost 2 * 8 bytes per output per chunk: a small penalty per extra chunk
            const double eff = (double)units / (double)(waves * sms) - 0.0005 * (double)cnd;
            if (eff > best + 1e-12) { best = eff; bestc = cnd; }
        }
        want = bestc;
    }
    int64_t cr = (m + want - 1) / want;
    cr = (cr + TN_BK - 1) / TN_BK * TN_BK;
    if (cr <= 0) cr = TN_BK;
    pl.chunk_rows = cr;
    pl.chunks = (int)((m + cr - 1) / cr);
    if (pl.chunks < 1) pl.chunks = 1;
    pl.ldp = n;
    pl.pstride = n * N;
    return pl;
}

template <int NT>
cudaError_t launch_tn(const GemmTN& p, const TNPlan& pl, double* ws, cudaStream_t st) {
    using Cfg = TNCfg<NT, TN_BK>;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(gemm_tn_kernel<NT, TN_BK>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM);
        if (e != cudaSuccess) return e;
        attr_set = true;
    }
    const int a_bulk = aligned16(p.A) && (p.lda % 2 == 0);
    const int q_bulk = aligned16(p.Q) && (p.ldq % 2 == 0);
    dim3 grid((unsigned)(pl.njb * pl.chunks), (unsigned)pl.nblkN);
    gemm_tn_kernel<NT, TN_BK><<<grid, NTHREADS, Cfg::SMEM, st>>>(p, ws, pl.ldp, pl.pstride, pl.chunks,
                                                               pl.chunk_rows, pl.njb, a_bulk, q_bulk);
    ++g_kernel_launches;
    return cudaGetLastError();
}

}  // namespace

static cudaError_t gemm_nn_dispatch(const GemmNN& p, int nblkN, int NT, cudaStream_t st) {
    switch (NT) {
        case 1: return launch_nn<1>(p, nblkN, st);
        case 2: return launch_nn<2>(p, nblkN, st);
        case 3: return launch_nn<3>(p, nblkN, st);
        case 4: return launch_nn<4>(p, nblkN, st);
        case 5: return launch_nn<5>(p, nblkN, st);
        case 6: return launch_nn<6>(p, nblkN, st);
        case 7: return launch_nn<7>(p, nblkN, st);
        default: return launch_nn<8>(p, nblkN, st);
    }
}

template <int MI, int NI>
static cudaError_t launch_nn_thin(const GemmNN& p, cudaStream_t st) {
    using Cfg = ThinCfg<MI, NI>;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(gemm_nn_thin_kernel<MI, NI>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM);
        if (e != cudaSuccess) return e;
        attr_set = true;
    }
    dim3 grid((unsigned)((p.m + Cfg::TBM - 1) / Cfg::TBM), 1, (unsigned)p.ksplit);
    gemm_nn_thin_kernel<MI, NI><<<grid, THIN_THREADS, Cfg::SMEM, st>>>(p);
    ++g_kernel_launches;
    return cudaGetLastError();
}

// split-K for the thin kernel: few, tall tiles (196 at m = 200 000), so the parts are what fills the SMs
int gemm_nn_thin_ksplit(int64_t m, int64_t K, int tbm, int sms) {
    const int64_t tiles = (m + tbm - 1) / tbm;
    const int KT = (int)((K + TBK - 1) / TBK);
    int best_s = 1; double best = -1.0;
    for (int s = 1; s <= 16; ++s) {
        if (s > 1 && KT / s < 128) break;
        const int64_t units = tiles * s, waves = (units + sms - 1) / sms;
        const double eff = (double)units / (double)(waves * sms) - (s > 1 ? 8.0 * s / (double)K : 0.0);
        if (eff > best + 1e-9) { best = eff; best_s = s; }
    }
    return best_s;
}

static cudaError_t gemm_nn_thin(GemmNN p, int sms, cudaStream_t st) {
    const bool wide = p.N > 16;
    const int tbm = wide ? ThinCfg<8, 3>::TBM : ThinCfg<16, 2>::TBM;
    p.ksplit = gemm_nn_thin_ksplit(p.m, p.K, tbm, sms);
    const int KT = (int)((p.K + TBK - 1) / TBK);
    p.kt_per = (KT + p.ksplit - 1) / p.ksplit;
    p.pstride = p.m * p.N;
    cudaError_t e = cudaSuccess;
    if (p.ksplit > 1) {
        e = cudaMallocAsync(reinterpret_cast<void**>(&p.P), (size_t)p.ksplit * (size_t)p.pstride * sizeof(double), st);
        if (e != cudaSuccess) return e;
    }
    if (p.N <= 8) e = launch_nn_thin<16, 1>(p, st);
    else if (p.N <= 16) e = launch_nn_thin<16, 2>(p, st);
    else e = launch_nn_thin<8, 3>(p, st);
    if (p.ksplit > 1) {
        if (e == cudaSuccess) {
            const int64_t total = p.m * p.N;
            int blocks = (int)((total + 255) / 256);
            if (blocks > sms * 8) blocks = sms * 8;
            nn_reduce_kernel<<<blocks, 256, 0, st>>>(p.P, p.pstride, p.ksplit, p.m, p.N, p.C, p.ldc);
            ++g_kernel_launches;
            e = cudaGetLastError();
        }
        cudaFreeAsync(p.P, st);
    }
    return e;
}

// host logic of gemm_nn's split-K choice, separated so that it can be checked without a GPU
int gemm_nn_ksplit(int64_t m, int64_t K, int64_t N, int sms) {
    const int nblkN = (int)((N + 127) / 128);
    const int64_t tiles = ((m + BM - 1) / BM) * nblkN;
    const int KT = (int)((K + NN_BK - 1) / NN_BK);
    int best_s = 1;
    double best = -1.0;
    for (int s = 1; s <= 4; ++s) {
        if (s > 1 && KT / s < 64) break;
        const int64_t units = tiles * s, waves = (units + sms - 1) / sms;
        const double eff = (double)units / (double)(waves * sms) - (s > 1 ? 42.0 * s / (double)K + 0.002 : 0.0);
        if (eff > best + 1e-9) { best = eff; best_s = s; }
    }
    return best_s;
}
void gemm_tn_plan_info(int64_t m, int64_t n, int64_t N, int sms, int* chunks, int64_t* chunk_rows, int* tiles) {
    const TNPlan pl = plan_tn(m, n, N, sms);
    *chunks = pl.chunks; *chunk_rows = pl.chunk_rows; *tiles = pl.njb * pl.nblkN;
}

cudaError_t gemm_nn(const GemmNN& p0, cudaStream_t st) {
    if (p0.m <= 0 || p0.N <= 0) return cudaSuccess;
    if (p0.gen && (p0.k_off & 3)) return cudaErrorInvalidValue;
    GemmNN p = p0;
    const int nblkN = (int)((p.N + 127) / 128);
    const int64_t per = (p.N + nblkN - 1) / nblkN;
    const int NT = (int)((per + 15) / 16);
    // One CTA per SM and one 128-row tile per CTA: 1563 tiles on 148 SMs are 10.56 waves, i.e. 4 % of the machine idles
    // in the last one (12 % at m = 50 000).  Splitting the K loop in s parts multiplies the number of (shorter) CTAs; take
    // the s whose waves are fullest, net of the partial tiles' extra traffic (2 * 8 bytes per output and part).
    static int sms = 0;
    if (sms == 0) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
    }
    if (!p.gen && p.N <= THIN_MAX_N && p.K >= 1024 && p.m >= 8192 && aligned16(p.A) && (p.lda % 2 == 0)) return gemm_nn_thin(p, sms, st);
    const int KT = (int)((p.K + NN_BK - 1) / NN_BK);
    const int best_s = gemm_nn_ksplit(p.m, p.K, p.N, sms);
    if (best_s == 1) return gemm_nn_dispatch(p, nblkN, NT, st);
    p.ksplit = best_s;
    p.kt_per = (KT + best_s - 1) / best_s;
    p.pstride = p.m * p.N;
    cudaError_t e = cudaMallocAsync(reinterpret_cast<void**>(&p.P), (size_t)best_s * (size_t)p.pstride * sizeof(double), st);
    if (e != cudaSuccess) return e;
    e = gemm_nn_dispatch(p, nblkN, NT, st);
    if (e == cudaSuccess) {
        const int64_t total = p.m * p.N;
        int blocks = (int)((total + 255) / 256);
        if (blocks > sms * 8) blocks = sms * 8;
        nn_reduce_kernel<<<blocks, 256, 0, st>>>(p.P, p.pstride, p.ksplit, p.m, p.N, p.C, p.ldc);
        ++g_kernel_launches;
        e = cudaGetLastError();
    }
    cudaFreeAsync(p.P, st);
    return e;
}

// ------------------------------------------------------------------------------------------
//  TN, thin right operand (N <= 24): the HBM-bound regime.  Same recipe as gemm_nn_thin_kernel: a CTA takes TCN = 32
//  columns of A and a chunk of rows, stages A as 32 column segments of TRK = 256 rows (2 KB pieces, 64 KB per stage, 3
//  stages), the producer warp only issues copies, and the thin operand Q (m x N, L2-resident) is fetched by the
//  consumers one stage ahead into registers.  The 8 consumer warps split the rows of a stage (32 each); their 32 x N
//  partial tiles are added in warp order through shared memory at the end, then the chunks by tn_reduce_kernel.
// ------------------------------------------------------------------------------------------
constexpr int TCN = 32;
constexpr int TRK = 256;
constexpr int TRKP = TRK + 4;
constexpr int TN_THIN_STAGES = 3;
constexpr int TN_THIN_SMEM = TN_THIN_STAGES * TCN * TRKP * 8 + 2 * TN_THIN_STAGES * 8;

template <int NI>
__global__ void __launch_bounds__(THIN_THREADS, 1)
gemm_tn_thin_kernel(const GemmTN p, double* __restrict__ P, int64_t ldp, int64_t pstride, int ncg, int64_t chunk_rows) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double* tiles = reinterpret_cast<double*>(smem_raw);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + (size_t)TN_THIN_STAGES * TCN * TRKP * 8);
    uint64_t* empty = full + TN_THIN_STAGES;
    constexpr int STAGE = TCN * TRKP;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int cg = blockIdx.x % ncg, chunk = blockIdx.x / ncg;
    const int64_t c0 = (int64_t)cg * TCN;
    const int cols_valid = (int)min((int64_t)TCN, p.n - c0);
    const int64_t r0 = (int64_t)chunk * chunk_rows, r1 = min(p.m, r0 + chunk_rows);
    const int KT = (int)((r1 - r0 + TRK - 1) / TRK);
    if (threadIdx.x == 0) {
        for (int s = 0; s < TN_THIN_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], THIN_CONS); }
        mbar_fence_init();
    }
    __syncthreads();
    if (warp == THIN_CONS) {
        // ================= producer warp =================
        for (int kt = 0; kt < KT; ++kt) {
            const int s = kt % TN_THIN_STAGES;
            mbar_wait(&empty[s], ((uint32_t)(kt / TN_THIN_STAGES) & 1u) ^ 1u);
            double* As = tiles + (size_t)s * STAGE;
            const int64_t k0 = r0 + (int64_t)kt * TRK;
            const int kv = (int)min((int64_t)TRK, r1 - k0);
            const uint32_t bytes = (uint32_t)(kv & ~1) * 8u;
            if (kv < TRK || cols_valid < TCN) {
                // ragged end of the chunk / of the matrix: what the copies do not bring is zero (0 * garbage would be NaN)
                for (int idx = lane; idx < TCN * TRK; idx += 32) {
                    const int c = idx / TRK, k = idx - c * TRK;
                    if (c >= cols_valid || k >= (kv & ~1)) As[c * TRKP + k] = (c < cols_valid && k < kv) ? p.A[k0 + k + (c0 + c) * p.lda] : 0.0;
                }
                __syncwarp();
            }
            if (lane == 0) mbar_arrive_expect_tx(&full[s], bytes * (uint32_t)cols_valid);
            __syncwarp();
            if (bytes && lane < cols_valid) bulk_g2s(As + lane * TRKP, p.A + k0 + (c0 + lane) * p.lda, bytes, &full[s]);
        }
    } else {
        // ================= consumers: warp w takes rows [32 w, 32 w + 32) of every stage =================
        const int g = lane >> 2, t = lane & 3;
        double acc[4][NI][2];
#pragma unroll
        for (int mt = 0; mt < 4; ++mt)
#pragma unroll
            for (int ni = 0; ni < NI; ++ni) { acc[mt][ni][0] = 0.0; acc[mt][ni][1] = 0.0; }
        constexpr int KS = TRK / THIN_CONS / 4;           // k-steps per warp and stage
        // b[ks][ni] = Q(k0 + 32 w + 4 ks + t, 8 ni + g), zero outside the chunk / the matrix
        auto load_b = [&](int kt, double (*b)[NI]) {
            const int64_t kb = r0 + (int64_t)kt * TRK + warp * (TRK / THIN_CONS) + t;
#pragma unroll
            for (int ks = 0; ks < KS; ++ks)
#pragma unroll
                for (int ni = 0; ni < NI; ++ni) {
                    const int64_t k = kb + 4 * ks;
                    const int col = ni * 8 + g;
                    b[ks][ni] = (kt < KT && k < r1 && col < p.N) ? __ldg(p.Q + k + (int64_t)col * p.ldq) : 0.0;
                }
        };
        double bn[KS][NI];
        load_b(0, bn);
        const int a_off = g * TRKP + warp * (TRK / THIN_CONS) + t;
        for (int kt = 0; kt < KT; ++kt) {
            const int s = kt % TN_THIN_STAGES;
            double b[KS][NI];
#pragma unroll
            for (int ks = 0; ks < KS; ++ks)
#pragma unroll
                for (int ni = 0; ni < NI; ++ni) b[ks][ni] = bn[ks][ni];
            load_b(kt + 1, bn);
            mbar_wait(&full[s], (uint32_t)(kt / TN_THIN_STAGES) & 1u);
            const double* As = tiles + (size_t)s * STAGE + a_off;
#pragma unroll
            for (int ks = 0; ks < KS; ++ks) {
#pragma unroll
                for (int mt = 0; mt < 4; ++mt) {
                    const double a = As[mt * 8 * TRKP + 4 * ks];          // A(k, c0 + 8 mt + g) = (A^T)(8 mt + g, k)
#pragma unroll
                    for (int ni = 0; ni < NI; ++ni) dmma884(acc[mt][ni][0], acc[mt][ni][1], a, b[ks][ni]);
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[s]);
        }
        // add the warps' partial tiles in warp order (the ring is drained: every full barrier was waited on)
        asm volatile("bar.sync 1, %0;" ::"n"(THIN_CONS * 32));
        double* red = tiles;                               // [warp][32 x 8 NI]
#pragma unroll
        for (int mt = 0; mt < 4; ++mt)
#pragma unroll
            for (int ni = 0; ni < NI; ++ni) {
                red[(size_t)warp * TCN * 8 * NI + (mt * 8 + g) + (ni * 8 + 2 * t) * TCN] = acc[mt][ni][0];
                red[(size_t)warp * TCN * 8 * NI + (mt * 8 + g) + (ni * 8 + 2 * t + 1) * TCN] = acc[mt][ni][1];
            }
        asm volatile("bar.sync 1, %0;" ::"n"(THIN_CONS * 32));
        for (int idx = threadIdx.x; idx < TCN * 8 * NI; idx += THIN_CONS * 32) {
            const int i = idx % TCN, j = idx / TCN;
            double sum = 0.0;
#pragma unroll
            for (int w = 0; w < THIN_CONS; ++w) sum += red[(size_t)w * TCN * 8 * NI + idx];
            if (i < cols_valid && j < p.N) P[(int64_t)chunk * pstride + (c0 + i) + (int64_t)j * ldp] = sum;
        }
    }
}

struct TNThinPlan { int ncg, chunks; int64_t chunk_rows; };
TNThinPlan plan_tn_thin(int64_t m, int64_t n, int sms) {
    TNThinPlan pl;
    pl.ncg = (int)((n + TCN - 1) / TCN);
    const int64_t maxc = std::max<int64_t>(1, m / (8 * TRK));
    int64_t want = std::min<int64_t>(maxc, std::max<int64_t>(1, (4LL * sms + pl.ncg - 1) / pl.ncg));
    double best = -1.0; int64_t bestc = want;
    for (int64_t cnd = std::max<int64_t>(1, want / 2); cnd <= std::min(maxc, want * 2 + 1); ++cnd) {
        const int64_t units = (int64_t)pl.ncg * cnd, waves = (units + sms - 1) / sms;
        const double eff = (double)units / (double)(waves * sms) - 0.0005 * (double)cnd;
        if (eff > best + 1e-12) { best = eff; bestc = cnd; }
    }
    int64_t cr = (m + bestc - 1) / bestc;
    cr = (cr + TRK - 1) / TRK * TRK;
    pl.chunk_rows = cr;
    pl.chunks = (int)((m + cr - 1) / cr);
    return pl;
}
inline bool tn_thin_shape(int64_t m, int64_t n, int64_t N) { return N <= THIN_MAX_N && m >= 16384 && n >= 64; }

size_t gemm_tn_workspace_bytes(int64_t m, int64_t n, int64_t N, int sms) {
    const TNPlan pl = plan_tn(m, n, N, sms);
    size_t bytes = pl.chunks > 1 ? (size_t)pl.chunks * (size_t)pl.pstride * sizeof(double) : 0;
    if (tn_thin_shape(m, n, N)) bytes = std::max(bytes, (size_t)plan_tn_thin(m, n, sms).chunks * (size_t)(n * N) * sizeof(double));
    return bytes;
}

cudaError_t gemm_tn(const GemmTN& p, double* workspace, size_t workspace_bytes, int sms, cudaStream_t st) {
    if (p.n <= 0 || p.N <= 0) return cudaSuccess;
    if (tn_thin_shape(p.m, p.n, p.N) && aligned16(p.A) && (p.lda % 2 == 0)) {
        const TNThinPlan tp = plan_tn_thin(p.m, p.n, sms);
        const int64_t pstride = p.n * p.N;
        if (workspace_bytes < (size_t)tp.chunks * (size_t)pstride * sizeof(double)) return cudaErrorInvalidValue;
        static bool attr_set = false;
        if (!attr_set) {
            cudaError_t e = cudaFuncSetAttribute(gemm_tn_thin_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, TN_THIN_SMEM);
            if (e == cudaSuccess) e = cudaFuncSetAttribute(gemm_tn_thin_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, TN_THIN_SMEM);
            if (e == cudaSuccess) e = cudaFuncSetAttribute(gemm_tn_thin_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, TN_THIN_SMEM);
            if (e != cudaSuccess) return e;
            attr_set = true;
        }
        const unsigned grid = (unsigned)(tp.ncg * tp.chunks);
        const int NI = (int)((p.N + 7) / 8);
        switch (NI) {
            case 1: gemm_tn_thin_kernel<1><<<grid, THIN_THREADS, TN_THIN_SMEM, st>>>(p, workspace, p.n, pstride, tp.ncg, tp.chunk_rows); break;
            case 2: gemm_tn_thin_kernel<2><<<grid, THIN_THREADS, TN_THIN_SMEM, st>>>(p, workspace, p.n, pstride, tp.ncg, tp.chunk_rows); break;
            default: gemm_tn_thin_kernel<3><<<grid, THIN_THREADS, TN_THIN_SMEM, st>>>(p, workspace, p.n, pstride, tp.ncg, tp.chunk_rows); break;
        }
        ++g_kernel_launches;
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return e;
        const int64_t total = p.n * p.N;
        int blocks = (int)((total + 255) / 256);
        if (blocks > sms * 8) blocks = sms * 8;
        tn_reduce_kernel<<<blocks, 256, 0, st>>>(workspace, p.n, pstride, tp.chunks, p.Z, p.ldz, p.n, p.N, p.accumulate);
        ++g_kernel_launches;
        return cudaGetLastError();
    }
    const TNPlan pl = plan_tn(p.m, p.n, p.N, sms);
    if (pl.chunks > 1 && workspace_bytes < (size_t)pl.chunks * (size_t)pl.pstride * sizeof(double))
        return cudaErrorInvalidValue;
    cudaError_t e;
    switch (pl.NT) {
        case 1: e = launch_tn<1>(p, pl, workspace, st); break;
        case 2: e = launch_tn<2>(p, pl, workspace, st); break;
        case 3: e = launch_tn<3>(p, pl, workspace, st); break;
        case 4: e = launch_tn<4>(p, pl, workspace, st); break;
        case 5: e = launch_tn<5>(p, pl, workspace, st); break;
        case 6: e = launch_tn<6>(p, pl, workspace, st); break;
        case 7: e = launch_tn<7>(p, pl, workspace, st); break;
        default: e = launch_tn<8>(p, pl, workspace, st); break;
    }
    if (e != cudaSuccess) return e;
    if (pl.chunks > 1) {
        const int64_t total = p.n * p.N;
        if (pl.chunks >= 16 && total <= (int64_t)sms * 8 * 32 * 4) {
            tn_reduce_grouped_kernel<<<(unsigned)((total + 31) / 32), 256, 0, st>>>(workspace, pl.ldp, pl.pstride, pl.chunks, p.Z, p.ldz, p.n, p.N, p.accumulate);
        } else {
            int blocks = (int)((total + 255) / 256);
            if (blocks > sms * 8) blocks = sms * 8;
            tn_reduce_kernel<<<blocks, 256, 0, st>>>(workspace, pl.ldp, pl.pstride, pl.chunks, p.Z, p.ldz, p.n, p.N, p.accumulate);
        }
        ++g_kernel_launches;
        e = cudaGetLastError();
    }
    return e;
}

}  // namespace rnla
