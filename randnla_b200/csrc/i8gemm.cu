// Range-finder passes on the INT8 tensor cores (tcgen05.mma kind::i8, accumulators in TMEM): an Ozaki-style fixed-point
// splitting of the FP64 operands, for the passes of the power iteration whose result only has to span the right subspace
// (Y = A Omega, S = A^T Y, Y = A S; reference src/lora_helpers.rs:71-95 / :41).  The pass that determines the singular
// values, B = Q^T A (src/lora_helpers.rs:21), stays on the FP64 DMMA kernels (gemm.cu).
//
//   a_ij = 2^{e_i} * sum_{t<4} d_t(i,j) 2^{-7(t+1)} + O(2^{e_i - 29}),   d_t in [-64, 64] (balanced digits)   (e_i: exponent of the row maximum)
//
// A is split ONCE per driver call into four int8 digit planes, stored pre-tiled as the exact shared-memory images the MMA
// descriptors expect (8 x 16-byte core matrices, no swizzle), in two arrangements: row-block major for A S (contraction over
// columns) and column-block major for A^T Y (contraction over rows; the row scale 2^{e_i} is folded into Y before Y is split).
// A stage of either pass is then ONE contiguous 32 KB bulk copy per operand (cp.async.bulk), 4 bytes of HBM traffic per element
// of A instead of 8, and ten 128 x 128 x 32 integer MMAs per 32 columns: digit pairs (ta, tb) with ta + tb = g accumulate
// exactly in int32 into accumulator g (4 x 128 TMEM columns = all 512), and the epilogue forms sum_g D_g 2^{-7(g+2)} exactly
// in FP64.  Products with ta + tb >= 4 are dropped: relative accuracy 2^-28 of (row max) x (column max), enough for a basis.
//
// Accumulation bound: 4 pairs x 64^2 x K < 2^31  =>  K <= 131072 per accumulation (A S needs n <= 131072; A^T Y is chunked).
#include "drivers.cuh"
#include "gemm.cuh"
#include "panel.cuh"
#include "ptx.cuh"
#include <algorithm>
#include <cmath>
#include <vector>

namespace rnla {

namespace {

constexpr int PL = 4;                       // digit planes
constexpr int BM = 128, BN = 128, BK = 64;  // CTA tile: 128 x 128 outputs, 64 contraction indices per stage
constexpr int CHUNK = PL * BM * BK;         // 32 KB: one stage of one operand (all planes)
constexpr int PLANE = BM * BK;              // 8 KB
constexpr int PLH = 3;                      // extra digit planes of the 49-bit split
constexpr int CHUNK_HI = PLH * PLANE;       // 24 KB
constexpr int STAGES = 3;
constexpr int MMA_THREADS = 192;            // warp 0 producer, warp 1 MMA issuer, warps 2-5 epilogue
constexpr size_t MMA_SMEM = (size_t)STAGES * 2 * CHUNK + 1024 + 256;
constexpr int STAGE_HI = 2 * CHUNK + 2 * CHUNK_HI;                          // 112 KB: both operands, all seven planes
constexpr size_t MMA_SMEM_HI = (size_t)2 * STAGE_HI + 1024 + 256;
constexpr int64_t K_ACC_MAX = 131072;      // 4 pairs x 64^2 x K < 2^31

// ---------------------------------------------------------------------------------------------- tcgen05 wrappers
__device__ __forceinline__ void tc_alloc(uint32_t* smem_dst, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tc_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc], int8 x int8 -> int32
__device__ __forceinline__ void tc_mma_i8(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tc_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// shared-memory matrix descriptor, no swizzle (cute::UMMA::SmemDescriptor: start >> 4 | LBO >> 4 << 16 | SBO >> 4 << 32 | version 1 << 46)
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((addr >> 4) & 0x3fffu) | ((uint64_t)((lbo_bytes >> 4) & 0x3fffu) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3fffu) << 32) | (1ull << 46);
}
// instruction descriptor (cute::UMMA::InstrDescriptor): D = S32, A = B = signed int8, M = 128, N = 128
// N = nmma (a multiple of 16, <= 128): only the thin operand's columns that exist are multiplied
__host__ __device__ constexpr uint32_t instr_desc(bool a_mn_major, bool b_mn_major, int nmma) {
    return (2u << 4) | (1u << 7) | (1u << 10) | ((a_mn_major ? 1u : 0u) << 15) | ((b_mn_major ? 1u : 0u) << 16) |
           ((uint32_t)(nmma >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
}

// ---------------------------------------------------------------------------------------------- splitting
// Balanced 7-bit digits without a single conversion instruction: y + 1.5 * 2^52 holds rint(y) in the low mantissa bits
// (|y| < 2^27), subtracting the magic number again gives rint(y) as a double, and the low 7 bits of an integer, sign-extended,
// are its balanced digit in [-64, 63] (v - d is then divisible by 128).
__device__ __forceinline__ int sext7(int v) { return (v << 25) >> 25; }
__device__ __forceinline__ int rint_bits(double y, double* r) {
    const double MAGIC = 6755399441055744.0;                // 1.5 * 2^52
    const double t = y + MAGIC;
    *r = t - MAGIC;
    return (int)(unsigned)__double_as_longlong(t);          // low 32 bits: two's complement of rint(y)
}
// x * scale (|.| < 2^27) -> v = rint -> v = d0 2^21 + d1 2^14 + d2 2^7 + d3, d1..d3 in [-64, 63], |d0| <= 64
__device__ __forceinline__ void digits4(double x, double scale, int (&d)[4]) {
    double r;
    int v = rint_bits(x * scale, &r);
    d[3] = sext7(v); v = (v - d[3]) >> 7;
    d[2] = sext7(v); v = (v - d[2]) >> 7;
    d[1] = sext7(v); v = (v - d[1]) >> 7;
    d[0] = v;
}
// seven digits of a 49-bit fixed-point value, in 32-bit arithmetic: vh = rint(y) gives the four leading digits exactly as
// digits4 does, the remainder y - vh (exact in FP64, |.| <= 1/2) times 2^21 the three trailing ones:
// x * scale * 2^21 = vh 2^21 + d4 2^14 + d5 2^7 + d6 (+- 1/2)
__device__ __forceinline__ void digits7(double x, double scale, int (&d)[7]) {
    const double y = x * scale;
    double yr, r2;
    int vh = rint_bits(y, &yr);
    int vl = rint_bits((y - yr) * 2097152.0, &r2);
    d[6] = sext7(vl); vl = (vl - d[6]) >> 7;
    d[5] = sext7(vl); vl = (vl - d[5]) >> 7;
    d[4] = vl;
    d[3] = sext7(vh); vh = (vh - d[3]) >> 7;
    d[2] = sext7(vh); vh = (vh - d[2]) >> 7;
    d[1] = sext7(vh); vh = (vh - d[1]) >> 7;
    d[0] = vh;
}
// exponent bookkeeping from the bit pattern of a maximum: up = 2^(e+1) with max < 2^e, down = 2^(28 - (e+1)), so that
// |x * down| < 2^27; zero / tiny rows -> 0
__device__ __forceinline__ void scales_from_max_bits(unsigned long long bits, double* up, double* down) {
    const int E = (int)(bits >> 52) & 0x7ff;
    if (E < 64 || E >= 2045) { *up = 0.0; *down = 0.0; return; }
    *up = __longlong_as_double((long long)(E + 2) << 52);
    *down = __longlong_as_double((long long)(2072 - E) << 52);
}

__global__ void __launch_bounds__(256)
rowmax_kernel(const double* __restrict__ A, int64_t lda, int64_t m, int64_t n, int64_t cols_per, unsigned long long* __restrict__ bits) {
    const int64_t i = ((int64_t)blockIdx.x * 256 + threadIdx.x) * 2;
    if (i >= m) return;
    const int64_t j0 = (int64_t)blockIdx.y * cols_per, j1 = min(n, j0 + cols_per);
    const bool two = i + 1 < m;
    const bool vec = two && ((reinterpret_cast<uintptr_t>(A) & 15) == 0) && (lda % 2 == 0);
    double m0 = 0.0, m1 = 0.0;
    if (vec) {
#pragma unroll 8
        for (int64_t j = j0; j < j1; ++j) {
            double2 v;
            asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(A + i + j * lda));
            m0 = fmax(m0, fabs(v.x)); m1 = fmax(m1, fabs(v.y));
        }
    } else {
        for (int64_t j = j0; j < j1; ++j) {
            m0 = fmax(m0, fabs(ldg_stream(A + i + j * lda)));
            if (two) m1 = fmax(m1, fabs(ldg_stream(A + i + 1 + j * lda)));
        }
    }
    atomicMax(bits + i, (unsigned long long)__double_as_longlong(m0));
    if (two) atomicMax(bits + i + 1, (unsigned long long)__double_as_longlong(m1));
}
__global__ void scales_kernel(const unsigned long long* __restrict__ bits, int64_t cnt, double* __restrict__ up, double* __restrict__ down) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < cnt) scales_from_max_bits(bits[i], up + i, down + i);
}

// The digit split of A.  It writes row-block-major images (NN: [plane][J 8][I 8][8 x 16 B] per 128 rows x 64 columns) and
// column-block-major images (TN: [plane][J 16][I 4][8 x 16 B] per 128 columns x 64 rows).  A core matrix holds 16 rows x 8
// columns of A as 8 rows (columns of A) of 16 bytes (rows of A): the same 128 bytes serve as an MN-major core matrix of A S and
// as a K-major one of A^T Y.
// staging-buffer swizzle (the 16-byte row inside a core matrix is XORed with the index of the core matrix): lanes that write
// the same row of neighbouring core matrices hit different banks; undone by the copy-out
__device__ __forceinline__ int stage_swz(int off) { return off ^ ((((off >> 7) ^ (off >> 12)) & 7) << 4); }
__device__ __forceinline__ unsigned pack4(int a, int b, int c, int d) {
    return __byte_perm(__byte_perm((unsigned)a, (unsigned)b, 0x0040), __byte_perm((unsigned)c, (unsigned)d, 0x0040), 0x5410);
}

// P7: seven digits per element; planes 4..6 go to a third image (column-block major only: [plane 3][J 16][I 4][8 x 16 B]).
// One CTA per 128 rows x 32 columns (half of a 64-column image block): 16 elements per thread keep the register count low
// enough for three or four resident CTAs per SM, which is what keeps loads in flight while other CTAs form digits and store.
constexpr int SL_COLS = 32;
constexpr int SL_NN = PL * 128 * SL_COLS;          // 16 KB: [plane][J 4][I 8][128 B]
constexpr int SL_TN = PL * 128 * SL_COLS;          // 16 KB: [half][plane][J 4][I 4][128 B]
constexpr int SL_HI = PLH * 128 * SL_COLS;         // 12 KB: [half][plane 3][J 4][I 4][128 B]
template <bool P7>
__global__ void __launch_bounds__(256, P7 ? 3 : 4)
slice_a_kernel(const double* __restrict__ A, int64_t lda, int64_t m, int64_t n, const double* __restrict__ down,
               uint8_t* __restrict__ nn, int64_t kb_total, uint8_t* __restrict__ tn, int64_t kr_total, uint8_t* __restrict__ tnhi,
               int64_t rb0 /* first 128-row block of this launch */) {
    extern __shared__ __align__(16) uint8_t img[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t hb_total = 2 * kb_total;
    const int64_t rb = rb0 + blockIdx.x / hb_total, hb = blockIdx.x % hb_total;
    const int64_t kb = hb >> 1;
    const int half = (int)(hb & 1);                         // which 32 columns of the 64-column image block
    const int64_t R0 = rb * 128, C0 = hb * SL_COLS;
    const int il = 4 * lane;                                // local rows il .. il + 3: a warp covers the 128 rows of one column
    const int64_t i = R0 + il;
    const bool vec = ((reinterpret_cast<uintptr_t>(A) & 15) == 0) && (lda % 2 == 0) && i + 3 < m;
    double x[4][4], sc[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) sc[e] = i + e < m ? down[i + e] : 0.0;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const int64_t j = C0 + warp + 8 * r;
#pragma unroll
        for (int e = 0; e < 4; ++e) x[r][e] = 0.0;
        if (j < n) {
            if (vec) {
                asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0,%1}, [%2];" : "=d"(x[r][0]), "=d"(x[r][1]) : "l"(A + i + j * lda));
                asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0,%1}, [%2];" : "=d"(x[r][2]), "=d"(x[r][3]) : "l"(A + i + 2 + j * lda));
            } else {
#pragma unroll
                for (int e = 0; e < 4; ++e) if (i + e < m) x[r][e] = ldg_stream(A + i + e + j * lda);
            }
        }
    }
    const int h = il >> 6, i64 = il & 63;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const int jl = warp + 8 * r;                        // 0 .. 31
        constexpr int ND = P7 ? 7 : 4;
        int d[4][ND];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            if (P7) digits7(x[r][e], sc[e], reinterpret_cast<int (&)[7]>(d[e]));
            else digits4(x[r][e], sc[e], reinterpret_cast<int (&)[4]>(d[e]));
        }
        const int intra_nn = (jl >> 3) * 1024 + (il >> 4) * 128 + (jl & 7) * 16 + (il & 15);
        const int intra_tn = (jl >> 3) * 512 + (i64 >> 4) * 128 + (jl & 7) * 16 + (i64 & 15);
#pragma unroll
        for (int t = 0; t < ND; ++t) {
            const unsigned word = pack4(d[0][t], d[1][t], d[2][t], d[3][t]);
            if (t < PL) {
                *reinterpret_cast<unsigned*>(img + stage_swz(t * (SL_NN / PL) + intra_nn)) = word;
                *reinterpret_cast<unsigned*>(img + stage_swz(SL_NN + h * (SL_TN / 2) + t * (SL_TN / 2 / PL) + intra_tn)) = word;
            } else {
                *reinterpret_cast<unsigned*>(img + stage_swz(SL_NN + SL_TN + h * (SL_HI / 2) + (t - PL) * (SL_HI / 2 / PLH) + intra_tn)) = word;
            }
        }
    }
    __syncthreads();
    const uint4* src = reinterpret_cast<const uint4*>(img);
    // NN: block (rb, kb), column groups J = 4 half .. + 3 of every plane: 4 KB per plane
    for (int q = threadIdx.x; q < SL_NN / 16; q += 256) {
        const int t = q >> 8, w = q & 255;
        uint4* dst = reinterpret_cast<uint4*>(nn + (rb * kb_total + kb) * (int64_t)CHUNK + t * PLANE + half * (PLANE / 2));
        dst[w] = src[stage_swz(q * 16) >> 4];
    }
    // TN: block (cb = kb / 2, kr = 2 rb + hh), column groups J = 8 (kb & 1) + 4 half .. + 3 of every plane: 2 KB per plane
    for (int q = threadIdx.x; q < SL_TN / 16; q += 256) {
        const int hh = q >> 9, t = (q >> 7) & 3, w = q & 127;              // 512 uint4 per half, 128 per plane piece
        uint4* dst = reinterpret_cast<uint4*>(tn + ((kb >> 1) * kr_total + 2 * rb + hh) * (int64_t)CHUNK + t * PLANE +
                                              ((kb & 1) * 8 + half * 4) * 512);
        dst[w] = src[stage_swz(SL_NN + q * 16) >> 4];
    }
    if (P7) {
        for (int q = threadIdx.x; q < SL_HI / 16; q += 256) {
            const int hh = q / 384, t = (q % 384) >> 7, w = q & 127;       // 384 uint4 per half, 128 per plane piece
            uint4* dst = reinterpret_cast<uint4*>(tnhi + ((kb >> 1) * kr_total + 2 * rb + hh) * (int64_t)CHUNK_HI + t * PLANE +
                                                  ((kb & 1) * 8 + half * 4) * 512);
            dst[w] = src[stage_swz(SL_NN + SL_TN + q * 16) >> 4];
        }
    }
}

// per-column maxima of X (K x N), optionally with the row scale folded in (X(k, c) * rs[k])
__global__ void __launch_bounds__(256)
colmax_kernel(const double* __restrict__ X, int64_t ldx, int64_t K, int N, const double* __restrict__ rs, int64_t rows_per,
              unsigned long long* __restrict__ bits) {
    __shared__ double red[8];
    const int c = blockIdx.x;
    const int64_t k0 = (int64_t)blockIdx.y * rows_per, k1 = min(K, k0 + rows_per);
    double mx = 0.0;
    for (int64_t k = k0 + threadIdx.x; k < k1; k += 256) mx = fmax(mx, fabs(X[k + (int64_t)c * ldx] * (rs ? rs[k] : 1.0)));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; ++w) mx = fmax(mx, red[w]);
        atomicMax(bits + c, (unsigned long long)__double_as_longlong(mx));
    }
}
// B operand images: [k block of 64][plane][k group of 8][n block of 16][8 x 16 B]; columns >= N and rows >= K are zero
template <bool P7>
__global__ void __launch_bounds__(256)
slice_b_kernel(const double* __restrict__ X, int64_t ldx, int64_t K, int N, const double* __restrict__ rs, const double* __restrict__ cdown,
               uint8_t* __restrict__ out, uint8_t* __restrict__ out_hi) {
    const int64_t kb = blockIdx.x;
    const int kl = threadIdx.x & 63;
    const int64_t k = kb * 64 + kl;
    const double rsk = (k < K) ? (rs ? rs[k] : 1.0) : 0.0;
    for (int cq = threadIdx.x >> 6; cq < 32; cq += 4) {
        const int c0 = 4 * cq;
        constexpr int ND = P7 ? 7 : 4;
        unsigned w[ND];
#pragma unroll
        for (int t = 0; t < ND; ++t) w[t] = 0u;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int c = c0 + e;
            int d[ND];
#pragma unroll
            for (int t = 0; t < ND; ++t) d[t] = 0;
            if (k < K && c < N) {
                if (P7) digits7(X[k + (int64_t)c * ldx] * rsk, cdown[c], reinterpret_cast<int (&)[7]>(d));
                else digits4(X[k + (int64_t)c * ldx] * rsk, cdown[c], reinterpret_cast<int (&)[4]>(d));
            }
#pragma unroll
            for (int t = 0; t < ND; ++t) w[t] |= ((unsigned)d[t] & 0xffu) << (8 * e);
        }
        const int intra = (kl >> 3) * 1024 + (c0 >> 4) * 128 + (kl & 7) * 16 + (c0 & 15);
#pragma unroll
        for (int t = 0; t < ND; ++t) {
            if (t < PL) *reinterpret_cast<unsigned*>(out + kb * (int64_t)CHUNK + t * PLANE + intra) = w[t];
            else *reinterpret_cast<unsigned*>(out_hi + kb * (int64_t)CHUNK_HI + (t - PL) * PLANE + intra) = w[t];
        }
    }
}

// ---------------------------------------------------------------------------------------------- the MMA kernel
// TN = false:  C(rows rb*128.., :) = rs_up(i) cs_up(c) * sum_g 2^{-7(g+2)} D_g,   contraction over all kb blocks of 64 columns
// TN = true :  P[chunk](cols cb*128.., :) = sum_g 2^{-7(g+2)} D_g,                contraction over this chunk's blocks of 64 rows
// HI = true: second sweep of A S, digit pairs with ta + tb in {4, 5, 6} into accumulators 0..2, ADDED to C: together with the
// first sweep all 16 pairs, i.e. the exact product of the two 28-bit representations.
template <bool TN, bool HI>
__global__ void __launch_bounds__(MMA_THREADS, 1)
i8_mma_kernel(const uint8_t* __restrict__ Aimg, int64_t a_blocks_per_tile, const uint8_t* __restrict__ Bimg, int64_t kblocks_total,
              int64_t kblocks_per_chunk, double* __restrict__ C, int64_t ldc, int64_t rows, int ncols, const double* __restrict__ rs_up,
              const double* __restrict__ cs_up, int64_t chunk_stride) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + (size_t)STAGES * 2 * CHUNK);
    uint64_t* empty = full + STAGES;
    uint64_t* accum = empty + STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum + 1);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t tile = blockIdx.x;
    const int64_t kb0 = (int64_t)blockIdx.y * kblocks_per_chunk;
    const int64_t kb1 = min(kblocks_total, kb0 + kblocks_per_chunk);
    const int nk = (int)(kb1 - kb0);

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, 1); }
        mbar_init(accum, 1);
        mbar_fence_init();
    }
    if (warp == 1) tc_alloc(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            const uint8_t* a_src = Aimg + (tile * a_blocks_per_tile + kb0) * (int64_t)CHUNK;
            const uint8_t* b_src = Bimg + kb0 * (int64_t)CHUNK;
            for (int it = 0; it < nk; ++it) {
                const int s = it % STAGES;
                if (it >= STAGES) mbar_wait(empty + s, ((it / STAGES) - 1) & 1);
                mbar_arrive_expect_tx(full + s, 2 * CHUNK);
                bulk_g2s(smem + (size_t)s * 2 * CHUNK, a_src + (int64_t)it * CHUNK, CHUNK, full + s);
                bulk_g2s(smem + (size_t)s * 2 * CHUNK + CHUNK, b_src + (int64_t)it * CHUNK, CHUNK, full + s);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc = instr_desc(!TN, true, (ncols + 15) & ~15);
            for (int it = 0; it < nk; ++it) {
                const int s = it % STAGES;
                mbar_wait(full + s, (it / STAGES) & 1);
                tc_fence_after();
                const uint32_t a_base = smem_u32(smem + (size_t)s * 2 * CHUNK);
                const uint32_t b_base = a_base + CHUNK;
#pragma unroll
                for (int ks = 0; ks < BK / 32; ++ks) {
#pragma unroll
                    for (int ta = 0; ta < PL; ++ta) {
#pragma unroll
                        for (int tb = 0; tb < PL; ++tb) {
                            constexpr int G0 = HI ? 4 : 0;
                            const int g = ta + tb;
                            if (g < G0 || g >= G0 + 4) continue;
                            // A S: MN-major core matrices, 16-row blocks 128 B apart (SBO), 8-column groups 1024 B apart (LBO), 32 columns = 4096 B
                            // A^T Y: K-major, 8-column blocks 512 B apart (SBO), the two 16-row halves 128 B apart (LBO), 32 rows = 256 B
                            const uint64_t ad = TN ? smem_desc(a_base + ta * PLANE + ks * 256, 128, 512)
                                                   : smem_desc(a_base + ta * PLANE + ks * 4096, 1024, 128);
                            const uint64_t bd = smem_desc(b_base + tb * PLANE + ks * 4096, 1024, 128);
                            const int ta_first = g > PL - 1 ? g - (PL - 1) : 0;                // first pair of group g in this loop order
                            const uint32_t acc = (it > 0 || ks > 0 || ta > ta_first) ? 1u : 0u;
                            tc_mma_i8(tmem + (uint32_t)(g - G0) * BN, ad, bd, idesc, acc);
                        }
                    }
                }
                tc_commit(empty + s);
            }
            tc_commit(accum);
        }
    } else {
        mbar_wait(accum, 0);
        tc_fence_after();
        const int quad = warp & 3;                          // TMEM lane quadrant this warp may read
        const int64_t r = tile * BM + quad * 32 + lane;     // output row (A S) / output row = column of A (A^T Y)
        const double rsc = (!TN && r < rows) ? rs_up[r] : 1.0;
        double* out = C + (TN ? (int64_t)blockIdx.y * chunk_stride : 0);
        for (int c0 = 0; c0 < ncols; c0 += 16) {
            uint32_t d[PL][16];
#pragma unroll
            for (int g = 0; g < (HI ? 3 : PL); ++g) tc_ld16(tmem + ((uint32_t)(quad * 32) << 16) + (uint32_t)(g * BN + c0), d[g]);
            tc_wait_ld();
            if (r < rows) {
#pragma unroll
                for (int e = 0; e < 16; ++e) {
                    const int c = c0 + e;
                    if (c < ncols) {
                        // exact in FP64: |D_g| < 2^31 and the four terms span 21 more bits
                        double v;
                        if (HI) {
                            v = (double)(int)d[2][e];
                            v = v * 0.0078125 + (double)(int)d[1][e];
                            v = v * 0.0078125 + (double)(int)d[0][e];
                            v *= 2.2737367544323206e-13;        // 2^-42: groups 4, 5, 6
                            out[r + (int64_t)c * ldc] += v * (rsc * cs_up[c]);
                        } else {
                            v = (double)(int)d[3][e];
                            v = v * 0.0078125 + (double)(int)d[2][e];
                            v = v * 0.0078125 + (double)(int)d[1][e];
                            v = v * 0.0078125 + (double)(int)d[0][e];
                            v *= 6.103515625e-05;               // 2^-14
                            if (!TN) v *= rsc * cs_up[c];
                            out[r + (int64_t)c * ldc] = v;
                        }
                    }
                }
            }
        }
        tc_fence_before();
    }
    __syncthreads();
    if (warp == 1) { tc_fence_after(); tc_dealloc(tmem, 512); }
}

// Second sweep of the 49-bit A^T Q: digit pairs with ta + tb in {4, 5, 6} (18 of them, all seven planes of both operands) into
// accumulators 0..2; P2[chunk] = sum_g 2^{-7(g+2)} D_g.  A stage holds both operands' seven planes (112 KB), two stages.
__global__ void __launch_bounds__(MMA_THREADS, 1)
i8_mma_tn_hi_kernel(const uint8_t* __restrict__ Alo, const uint8_t* __restrict__ Ahi, const uint8_t* __restrict__ Blo,
                    const uint8_t* __restrict__ Bhi, int64_t kr_total, int64_t kblocks_per_chunk, double* __restrict__ P2, int64_t ldc,
                    int64_t rows, int ncols, int64_t chunk_stride) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + (size_t)2 * STAGE_HI);
    uint64_t* empty = full + 2;
    uint64_t* accum = empty + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum + 1);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t tile = blockIdx.x;
    const int64_t kb0 = (int64_t)blockIdx.y * kblocks_per_chunk;
    const int64_t kb1 = min(kr_total, kb0 + kblocks_per_chunk);
    const int nk = (int)(kb1 - kb0);
    if (threadIdx.x == 0) {
        for (int s = 0; s < 2; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, 1); }
        mbar_init(accum, 1);
        mbar_fence_init();
    }
    if (warp == 1) tc_alloc(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    if (warp == 0) {
        if (lane == 0) {
            for (int it = 0; it < nk; ++it) {
                const int s = it & 1;
                if (it >= 2) mbar_wait(empty + s, ((it >> 1) - 1) & 1);
                mbar_arrive_expect_tx(full + s, STAGE_HI);
                uint8_t* st = smem + (size_t)s * STAGE_HI;
                bulk_g2s(st, Alo + (tile * kr_total + kb0 + it) * (int64_t)CHUNK, CHUNK, full + s);
                bulk_g2s(st + CHUNK, Blo + (kb0 + it) * (int64_t)CHUNK, CHUNK, full + s);
                bulk_g2s(st + 2 * CHUNK, Ahi + (tile * kr_total + kb0 + it) * (int64_t)CHUNK_HI, CHUNK_HI, full + s);
                bulk_g2s(st + 2 * CHUNK + CHUNK_HI, Bhi + (kb0 + it) * (int64_t)CHUNK_HI, CHUNK_HI, full + s);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc = instr_desc(false, true, (ncols + 15) & ~15);
            for (int it = 0; it < nk; ++it) {
                const int s = it & 1;
                mbar_wait(full + s, (it >> 1) & 1);
                tc_fence_after();
                const uint32_t base = smem_u32(smem + (size_t)s * STAGE_HI);
#pragma unroll
                for (int ks = 0; ks < BK / 32; ++ks) {
#pragma unroll
                    for (int ta = 0; ta < PL + PLH; ++ta) {
#pragma unroll
                        for (int tb = 0; tb < PL + PLH; ++tb) {
                            const int g = ta + tb;
                            if (g < 4 || g > 6) continue;
                            const uint32_t a_plane = ta < PL ? base + ta * PLANE : base + 2 * CHUNK + (ta - PL) * PLANE;
                            const uint32_t b_plane = tb < PL ? base + CHUNK + tb * PLANE : base + 2 * CHUNK + CHUNK_HI + (tb - PL) * PLANE;
                            const uint64_t ad = smem_desc(a_plane + ks * 256, 128, 512);
                            const uint64_t bd = smem_desc(b_plane + ks * 4096, 1024, 128);
                            const uint32_t acc = (it > 0 || ks > 0 || ta > 0) ? 1u : 0u;      // first pair of group g is (0, g)
                            tc_mma_i8(tmem + (uint32_t)(g - 4) * BN, ad, bd, idesc, acc);
                        }
                    }
                }
                tc_commit(empty + s);
            }
            tc_commit(accum);
        }
    } else {
        mbar_wait(accum, 0);
        tc_fence_after();
        const int quad = warp & 3;
        const int64_t r = tile * BM + quad * 32 + lane;
        double* out = P2 + (int64_t)blockIdx.y * chunk_stride;
        for (int c0 = 0; c0 < ncols; c0 += 16) {
            uint32_t d[3][16];
#pragma unroll
            for (int g = 0; g < 3; ++g) tc_ld16(tmem + ((uint32_t)(quad * 32) << 16) + (uint32_t)(g * BN + c0), d[g]);
            tc_wait_ld();
            if (r < rows) {
#pragma unroll
                for (int e = 0; e < 16; ++e) {
                    const int c = c0 + e;
                    if (c < ncols) {
                        double v = (double)(int)d[2][e];
                        v = v * 0.0078125 + (double)(int)d[1][e];
                        v = v * 0.0078125 + (double)(int)d[0][e];
                        out[r + (int64_t)c * ldc] = v * 2.2737367544323206e-13;       // 2^-42
                    }
                }
            }
        }
        tc_fence_before();
    }
    __syncthreads();
    if (warp == 1) { tc_fence_after(); tc_dealloc(tmem, 512); }
}

// Z(j, c) = cs_up(c) * sum over chunks, fixed order
__global__ void __launch_bounds__(256)
i8_tn_reduce_kernel(const double* __restrict__ P, const double* __restrict__ P2, int nchunks, int64_t chunk_stride, int64_t n, int ncols,
                    const double* __restrict__ cs_up, double* __restrict__ Z, int64_t ldz) {
    const int64_t total = n * ncols;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
        const int64_t c = idx / n, j = idx - c * n;
        double s = 0.0;
        for (int k = 0; k < nchunks; ++k) s += P[(int64_t)k * chunk_stride + j + c * n];
        if (P2) {
            double s2 = 0.0;                  // the low-order digit pairs, summed separately and added last
            for (int k = 0; k < nchunks; ++k) s2 += P2[(int64_t)k * chunk_stride + j + c * n];
            s += s2;
        }
        Z[j + c * ldz] = s * cs_up[c];
    }
}

// The digit-plane images are tens of GB: they live in a workspace that persists across driver calls (grown on demand, released
// by rnla_release_workspace / rnla_shutdown) instead of going through the stream-ordered pool on every call -- re-carving 44 GB
// out of the pool per call cost 60 ms to 1 s of host time (measured), more than all the passes together.
struct Persist {
    void* p = nullptr; size_t cap = 0;
    cudaError_t ensure(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) { cudaError_t e = cudaFree(p); p = nullptr; cap = 0; if (e != cudaSuccess) return e; }
        cudaError_t e = cudaMalloc(&p, bytes);
        if (e == cudaSuccess) cap = bytes;
        return e;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
    double* d() const { return static_cast<double*>(p); }
    template <class T> T* as() const { return static_cast<T*>(p); }
};
struct Sliced {
    const double* A = nullptr; int64_t lda = 0, m = 0, n = 0;
    int64_t rblocks = 0, cblocks = 0, kb_total = 0, kr_total = 0;
    Persist nn, tn, tnhi, up, down, bits, bimg, bimg_hi, cbits, cup, cdown, part, part2;
    bool ready = false, p7 = false;
};
Sliced g_sl;
bool g_active = false;
bool g_precise = false;       // A S with all 16 digit pairs (two sweeps)
bool g_full = false;          // A^T Q on the 49-bit split (two sweeps, 28 digit pairs)

rnla_status slice_b(const double* X, int64_t ldx, int64_t K, int N, const double* rs, int64_t kblocks, bool p7 = false) {
    Ctx& c = ctx();
    RNLA_CUDA(cudaMemsetAsync(g_sl.cbits.p, 0, 128 * 8, c.stream));
    const int64_t rows_per = 32768;
    colmax_kernel<<<dim3((unsigned)N, (unsigned)((K + rows_per - 1) / rows_per)), 256, 0, c.stream>>>(X, ldx, K, N, rs, rows_per,
                                                                                                      g_sl.cbits.as<unsigned long long>());
    scales_kernel<<<1, 128, 0, c.stream>>>(g_sl.cbits.as<unsigned long long>(), 128, g_sl.cup.d(), g_sl.cdown.d());
    if (p7) slice_b_kernel<true><<<(unsigned)kblocks, 256, 0, c.stream>>>(X, ldx, K, N, rs, g_sl.cdown.d(), g_sl.bimg.as<uint8_t>(), g_sl.bimg_hi.as<uint8_t>());
    else slice_b_kernel<false><<<(unsigned)kblocks, 256, 0, c.stream>>>(X, ldx, K, N, rs, g_sl.cdown.d(), g_sl.bimg.as<uint8_t>(), nullptr);
    g_kernel_launches += 3;
    RNLA_CUDA(cudaGetLastError());
    return RNLA_OK;
}

}  // namespace

// thin operands wider than one 128-column MMA tile are processed tile by tile (each tile is another sweep over the planes of A)
constexpr int I8_MAX_N = 256;
bool i8_supported(int64_t m, int64_t n, int l) {
    return l >= 1 && l <= I8_MAX_N && n >= 1 && m >= 1 && ((n + 127) / 128) * 128 <= K_ACC_MAX && m * n >= (int64_t)1 << 22;
}
bool i8_active_for(const double* A, int64_t lda, int64_t m, int64_t n, int64_t N) {
    return g_active && g_sl.ready && g_sl.A == A && g_sl.lda == lda && g_sl.m == m && g_sl.n == n && N <= I8_MAX_N;
}
void i8_deactivate() { g_active = false; g_precise = false; g_full = false; }
void i8_set_full(bool on) { g_full = on; }
void i8_set_precise(bool on) { g_precise = on; }
void i8_release() { g_active = false; g_sl.ready = false; }      // the workspace persists (see Persist)
void i8_free_workspace() {
    g_active = false; g_sl.ready = false;
    Persist* all[] = {&g_sl.nn, &g_sl.tn, &g_sl.tnhi, &g_sl.up, &g_sl.down, &g_sl.bits, &g_sl.bimg, &g_sl.bimg_hi, &g_sl.cbits, &g_sl.cup, &g_sl.cdown,
                      &g_sl.part, &g_sl.part2};
    for (Persist* b : all) b->release();
}

// split A (m x n, lda) into the two tiled int8 images; afterwards dev_gemm_nn / dev_gemm_tn with this A and N <= 128 run on
// the integer tensor cores until i8_deactivate().  Three steps so that the host-buffer entry point can split each row block of A
// as it lands over PCIe (row maxima are per row: a block of whole rows is self-contained): begin, rows(r0, count)*, end.
rnla_status i8_prepare_begin(const double* A, int64_t lda, int64_t m, int64_t n, bool p7) {
    Ctx& c = ctx();
    Sliced& s = g_sl;
    s.ready = false; g_active = false;
    s.A = A; s.lda = lda; s.m = m; s.n = n; s.p7 = p7;
    s.rblocks = (m + 127) / 128; s.cblocks = (n + 127) / 128;
    s.kb_total = 2 * s.cblocks; s.kr_total = 2 * s.rblocks;
    const size_t img_bytes = (size_t)s.rblocks * s.cblocks * 2 * CHUNK;
    const int64_t kmax = std::max(s.kb_total, s.kr_total);
    if (s.nn.ensure(img_bytes) != cudaSuccess || s.tn.ensure(img_bytes) != cudaSuccess || (p7 && s.tnhi.ensure(img_bytes / PL * PLH) != cudaSuccess)) {
        cudaGetLastError();
        i8_free_workspace();
        return fail(RNLA_ERR_COMPUTATION, "int8 passes: no room for the digit-plane workspace");
    }
    RNLA_CUDA(s.up.ensure((size_t)m * 8)); RNLA_CUDA(s.down.ensure((size_t)m * 8)); RNLA_CUDA(s.bits.ensure((size_t)m * 8));
    RNLA_CUDA(s.bimg.ensure((size_t)kmax * CHUNK));
    if (p7) { RNLA_CUDA(s.tnhi.ensure(img_bytes / PL * PLH)); RNLA_CUDA(s.bimg_hi.ensure((size_t)kmax * CHUNK_HI)); }
    RNLA_CUDA(s.cbits.ensure(128 * 8)); RNLA_CUDA(s.cup.ensure(128 * 8)); RNLA_CUDA(s.cdown.ensure(128 * 8));
    RNLA_CUDA(cudaMemsetAsync(s.bits.p, 0, (size_t)m * 8, c.stream));
    static bool attr = false;
    if (!attr) {
        RNLA_CUDA(cudaFuncSetAttribute(slice_a_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SL_NN + SL_TN));
        RNLA_CUDA(cudaFuncSetAttribute(slice_a_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SL_NN + SL_TN + SL_HI));
        RNLA_CUDA(cudaFuncSetAttribute(i8_mma_tn_hi_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)MMA_SMEM_HI));
        RNLA_CUDA(cudaFuncSetAttribute(i8_mma_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)MMA_SMEM));
        RNLA_CUDA(cudaFuncSetAttribute(i8_mma_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)MMA_SMEM));
        RNLA_CUDA(cudaFuncSetAttribute(i8_mma_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)MMA_SMEM));
        attr = true;
    }
    return RNLA_OK;
}
// rows [r0, r0 + count) of the matrix given to i8_prepare_begin; r0 must be a multiple of 128 (whole image row blocks)
rnla_status i8_prepare_rows(int64_t r0, int64_t count, bool phases) {
    Ctx& c = ctx();
    Sliced& s = g_sl;
    if (count <= 0) return RNLA_OK;
    if (r0 % 128 != 0 || r0 + count > s.m) return fail(RNLA_ERR_COMPUTATION, "int8 split: row block must start on a multiple of 128");
    const double* A = s.A; const int64_t lda = s.lda, m = s.m, n = s.n;
    if (phases) phase_begin("i8:rowmax(A)");
    const int64_t cols_per = std::max<int64_t>(256, (n + 7) / 8);
    rowmax_kernel<<<dim3((unsigned)((count + 511) / 512), (unsigned)((n + cols_per - 1) / cols_per)), 256, 0, c.stream>>>(
        A + r0, lda, count, n, cols_per, s.bits.as<unsigned long long>() + r0);
    scales_kernel<<<(unsigned)((count + 255) / 256), 256, 0, c.stream>>>(s.bits.as<unsigned long long>() + r0, count, s.up.d() + r0, s.down.d() + r0);
    if (phases) { phase_end(); phase_begin("i8:split(A)"); }
    const int64_t rb0 = r0 / 128, nrb = (count + 127) / 128;
    if (s.p7)
        slice_a_kernel<true><<<(unsigned)(nrb * s.kb_total * 2), 256, SL_NN + SL_TN + SL_HI, c.stream>>>(
            A, lda, m, n, s.down.d(), s.nn.as<uint8_t>(), s.kb_total, s.tn.as<uint8_t>(), s.kr_total, s.tnhi.as<uint8_t>(), rb0);
    else
        slice_a_kernel<false><<<(unsigned)(nrb * s.kb_total * 2), 256, SL_NN + SL_TN, c.stream>>>(
            A, lda, m, n, s.down.d(), s.nn.as<uint8_t>(), s.kb_total, s.tn.as<uint8_t>(), s.kr_total, nullptr, rb0);
    if (phases) phase_end();
    g_kernel_launches += 3;
    RNLA_CUDA(cudaGetLastError());
    return RNLA_OK;
}
void i8_prepare_end() { g_sl.ready = true; g_active = true; }
rnla_status i8_prepare(const double* A, int64_t lda, int64_t m, int64_t n, bool p7) {
    RNLA_TRY(i8_prepare_begin(A, lda, m, n, p7));
    RNLA_TRY(i8_prepare_rows(0, m, true));
    i8_prepare_end();
    return RNLA_OK;
}

// C (m x N) = A * B (n x N) on the split A
static rnla_status i8_gemm_nn_tile(const double* B, int64_t ldb, int64_t N, double* C, int64_t ldc) {
    Ctx& c = ctx();
    Sliced& s = g_sl;
    RNLA_TRY(slice_b(B, ldb, s.n, (int)N, nullptr, s.kb_total));
    i8_mma_kernel<false, false><<<dim3((unsigned)s.rblocks, 1), MMA_THREADS, MMA_SMEM, c.stream>>>(
        s.nn.as<uint8_t>(), s.kb_total, s.bimg.as<uint8_t>(), s.kb_total, s.kb_total, C, ldc, s.m, (int)N, s.up.d(), s.cup.d(), 0);
    ++g_kernel_launches;
    if (g_precise) {
        // second sweep over the same images: the digit pairs the first one dropped
        i8_mma_kernel<false, true><<<dim3((unsigned)s.rblocks, 1), MMA_THREADS, MMA_SMEM, c.stream>>>(
            s.nn.as<uint8_t>(), s.kb_total, s.bimg.as<uint8_t>(), s.kb_total, s.kb_total, C, ldc, s.m, (int)N, s.up.d(), s.cup.d(), 0);
        ++g_kernel_launches;
    }
    RNLA_CUDA(cudaGetLastError());
    return RNLA_OK;
}
rnla_status i8_gemm_nn(const double* B, int64_t ldb, int64_t N, double* C, int64_t ldc) {
    for (int64_t c0 = 0; c0 < N; c0 += BN)
        RNLA_TRY(i8_gemm_nn_tile(B + c0 * ldb, ldb, std::min<int64_t>(BN, N - c0), C + c0 * ldc, ldc));
    return RNLA_OK;
}

// Z (n x N) = A^T * Q (m x N) on the split A (local rows only; the caller all-reduces)
static rnla_status i8_gemm_tn_tile(const double* Q, int64_t ldq, int64_t N, double* Z, int64_t ldz) {
    Ctx& c = ctx();
    Sliced& s = g_sl;
    const bool full = g_full && s.p7;
    RNLA_TRY(slice_b(Q, ldq, s.m, (int)N, s.up.d(), s.kr_total, full));
    const int64_t max_per = full ? 1024 : K_ACC_MAX / 64 / 2 * 2;        // blocks of 64 rows per accumulation (7 pairs x 64^2 x K < 2^31 when full)
    int64_t nchunks = std::max<int64_t>((s.kr_total + max_per - 1) / max_per, (4LL * c.sms + s.cblocks - 1) / s.cblocks);
    nchunks = std::max<int64_t>(1, std::min<int64_t>(nchunks, s.kr_total));
    const int64_t per = (s.kr_total + nchunks - 1) / nchunks;
    nchunks = (s.kr_total + per - 1) / per;
    const int64_t stride = s.n * N;
    Persist& P = s.part; Persist& P2 = s.part2;
    RNLA_CUDA(P.ensure((size_t)nchunks * stride * 8));
    if (full) RNLA_CUDA(P2.ensure((size_t)nchunks * stride * 8));
    i8_mma_kernel<true, false><<<dim3((unsigned)s.cblocks, (unsigned)nchunks), MMA_THREADS, MMA_SMEM, c.stream>>>(
        s.tn.as<uint8_t>(), s.kr_total, s.bimg.as<uint8_t>(), s.kr_total, per, P.d(), s.n, s.n, (int)N, nullptr, nullptr, stride);
    if (full) {
        i8_mma_tn_hi_kernel<<<dim3((unsigned)s.cblocks, (unsigned)nchunks), MMA_THREADS, MMA_SMEM_HI, c.stream>>>(
            s.tn.as<uint8_t>(), s.tnhi.as<uint8_t>(), s.bimg.as<uint8_t>(), s.bimg_hi.as<uint8_t>(), s.kr_total, per, P2.d(), s.n, s.n, (int)N, stride);
        ++g_kernel_launches;
    }
    i8_tn_reduce_kernel<<<(unsigned)std::min<int64_t>(148 * 8, (stride + 255) / 256), 256, 0, c.stream>>>(P.d(), full ? P2.d() : nullptr, (int)nchunks,
                                                                                                           stride, s.n, (int)N, s.cup.d(), Z, ldz);
    g_kernel_launches += 2;
    RNLA_CUDA(cudaGetLastError());
    return RNLA_OK;
}
rnla_status i8_gemm_tn(const double* Q, int64_t ldq, int64_t N, double* Z, int64_t ldz) {
    for (int64_t c0 = 0; c0 < N; c0 += BN)
        RNLA_TRY(i8_gemm_tn_tile(Q + c0 * ldq, ldq, std::min<int64_t>(BN, N - c0), Z + c0 * ldz, ldz));
    return RNLA_OK;
}

}  // namespace rnla
