// The passes over A on the INT8 tensor cores (tcgen05.mma kind::i8, accumulators in TMEM): an Ozaki-style fixed-point splitting of
// the FP64 operands with EXACT int32 accumulation (reference passes: src/lora_helpers.rs:21, :41, :71-95).
//
//   a_ij = up_i * sum_{t<P} d_t(i,j) 2^{-(7+8t)}  (+ at most 2^{-(8P-1)} up_i),    up_i = 2^{e_i+1},  |a_ij| < 2^{e_i} (row maximum)
//
// P digit planes of widths [7, 8, 8, ...] bits: d_0 in [-64, 64], d_t in [-128, 127] (balanced digits: v - d divisible by 256).
// P = 4 is a 31-bit representation, P = 6 a 47-bit one, P = 7 a 55-bit one -- more bits than an FP64 mantissa, relative to the row
// maximum.  The thin operand (Omega, Y, S or Q) is split the same way per column; for the passes that contract over rows
// (A^T Y, Q^T A) the row scale up_i is folded into the thin operand before it is split, so ONE split of A serves both directions.
// Digit pair (ta, tb) has weight 2^{-(14+8(ta+tb))}: all pairs with ta + tb = g accumulate exactly into TMEM accumulator g.
//
// A is split ONCE per driver call into int8 planes stored pre-tiled as shared-memory images of the MMA operands: per 128 x 128
// block of A and per plane [I 8][J 16][8 x 16 B] (I: 16-row group, J: 8-column group; a core matrix holds 16 rows x 8 columns of A
// as 8 rows (columns of A) of 16 bytes (rows of A)).  The same 128 bytes are an MN-major core matrix of A S (contraction over
// columns) and a K-major one of A^T Y (contraction over rows); because the descriptors take both strides, ONE arrangement serves
// both passes, and one tiled TMA load (cp.async.bulk.tensor over a 5-D view of the images) fetches a step of either.
//
// The sweeps run as CTA PAIRS (tcgen05.mma.cta_group::2, M = 256 over the two SMs of a TPC): each CTA holds its own 128-row tile of A
// and half of the thin operand's columns.  TMEM holds four 112-column int32 accumulators per CTA at l = 110, so a product is at most two
// sweeps over the images:
//   4 planes: groups 0..3 (10 pairs; 31-bit operands, result to ~2^-27), optionally groups 4..6 (6 pairs)
//   6 planes: groups 0..3 on planes 0..3 + groups 4, 5 on planes 0..5 (11 pairs)
//   7 planes: groups 0..2 on planes 0..2 (6 pairs) + groups 3..6 on all seven planes (22 pairs): the product of the 55-bit
//             representations up to 2^-54 of (row max) x (column max): FP64-grade.
// Accumulation bound: the largest group (g = 6: 2 x 64*128 + 5 x 128^2 per index) stays below 2^31 for 21845 contraction indices;
// the kernels drain the accumulators into the FP64 output every `flush` stages, so n and m are unlimited.
#include "drivers.cuh"
#include "gemm.cuh"
#include "panel.cuh"
#include "ptx.cuh"
#include "rng.cuh"
#include <cuda.h>
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <string>
#include <vector>

namespace rnla {

namespace {

constexpr int BM = 128, BN = 128;            // CTA tile: 128 x (<= 128) outputs; the work split and `flush` count blocks of 64 contraction indices
constexpr int APLANE = 16384;                // one digit plane of a 128 x 128 block of A
constexpr int MMA_THREADS = 224;             // warp 0 A producer, warp 1 MMA issuer, warps 2-5 epilogue, warp 6 B producer
// stages of 64 contraction indices per int32 accumulation (worst group of the sweep, |d_0| <= 64, |d_t| <= 128)
constexpr int FLUSH_P4 = 682, FLUSH_P6 = 408, FLUSH_P7 = 340;

// ---------------------------------------------------------------------------------------------- tcgen05 wrappers
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---- CTA pair (tcgen05 cta_group::2): two CTAs of a cluster on the two SMs of a TPC run ONE MMA of M = 256; each CTA supplies its own
// 128 rows of the A operand and HALF of the columns of the B operand from its own shared memory, and holds the accumulators of its
// own 128 rows in its own TMEM.  The leader (cluster rank 0) issues; commits are multicast to the same barrier in both CTAs.
__device__ __forceinline__ uint32_t cluster_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrive on the barrier at the same shared-memory offset in CTA `cta` of the cluster
__device__ __forceinline__ void mbar_arrive_remote(uint64_t* bar, uint32_t cta) {
    asm volatile("{\n\t.reg .b32 ra;\n\tmapa.shared::cluster.u32 ra, %0, %1;\n\t"
                 "mbarrier.arrive.shared::cluster.b64 _, [ra];\n\t}" ::"r"(smem_u32(bar)), "r"(cta) : "memory");
}
// address of the barrier at the same offset in the leader's (rank 0) shared memory
__device__ __forceinline__ uint32_t leader_addr(const void* p) {
    uint32_t ra;
    asm volatile("mapa.shared::cluster.u32 %0, %1, 0;" : "=r"(ra) : "r"(smem_u32(p)));
    return ra;
}
// tiled TMA loads of the pair: the bytes land in the executing CTA's shared memory, the transaction count is reported to a barrier
// that may live in the peer (the leader's `full` barrier expects the bytes of both CTAs)
__device__ __forceinline__ void tma2_load_5d(void* dst, const CUtensorMap* tm, uint32_t bar_cluster_addr, int c0, int c1, int c2, int c3, int c4) {
    asm volatile("cp.async.bulk.tensor.5d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
                 ::"r"(smem_u32(dst)), "l"(tm), "r"(bar_cluster_addr), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4) : "memory");
}
__device__ __forceinline__ void tma_prefetch_l2_5d(const CUtensorMap* tm, int c0, int c1, int c2, int c3, int c4) {
    asm volatile("cp.async.bulk.prefetch.tensor.5d.L2.global [%0, {%1, %2, %3, %4, %5}];" ::"l"(tm), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4) : "memory");
}
__device__ __forceinline__ void tma2_load_3d(void* dst, const CUtensorMap* tm, uint32_t bar_cluster_addr, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                 ::"r"(smem_u32(dst)), "l"(tm), "r"(bar_cluster_addr), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tc_alloc2(uint32_t* smem_dst, uint32_t ncols) {         // the same warp of BOTH CTAs executes this
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tc_dealloc2(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_commit2(uint64_t* bar) {                               // arrives in both CTAs of the pair
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ void tc_mma2_i8(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::i8 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}

// shared-memory matrix descriptor, no swizzle (cute::UMMA::SmemDescriptor: start >> 4 | LBO >> 4 << 16 | SBO >> 4 << 32 | version 1 << 46)
// LBO: stride between core matrices along the contraction dimension, SBO: along the M / N dimension
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((addr >> 4) & 0x3fffu) | ((uint64_t)((lbo_bytes >> 4) & 0x3fffu) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3fffu) << 32) | (1ull << 46);
}
// instruction descriptor (cute::UMMA::InstrDescriptor): D = S32, A = B = signed int8, M = 128, N = nmma (a multiple of 16)
__host__ __device__ constexpr uint32_t instr_desc(bool a_mn_major, bool b_mn_major, int nmma, int mmma = BM) {
    return (2u << 4) | (1u << 7) | (1u << 10) | ((a_mn_major ? 1u : 0u) << 15) | ((b_mn_major ? 1u : 0u) << 16) |
           ((uint32_t)(nmma >> 3) << 17) | ((uint32_t)(mmma >> 4) << 24);
}

// ---------------------------------------------------------------------------------------------- splitting
// y + 1.5 * 2^52 holds rint(y) in the low mantissa bits (|y| < 2^31); subtracting the magic number again gives rint(y) as a double
__device__ __forceinline__ int sext8(int v) { return (v << 24) >> 24; }
__device__ __forceinline__ int rint_bits(double y, double* r) {
    const double MAGIC = 6755399441055744.0;                // 1.5 * 2^52
    const double t = y + MAGIC;
    *r = t - MAGIC;
    return (int)(unsigned)__double_as_longlong(t);          // low 32 bits: two's complement of rint(y)
}
// x * down = y, |y| < 2^30:  vh = rint(y) carries planes 0..3 (7 + 8 + 8 + 8 bits); the remainder y - vh (exact, |.| <= 1/2) times
// 2^{8(P-4)} the trailing planes.  The trailing part's top digit can come out as +128: the carry goes into vh.
template <int P>
__device__ __forceinline__ void digits(double x, double down, int (&d)[P]) {
    const double y = x * down;
    double yr;
    int vh = rint_bits(y, &yr);
    if constexpr (P > 4) {
        double r2;
        int vl = rint_bits((y - yr) * (P == 6 ? 65536.0 : 16777216.0), &r2);
#pragma unroll
        for (int t = P - 1; t > 4; --t) { d[t] = sext8(vl); vl = (vl - d[t]) >> 8; }
        const int c = (vl + 128) >> 8;
        d[4] = vl - (c << 8);
        vh += c;
    }
    d[3] = sext8(vh); vh = (vh - d[3]) >> 8;
    d[2] = sext8(vh); vh = (vh - d[2]) >> 8;
    d[1] = sext8(vh); vh = (vh - d[1]) >> 8;
    d[0] = vh;
}
// The digits of four elements at once, packed for the images: byte e of w[t] = digit t of element e.  Balanced digits are the
// bytes of a biased integer: v = sum d_t 256^t with d_t in [-128, 127]  <=>  the unsigned bytes of v + sum 128 * 256^t are d_t + 128, and
// d_t + 128 as an unsigned byte is d_t as a signed byte with the top bit flipped.  So (v + 0x00808080) ^ 0x00808080 holds planes 3, 2, 1
// in its low bytes and plane 0 (no bias: the signed remainder) in its top byte; likewise the trailing planes, whose overflow (the top
// digit coming out as +128) is the byte above them and is carried into v.  Two integer operations per element instead of a shift /
// sign-extend / subtract chain per digit, then a 4 x 4 byte transpose (eight PRMT) across the four elements.  Same digits as digits<P>
// (the representation is unique), which the images of the thin operand and the CPU emulation use.
template <int P>
__device__ __forceinline__ void digits4_packed(const double (&x)[4], const double (&sc)[4], unsigned (&w)[P]) {
    unsigned wh[4], wl[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const double y = x[e] * sc[e];
        double yr;
        int vh = rint_bits(y, &yr);
        if constexpr (P > 4) {
            double r2;
            const int vl = rint_bits((y - yr) * (P == 6 ? 65536.0 : 16777216.0), &r2);
            const unsigned bias = P == 6 ? 0x00008080u : 0x00808080u;
            const unsigned ul = (unsigned)vl + bias;
            vh += (int)(ul >> (P == 6 ? 16 : 24));               // 0 or 1
            wl[e] = ul ^ bias;
        }
        wh[e] = ((unsigned)vh + 0x00808080u) ^ 0x00808080u;
    }
    {
        const unsigned t0 = __byte_perm(wh[0], wh[1], 0x5140), t1 = __byte_perm(wh[0], wh[1], 0x7362);
        const unsigned t2 = __byte_perm(wh[2], wh[3], 0x5140), t3 = __byte_perm(wh[2], wh[3], 0x7362);
        w[3] = __byte_perm(t0, t2, 0x5410); w[2] = __byte_perm(t0, t2, 0x7632);
        w[1] = __byte_perm(t1, t3, 0x5410); w[0] = __byte_perm(t1, t3, 0x7632);
    }
    if constexpr (P > 4) {
        const unsigned t0 = __byte_perm(wl[0], wl[1], 0x5140), t2 = __byte_perm(wl[2], wl[3], 0x5140);
        w[P - 1] = __byte_perm(t0, t2, 0x5410); w[P - 2] = __byte_perm(t0, t2, 0x7632);
        if constexpr (P == 7) {
            const unsigned t1 = __byte_perm(wl[0], wl[1], 0x7362), t3 = __byte_perm(wl[2], wl[3], 0x7362);
            w[4] = __byte_perm(t1, t3, 0x5410);
        }
    }
}
// exponent bookkeeping from the bit pattern of a maximum (biased exponent E: max < 2^(E-1022)): up = 2^(E-1021), down = 2^31 / up.
// flags: 1 = Inf / NaN seen, 2 = a non-zero maximum too small to scale (below 2^-959)
__device__ __forceinline__ void scales_from_max_bits(unsigned long long bits, double* up, double* down, int* flags) {
    const int E = (int)(bits >> 52) & 0x7ff;
    if (E == 0x7ff) { *up = 0.0; *down = 0.0; atomicOr(flags, 1); return; }
    if (E < 64 || E >= 2045) { *up = 0.0; *down = 0.0; if (bits != 0ull) atomicOr(flags, 2); return; }
    *up = __longlong_as_double((long long)(E + 2) << 52);
    *down = __longlong_as_double((long long)(2075 - E) << 52);
}

// row maxima as integer maxima of the bit patterns of |a_ij| (ordered like the values; a NaN beats everything and is reported)
__global__ void __launch_bounds__(256)
rowmax_kernel(const double* __restrict__ A, int64_t lda, int64_t m, int64_t n, int64_t cols_per, unsigned long long* __restrict__ bits) {
    const int64_t i = ((int64_t)blockIdx.x * 256 + threadIdx.x) * 2;
    if (i >= m) return;
    const int64_t j0 = (int64_t)blockIdx.y * cols_per, j1 = min(n, j0 + cols_per);
    const bool two = i + 1 < m;
    const bool vec = two && ((reinterpret_cast<uintptr_t>(A) & 15) == 0) && (lda % 2 == 0);
    const unsigned long long ABS = 0x7fffffffffffffffull;
    unsigned long long m0 = 0ull, m1 = 0ull;
    if (vec) {
#pragma unroll 8
        for (int64_t j = j0; j < j1; ++j) {
            unsigned long long a, b;
            asm volatile("ld.global.nc.L1::no_allocate.v2.b64 {%0,%1}, [%2];" : "=l"(a), "=l"(b) : "l"(A + i + j * lda));
            m0 = max(m0, a & ABS); m1 = max(m1, b & ABS);
        }
    } else {
        for (int64_t j = j0; j < j1; ++j) {
            m0 = max(m0, (unsigned long long)__double_as_longlong(ldg_stream(A + i + j * lda)) & ABS);
            if (two) m1 = max(m1, (unsigned long long)__double_as_longlong(ldg_stream(A + i + 1 + j * lda)) & ABS);
        }
    }
    atomicMax(bits + i, m0);
    if (two) atomicMax(bits + i + 1, m1);
}
__global__ void scales_kernel(const unsigned long long* __restrict__ bits, int64_t cnt, double* __restrict__ up, double* __restrict__ down,
                              int* __restrict__ flags) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < cnt) scales_from_max_bits(bits[i], up + i, down + i, flags);
}

// staging-buffer swizzle (the 16-byte row inside a core matrix is XORed with the index I of its 16-row group): lanes that write the
// same row of the eight core matrices of one column hit different banks; undone by the copy-out
__device__ __forceinline__ int stage_swz(int off) { return off ^ (((off >> 9) & 7) << 4); }

// The digit split of A: one CTA per 128 rows x 32 columns (a quarter of an image block: 16 elements per thread keep the register
// count low enough for several resident CTAs per SM, which is what keeps loads in flight while other CTAs form digits and store).
// Shared staging per plane: [I 8][Jl 4][8 x 16 B] = 4 KB; copied out as eight 512-byte runs per plane.
constexpr int SL_COLS = 32;
constexpr int SL_PLANE = 128 * SL_COLS;            // 4 KB
template <int P>
__global__ void __launch_bounds__(256, P == 4 ? 4 : 3)
slice_a_kernel(const double* __restrict__ A, int64_t lda, int64_t m, int64_t n, const double* __restrict__ down,
               uint8_t* __restrict__ img, int64_t cblocks, int64_t rb0 /* first 128-row block of this launch */) {
    extern __shared__ __align__(16) uint8_t stg[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t hb_total = 4 * cblocks;
    const int64_t rb = rb0 + blockIdx.x / hb_total, hb = blockIdx.x % hb_total;
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
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const int jl = warp + 8 * r;                        // 0 .. 31
        unsigned wd[P];
        digits4_packed<P>(x[r], sc, wd);
        const int intra = (il >> 4) * 512 + (jl >> 3) * 128 + (jl & 7) * 16 + (il & 15);
#pragma unroll
        for (int t = 0; t < P; ++t) *reinterpret_cast<unsigned*>(stg + stage_swz(t * SL_PLANE + intra)) = wd[t];
    }
    __syncthreads();
    const uint4* src = reinterpret_cast<const uint4*>(stg);
    uint8_t* blk = img + ((rb * cblocks + (hb >> 2)) * P) * (int64_t)APLANE + (hb & 3) * 512;
    for (int q = threadIdx.x; q < P * (SL_PLANE / 16); q += 256) {
        const int t = q >> 8, w = q & 255;                  // 256 uint4 per plane: I = w >> 5, 32 uint4 = one 512-byte run
        *reinterpret_cast<uint4*>(blk + t * APLANE + (w >> 5) * 2048 + (w & 31) * 16) = src[stage_swz(q * 16) >> 4];
    }
}

// Where the thin operand comes from: a matrix X in memory, or -- for the first product of the range finder, Y = A Omega -- the
// counter-based generator itself (rng.cuh: omega(k, c) is a pure function of (seed, stream, k, c)), so that Omega's FP64 values are
// never materialised: its digit planes are formed straight from the Philox blocks, and the column maxima the split needs come from a
// first evaluation of the same function (2.2 M Gaussians at the headline size: microseconds).  Bit-identical to materialising Omega
// and splitting it.
struct ThinSrc {
    const double* X; int64_t ldx;
    int dist;                                   // < 0: read X; else generate with (dist, seed, stream, column offset c0)
    uint64_t seed; uint32_t stream; uint32_t c0;
};
__device__ __forceinline__ double thin_value(const ThinSrc& t, int64_t k, int c) {
    return t.dist < 0 ? t.X[k + (int64_t)c * t.ldx] : omega_entry(t.dist, t.seed, t.stream, (uint64_t)k, t.c0 + (uint32_t)c);
}
// per-column maxima of X (K x N), optionally with the row scale folded in (X(k, c) * rs[k])
__global__ void __launch_bounds__(256)
colmax_kernel(const ThinSrc src, int64_t K, int N, const double* __restrict__ rs, int64_t rows_per,
              unsigned long long* __restrict__ bits) {
    __shared__ double red[8];
    const int c = blockIdx.x;
    const int64_t k0 = (int64_t)blockIdx.y * rows_per, k1 = min(K, k0 + rows_per);
    double mx = 0.0;
    for (int64_t k = k0 + threadIdx.x; k < k1; k += 256) mx = fmax(mx, fabs(thin_value(src, k, c) * (rs ? rs[k] : 1.0)));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; ++w) mx = fmax(mx, red[w]);
        atomicMax(bits + c, (unsigned long long)__double_as_longlong(mx));
    }
}
// ---------------------------------------------------------------------------------------------- the CTA-pair MMA kernel
// The single-CTA sweeps above are bound by what the SMs can take in from L2, and half of that traffic is the thin
// operand: every 128-row tile of A re-reads all of it.  Here two CTAs of a cluster (the two SMs of a TPC) work on two neighbouring
// tiles with ONE tcgen05.mma.cta_group::2 of M = 256 per digit pair: each CTA loads the planes of its own tile and only HALF of the
// columns of the thin operand (23 % fewer bytes into the SMs per product at l = 110).
//
// Thin-operand images for the pair, K-major: [block of 32 k][rank 2][plane P][n group w8/8][k chunk 2][8 columns x 16 B of k].
// Rank 0 holds columns [0, w8), rank 1 columns [w8, N), w8 = 8 ceil(N / 16).  A plane is w8 x 32 bytes, the MMA reads w = N_mma / 2 >= w8
// columns of it (N_mma a multiple of 32): the n groups beyond w8 overlap the next plane -- whatever they hold only reaches the
// accumulator columns [w8, w) and [w + w8, 2 w), which nobody reads.
template <int P>
__global__ void __launch_bounds__(256)
slice_b2_kernel(const ThinSrc src, int64_t K, int N, int w8, const double* __restrict__ rs, const double* __restrict__ cdown,
                uint8_t* __restrict__ out) {
    const int64_t kb = blockIdx.x;                          // 64 contraction indices = two image blocks
    const int kq = threadIdx.x & 15;                        // four consecutive k: one 4-byte piece of a 16-byte row of a core matrix
    const int kk = 4 * kq;
    const int64_t k0 = kb * 64 + kk;
    double rsk[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) rsk[e] = (k0 + e < K) ? (rs ? rs[k0 + e] : 1.0) : 0.0;
    const int plane = w8 * 32;
    for (int c = threadIdx.x >> 4; c < 2 * w8; c += 16) {
        const int rank = c >= w8 ? 1 : 0, lc = c - rank * w8;
        unsigned wd[P];
#pragma unroll
        for (int t = 0; t < P; ++t) wd[t] = 0u;
        if (c < N) {
            const double cd = cdown[c];
            double xv[4];
            if (src.dist < 0) {
#pragma unroll
                for (int e = 0; e < 4; ++e) xv[e] = k0 + e < K ? src.X[k0 + e + (int64_t)c * src.ldx] : 0.0;
            } else {
                // k0 is a multiple of 4: the four values are the four words of ONE Philox block (rng.cuh omega_block)
                const u32x4 b = omega_block(src.seed, src.stream, (uint64_t)k0 >> 2, src.c0 + (uint32_t)c);
                xv[0] = sample_from_u32(src.dist, b.x); xv[1] = sample_from_u32(src.dist, b.y);
                xv[2] = sample_from_u32(src.dist, b.z); xv[3] = sample_from_u32(src.dist, b.w);
            }
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                int d[P];
#pragma unroll
                for (int t = 0; t < P; ++t) d[t] = 0;
                if (k0 + e < K) digits<P>(xv[e] * rsk[e], cd, d);
#pragma unroll
                for (int t = 0; t < P; ++t) wd[t] |= ((unsigned)d[t] & 0xffu) << (8 * e);
            }
        }
        uint8_t* dst = out + (((kb * 2 + (kk >> 5)) * 2 + rank) * P) * (int64_t)plane + (lc >> 3) * 256 + ((kk & 31) >> 4) * 128 + (lc & 7) * 16 + (kk & 15);
#pragma unroll
        for (int t = 0; t < P; ++t) *reinterpret_cast<unsigned*>(dst + t * plane) = wd[t];
    }
}

// all digit pairs of a sweep, one MMA each, on one step of 32 contraction indices; FRESH: the first step of an accumulation (the
// first MMA into an accumulator overwrites it)
constexpr int ASTEP = 4096;                  // one digit plane of a step (128 x 32)
constexpr int MAX_RING2 = 24;                // slots per ring (barrier arrays)
template <bool TN, int PU, int G0, int NG, bool WIDE, bool FRESH>
__device__ __forceinline__ void issue_step2(uint32_t tmem, uint32_t a_base, uint32_t b_base, uint32_t w, uint32_t bplane, uint32_t idesc) {
    static_assert(!WIDE || (NG % 2 == 0 && G0 + NG <= PU), "a pair of groups needs the partner plane tb + 1 < PU");
    unsigned touched = FRESH ? 0u : 0xffu;
#pragma unroll
    for (int ta = 0; ta < PU; ++ta) {
        // A S:   MN-major, step plane [I 8][J 4][128 B]:  K stride (J) 128, M stride (I) 512
        // A^T Y: K-major,  step plane [I 2][J 16][128 B]: K stride (I) 2048, M stride (J) 128
        const uint64_t ad = TN ? smem_desc(a_base + ta * ASTEP, 2048, 128) : smem_desc(a_base + ta * ASTEP, 128, 512);
        if constexpr (WIDE) {
            // One MMA of double width per (plane of A, PAIR of groups 2 p, 2 p + 1): the thin operand's planes tb, tb + 1 are neighbours
            // along N in shared memory, and each CTA's half of the N = 4 w columns is [plane tb: w][plane tb + 1: w], so the accumulators
            // of a pair of groups sit as [rank 0: g, g + 1][rank 1: g, g + 1].  A is read from shared memory 12 instead of 22 times per
            // step.  A group without a partner plane (tb = -1) multiplies the zero plane that precedes plane 0 in every slot.
#pragma unroll
            for (int pi = 0; pi < NG / 2; ++pi) {
                const int tb = G0 + 2 * pi - ta;                 // plane of group 2 pi; group 2 pi + 1 takes plane tb + 1
                if (tb + 1 < 0 || tb >= PU) continue;
                const uint64_t bd = smem_desc(b_base + (tb + 1) * bplane, 128, 256);     // slot = [zero plane][plane 0] ... [plane PU - 1]
                tc_mma2_i8(tmem + (uint32_t)pi * 4u * w, ad, bd, idesc, (touched >> pi) & 1u);
                touched |= 1u << pi;
            }
        } else {
#pragma unroll
            for (int tb = 0; tb < PU; ++tb) {
                const int g = ta + tb - G0;
                if (g < 0 || g >= NG) continue;
                const uint64_t bd = smem_desc(b_base + (tb + 1) * bplane, 128, 256);     // thin operand, K-major: k chunks 128 B apart, n groups 256 B apart
                tc_mma2_i8(tmem + (uint32_t)g * 2u * w, ad, bd, idesc, (touched >> g) & 1u);
                touched |= 1u << g;
            }
        }
    }
}
// grid.x = 2 x pairs of tiles (cluster = blockIdx.x {2p, 2p + 1}); `tiles` = valid tiles (an odd count leaves a phantom that loads
// the last tile again and stores nothing); w = columns of the thin operand each CTA feeds to the MMA (N_mma = 2 w), w8 = columns per
// rank in the images.  Both operands advance in steps of 32 contraction indices (one MMA K step) through their own rings: what covers
// the latency of a refill is the number of bytes in flight, (depth - 1) / depth of a ring, so the steps are as small as the MMA allows.
// Barriers: the LEADER's fullA / fullB count the bytes of both CTAs' tiled TMA loads (cp.async.bulk.tensor with cta_group::2 may
// report to a barrier in the peer); emptyA / emptyB / accum are committed by the leader's MMAs in both CTAs; the leader's `drained`
// collects the epilogue warps of both.
// what a sweep needs besides its tensor maps
struct SweepArgs {
    int64_t cblocks, tiles;                 // column blocks of the images; valid tiles along grid.x
    int w8;                                 // thin-operand columns per rank
    int pfd;                                // experiments: L2 prefetch distance in steps (RNLA_I8_PFD)
    int64_t kblocks_total, kblocks_per_chunk;
    double* C; int64_t ldc, rows; int ncols;
    const double* rs_up; const double* cs_up;
    int64_t chunk_stride;
};
constexpr int BAR_SET_BYTES = 1024;          // one set of ring barriers (4 x MAX_RING2 + accum + drained)
constexpr size_t SMEM_MAX = 232448;
constexpr size_t RING_BYTES = SMEM_MAX - 128 - 2 * BAR_SET_BYTES - 64;       // alignment slack, two barrier sets, the TMEM slot

// One sweep of a CTA pair over its tile: digit pairs (ta, tb), ta, tb < PU, with G0 <= ta + tb < G0 + NG into NG TMEM accumulators.
// `set` selects the barrier set (a kernel that runs several sweeps back to back gives each its own); `add`: add to what is in C.
// (Both sweeps of a 55-bit product in one launch, pairs alternating which one they run first so that the chip always holds a mix of
// the light and the heavy sweep, was measured: 10.1 - 10.9 ms against 10.1 - 10.2 ms for two launches -- the time of a product is the
// sum of its sweeps' energies under the power cap, not a matter of which resource idles; profiles/r02_pair_both_vs_separate.log.)
template <bool TN, int PU, int G0, int NG, bool WIDE>
__device__ __forceinline__ void pair_sweep(const CUtensorMap* tmA, const CUtensorMap* tmB, const SweepArgs& a, uint8_t* smem, int set, uint32_t tmem,
                                           int na, int nbs, int flush, bool add) {
    const int w8 = a.w8;
    const uint32_t bplane = (uint32_t)w8 * 32;                 // one plane of a step of this rank's columns
    const uint32_t a_bytes = PU * ASTEP, b_bytes = PU * bplane, b_slot = (PU + 1) * bplane;   // slot of the thin operand: [zero plane][PU planes]
    const int w = w8;                                          // columns per rank the MMA reads: N_mma = 2 w (4 w for a pair of groups)
    uint8_t* bring = smem + (size_t)na * a_bytes;
    uint64_t* fullA = reinterpret_cast<uint64_t*>(smem + RING_BYTES + (size_t)set * BAR_SET_BYTES);
    uint64_t* emptyA = fullA + MAX_RING2;
    uint64_t* fullB = emptyA + MAX_RING2;
    uint64_t* emptyB = fullB + MAX_RING2;
    uint64_t* accum = emptyB + MAX_RING2;
    uint64_t* drained = accum + 1;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_rank();
    const int64_t tile_raw = blockIdx.x, tile = min(tile_raw, a.tiles - 1);
    const int64_t cblocks = a.cblocks;
    const int64_t kb0 = (int64_t)blockIdx.y * a.kblocks_per_chunk;
    const int64_t kb1 = min(a.kblocks_total, kb0 + a.kblocks_per_chunk);
    const int nk = (int)max((int64_t)0, kb1 - kb0);            // blocks of 64 contraction indices (the unit of the work split and of `flush`)
    const int ns = 2 * nk;                                     // steps of 32
    const int64_t gs0 = 2 * kb0;                               // first step
    const int nflush = (nk + flush - 1) / flush;
    const int pfd = a.pfd;

    if (threadIdx.x == 0) {
        for (int s = 0; s < na; ++s) { mbar_init(fullA + s, 1); mbar_init(emptyA + s, 1); }
        for (int s = 0; s < nbs; ++s) { mbar_init(fullB + s, 1); mbar_init(emptyB + s, 1); }
        mbar_init(accum, 1);
        mbar_init(drained, 8);
        mbar_fence_init();
    }
    for (int s = 0; s < nbs; ++s)                              // the zero planes (the TMA loads land behind them)
        for (uint32_t q = threadIdx.x * 16; q < bplane; q += MMA_THREADS * 16) *reinterpret_cast<uint4*>(bring + (size_t)s * b_slot + q) = make_uint4(0, 0, 0, 0);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic-proxy stores -> visible to the tensor core's reads
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();                                        // the peer's barriers exist before anything is signalled across
    tc_fence_after();

    if (warp == 0) {
        if (lane == 0) {
            // one tiled TMA load per step: all PU planes of this CTA's tile (A S: eight 512-byte runs per plane, A^T Y: 4 KB per plane)
            int s = 0; uint32_t ph = 1;                         // parity of the `empty` phase to wait for: the first lap passes at once
            for (int j = 0; j < ns; ++j) {
                mbar_wait(emptyA + s, ph);
                const int64_t gs = gs0 + j, blk2 = gs >> 2;
                const int q = (int)(gs & 3);
                if (rank == 0) mbar_arrive_expect_tx(fullA + s, 2 * a_bytes);         // the leader's barrier counts both CTAs' bytes
                uint8_t* st = smem + (size_t)s * a_bytes;
                if (TN) tma2_load_5d(st, tmA, leader_addr(fullA + s), 0, 0, 2 * q, 0, (int)(blk2 * cblocks + tile));
                else tma2_load_5d(st, tmA, leader_addr(fullA + s), 0, q, 0, 0, (int)(tile * cblocks + blk2));
                if (j + pfd < ns) {
                    const int64_t gp = gs + pfd, bp = gp >> 2;
                    const int qp = (int)(gp & 3);
                    if (TN) tma_prefetch_l2_5d(tmA, 0, 0, 2 * qp, 0, (int)(bp * cblocks + tile));
                    else tma_prefetch_l2_5d(tmA, 0, qp, 0, 0, (int)(tile * cblocks + bp));
                }
                if (++s == na) { s = 0; ph ^= 1; }
            }
        }
    } else if (warp == 6) {
        if (lane == 0) {
            int s = 0; uint32_t ph = 1;
            for (int j = 0; j < ns; ++j) {
                mbar_wait(emptyB + s, ph);
                if (rank == 0) mbar_arrive_expect_tx(fullB + s, 2 * b_bytes);
                tma2_load_3d(bring + (size_t)s * b_slot + bplane, tmB, leader_addr(fullB + s), 0, 0, (int)((gs0 + j) * 2 + rank));   // the leading PU planes
                if (++s == nbs) { s = 0; ph ^= 1; }
            }
        }
    } else if (warp == 1) {
        if (lane == 0 && rank == 0) {
            const uint32_t idesc = instr_desc(!TN, false, (WIDE ? 4 : 2) * w, 2 * BM);
            const int fsteps = 2 * flush;
            int sa = 0, sb = 0; uint32_t pha = 0, phb = 0;
            int since = 0, drains = 0;                          // steps since the last drain of the accumulators
            for (int j = 0; j < ns; ++j) {
                const bool fresh = since == 0;                  // first step of an accumulation: overwrite the accumulators
                if (fresh && j > 0) {
                    tc_commit2(accum);                          // everything issued so far -> the epilogue warps of both CTAs drain
                    mbar_wait(drained, drains & 1);
                    ++drains;
                    tc_fence_after();
                }
                if (++since == fsteps) since = 0;
                mbar_wait(fullA + sa, pha);
                mbar_wait(fullB + sb, phb);
                tc_fence_after();
                const uint32_t a_base = smem_u32(smem + (size_t)sa * a_bytes), b_base = smem_u32(bring + (size_t)sb * b_slot);
                if (fresh) issue_step2<TN, PU, G0, NG, WIDE, true>(tmem, a_base, b_base, (uint32_t)w, bplane, idesc);
                else issue_step2<TN, PU, G0, NG, WIDE, false>(tmem, a_base, b_base, (uint32_t)w, bplane, idesc);
                tc_commit2(emptyA + sa);
                tc_commit2(emptyB + sb);
                if (++sa == na) { sa = 0; pha ^= 1; }
                if (++sb == nbs) { sb = 0; phb ^= 1; }
            }
            tc_commit2(accum);
        }
    } else {
        const int quad = warp & 3;                          // TMEM lane quadrant this warp may read
        const int64_t r = tile_raw * BM + quad * 32 + lane;     // output row (A S) / output row = column of A (A^T Y)
        const int ncols = a.ncols;
        const int64_t rows = a.rows, ldc = a.ldc;
        const double rsc = (!TN && r < rows) ? a.rs_up[r] : 1.0;
        double* out = a.C + (TN ? (int64_t)blockIdx.y * a.chunk_stride : 0);
        if (nk <= 0 && !add && r < rows) for (int c = 0; c < ncols; ++c) out[r + (int64_t)c * ldc] = 0.0;
        for (int f = 0; f < nflush; ++f) {
            mbar_wait(accum, f & 1);
            tc_fence_after();
            for (int seg = 0; seg < 2; ++seg) {
                const int cbase = seg * w8, cnt = min(ncols - cbase, w8);     // output columns [cbase, cbase + cnt) <- this rank's accumulator columns
                for (int c0 = 0; c0 < cnt; c0 += 16) {
                    uint32_t d[NG][16];
#pragma unroll
                    for (int g = 0; g < NG; ++g)       // WIDE: pairs of groups as [rank 0: g, g + 1][rank 1: g, g + 1]; else [rank 0: g][rank 1: g]
                        tc_ld16(tmem + ((uint32_t)(quad * 32) << 16) + (uint32_t)((WIDE ? (g >> 1) * 4 * w + seg * 2 * w + (g & 1) * w : g * 2 * w + seg * w) + c0), d[g]);
                    tc_wait_ld();
                    if (r < rows) {
#pragma unroll
                        for (int e = 0; e < 16; ++e) {
                            if (c0 + e < cnt) {
                                const int c = cbase + c0 + e;
                                double v = (double)(int)d[NG - 1][e];
#pragma unroll
                                for (int g = NG - 2; g >= 0; --g) v = v * 0.00390625 + (double)(int)d[g][e];
                                v *= 1.0 / (double)(1ull << (14 + 8 * G0));                     // 2^-(14 + 8 G0): weight of the sweep's first group
                                if (!TN) v *= rsc * a.cs_up[c];
                                double* o = out + r + (int64_t)c * ldc;
                                if (add || f > 0) *o += v; else *o = v;
                            }
                        }
                    }
                }
            }
            tc_fence_before();
            if (f + 1 < nflush) {
                __syncwarp();
                if (lane == 0) { if (rank == 0) mbar_arrive(drained); else mbar_arrive_remote(drained, 0); }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();                                        // every MMA of the pair has completed, every load has been consumed
    tc_fence_after();
}

// grid.x = 2 x pairs of tiles (cluster = blockIdx.x {2p, 2p + 1}); `tiles` = valid tiles (an odd count leaves a phantom that loads
// the last tile again and stores nothing).  Both operands advance in steps of 32 contraction indices (one MMA K step) through their
// own rings.  Barriers: the LEADER's fullA / fullB count the bytes of both CTAs' tiled TMA loads (cp.async.bulk.tensor with
// cta_group::2 may report to a barrier in the peer); emptyA / emptyB / accum are committed by the leader's MMAs in both CTAs; the
// leader's `drained` collects the epilogue warps of both.
template <bool TN, int PU, int G0, int NG, bool ADD, bool WIDE>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(MMA_THREADS, 1)
i8_mma2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const SweepArgs a, int na, int nbs, int flush) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~(uintptr_t)127);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + RING_BYTES + 2 * BAR_SET_BYTES);
    if ((threadIdx.x >> 5) == 1) tc_alloc2(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    pair_sweep<TN, PU, G0, NG, WIDE>(&tmA, &tmB, a, smem, 0, tmem, na, nbs, flush, ADD);
    if ((threadIdx.x >> 5) == 1) tc_dealloc2(tmem, 512);
}
// Z(j, c) = cs_up(c) * sum over chunks, fixed order
__global__ void __launch_bounds__(256)
i8_tn_reduce_kernel(const double* __restrict__ P, int nchunks, int64_t chunk_stride, int64_t n, int ncols,
                    const double* __restrict__ cs_up, double* __restrict__ Z, int64_t ldz) {
    const int64_t total = n * ncols;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
        const int64_t c = idx / n, j = idx - c * n;
        double s = 0.0;
        for (int k = 0; k < nchunks; ++k) s += P[(int64_t)k * chunk_stride + j + c * n];
        Z[j + c * ldz] = s * cs_up[c];
    }
}

// The digit-plane images are tens of GB: they live in a workspace that persists across driver calls (grown on demand, released
// by rnla_release_workspace / rnla_shutdown) instead of going through the stream-ordered pool on every call -- re-carving 28 GB
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
    int64_t rblocks = 0, cblocks = 0;
    int planes = 0;                                 // digit planes stored for A
    Persist img, up, down, bits, bimg, cbits, cup, cdown, part, flags;
    bool ready = false;
};
Sliced g_sl;
bool g_active = false;
int g_planes = 4;             // precision of the next products: digit planes of both operands ...
bool g_all_pairs = false;     // ... and, for 4 planes, whether the second sweep (groups 4..6) runs
int g_flush_override = 0;     // tests: drain the accumulators every this many stages (0: the exactness bound FLUSH_P*)

inline int pair_w8(int N) { return 8 * ((N + 15) / 16); }                 // thin-operand columns per rank in the images
// tcgen05.mma.cta_group::2.kind::i8 takes N in steps of 16: the MMA reads exactly the w8 columns per rank that the images hold
// (N_mma = 112 at l = 110; CuTe's static_assert of N % 32 is a library limit -- the products are bit-identical to the CPU emulation
// at N_mma = 16, 48, 80, 112, ... on the B200, tests/test_gpu_parity.py).
rnla_status slice_b(const ThinSrc& src, int64_t K, int N, const double* rs, int64_t kblocks, int planes) {
    Ctx& c = ctx();
    const int w8 = pair_w8(N);
    RNLA_CUDA(g_sl.bimg.ensure((size_t)kblocks * planes * w8 * 128));       // kblocks blocks of 64 = 2 kblocks steps of 32, two ranks each
    RNLA_CUDA(cudaMemsetAsync(g_sl.cbits.p, 0, 128 * 8, c.stream));
    const int64_t rows_per = 32768;
    colmax_kernel<<<dim3((unsigned)N, (unsigned)((K + rows_per - 1) / rows_per)), 256, 0, c.stream>>>(src, K, N, rs, rows_per,
                                                                                                      g_sl.cbits.as<unsigned long long>());
    scales_kernel<<<1, 128, 0, c.stream>>>(g_sl.cbits.as<unsigned long long>(), 128, g_sl.cup.d(), g_sl.cdown.d(), g_sl.flags.as<int>() + 1);
    uint8_t* out = g_sl.bimg.as<uint8_t>();
    if (planes == 4) slice_b2_kernel<4><<<(unsigned)kblocks, 256, 0, c.stream>>>(src, K, N, w8, rs, g_sl.cdown.d(), out);
    else if (planes == 6) slice_b2_kernel<6><<<(unsigned)kblocks, 256, 0, c.stream>>>(src, K, N, w8, rs, g_sl.cdown.d(), out);
    else slice_b2_kernel<7><<<(unsigned)kblocks, 256, 0, c.stream>>>(src, K, N, w8, rs, g_sl.cdown.d(), out);
    g_kernel_launches += 3;
    RNLA_CUDA(cudaGetLastError());
    return RNLA_OK;
}

// Ring sizes: both operands advance in steps of 32 contraction indices.  The thin operand (L2-resident) gets four slots of
// (PU + 1) x w8 x 32 bytes, the planes of A the rest of the 227 KB of dynamic shared memory (6 steps of 28 KB in the 7-plane sweep).
// Measured at the headline size (profiles/r02_pair_experiments_last.log): 3, 4 or 6 slots for the thin operand, steps of 64 instead of
// 32 for A, and L2 prefetch 3 .. 12 steps ahead all leave the 7-plane sweep within 2 % -- under the power cap it is the energy of a
// product that sets its time, not the latency the rings cover.
inline void mma2_rings(int pu, int w8, int* na, int* nbs) {
    const size_t a = (size_t)pu * ASTEP, b = (size_t)(pu + 1) * w8 * 32;
    int s = 4;
    if (const char* ov = getenv("RNLA_I8_NBS2")) { const int v = atoi(ov); if (v >= 2 && v <= MAX_RING2 && (RING_BYTES - (size_t)v * b) / a >= 2) s = v; }
    *nbs = s;
    *na = (int)std::min<size_t>(MAX_RING2, (RING_BYTES - (size_t)s * b) / a);
}
// Tensor maps over the images (elements: 8-byte words).  A: [block][plane][I 8][quarter 4][512 B]; a step (32 contraction indices) of
// A S is the box (512 B, one quarter, all I, PU planes), a step of A^T Y the box (512 B, all quarters, two I, PU planes): ONE copy
// instruction per step either way (separate 1 KB bulk copies are limited to 30 GB/s per SM by their per-instruction cost,
// profiles/r02_ingest_rate.json).
typedef CUresult (*TensorMapEncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                      const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
inline TensorMapEncodeFn tensor_map_encoder() {
    static TensorMapEncodeFn fn = [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) p = nullptr;
        return reinterpret_cast<TensorMapEncodeFn>(p);
    }();
    return fn;
}
inline rnla_status make_tensor_map(CUtensorMap* tm, void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides_bytes, const cuuint32_t* box) {
    TensorMapEncodeFn enc = tensor_map_encoder();
    if (!enc) return fail(RNLA_ERR_COMPUTATION, "int8 passes: cuTensorMapEncodeTiled is not available from the driver");
    const cuuint32_t ones[5] = {1, 1, 1, 1, 1};
    const CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_UINT64, (cuuint32_t)rank, base, dims, strides_bytes, box, ones, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(RNLA_ERR_COMPUTATION, "int8 passes: cuTensorMapEncodeTiled failed (" + std::to_string((int)r) + ")");
    return RNLA_OK;
}
inline rnla_status sweep_tensor_maps(bool tn, int pu, int w8, int64_t kblocks_total, CUtensorMap* tmA, CUtensorMap* tmB) {
    Sliced& s = g_sl;
    const cuuint64_t dims[5] = {64, 4, 8, (cuuint64_t)s.planes, (cuuint64_t)(s.rblocks * s.cblocks)};
    const cuuint64_t strides[4] = {512, 2048, (cuuint64_t)APLANE, (cuuint64_t)s.planes * APLANE};
    const cuuint32_t box_nn[5] = {64, 1, 8, (cuuint32_t)pu, 1}, box_tn[5] = {64, 4, 2, (cuuint32_t)pu, 1};
    RNLA_TRY(make_tensor_map(tmA, s.img.p, 5, dims, strides, tn ? box_tn : box_nn));
    // thin operand: [block of 32 k x rank][plane][w8 x 32 bytes]
    const cuuint64_t bdims[3] = {(cuuint64_t)w8 * 4, (cuuint64_t)g_planes, (cuuint64_t)(4 * kblocks_total)};
    const cuuint64_t bstrides[2] = {(cuuint64_t)w8 * 32, (cuuint64_t)g_planes * w8 * 32};
    const cuuint32_t bbox[3] = {(cuuint32_t)w8 * 4, (cuuint32_t)pu, 1};
    return make_tensor_map(tmB, s.bimg.p, 3, bdims, bstrides, bbox);
}
inline SweepArgs sweep_args(dim3* grid, int N, int64_t kblocks_total, int64_t per, double* C, int64_t ldc, int64_t rows, int ncols, const double* rs_up,
                            int64_t chunk_stride) {
    SweepArgs a;
    a.cblocks = g_sl.cblocks; a.tiles = grid->x; a.w8 = pair_w8(N);
    a.pfd = 1 << 30;
    if (const char* e = getenv("RNLA_I8_PFD")) { const int v = atoi(e); if (v > 0) a.pfd = v; }
    a.kblocks_total = kblocks_total; a.kblocks_per_chunk = per;
    a.C = C; a.ldc = ldc; a.rows = rows; a.ncols = ncols; a.rs_up = rs_up; a.cs_up = g_sl.cup.d(); a.chunk_stride = chunk_stride;
    grid->x = (unsigned)(2 * ((a.tiles + 1) / 2));
    return a;
}
template <bool TN, int PU, int G0, int NG, bool ADD>
rnla_status launch_mma2(dim3 grid, int N, int64_t kblocks_total, int64_t per, int flush, double* C, int64_t ldc, int64_t rows, int ncols,
                        const double* rs_up, int64_t chunk_stride) {
    Ctx& c = ctx();
    // pairs of groups in one MMA of double width where the sweep has an even number of groups and the accumulators fit (4 w8 <= 256)
    constexpr bool WIDE = NG % 2 == 0;
    static bool attr = false;
    if (!attr) {
        RNLA_CUDA(cudaFuncSetAttribute(i8_mma2_kernel<TN, PU, G0, NG, ADD, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_MAX));
        RNLA_CUDA(cudaFuncSetAttribute(i8_mma2_kernel<TN, PU, G0, NG, ADD, WIDE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_MAX));
        attr = true;
    }
    const SweepArgs a = sweep_args(&grid, N, kblocks_total, per, C, ldc, rows, ncols, rs_up, chunk_stride);
    int na, nbs;
    mma2_rings(PU, a.w8, &na, &nbs);
    static const bool wide_on = [] { const char* e = getenv("RNLA_I8_WIDE"); return !(e && e[0] == '0'); }();
    const bool wide = WIDE && wide_on && 4 * a.w8 <= 256;
    static const std::string kname = std::string("k:i8_mma<") + (TN ? "A^T B" : "A B") + ", planes " + std::to_string(PU) + ", groups " +
                                     std::to_string(G0) + ".." + std::to_string(G0 + NG - 1) + ">";
    CUtensorMap tmA, tmB;
    RNLA_TRY(sweep_tensor_maps(TN, PU, a.w8, kblocks_total, &tmA, &tmB));
    kernel_phase_begin(kname.c_str());
    if (wide) i8_mma2_kernel<TN, PU, G0, NG, ADD, WIDE><<<grid, MMA_THREADS, SMEM_MAX, c.stream>>>(tmA, tmB, a, na, nbs, flush);
    else i8_mma2_kernel<TN, PU, G0, NG, ADD, false><<<grid, MMA_THREADS, SMEM_MAX, c.stream>>>(tmA, tmB, a, na, nbs, flush);
    kernel_phase_end();
    ++g_kernel_launches;
    RNLA_CUDA(cudaGetLastError());
    return RNLA_OK;
}
// the sweeps of one product at the current precision
template <bool TN>
rnla_status run_sweeps(dim3 grid, int64_t kblocks_total, int64_t per, double* C, int64_t ldc, int64_t rows, int ncols,
                       const double* rs_up, int64_t chunk_stride) {
    const int f4 = g_flush_override ? g_flush_override : FLUSH_P4, f6 = g_flush_override ? g_flush_override : FLUSH_P6,
              f7 = g_flush_override ? g_flush_override : FLUSH_P7;
    if (g_planes == 7) {
        // 28 pairs: groups 0..2 need planes 0..2 only (6 pairs), groups 3..6 all seven (22 pairs): 3 + 7 planes of each operand enter
        // the SMs per product instead of 4 + 7 -- the kernels are bound by the L2 -> SM traffic
        RNLA_TRY((launch_mma2<TN, 3, 0, 3, false>(grid, ncols, kblocks_total, per, f4, C, ldc, rows, ncols, rs_up, chunk_stride)));
        return launch_mma2<TN, 7, 3, 4, true>(grid, ncols, kblocks_total, per, f7, C, ldc, rows, ncols, rs_up, chunk_stride);
    }
    RNLA_TRY((launch_mma2<TN, 4, 0, 4, false>(grid, ncols, kblocks_total, per, f4, C, ldc, rows, ncols, rs_up, chunk_stride)));
    if (g_planes == 6) return launch_mma2<TN, 6, 4, 2, true>(grid, ncols, kblocks_total, per, f6, C, ldc, rows, ncols, rs_up, chunk_stride);
    if (g_all_pairs) return launch_mma2<TN, 4, 4, 3, true>(grid, ncols, kblocks_total, per, f4, C, ldc, rows, ncols, rs_up, chunk_stride);
    return RNLA_OK;
}

}  // namespace

// thin operands wider than one 128-column MMA tile are processed tile by tile (each tile is another sweep over the planes of A)
constexpr int I8_MAX_N = 256;
bool i8_supported(int64_t m, int64_t n, int l) { return l >= 1 && l <= I8_MAX_N && n >= 1 && m >= 1 && m * n >= (int64_t)1 << 22; }
bool i8_active_for(const double* A, int64_t lda, int64_t m, int64_t n, int64_t N) {
    return g_active && g_sl.ready && g_sl.A == A && g_sl.lda == lda && g_sl.m == m && g_sl.n == n && N <= I8_MAX_N;
}
void i8_deactivate() { g_active = false; g_planes = 4; g_all_pairs = false; }
// precision of the products that follow: 4 planes (31-bit operands; all_pairs adds the second sweep, groups 4..6), 6 (47-bit) or
// 7 (55-bit, FP64-grade); capped by the planes the split of A stored
void i8_set_precision(int planes, bool all_pairs) {
    g_planes = std::min(planes >= 7 ? 7 : planes >= 6 ? 6 : 4, g_sl.planes >= 7 ? 7 : g_sl.planes >= 6 ? 6 : 4);
    g_all_pairs = all_pairs;
}
void i8_debug_flush(int stages) { g_flush_override = stages > 0 ? stages : 0; }
void i8_release() { g_active = false; g_sl.ready = false; }      // the workspace persists (see Persist)
void i8_free_workspace() {
    g_active = false; g_sl.ready = false;
    Persist* all[] = {&g_sl.img, &g_sl.up, &g_sl.down, &g_sl.bits, &g_sl.bimg, &g_sl.cbits, &g_sl.cup, &g_sl.cdown, &g_sl.part, &g_sl.flags};
    for (Persist* b : all) b->release();
}

// split A (m x n, lda) into `planes` (4, 6 or 7) tiled int8 digit planes; afterwards dev_gemm_nn / dev_gemm_tn with this A run on
// the integer tensor cores until i8_deactivate().  Three steps so that the host-buffer entry point can split each row block of A
// as it lands over PCIe (row maxima are per row: a block of whole rows is self-contained): begin, rows(r0, count)*, end.
rnla_status i8_prepare_begin(const double* A, int64_t lda, int64_t m, int64_t n, int planes) {
    Ctx& c = ctx();
    Sliced& s = g_sl;
    s.ready = false; g_active = false;
    planes = planes >= 7 ? 7 : planes >= 6 ? 6 : 4;
    s.A = A; s.lda = lda; s.m = m; s.n = n; s.planes = planes;
    s.rblocks = (m + 127) / 128; s.cblocks = (n + 127) / 128;
    const size_t img_bytes = (size_t)s.rblocks * s.cblocks * planes * APLANE;
    if (img_bytes > s.img.cap) {
        // Look before allocating: a failed 100 GB cudaMalloc (and the pool churn after it) costs hundreds of milliseconds per call
        // (measured on the 2-GPU shard of BASELINE config 3: 160 GB of A per GPU, no room for 140 GB of planes).  The images need
        // their own size plus headroom for the panels of the passes; cached blocks of the stream-ordered pool are given back first.
        const size_t headroom = (size_t)2 << 30;
        size_t freeb = 0, total = 0;
        RNLA_CUDA(cudaMemGetInfo(&freeb, &total));
        static size_t refused_bytes = 0, refused_free = 0;     // the last request that did not fit, and the free memory it saw after trimming
        if (img_bytes >= refused_bytes && refused_bytes && freeb <= refused_free + ((size_t)1 << 30))
            return fail(RNLA_ERR_COMPUTATION, "int8 passes: no room for the digit-plane workspace");
        if (freeb + s.img.cap < img_bytes + headroom) {
            cudaMemPool_t pool;
            if (cudaDeviceGetDefaultMemPool(&pool, c.device) == cudaSuccess) {
                cudaStreamSynchronize(c.stream);
                cudaMemPoolTrimTo(pool, 0);
            }
            RNLA_CUDA(cudaMemGetInfo(&freeb, &total));
        }
        if (freeb + s.img.cap < img_bytes + headroom) {
            refused_bytes = img_bytes; refused_free = freeb;
            return fail(RNLA_ERR_COMPUTATION, "int8 passes: no room for the digit-plane workspace");
        }
        refused_bytes = 0;
    }
    if (s.img.ensure(img_bytes) != cudaSuccess) {
        cudaGetLastError();
        i8_free_workspace();
        return fail(RNLA_ERR_COMPUTATION, "int8 passes: no room for the digit-plane workspace");
    }
    RNLA_CUDA(s.up.ensure((size_t)m * 8)); RNLA_CUDA(s.down.ensure((size_t)m * 8)); RNLA_CUDA(s.bits.ensure((size_t)m * 8));
    RNLA_CUDA(s.cbits.ensure(128 * 8)); RNLA_CUDA(s.cup.ensure(128 * 8)); RNLA_CUDA(s.cdown.ensure(128 * 8)); RNLA_CUDA(s.flags.ensure(16));
    RNLA_CUDA(cudaMemsetAsync(s.bits.p, 0, (size_t)m * 8, c.stream));
    RNLA_CUDA(cudaMemsetAsync(s.flags.p, 0, 16, c.stream));
    static bool attr = false;
    if (!attr) {
        RNLA_CUDA(cudaFuncSetAttribute(slice_a_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * SL_PLANE));
        RNLA_CUDA(cudaFuncSetAttribute(slice_a_kernel<6>, cudaFuncAttributeMaxDynamicSharedMemorySize, 6 * SL_PLANE));
        RNLA_CUDA(cudaFuncSetAttribute(slice_a_kernel<7>, cudaFuncAttributeMaxDynamicSharedMemorySize, 7 * SL_PLANE));
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
    scales_kernel<<<(unsigned)((count + 255) / 256), 256, 0, c.stream>>>(s.bits.as<unsigned long long>() + r0, count, s.up.d() + r0, s.down.d() + r0,
                                                                         s.flags.as<int>());
    if (phases) { phase_end(); phase_begin("i8:split(A)"); }
    const int64_t rb0 = r0 / 128, nrb = (count + 127) / 128;
    const unsigned grid = (unsigned)(nrb * s.cblocks * 4);
    if (s.planes == 7) slice_a_kernel<7><<<grid, 256, 7 * SL_PLANE, c.stream>>>(A, lda, m, n, s.down.d(), s.img.as<uint8_t>(), s.cblocks, rb0);
    else if (s.planes == 6) slice_a_kernel<6><<<grid, 256, 6 * SL_PLANE, c.stream>>>(A, lda, m, n, s.down.d(), s.img.as<uint8_t>(), s.cblocks, rb0);
    else slice_a_kernel<4><<<grid, 256, 4 * SL_PLANE, c.stream>>>(A, lda, m, n, s.down.d(), s.img.as<uint8_t>(), s.cblocks, rb0);
    if (phases) phase_end();
    g_kernel_launches += 3;
    RNLA_CUDA(cudaGetLastError());
    return RNLA_OK;
}
// Ends the split.  The fixed-point representation needs finite rows whose maxima can be scaled: Inf / NaN anywhere in A, or a
// non-zero row maximum below 2^-959, are reported by the scale kernel; such a matrix keeps the FP64 kernels (which report
// non-finite input themselves).  One 4-byte read-back per driver call.
rnla_status i8_prepare_end(bool* usable) {
    Ctx& c = ctx();
    int h[2] = {0, 0};
    RNLA_CUDA(cudaMemcpyAsync(h, g_sl.flags.p, 4, cudaMemcpyDeviceToHost, c.stream));
    RNLA_CUDA(cudaStreamSynchronize(c.stream));
    const bool ok = h[0] == 0;
    g_sl.ready = ok; g_active = ok;
    if (usable) *usable = ok;
    return RNLA_OK;
}
rnla_status i8_prepare(const double* A, int64_t lda, int64_t m, int64_t n, int planes, bool* usable) {
    if (usable) *usable = false;
    RNLA_TRY(i8_prepare_begin(A, lda, m, n, planes));
    RNLA_TRY(i8_prepare_rows(0, m, true));
    return i8_prepare_end(usable);
}

// C (m x N) = A * B (n x N) on the split A
static rnla_status i8_gemm_nn_tile(const ThinSrc& src, int64_t N, double* C, int64_t ldc) {
    Sliced& s = g_sl;
    const int64_t kblocks = 2 * s.cblocks;
    RNLA_TRY(slice_b(src, s.n, (int)N, nullptr, kblocks, g_planes));
    return run_sweeps<false>(dim3((unsigned)s.rblocks, 1), kblocks, kblocks, C, ldc, s.m, (int)N, s.up.d(), 0);
}
rnla_status i8_gemm_nn(const double* B, int64_t ldb, int64_t N, double* C, int64_t ldc) {
    for (int64_t c0 = 0; c0 < N; c0 += BN)
        RNLA_TRY(i8_gemm_nn_tile(ThinSrc{B + c0 * ldb, ldb, -1, 0, 0, 0}, std::min<int64_t>(BN, N - c0), C + c0 * ldc, ldc));
    return RNLA_OK;
}
// C (m x N) = A * Omega (n x N), Omega(k, c) = the Philox entry map of rng.cuh: Omega is never materialised in FP64
rnla_status i8_gemm_nn_omega(int dist, uint64_t seed, uint32_t stream, int64_t N, double* C, int64_t ldc) {
    for (int64_t c0 = 0; c0 < N; c0 += BN)
        RNLA_TRY(i8_gemm_nn_tile(ThinSrc{nullptr, 0, dist, seed, stream, (uint32_t)c0}, std::min<int64_t>(BN, N - c0), C + c0 * ldc, ldc));
    return RNLA_OK;
}

// Row chunks of A^T Q for the pair kernel: (pairs of column tiles) x chunks units run on sms / 2 pair slots, every wave as long as one
// chunk; pick the chunk count whose last wave is fullest, net of a per-chunk cost of about six stages (pipeline fill, epilogue).
// Headline shape: 79 pairs x 14 chunks = 1106 units = 14.95 waves of 74.
static int64_t pair_chunks(int64_t pairs, int64_t kblocks, int sms) {
    const int64_t slots = std::max(1, sms / 2);
    int64_t best = 1;
    double best_eff = 0.0;
    for (int64_t nc = 1; nc <= std::min<int64_t>(kblocks, 64); ++nc) {
        const int64_t per = (kblocks + nc - 1) / nc;
        if ((kblocks + per - 1) / per != nc) continue;
        const int64_t waves = (pairs * nc + slots - 1) / slots;
        const double eff = (double)(pairs * kblocks) / ((double)waves * (double)per * (double)slots) * (double)per / ((double)per + 6.0);
        if (eff > best_eff * 1.005) { best_eff = eff; best = nc; }
    }
    return best;
}
// Z (n x N) = A^T * Q (m x N) on the split A (local rows only; the caller all-reduces)
static rnla_status i8_gemm_tn_tile(const double* Q, int64_t ldq, int64_t N, double* Z, int64_t ldz) {
    Ctx& c = ctx();
    Sliced& s = g_sl;
    const int64_t kblocks = 2 * s.rblocks;
    RNLA_TRY(slice_b(ThinSrc{Q, ldq, -1, 0, 0, 0}, s.m, (int)N, s.up.d(), kblocks, g_planes));
    // row chunks: enough (column block, chunk) units to fill the SMs a few times over; the partials are summed in chunk order
    int64_t nchunks = pair_chunks((s.cblocks + 1) / 2, kblocks, c.sms);
    const int64_t per = (kblocks + nchunks - 1) / nchunks;
    nchunks = (kblocks + per - 1) / per;
    const int64_t stride = s.n * N;
    RNLA_CUDA(s.part.ensure((size_t)nchunks * stride * 8));
    RNLA_TRY(run_sweeps<true>(dim3((unsigned)s.cblocks, (unsigned)nchunks), kblocks, per, s.part.d(), s.n, s.n, (int)N, nullptr, stride));
    i8_tn_reduce_kernel<<<(unsigned)std::min<int64_t>(148 * 8, (stride + 255) / 256), 256, 0, c.stream>>>(s.part.d(), (int)nchunks, stride, s.n, (int)N,
                                                                                                           s.cup.d(), Z, ldz);
    ++g_kernel_launches;
    RNLA_CUDA(cudaGetLastError());
    return RNLA_OK;
}
rnla_status i8_gemm_tn(const double* Q, int64_t ldq, int64_t N, double* Z, int64_t ldz) {
    for (int64_t c0 = 0; c0 < N; c0 += BN)
        RNLA_TRY(i8_gemm_tn_tile(Q + c0 * ldq, ldq, std::min<int64_t>(BN, N - c0), Z + c0 * ldz, ldz));
    return RNLA_OK;
}

}  // namespace rnla
