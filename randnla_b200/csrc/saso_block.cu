// K6b: block sparse-sign sketch  A_sk (d x n) = S A,  S (d x m) with exactly `zeta` non-zeros +-1/sqrt(zeta) per column.
//
// Replaces the dense `&s * a` of the reference's sketch step (src/sketch_and_precondition.rs:50-52, 106-107,
// 173-176); the reference has no sparse operator, so the operator is defined here.
//
// Why not the textbook SASO (zeta independent row indices per column, kind RNLA_SKETCH_SASO)?  Per element of A it
// needs zeta read-modify-writes of shared-memory accumulators; f64 shared atomics are CAS loops (ATOMS.CAST.SPIN)
// and even conflict-free RMWs would cost 2*zeta shared accesses per 8 bytes streamed -- 3x the HBM time at zeta = 8.
// The operator below keeps "zeta non-zeros per column, signs i.i.d." but places them so the accumulators live in
// REGISTERS, nothing is atomic and the summation order is fixed:
//
//   * zeta = g * w: the non-zeros of a column come as g groups of w consecutive rows.  The d output rows are cut into
//     g stripes (the OSNAP block construction), each stripe into nbs = d / zeta blocks of w rows;
//   * rows of A are cut into chunks of SB_R = 2048 consecutive GLOBAL rows, chunk q = gr / SB_R, x = gr % SB_R;
//   * per (chunk, stripe) a keyed bijection sigma on [0, SB_R) deals the rows to the blocks like a shuffled deck: slot y
//     holds row x = sigma(y) and belongs to block (y + off) mod nbs of the stripe.  Every block receives floor or ceil of
//     SB_R / nbs rows of every chunk: the load of the owning threads is balanced by construction (no Poisson tail);
//   * sigma(16 yh + yl) = 16 tau(yh) + (mo * yl + rho(yh)) mod 16 with tau a three-round multiply-xorshift bijection on
//     7 bits: sixteen consecutive slots read sixteen different shared-memory banks;
//   * signs: bit (t*w + r) of one Philox word per global row.
//
// w = 1 is the OSNAP block construction itself (statistically the textbook SASO); larger w reads each element of A from
// shared memory only g = zeta / w times.  Second moments are those of the textbook operator for every w; the tails are
// not -- w = 8 (one block per column) loses rank on coherent inputs at d = 4n and is not offered; w <= 4 holds
// (tests/test_oracle_pinning.py::test_block_sparse_sign_embedding_quality).
//
// Kernel: one CTA per work-list entry = (column group of CB columns) x (chunk range), 512 threads.  Thread 0 streams, per chunk, the CB
// column segments (16 KB each, contiguous) and the chunk's slot tables (4 KB per stripe) with cp.async.bulk (TMA
// engine) into a ring of stages with one `full` mbarrier each; the last warp to finish a chunk refills its stage (a
// dedicated producer warp would cap the kernel at 96 registers per thread); all 512 threads each own BPT blocks (BPT * w * CB f64 accumulators
// in registers), walk their slots of the chunk, read A(x, c) from shared memory and apply the w signed updates as DFMAs.
// HBM traffic: A exactly once (8 m n bytes) + d n 8 written; the slot tables (2 g bytes per row) are L2-resident.
#include "drivers.cuh"
#include "gemm.cuh"
#include "panel.cuh"
#include "ptx.cuh"
#include "rng.cuh"
#include <algorithm>
#include <cmath>
#include <type_traits>
#include <vector>

namespace rnla {

namespace {

constexpr int SB_R = 2048;            // rows per chunk
constexpr int SB_T = 512;             // consumer threads
constexpr int SB_TABW = SB_R + 8;     // uint16 entries per (chunk, stripe) record of the slot table (4112 bytes, 16-byte aligned)
typedef uint16_t sbtab_t;             // entry = x (11 bits) | signs (w <= 4 bits) << 11 ; entry SB_R of the record = block offset
constexpr uint32_t STREAM_SASO_BLOCK = 5u;
constexpr int SB_MAXACC = 32;         // f64 accumulators per thread
constexpr int SB_MAXSTAGES = 4;
constexpr int SB_SMEM = 200 * 1024;

struct SbKey { uint32_t a0, c0, a1, c1, a2, c2, off, mo, rk; };

__device__ __forceinline__ SbKey sb_key(uint64_t seed, uint64_t q, int t, uint32_t nbs) {
    const u32x4 k = philox4x32_10((uint32_t)q, (uint32_t)(q >> 32), (uint32_t)(2 * t), STREAM_SASO_BLOCK, (uint32_t)seed, (uint32_t)(seed >> 32));
    const u32x4 k2 = philox4x32_10((uint32_t)q, (uint32_t)(q >> 32), (uint32_t)(2 * t + 1), STREAM_SASO_BLOCK, (uint32_t)seed, (uint32_t)(seed >> 32));
    SbKey s;
    s.a0 = k.x | 1u; s.c0 = k.x >> 16;
    s.a1 = k.y | 1u; s.c1 = k.y >> 16;
    s.a2 = k.z | 1u; s.c2 = k.z >> 16;
    s.off = ((k.w % nbs) >> 4) << 4;
    s.mo = (k2.x & 15u) | 1u;
    s.rk = k2.y | 1u;
    return s;
}
__device__ __forceinline__ uint32_t sb_sigma(const SbKey& k, uint32_t y) {
    const uint32_t yh = y >> 4, yl = y & 15u;
    uint32_t h = yh;
    h = (h * k.a0 + k.c0) & 127u; h ^= h >> 3;
    h = (h * k.a1 + k.c1) & 127u; h ^= h >> 4;
    h = (h * k.a2 + k.c2) & 127u; h ^= h >> 2;
    const uint32_t rho = (((yh + 1u) * k.rk) >> 11) & 15u;
    return (h << 4) | ((yl * k.mo + rho) & 15u);
}

// slot table: tab[((q - q_first) * g + t) * SB_TABW + y] = x | signs << 11 ; entry SB_R of the record = off
__global__ void __launch_bounds__(256)
sb_table_kernel(uint64_t seed, int64_t q_first, int64_t nchunks, uint32_t nbs, int g, int w, sbtab_t* __restrict__ tab) {
    const int64_t total = nchunks * g * SB_R;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t rec = i / SB_R;
        const uint32_t y = (uint32_t)(i - rec * SB_R);
        const int64_t ql = rec / g;
        const int t = (int)(rec - ql * g);
        const uint64_t q = (uint64_t)(q_first + ql);
        const SbKey k = sb_key(seed, q, t, nbs);
        const uint32_t x = sb_sigma(k, y);
        const uint64_t gr = q * SB_R + x;
        const u32x4 sg = philox4x32_10((uint32_t)gr, (uint32_t)(gr >> 32), 1u, STREAM_SASO_BLOCK, (uint32_t)seed, (uint32_t)(seed >> 32));
        tab[rec * SB_TABW + y] = (sbtab_t)(x | (((sg.x >> (t * w)) & ((1u << w) - 1u)) << 11));
        if (y < 8) tab[rec * SB_TABW + SB_R + y] = (sbtab_t)k.off;
    }
}

struct SbArgs {
    const double* A; int64_t lda; int64_t m_local; int64_t n; int64_t row_off;
    const sbtab_t* tab; int64_t q_first; int64_t nchunks;
    const int4* desc;             // per CTA: column group, first chunk, end chunk, partial slot (-1: writes A_sk directly)
    int nbs; int g; int parts;    // parts > 1 only when g * nbs < SB_T
    int stages;
    double* out; int64_t ldo;     // A_sk
    double* partial; int64_t d_used;   // slot s: d_used x CB doubles
    int ncg;                      // column groups
    double scale;
};

// The W sign bits of a slot (bits 11.. of the table entry) -> the top bytes 0x3F / 0xBF of +-1.0, one per byte of a
// register: nib * 0x10204080 moves bit i to bit 8i+7 (the partial products land on distinct bit positions, so nothing
// carries), then one LOP3 masks and ors.  sign_hi() picks byte r into the high word of a double (0x3FF00000 / 0xBFF00000).
__device__ __forceinline__ uint32_t sign_bytes(uint32_t e) {
    const uint32_t t = (e >> 11) * 0x10204080u;
    uint32_t r;
    asm("lop3.b32 %0, %1, %2, %3, 0xEA;" : "=r"(r) : "r"(t), "r"(0x80808080u), "r"(0x3F3F3F3Fu));
    return r;
}
template <int R> __device__ __forceinline__ double sign_f64(uint32_t sb) {
    // bytes of the result, high to low: sb.byte[R], 0xF0, 0x00, 0x00
    const uint32_t hi = __byte_perm(sb, 0x00F00000u, (R << 12) | 0x0644);
    return __hiloint2double((int)hi, 0);
}

template <int OFF> __device__ __forceinline__ double lds_f64(uint32_t addr) {       // ld.shared with an immediate offset
    double v;
    asm volatile("ld.shared.f64 %0, [%1+%2];" : "=d"(v) : "r"(addr), "n"(OFF));
    return v;
}

template <int W, int BPT, int CB, bool TMA>
__global__ void __launch_bounds__(SB_T, 1)
saso_block_kernel(const SbArgs a) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ uint64_t full_bar[SB_MAXSTAGES];
    __shared__ int done_cnt[SB_MAXSTAGES];      // warps that have finished the chunk in a stage; the last one refills it
    const int tab_bytes = a.g * SB_TABW * (int)sizeof(sbtab_t);
    const int stage_bytes = SB_R * 8 * CB + tab_bytes;
    const int stages = a.stages;

    const int tid = threadIdx.x;
    const int4 dsc = a.desc[blockIdx.x];
    const int cg = dsc.x;
    const int64_t c0 = (int64_t)cg * CB;
    const int cbv = (int)min((int64_t)CB, a.n - c0);
    const int64_t ch0 = dsc.y;
    const int nch = dsc.z - dsc.y;

    // columns of the tile that have no source stay zero for the whole kernel
    if (cbv < CB) {
        for (int s = 0; s < stages; ++s) {
            double* tile = reinterpret_cast<double*>(smem_raw + (size_t)s * stage_bytes);
            for (int i = tid; i < (CB - cbv) * SB_R; i += blockDim.x) tile[cbv * SB_R + i] = 0.0;
        }
    }
    if (TMA) {
        if (tid == 0) {
            for (int s = 0; s < stages; ++s) { mbar_init(&full_bar[s], 1); done_cnt[s] = 0; }
            mbar_fence_init();
        }
    }
    __syncthreads();

    // valid local rows of chunk ql: global rows [q*R, q*R + R) intersected with [row_off, row_off + m_local)
    auto chunk_rows = [&](int64_t ql, int& xlo, int& xhi, int64_t& lr0) {
        const int64_t gbase = (a.q_first + ql) * SB_R;
        const int64_t g0 = max(gbase, a.row_off), g1 = min(gbase + SB_R, a.row_off + a.m_local);
        xlo = (int)(g0 - gbase); xhi = (int)(g1 - gbase); lr0 = g0 - a.row_off;
    };

    // producer role: chunk ci goes to stage ci % stages.  Thread 0 fills the ring once; afterwards the LAST warp to finish
    // a chunk (shared-memory counter) issues the copy that refills the stage, so nobody ever waits for a free stage.
    auto issue = [&](int ci) {
        const int s = ci % stages;
        int xlo, xhi; int64_t lr0;
        chunk_rows(ch0 + ci, xlo, xhi, lr0);
        unsigned char* st = smem_raw + (size_t)s * stage_bytes;
        double* tile = reinterpret_cast<double*>(st);
        const uint32_t colbytes = (uint32_t)(xhi - xlo) * 8u;
        mbar_arrive_expect_tx(&full_bar[s], colbytes * (uint32_t)cbv + (uint32_t)tab_bytes);
        for (int c = 0; c < cbv; ++c)
            bulk_g2s(tile + (size_t)c * SB_R + xlo, a.A + (c0 + c) * a.lda + lr0, colbytes, &full_bar[s]);
        bulk_g2s(st + (size_t)SB_R * 8 * CB, a.tab + (ch0 + ci) * a.g * SB_TABW, (uint32_t)tab_bytes, &full_bar[s]);
    };
    if (TMA && tid == 0)
        for (int ci = 0; ci < stages && ci < nch; ++ci) issue(ci);

    // ---------------- consumers
    double acc[BPT][W][CB];
#pragma unroll
    for (int i = 0; i < BPT; ++i)
#pragma unroll
        for (int r = 0; r < W; ++r)
#pragma unroll
            for (int c = 0; c < CB; ++c) acc[i][r][c] = 0.0;

    const int nbs = a.nbs;
    const int nbt = a.g * nbs;                     // blocks over all stripes
    int part = 0, step = nbs;
    bool active = true;
    int B0 = tid;
    if (a.parts > 1) {
        B0 = tid % nbt; part = tid / nbt; active = part < a.parts; step = a.parts * nbs;
        if (!active) part = 0;                      // idle threads shadow part 0; their sums are never written
    }
    int tabo[BPT], bb[BPT];                        // table section offset (words) and block index inside the stripe
#pragma unroll
    for (int i = 0; i < BPT; ++i) {
        const int B = B0 + i * SB_T;
        const int t = min(B / nbs, a.g - 1);
        tabo[i] = t * SB_TABW; bb[i] = min(B - t * nbs, nbs - 1);   // blocks past the end shadow the last one (never written)
    }

    const int nfull = SB_R / step;               // rounds in which every lane has a slot; one partial round follows
    int s = 0; uint32_t ph = 0;
    for (int ci = 0; ci < nch; ++ci) {
        unsigned char* st = smem_raw + (size_t)s * stage_bytes;
        double* tile = reinterpret_cast<double*>(st);
        sbtab_t* tabs = reinterpret_cast<sbtab_t*>(st + (size_t)SB_R * 8 * CB);
        int xlo, xhi; int64_t lr0;
        chunk_rows(ch0 + ci, xlo, xhi, lr0);
        if (TMA) {
            mbar_wait(&full_bar[s], ph);
        } else {
            // cooperative loads for inputs the bulk copy cannot take (odd leading dimension / offsets)
            asm volatile("bar.sync 1, %0;" ::"n"(SB_T));
            for (int c = 0; c < cbv; ++c)
                for (int x = tid; x < SB_R; x += SB_T)
                    tile[(size_t)c * SB_R + x] = (x >= xlo && x < xhi) ? ldg_stream(a.A + (c0 + c) * a.lda + lr0 + (x - xlo)) : 0.0;
            for (int i = tid; i < a.g * SB_TABW; i += SB_T) tabs[i] = a.tab[(ch0 + ci) * a.g * SB_TABW + i];
            asm volatile("bar.sync 1, %0;" ::"n"(SB_T));
        }
        if (xlo != 0 || xhi != SB_R) {
            // first / last chunk of a shard: rows outside the shard contribute zero (no range checks in the slot loop)
            if (TMA) {
                for (int c = 0; c < cbv; ++c)
                    for (int x = tid; x < SB_R; x += SB_T)
                        if (x < xlo || x >= xhi) tile[(size_t)c * SB_R + x] = 0.0;
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            }
            asm volatile("bar.sync 1, %0;" ::"n"(SB_T));
        }
        {
            // slots of this chunk: block i walks y_i, y_i + step, ...  The first nfull rounds are valid for every lane
            // (threads without a block accumulate into registers that are never written back), so they carry no
            // predicates and the BPT independent chains interleave; one tail round takes the remaining slots.
            // The loop is issue-bound at d = 8000 (48 accumulators, 16 warps per SM), so a slot is kept to the instructions it
            // needs: the table is walked through a running shared-memory address per block (one add per round), the element offset is
            // one shift-and-mask of the entry, the three loads of a slot use immediate offsets from one address.
            const uint32_t tile_a = smem_u32(tile), tabs_a = smem_u32(tabs);
            uint32_t ta[BPT];                      // shared address of the next table entry of block i
            unsigned tailv = 0u;                   // bit i: block i has a slot in the tail round
#pragma unroll
            for (int i = 0; i < BPT; ++i) {
                int yy = bb[i] - (int)tabs[tabo[i] + SB_R]; if (yy < 0) yy += nbs;
                yy += part * nbs;
                ta[i] = tabs_a + 2u * (uint32_t)(tabo[i] + yy);
                if (yy + nfull * step < SB_R) tailv |= 1u << i;
            }
            const uint32_t step2 = 2u * (uint32_t)step;
            auto slot = [&](int i) {
                uint32_t e;
                asm volatile("ld.shared.u16 %0, [%1];" : "=r"(e) : "r"(ta[i]));
                const uint32_t src = tile_a + ((e << 3) & (uint32_t)((SB_R - 1) * 8));
                double v[CB];
                if constexpr (CB > 0) v[0] = lds_f64<0>(src);
                if constexpr (CB > 1) v[1 % CB] = lds_f64<SB_R * 8>(src);
                if constexpr (CB > 2) v[2 % CB] = lds_f64<2 * SB_R * 8>(src);
                if constexpr (CB > 3) v[3 % CB] = lds_f64<3 * SB_R * 8>(src);
                const uint32_t sb = sign_bytes(e);
                if constexpr (W >= 1) {
                    const double sg = sign_f64<0>(sb);
#pragma unroll
                    for (int c = 0; c < CB; ++c) acc[i][0][c] = fma(sg, v[c], acc[i][0][c]);
                }
                if constexpr (W >= 2) {
                    const double sg = sign_f64<1>(sb);
#pragma unroll
                    for (int c = 0; c < CB; ++c) acc[i][1 % W][c] = fma(sg, v[c], acc[i][1 % W][c]);
                }
                if constexpr (W >= 4) {
                    const double s2 = sign_f64<2>(sb), s3 = sign_f64<3>(sb);
#pragma unroll
                    for (int c = 0; c < CB; ++c) {
                        acc[i][2 % W][c] = fma(s2, v[c], acc[i][2 % W][c]);
                        acc[i][3 % W][c] = fma(s3, v[c], acc[i][3 % W][c]);
                    }
                }
                ta[i] += step2;
            };
#pragma unroll 1
            for (int it = 0; it < nfull; ++it) {
#pragma unroll
                for (int i = 0; i < BPT; ++i) slot(i);
            }
            if (__any_sync(0xffffffffu, tailv != 0u)) {
#pragma unroll
                for (int i = 0; i < BPT; ++i)
                    if (tailv & (1u << i)) slot(i);
            }
        }
        if (TMA) {
            __syncwarp();
            if ((tid & 31) == 0) {
                __threadfence_block();
                if (atomicAdd(&done_cnt[s], 1) == SB_T / 32 - 1) {
                    done_cnt[s] = 0;
                    if (ci + stages < nch) issue(ci + stages);
                }
            }
        }
        if (++s == stages) { s = 0; ph ^= 1u; }
    }

    // a CTA that covers all chunks of its column group writes A_sk; a fragment writes its slot (sb_fixup_kernel adds them)
    double* out = dsc.w < 0 ? a.out + c0 * a.ldo : a.partial + (int64_t)dsc.w * a.d_used * CB;
    const int64_t ldo = dsc.w < 0 ? a.ldo : a.d_used;
    if (a.parts > 1) {
        // fixed-order reduction over the parts through shared memory (the ring is drained: every full barrier was waited on)
        asm volatile("bar.sync 1, %0;" ::"n"(SB_T));
        double* red = reinterpret_cast<double*>(smem_raw);
        if (active) {
#pragma unroll
            for (int r = 0; r < W; ++r)
#pragma unroll
                for (int c = 0; c < CB; ++c) red[((size_t)(part * nbt + B0) * W + r) * CB + c] = acc[0][r][c];
        }
        asm volatile("bar.sync 1, %0;" ::"n"(SB_T));
        if (active && part == 0) {
#pragma unroll
            for (int r = 0; r < W; ++r)
#pragma unroll
                for (int c = 0; c < CB; ++c) {
                    double sum = 0.0;
                    for (int p = 0; p < a.parts; ++p) sum += red[((size_t)(p * nbt + B0) * W + r) * CB + c];
                    if (c < cbv) out[(int64_t)B0 * W + r + c * ldo] = sum * a.scale;
                }
        }
    } else {
#pragma unroll
        for (int i = 0; i < BPT; ++i) {
            const int B = B0 + i * SB_T;
            if (B < nbt) {
#pragma unroll
                for (int c = 0; c < CB; ++c)
                    if (c < cbv) {
#pragma unroll
                        for (int r = 0; r < W; ++r) out[(int64_t)B * W + r + c * ldo] = acc[i][r][c] * a.scale;
                    }
            }
        }
    }
}

// A column group whose chunks were shared by several CTAs: add their fragments in chunk order (= slot list order)
__global__ void __launch_bounds__(256)
sb_fixup_kernel(const double* __restrict__ partial, int64_t d_used, int CB, const int* __restrict__ fix_cg,
                const int* __restrict__ fix_off, const int* __restrict__ fix_cnt, const int* __restrict__ slots,
                double* __restrict__ out, int64_t ldo, int64_t n) {
    const int e = blockIdx.y;
    const int cg = fix_cg[e], off = fix_off[e], cnt = fix_cnt[e];
    const int cbv = (int)min((int64_t)CB, n - (int64_t)cg * CB);
    const int64_t total = d_used * cbv;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        double s = 0.0;
        for (int k = 0; k < cnt; ++k) s += partial[(int64_t)slots[off + k] * d_used * CB + i];
        out[(i % d_used) + ((int64_t)cg * CB + i / d_used) * ldo] = s;
    }
}

template <int W, int BPT, int CB>
cudaError_t sb_launch(SbArgs a, int grid, bool tma, cudaStream_t st) {
    const int stage_bytes = SB_R * 8 * CB + a.g * SB_TABW * (int)sizeof(sbtab_t);
    a.stages = std::min(SB_MAXSTAGES, SB_SMEM / stage_bytes);
    if (a.stages < 2) return cudaErrorInvalidValue;
    const size_t smem = (size_t)a.stages * stage_bytes;
    cudaError_t e;
    if (tma) {
        auto kern = saso_block_kernel<W, BPT, CB, true>;
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SB_SMEM);
        if (e != cudaSuccess) return e;
        kern<<<grid, SB_T, smem, st>>>(a);
    } else {
        auto kern = saso_block_kernel<W, BPT, CB, false>;
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SB_SMEM);
        if (e != cudaSuccess) return e;
        kern<<<grid, SB_T, smem, st>>>(a);
    }
    ++g_kernel_launches;
    return cudaGetLastError();
}

template <int W, int BPT>
cudaError_t sb_dispatch_cb(const SbArgs& a, int cb, int grid, bool tma, cudaStream_t st) {
    constexpr int CBMAX = SB_MAXACC / (W * BPT);
    if constexpr (CBMAX >= 4) { if (cb == 4) return sb_launch<W, BPT, 4>(a, grid, tma, st); }
    if constexpr (W == 4 && BPT == 4) { if (cb == 3) return sb_launch<W, BPT, 3>(a, grid, tma, st); }   // 48 accumulators: 126 registers
    if constexpr (CBMAX >= 2) { if (cb == 2) return sb_launch<W, BPT, 2>(a, grid, tma, st); }
    if (cb == 1) return sb_launch<W, BPT, 1>(a, grid, tma, st);
    return cudaErrorInvalidValue;
}

template <int W>
cudaError_t sb_dispatch_bpt(const SbArgs& a, int bpt, int cb, int grid, bool tma, cudaStream_t st) {
    if constexpr (W * 1 <= SB_MAXACC) { if (bpt == 1) return sb_dispatch_cb<W, 1>(a, cb, grid, tma, st); }
    if constexpr (W * 2 <= SB_MAXACC) { if (bpt == 2) return sb_dispatch_cb<W, 2>(a, cb, grid, tma, st); }
    if constexpr (W * 4 <= SB_MAXACC) { if (bpt == 4) return sb_dispatch_cb<W, 4>(a, cb, grid, tma, st); }
    if constexpr (W * 8 <= SB_MAXACC) { if (bpt == 8) return sb_dispatch_cb<W, 8>(a, cb, grid, tma, st); }
    if constexpr (W * 16 <= SB_MAXACC) { if (bpt == 16) return sb_dispatch_cb<W, 16>(a, cb, grid, tma, st); }
    if constexpr (W * 32 <= SB_MAXACC) { if (bpt == 32) return sb_dispatch_cb<W, 32>(a, cb, grid, tma, st); }
    return cudaErrorInvalidValue;
}

}  // namespace

// host logic of the launch, separated so that it can be checked without a GPU (rnla_plan_saso_block, tests/test_host_logic.py)
void saso_block_worklist(int ncg, int64_t nchunks, int sms_in, std::vector<SbFrag>& work, std::vector<int>& fix_cg,
                         std::vector<int>& fix_off, std::vector<int>& fix_cnt, std::vector<int>& slots) {
    const int sms = std::max(sms_in, 1);
    const int nwhole = nchunks >= 16 ? (ncg / sms) * sms : ncg;      // short inputs are not worth fragmenting
    for (int gidx = 0; gidx < nwhole; ++gidx) work.push_back({gidx, 0, (int)nchunks, -1});
    const int64_t left = (int64_t)(ncg - nwhole) * nchunks;
    std::vector<SbFrag> frags;
    for (int k = 0; k < sms && left > 0; ++k) {
        int64_t t = left * k / sms;
        const int64_t t1 = left * (k + 1) / sms;
        while (t < t1) {
            const int64_t gl = t / nchunks, end = std::min(t1, (gl + 1) * nchunks);
            const int cgi = nwhole + (int)gl, lo = (int)(t - gl * nchunks), hi = (int)(end - gl * nchunks);
            if (lo == 0 && hi == (int)nchunks) frags.push_back({cgi, lo, hi, -1});
            else {
                const int slot = (int)slots.size();
                if (fix_cg.empty() || fix_cg.back() != cgi) { fix_cg.push_back(cgi); fix_off.push_back(slot); fix_cnt.push_back(0); }
                slots.push_back(slot); ++fix_cnt.back();
                frags.push_back({cgi, lo, hi, slot});
            }
            t = end;
        }
    }
    std::stable_sort(frags.begin(), frags.end(), [](const SbFrag& x, const SbFrag& y) { return x.hi - x.lo > y.hi - y.lo; });
    work.insert(work.end(), frags.begin(), frags.end());
}

// shape decisions of saso_block_apply: returns false if (d, zeta, w) is not supported
bool saso_block_shape(int64_t d, int zeta, int w, int64_t n, int* bpt_out, int* cb_out, int* parts_out) {
    if (zeta != 1 && zeta != 2 && zeta != 4 && zeta != 8) return false;
    if (w == 0) w = std::min(zeta, 4);
    if ((w != 1 && w != 2 && w != 4) || w > zeta || d < zeta) return false;
    const int g = zeta / w;
    const int64_t nbt64 = (d / zeta) * g;
    int bpt = 1;
    while ((int64_t)bpt * SB_T < nbt64) bpt *= 2;
    if ((int64_t)bpt * w > SB_MAXACC) return false;
    int cb = SB_MAXACC / (bpt * w);
    cb = cb >= 4 ? 4 : cb >= 2 ? 2 : 1;
    if (bpt * w == 16 && w == 4 && n >= 3) cb = 3;   // w < 4 spills at 48 accumulators
    while (cb > 1 && cb / 2 >= n) cb /= 2;
    while (cb > 1 && SB_SMEM / (SB_R * 8 * cb + g * SB_TABW * (int)sizeof(sbtab_t)) < 2) cb /= 2;
    *bpt_out = bpt; *cb_out = cb; *parts_out = nbt64 < SB_T ? (int)(SB_T / nbt64) : 1;
    return true;
}

// A_sk (d x n, fully overwritten) = S A_local for the block sparse-sign operator with zeta = g * w non-zeros per column.
// Rows [0, m_local) of A are global rows [row_off, row_off + m_local).  w = 0 selects the default width min(zeta, 4).
rnla_status saso_block_apply(uint64_t seed, int64_t d, int zeta, int w, const double* A, int64_t lda, int64_t m_local,
                             int64_t n, int64_t row_off, double* Ask, int64_t ldk) {
    Ctx& c = ctx();
    cudaStream_t st = c.stream;
    if (zeta != 1 && zeta != 2 && zeta != 4 && zeta != 8)
        return fail(RNLA_ERR_INVALID_PARAMETERS, "block SASO: zeta must be 1, 2, 4 or 8");
    if (w == 0) w = std::min(zeta, 4);
    if ((w != 1 && w != 2 && w != 4) || w > zeta)
        return fail(RNLA_ERR_INVALID_PARAMETERS, "block SASO: block width must be 1, 2 or 4 and divide zeta");
    if (d < zeta) return fail(RNLA_ERR_INVALID_DIMENSIONS, "block SASO: sketch dimension d must be >= zeta");
    const int g = zeta / w;
    const int64_t nbs64 = d / zeta;               // blocks per stripe
    const int64_t nbt64 = nbs64 * g;              // blocks over all stripes
    int bpt = 1, cb = 1, parts = 1;
    if (!saso_block_shape(d, zeta, w, n, &bpt, &cb, &parts))
        return fail(RNLA_ERR_INVALID_DIMENSIONS, "block SASO: sketch dimension d must be <= 16384");
    const int ncg = (int)((n + cb - 1) / cb);
    const int64_t d_used = nbt64 * w;

    // rows d_used .. d-1 (d not a multiple of zeta) are structurally zero
    if (d_used != d || m_local <= 0)
        RNLA_CUDA(axpby_matrix(0.0, nullptr, 0, 0.0, nullptr, 0, Ask, ldk, d, n, st));
    if (m_local <= 0) return RNLA_OK;

    const int64_t q_first = row_off / SB_R;
    const int64_t q_last = (row_off + m_local - 1) / SB_R;
    const int64_t nchunks = q_last - q_first + 1;

    DevBuf tab, partial, meta;
    RNLA_CUDA(tab.alloc((size_t)nchunks * g * SB_TABW * sizeof(sbtab_t)));
    {
        const int64_t total = nchunks * g * SB_R;
        const int blocks = (int)std::min<int64_t>((total + 255) / 256, 148 * 8);
        sb_table_kernel<<<blocks, 256, 0, st>>>(seed, q_first, nchunks, (uint32_t)nbs64, g, w, tab.as<sbtab_t>());
        ++g_kernel_launches;
        RNLA_CUDA(cudaGetLastError());
    }
    // Work list, one CTA per entry, one CTA per SM at a time.  Whole column groups fill as many complete waves as they
    // can; the chunks of the remaining groups are cut into `sms` equal ranges and every range becomes one or two
    // fragments (column group, chunk range) with a partial-result slot.  Entries are ordered longest first, so the
    // hardware's in-order dispatch packs the fragments of the last wave tightly (longest-processing-time rule): every SM
    // streams nearly the same number of bytes, and only the fragments cost extra traffic (their slots).
    std::vector<SbFrag> work;
    std::vector<int> fix_cg, fix_off, fix_cnt, slots;
    saso_block_worklist(ncg, nchunks, c.sms, work, fix_cg, fix_off, fix_cnt, slots);
    const int nfix = (int)fix_cg.size(), nslots = (int)slots.size(), grid = (int)work.size();
    // one upload: descriptors (int4 per CTA), then the fix-up lists
    std::vector<int> host((size_t)4 * grid + 3 * (size_t)nfix + (size_t)nslots);
    for (int i = 0; i < grid; ++i) { host[4 * i] = work[i].cg; host[4 * i + 1] = work[i].lo; host[4 * i + 2] = work[i].hi; host[4 * i + 3] = work[i].slot; }
    std::copy(fix_cg.begin(), fix_cg.end(), host.begin() + 4 * (size_t)grid);
    std::copy(fix_off.begin(), fix_off.end(), host.begin() + 4 * (size_t)grid + nfix);
    std::copy(fix_cnt.begin(), fix_cnt.end(), host.begin() + 4 * (size_t)grid + 2 * (size_t)nfix);
    std::copy(slots.begin(), slots.end(), host.begin() + 4 * (size_t)grid + 3 * (size_t)nfix);
    RNLA_CUDA(meta.alloc(host.size() * 4));
    RNLA_CUDA(cudaMemcpyAsync(meta.p, host.data(), host.size() * 4, cudaMemcpyHostToDevice, st));   // pageable source: staged before return
    if (nslots > 0) RNLA_CUDA(partial.alloc((size_t)nslots * d_used * cb * 8));

    SbArgs a;
    a.A = A; a.lda = lda; a.m_local = m_local; a.n = n; a.row_off = row_off;
    a.tab = tab.as<sbtab_t>(); a.q_first = q_first; a.nchunks = nchunks; a.desc = meta.as<int4>();
    a.nbs = (int)nbs64; a.g = g; a.parts = parts; a.stages = 0; a.ncg = ncg; a.scale = 1.0 / std::sqrt((double)zeta);
    a.out = Ask; a.ldo = ldk; a.partial = partial.d(); a.d_used = d_used;
    // bulk copies need 16-byte aligned sources and sizes: even leading dimension, even row offsets and counts
    const bool tma = (reinterpret_cast<uintptr_t>(A) % 16 == 0) && (lda % 2 == 0) && (row_off % 2 == 0) && (m_local % 2 == 0);
    cudaError_t e;
    switch (w) {
        case 1: e = sb_dispatch_bpt<1>(a, bpt, cb, grid, tma, st); break;
        case 2: e = sb_dispatch_bpt<2>(a, bpt, cb, grid, tma, st); break;
        default: e = sb_dispatch_bpt<4>(a, bpt, cb, grid, tma, st); break;
    }
    RNLA_CUDA(e);
    if (nfix > 0) {
        const int* f = meta.as<int>() + 4 * (size_t)grid;
        const int bx = (int)std::min<int64_t>((d_used * cb + 255) / 256, 64);
        sb_fixup_kernel<<<dim3((unsigned)bx, (unsigned)nfix), 256, 0, st>>>(partial.d(), d_used, cb, f, f + nfix, f + 2 * nfix, f + 3 * nfix, Ask, ldk, n);
        ++g_kernel_launches;
        RNLA_CUDA(cudaGetLastError());
    }
    return RNLA_OK;
}

}  // namespace rnla
