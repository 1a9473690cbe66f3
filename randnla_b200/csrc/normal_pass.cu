// One pass over a tall column-major A for BOTH products of a least-squares iteration (reference: the two products of every
// cgls iteration, src/cg.rs:36 `a * &p` and :40 `a.transpose() * &r`; of every lsqr iteration, src/solvers.rs:188 and :196):
//
//   u  = cq * (A x) + cy * y          m_local values (y optional; u optionally stored; u = y goes through a temporary)
//   t  = A^T u                        n values
//   uu = u . u
//
// The two mat-vec kernels of solve.cu stream A twice per iteration (2 x 8 m n bytes, each at the HBM roof).  Here a CLUSTER of C
// CTAs holds a slab of 32 rows x n columns in its distributed shared memory (CTA c: columns [c ncb, (c + 1) ncb), ncb <= 256, one
// tiled TMA load of 32 x ncb doubles per slab and CTA, three stages), so A is read from HBM ONCE per iteration.  Per CTA:
//   warps 0-7    "row" group: lane = row of the slab, warp w owns the columns k 8 + w (x of those columns in 32 registers); each
//                thread sums its 32 products, the 8 warps' parts are added in shared memory and warp 0 sends the CTA's partial of
//                (A x)_row to EVERY CTA of the cluster with st.async (remote shared-memory store that completes on the receiver's
//                mbarrier: no cluster barrier, no fences) -- and goes on to the next slab at once.
//   warps 8-15   "column" group: waits for the C partials of a slab (its own mbarrier), sums them in rank order (u_row is complete),
//                and accumulates a_rc u_row into 32 REGISTER accumulators per thread (row class = lane, column k 8 + w) that live
//                across all slabs of the cluster; the reduction over the 32 row classes happens once, at the end of the kernel.
//                Lane 0 of its first warp refills a stage (one TMA load) as soon as the eight warps have released it.
// So each element is read from shared memory twice (conflict-free, lane = row), costs two DFMAs, and nothing else per element; the
// exchange latency only delays the column group behind the row group.  Every sum has a fixed order (columns of a thread, warps,
// cluster ranks, slabs of a cluster, row classes, clusters), so results are reproducible.
// Limit: n <= 8 x 256 (portable cluster size); beyond it callers keep the two streaming kernels.  Any leading dimension and any 8-byte
// aligned base are addressed through tensor maps (see normal_pass_supported).
#include "drivers.cuh"
#include "gemm.cuh"
#include "ptx.cuh"
#include <cuda.h>
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <string>

namespace rnla {

namespace {

constexpr int NP_R = 32;                 // rows of a slab: one per lane
constexpr int NP_GW = 8;                 // warps per group
constexpr int NP_THREADS = 2 * NP_GW * 32;
constexpr int NP_NCB = 256;              // columns per CTA (slots k 8 + w, k < 32, w < 8)
constexpr int NP_STAGES = 3;
constexpr int NP_CMAX = 8;               // portable cluster size
constexpr int NP_XS = 8;                 // exchange slots (a sender is never more than 6 slabs ahead of a receiver, see below)
constexpr int NP_RP = NP_R + 2;          // rows of a box when a column may start 8 bytes off a 16-byte boundary (see normal_pass_supported)
constexpr uint32_t NP_STAGE_BYTES = NP_RP * NP_NCB * 8 + 128;   // 68 KB (+ the odd columns' box starts on a 128-byte boundary)
constexpr size_t NP_SMEM = (size_t)NP_STAGES * NP_STAGE_BYTES + (size_t)NP_XS * NP_CMAX * 32 * 8 + 2 * NP_GW * 32 * 8 + 256 + 128;

struct NpArgs {
    int64_t m;                           // local rows
    int n, ncb;                          // columns, columns per CTA
    int64_t nslabs;
    const double* x; const double* y; double cq, cy; double* uout;
    int ne;                              // stage positions [0, ne): local columns 0, 2, 4, ... (map E); [ne, ncb): 1, 3, 5, ... (map O); ne = ncb: one map, identity
    int shift_e, shift_o;                // 1: the columns of that map start 8 bytes off; their boxes start one element early
    int pitch;                           // rows per column in a stage: 32, or 34 when some column starts 8 bytes off
    double* tpart;                       // [clusters][n]
    double* uupart;                      // [clusters]
};

__device__ __forceinline__ uint32_t np_cluster_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t np_cluster_size() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void np_cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// 8-byte store into the shared memory of CTA `cta` of the cluster, completing 8 bytes on the mbarrier `bar` of that CTA
__device__ __forceinline__ void np_st_async(uint32_t local_addr, uint32_t local_bar, uint32_t cta, double v) {
    uint32_t ra, rb;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(local_addr), "r"(cta));
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(rb) : "r"(local_bar), "r"(cta));
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b64 [%0], %1, [%2];"
                 ::"r"(ra), "l"(__double_as_longlong(v)), "r"(rb) : "memory");
}
__device__ __forceinline__ void np_tma_load_2d(void* dst, const CUtensorMap* tm, uint64_t* bar, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(smem_u32(dst)), "l"(tm), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}

// Flow control of the exchange slots: CTA X sends its partial of slab i once its stage of slab i is loaded, i.e. after X's column
// group released slab i - 3, which needed the partial of slab i - 3 from every CTA Y; Y sent that after ITS stage of slab i - 3 was
// loaded, i.e. after Y's column group released slab i - 6 (and had read the slot of slab i - 6).  Eight slots therefore never collide,
// and a completion can never land on an mbarrier phase that is still open for an older slab.
// PAD: some column starts 8 bytes off a 16-byte boundary (34-row boxes, two kinds of columns); otherwise the plain 32-row layout with
// compile-time strides
template <bool PAD>
__global__ void __launch_bounds__(NP_THREADS, 1)
normal_pass_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmO, const NpArgs a) {
    extern __shared__ uint8_t np_smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(np_smem_raw) + 127) & ~(uintptr_t)127);
    uint8_t* stages = smem;
    double* xbuf = reinterpret_cast<double*>(smem + (size_t)NP_STAGES * NP_STAGE_BYTES);     // [slot][rank][32]
    double* red = xbuf + NP_XS * NP_CMAX * 32;                                                  // [parity][warp][32]
    uint64_t* full = reinterpret_cast<uint64_t*>(red + 2 * NP_GW * 32);
    uint64_t* empty = full + NP_STAGES;
    uint64_t* xfull = empty + NP_STAGES;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t rank = np_cluster_rank(), csize = np_cluster_size();
    const int64_t G = gridDim.x / csize, g = blockIdx.x / csize;
    const int cnt = (int)((a.nslabs - g + G - 1) / G);         // slabs of this cluster: g, g + G, ...
    const int col0 = (int)rank * a.ncb;
    const bool split = a.ne < a.ncb;                           // odd leading dimension: even and odd columns through two tensor maps
    // stage position -> local column
    auto lcol = [&](int pos) { return !split ? pos : (pos < a.ne ? 2 * pos : 2 * (pos - a.ne) + 1); };
    constexpr int pitch = PAD ? NP_RP : NP_R;
    const int obase = (a.ne * pitch + 15) & ~15;               // doubles: where the second box of a stage starts (128-byte aligned)
    const int wq = (tid >> 5) & (NP_GW - 1);                   // this warp's column class: positions k 8 + wq
    const int kE = max(0, min(32, (a.ne - wq + 7) / 8));       // positions k 8 + wq with k < kE lie in the first box
    // element (row `lane`, position k 8 + wq) of a stage sits at (k < kE ? offE : offO) + k 8 pitch: a column that starts 8 bytes off a
    // 16-byte boundary was loaded from one element earlier (row offset 1)
    const int offE = wq * pitch + lane + a.shift_e;
    const int offO = obase + (wq - a.ne) * pitch + lane + a.shift_o;
    const uint32_t box_bytes = (uint32_t)a.ncb * (uint32_t)pitch * 8u;
    auto load_slab = [&](int i) {                              // one thread: slab i of this cluster into stage i % NP_STAGES
        const int st = i % NP_STAGES;
        const int r0 = (int)((g + (int64_t)i * G) * NP_R);      // (a view shifted by one element starts its box at the same coordinate:
        uint8_t* dst = stages + (size_t)st * NP_STAGE_BYTES;   //  one element early, on a 16-byte boundary)
        mbar_arrive_expect_tx(full + st, box_bytes);
        if (!split) np_tma_load_2d(dst, &tmA, full + st, r0, col0);
        else {
            np_tma_load_2d(dst, &tmA, full + st, r0, col0 / 2);
            np_tma_load_2d(dst + (size_t)obase * 8, &tmO, full + st, r0, col0 / 2);
        }
    };
    // the columns of the stages that no load ever writes (ncb < 256) must read as zeros
    for (int st = 0; st < NP_STAGES; ++st) {
        double* tail = reinterpret_cast<double*>(stages + (size_t)st * NP_STAGE_BYTES) + (((a.ne * (PAD ? NP_RP : NP_R)) + 15) & ~15) + (size_t)(a.ncb - a.ne) * (PAD ? NP_RP : NP_R);
        for (int q = tid; q < (NP_NCB - a.ncb) * (PAD ? NP_RP : NP_R); q += NP_THREADS) tail[q] = 0.0;
    }
    if (tid == 0) {
        for (int st = 0; st < NP_STAGES; ++st) { mbar_init(full + st, 1); mbar_init(empty + st, NP_GW); }
        for (int sl = 0; sl < NP_XS; ++sl) mbar_init(xfull + sl, 1);
        mbar_fence_init();
        for (int i = 0; i < NP_STAGES && i < cnt; ++i) load_slab(i);
    }
    __syncthreads();
    np_cluster_sync();                                         // every CTA's barriers exist before anyone sends to it

    if (warp < NP_GW) {
        // ---------------------------------------------------------------- row group: partials of (A x)
        const int w = warp;
        double p[32];
#pragma unroll
        for (int k = 0; k < 32; ++k) { const int pos = k * 8 + w, lc = lcol(pos); p[k] = (pos < a.ncb && col0 + lc < a.n) ? a.x[col0 + lc] : 0.0; }
        for (int i = 0; i < cnt; ++i) {
            const int st = i % NP_STAGES;
            mbar_wait(full + st, (uint32_t)(i / NP_STAGES) & 1u);
            const double* sg = reinterpret_cast<const double*>(stages + (size_t)st * NP_STAGE_BYTES);
            const double* spE = sg + offE; const double* spO = sg + offO;
            double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
#pragma unroll
            for (int k = 0; k < 32; k += 4) {
                s0 = fma(((!PAD || k < kE) ? spE : spO)[(k * 8) * pitch], p[k], s0);
                s1 = fma(((!PAD || k + 1 < kE) ? spE : spO)[((k + 1) * 8) * pitch], p[k + 1], s1);
                s2 = fma(((!PAD || k + 2 < kE) ? spE : spO)[((k + 2) * 8) * pitch], p[k + 2], s2);
                s3 = fma(((!PAD || k + 3 < kE) ? spE : spO)[((k + 3) * 8) * pitch], p[k + 3], s3);
            }
            double* rd = red + (i & 1) * NP_GW * 32;
            rd[w * 32 + lane] = (s0 + s1) + (s2 + s3);
            asm volatile("bar.sync 1, 256;" ::: "memory");
            if (w == 0) {
                double part = 0.0;
#pragma unroll
                for (int ww = 0; ww < NP_GW; ++ww) part += rd[ww * 32 + lane];
                const int slot = i % NP_XS;
                if (csize == 1) {
                    // a cluster of one CTA has no distributed shared memory: plain store, the warp's arrival completes the slot's phase
                    xbuf[slot * NP_CMAX * 32 + lane] = part;
                    __syncwarp();
                    if (lane == 0) mbar_arrive(xfull + slot);
                } else {
                    const uint32_t dst = smem_u32(xbuf + (slot * NP_CMAX + (int)rank) * 32 + lane), bar = smem_u32(xfull + slot);
                    for (uint32_t c = 0; c < csize; ++c) np_st_async(dst, bar, c, part);
                }
            }
        }
    } else {
        // ---------------------------------------------------------------- column group: u complete, t += a u
        const int w = warp - NP_GW;
        double acc[32];
#pragma unroll
        for (int k = 0; k < 32; ++k) acc[k] = 0.0;
        double uu = 0.0;
        const uint32_t xbytes = csize * 32u * 8u;
        const bool arm = csize > 1 && w == 0 && lane == 0;       // (one CTA: the row group's arrival completes the phase instead)
        if (cnt > 0 && arm) mbar_arrive_expect_tx(xfull + 0, xbytes);
        int64_t row = g * NP_R + lane;
        double yv = (cnt > 0 && a.y != nullptr && row < a.m) ? a.y[row] : 0.0;
        for (int i = 0; i < cnt; ++i) {
            const int st = i % NP_STAGES, slot = i % NP_XS;
            // next slab: arm its exchange barrier (its previous phase, slab i + 1 - 8, completed long ago) and fetch its y
            if (i + 1 < cnt && arm) mbar_arrive_expect_tx(xfull + (i + 1) % NP_XS, xbytes);
            const int64_t row_next = (g + (int64_t)(i + 1) * G) * NP_R + lane;
            const double y_next = (i + 1 < cnt && a.y != nullptr && row_next < a.m) ? a.y[row_next] : 0.0;
            mbar_wait(full + st, (uint32_t)(i / NP_STAGES) & 1u);
            mbar_wait(xfull + slot, (uint32_t)(i / NP_XS) & 1u);
            const double* xb = xbuf + slot * NP_CMAX * 32 + lane;
            double q = 0.0;
            for (uint32_t c = 0; c < csize; ++c) q += xb[c * 32];
            const double u = a.cq * q + a.cy * yv;
            if (w == 0 && rank == 0) {
                uu = fma(u, u, uu);
                if (a.uout != nullptr && row < a.m) a.uout[row] = u;
            }
            const double* sg = reinterpret_cast<const double*>(stages + (size_t)st * NP_STAGE_BYTES);
            const double* spE = sg + offE; const double* spO = sg + offO;
#pragma unroll
            for (int k = 0; k < 32; ++k) acc[k] = fma(((!PAD || k < kE) ? spE : spO)[(k * 8) * pitch], u, acc[k]);
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(empty + st);
                if (w == 0 && i + NP_STAGES < cnt) {           // producer: refill the stage once all eight warps have released it
                    mbar_wait(empty + st, (uint32_t)(i / NP_STAGES) & 1u);
                    load_slab(i + NP_STAGES);
                }
            }
            row = row_next; yv = y_next;
        }
        // sum over the 32 row classes: transposing butterfly, slot L (column L 8 + w) ends up in lane L
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1) {
            const bool up = (lane & off) != 0;
#pragma unroll
            for (int k = 0; k < off; ++k) {
                const double send = up ? acc[k] : acc[k + off];
                const double keep = up ? acc[k + off] : acc[k];
                acc[k] = keep + __shfl_xor_sync(0xffffffffu, send, off);
            }
        }
        const int pos = lane * 8 + w, lc = lcol(pos);
        if (pos < a.ncb && col0 + lc < a.n) a.tpart[g * a.n + col0 + lc] = acc[0];
        if (w == 0 && rank == 0) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) uu += __shfl_xor_sync(0xffffffffu, uu, o);
            if (lane == 0) a.uupart[g] = uu;
        }
    }
    __syncthreads();
    np_cluster_sync();                                         // nobody leaves while a store to it may be in flight
}

// t[j] = sum over clusters (fixed order), t[n] = u . u
__global__ void __launch_bounds__(256)
normal_pass_reduce_kernel(const double* __restrict__ tpart, const double* __restrict__ uupart, int G, int n, double* __restrict__ t) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j > n) return;
    double sum = 0.0;
    if (j < n) { for (int g = 0; g < G; ++g) sum += tpart[(int64_t)g * n + j]; }
    else { for (int g = 0; g < G; ++g) sum += uupart[g]; }
    t[j] = sum;
}

typedef CUresult (*NpEncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                               const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
inline NpEncodeFn np_encoder() {
    static NpEncodeFn fn = [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) p = nullptr;
        return reinterpret_cast<NpEncodeFn>(p);
    }();
    return fn;
}
inline int np_cluster_size_for(int64_t n) { int c = 1; while ((int64_t)c * NP_NCB < n) c *= 2; return c; }

}  // namespace

// whether the one-pass kernel takes this operand: n <= 2048.  A tensor map wants a 16-byte aligned base and strides that are multiples
// of 16 bytes, and every row of a box must start on a 16-byte boundary: a column that starts 8 bytes off is loaded from one element
// earlier (34-row boxes, read at row offset 1), and with an odd leading dimension -- columns alternate between the two cases -- the even
// and the odd columns get a map each.
bool normal_pass_supported(const double* A, int64_t lda, int64_t m_local, int64_t n) {
    if (const char* e = getenv("RNLA_ONEPASS")) { if (e[0] == '0') return false; }
    return n >= 1 && n <= (int64_t)NP_CMAX * NP_NCB && m_local >= 1 && m_local < ((int64_t)1 << 31) - 64 && lda >= m_local &&
           (reinterpret_cast<uintptr_t>(A) & 7) == 0 && np_encoder() != nullptr;
}

// Host logic of the launch, separated so that it can be checked without a GPU (rnla_plan_normal_pass, tests/test_host_logic.py):
// cluster size, columns per CTA, and how the columns of a CTA are served by the tensor maps.
NormalPassPlan normal_pass_plan(uint64_t base_address, int64_t lda, int64_t n) {
    NormalPassPlan p;
    p.cluster = np_cluster_size_for(n);
    const bool odd_ld = (lda & 1) != 0;
    p.ncb = (int)((n + p.cluster - 1) / p.cluster);
    if (odd_ld && p.cluster > 1) p.ncb += p.ncb & 1;           // every CTA starts on an even column
    // Column j starts at A + 8 j lda: on a 16-byte boundary iff (A / 8 + j lda) is even.  TMA wants every row of a box to start on one,
    // so the columns that do not are loaded from ONE ELEMENT EARLIER (a view whose base is A - 8, same coordinates) with a box of 34
    // rows, and read at row offset 1; with an odd leading dimension the two kinds alternate and get a tensor map each.
    const int par0 = (int)((base_address >> 3) & 1);
    p.ne = odd_ld ? (p.ncb + 1) / 2 : p.ncb;
    p.shift_e = par0; p.shift_o = odd_ld ? 1 - par0 : par0;
    p.pitch = (odd_ld || par0) ? NP_RP : NP_R;
    p.stage_bytes = (int)NP_STAGE_BYTES;
    return p;
}

// t (n + 1 doubles, device): t[0..n) = A^T u, t[n] = u . u with u = cq (A x) + cy y; all-reduced over the row shards.
// uout (optional, m_local) receives u and may be y itself.
rnla_status dev_normal_pass(const double* A, int64_t lda, int64_t m_local, int64_t n, const double* x, double cq, const double* y, double cy,
                            double* uout, double* t) {
    Ctx& c = ctx();
    if (!normal_pass_supported(A, lda, m_local, n)) return fail(RNLA_ERR_COMPUTATION, "normal_pass: operand not supported by the one-pass kernel");
    const NormalPassPlan plan = normal_pass_plan((uint64_t)reinterpret_cast<uintptr_t>(A), lda, n);
    const int C = plan.cluster, ncb = plan.ncb;
    const bool odd_ld = (lda & 1) != 0;
    static bool attr = false;
    if (!attr) {
        RNLA_CUDA(cudaFuncSetAttribute(normal_pass_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)NP_SMEM));
        RNLA_CUDA(cudaFuncSetAttribute(normal_pass_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)NP_SMEM));
        attr = true;
    }
    cudaLaunchConfig_t cfg = {};
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = (unsigned)C; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.blockDim = dim3(NP_THREADS); cfg.dynamicSmemBytes = NP_SMEM; cfg.stream = c.stream; cfg.attrs = at; cfg.numAttrs = 1;
    static int clusters_for[NP_CMAX + 1] = {0};
    if (clusters_for[C] == 0) {
        cfg.gridDim = dim3((unsigned)(C * c.sms));
        int nc = 0;
        RNLA_CUDA(cudaOccupancyMaxActiveClusters(&nc, normal_pass_kernel<true>, &cfg));
        clusters_for[C] = std::max(1, nc);
    }
    NpArgs a;
    a.m = m_local; a.n = (int)n; a.ncb = ncb; a.nslabs = (m_local + NP_R - 1) / NP_R;
    const int G = (int)std::min<int64_t>(clusters_for[C], a.nslabs);
    DevBuf part;
    RNLA_CUDA(part.alloc(((size_t)G * n + G) * 8));
    if (getenv("RNLA_NP_VERBOSE")) fprintf(stderr, "normal_pass: C = %d, ncb = %d, clusters = %d (max %d), slabs = %lld\n", C, ncb, G, clusters_for[C], (long long)a.nslabs);
    // u may be asked to overwrite y: every column warp of every CTA of a cluster reads y_row (up to three slabs apart from one another)
    // while ONE warp writes u_row, so an in-place update would race; it goes through a temporary and is copied back
    DevBuf utmp;
    double* uw = uout;
    if (uout != nullptr && uout == y) { RNLA_CUDA(utmp.alloc((size_t)m_local * 8)); uw = utmp.d(); }
    a.x = x; a.y = y; a.cq = cq; a.cy = cy; a.uout = uw; a.tpart = part.d(); a.uupart = part.d() + (size_t)G * n;
    a.ne = plan.ne; a.shift_e = plan.shift_e; a.shift_o = plan.shift_o; a.pitch = plan.pitch;
    CUtensorMap tm, tmo;
    auto encode = [&](CUtensorMap* out, const double* first_col, int shift, int64_t ncols, int64_t col_stride, int box_cols) -> rnla_status {
        const cuuint64_t dims[2] = {(cuuint64_t)(m_local + shift), (cuuint64_t)std::max<int64_t>(ncols, 1)};
        const cuuint64_t strides[1] = {(cuuint64_t)col_stride * 8};
        const cuuint32_t box[2] = {(cuuint32_t)a.pitch, (cuuint32_t)std::max(box_cols, 1)}, ones[2] = {1, 1};
        const CUresult r = np_encoder()(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<double*>(first_col - shift), dims, strides, box, ones,
                                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return fail(RNLA_ERR_COMPUTATION, "normal_pass: cuTensorMapEncodeTiled failed (" + std::to_string((int)r) + ")");
        return RNLA_OK;
    };
    if (!odd_ld) {
        RNLA_TRY(encode(&tm, A, a.shift_e, n, lda, ncb));
        tmo = tm;
    } else {
        RNLA_TRY(encode(&tm, A, a.shift_e, (n + 1) / 2, 2 * lda, a.ne));                // columns 0, 2, 4, ...
        RNLA_TRY(encode(&tmo, A + lda, a.shift_o, n / 2, 2 * lda, ncb - a.ne));        // columns 1, 3, 5, ...
    }
    cfg.gridDim = dim3((unsigned)(G * C));
    kernel_phase_begin("k:normal_pass");
    if (a.pitch == NP_RP) RNLA_CUDA(cudaLaunchKernelEx(&cfg, normal_pass_kernel<true>, tm, tmo, a));
    else RNLA_CUDA(cudaLaunchKernelEx(&cfg, normal_pass_kernel<false>, tm, tmo, a));
    kernel_phase_end();
    normal_pass_reduce_kernel<<<(unsigned)((n + 1 + 255) / 256), 256, 0, c.stream>>>(a.tpart, a.uupart, G, (int)n, t);
    g_kernel_launches += 2;
    RNLA_CUDA(cudaGetLastError());
    if (uw != uout) RNLA_CUDA(cudaMemcpyAsync(uout, uw, (size_t)m_local * 8, cudaMemcpyDeviceToDevice, c.stream));
    if (c.nranks > 1) RNLA_TRY(allreduce_sum_f64(t, (size_t)n + 1));
    return RNLA_OK;
}

}  // namespace rnla
