// Per-SM ingest rate into shared memory on sm_100a: cp.async.bulk (TMA engine) with different copy sizes, cp.async (LDGSTS, 16 B per
// thread), plain LDG + STS, and bulk + LDGSTS together; one CTA per SM on k SMs.  Decides what bounds the integer sweeps of
// randnla_b200/csrc/i8gemm.cu (measured there: ~55 GB/s per SM whatever the MMA density).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ingest_rate tools/ingest_rate.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("CUDA error %s at %d\n",cudaGetErrorString(e),__LINE__); return 1;}}while(0)
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(b)), "r"(c) : "memory"); }
__device__ __forceinline__ void mbar_expect(uint64_t* b, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(b)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t par) {
    uint32_t ok = 0;
    while (!ok) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(s32(b)), "r"(par) : "memory");
}
__device__ __forceinline__ void bulk(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(s32(dst)), "l"(src), "r"(bytes), "r"(s32(bar)) : "memory");
}
// mode 0: bulk copies of `piece` bytes, STAGES stages of `stage` bytes in flight; mode 1: cp.async 16 B by all threads; mode 2: both
// (half of the bytes each); mode 3: LDG.128 + STS.128
constexpr int STAGES = 4;
__global__ void __launch_bounds__(288, 1) ingest_kernel(const uint8_t* __restrict__ src, size_t per_cta, int stage, int piece, int mode, int iters,
                                                        long long* cycles) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t full[STAGES];
    const uint8_t* base = src + (size_t)blockIdx.x * per_cta;
    const int tid = threadIdx.x, warp = tid >> 5;
    if (tid == 0) { for (int s = 0; s < STAGES; ++s) mbar_init(full + s, 1); asm volatile("fence.mbarrier_init.release.cluster;"); }
    __syncthreads();
    const long long t0 = clock64();
    const int bulk_bytes = mode == 2 ? stage / 2 : stage;
    if (mode == 0 || mode == 2) {
        if (warp == 8) {
            // producer warp: keeps STAGES stages in flight; "consumption" is just the wait of the others (no reuse hazard modelled:
            // stages are re-filled as soon as they have landed, which is the upper bound of what the engine can deliver)
            const int lane = tid & 31;
            for (int it = 0; it < iters; ++it) {
                const int s = it % STAGES;
                if (it >= STAGES) mbar_wait(full + s, ((it / STAGES) - 1) & 1);
                if (lane == 0) mbar_expect(full + s, bulk_bytes);
                __syncwarp();
                const uint8_t* g = base + ((size_t)it * stage) % per_cta;
                for (int o = lane * piece; o < bulk_bytes; o += 32 * piece) bulk(smem + s * stage + o, g + o, piece, full + s);
            }
            for (int it = iters; it < iters + STAGES; ++it) mbar_wait(full + it % STAGES, ((it / STAGES) - 1) & 1);
        }
    }
    if ((mode == 1 || mode == 2) && warp < 8) {
        const int off0 = mode == 2 ? stage / 2 : 0, nbytes = mode == 2 ? stage / 2 : stage;
        for (int it = 0; it < iters; ++it) {
            const int s = it % STAGES;
            const uint8_t* g = base + ((size_t)it * stage) % per_cta + off0;
            for (int o = tid * 16; o < nbytes; o += 256 * 16)
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s32(smem + s * stage + off0 + o)), "l"(g + o) : "memory");
            asm volatile("cp.async.commit_group;" ::: "memory");
            asm volatile("cp.async.wait_group %0;" ::"n"(STAGES - 1) : "memory");
        }
        asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    if (mode == 3 && warp < 8) {
        for (int it = 0; it < iters; ++it) {
            const int s = it % STAGES;
            const uint8_t* g = base + ((size_t)it * stage) % per_cta;
            uint4 v[8];
            int cnt = 0;
            for (int o = tid * 16; o < stage; o += 256 * 16 * 8) {
#pragma unroll
                for (int q = 0; q < 8; ++q) if (o + q * 256 * 16 < stage) v[q] = __ldcg(reinterpret_cast<const uint4*>(g + o + q * 256 * 16));
#pragma unroll
                for (int q = 0; q < 8; ++q) if (o + q * 256 * 16 < stage) *reinterpret_cast<uint4*>(smem + s * stage + o + q * 256 * 16) = v[q];
                ++cnt;
            }
        }
    }
    __syncthreads();
    if (tid == 0) cycles[blockIdx.x] = clock64() - t0;
}
int main() {
    int sms = 0; CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    int khz = 0; CK(cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0));
    const size_t per_cta = (size_t)64 << 20;                   // 64 MB per CTA: streams from HBM
    uint8_t* buf; CK(cudaMalloc(&buf, per_cta * sms)); CK(cudaMemset(buf, 1, per_cta * sms));
    long long* d; CK(cudaMalloc(&d, sms * 8));
    CK(cudaFuncSetAttribute(ingest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int stage = 48 * 1024;
    printf("{\"rows\": [\n");
    bool first = true;
    struct Cfg { int mode, piece; const char* name; } cfgs[] = {{0, 1024, "bulk 1 KB"}, {0, 8192, "bulk 8 KB"}, {0, 49152 / 2, "bulk 24 KB"}, {1, 0, "cp.async 16 B x 256 threads"},
                                                             {3, 0, "LDG.128 + STS.128"}, {2, 8192, "bulk 8 KB (half) + cp.async (half)"}};
    for (auto& c : cfgs)
        for (int k : {1, 37, 148}) {
            const int iters = 600;
            float ms = 0;
            for (int rep = 0; rep < 2; ++rep) {
                cudaEventRecord(e0);
                ingest_kernel<<<k, 288, STAGES * stage, 0>>>(buf, per_cta, stage, c.piece ? c.piece : 16, c.mode, iters, d);
                cudaEventRecord(e1); CK(cudaDeviceSynchronize());
                cudaEventElapsedTime(&ms, e0, e1);
            }
            const double gbs = (double)iters * stage / (ms * 1e-3) * 1e-9;
            printf("%s {\"method\": \"%s\", \"sms\": %d, \"GBps_per_sm\": %.1f, \"TBps_total\": %.2f}", first ? "" : ",\n", c.name, k, gbs, gbs * k * 1e-3);
            first = false;
        }
    printf("\n]}\n");
    return 0;
}
