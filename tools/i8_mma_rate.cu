// Rate of tcgen05.mma kind::i8 (M = 128, K = 32) with both operands resident in shared memory (no loads): cycles per MMA for
// each combination of operand majors, N in {64, 112, 128, 256}, on every SM at once.  Decides what bounds the integer passes of
// randnla_b200/csrc/i8gemm.cu.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/i8_mma_rate tools/i8_mma_rate.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("CUDA error %s at %d\n",cudaGetErrorString(e),__LINE__); return 1;}}while(0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((addr >> 4) & 0x3fffu) | ((uint64_t)((lbo >> 4) & 0x3fffu) << 16) | ((uint64_t)((sbo >> 4) & 0x3fffu) << 32) | (1ull << 46);
}
__host__ __device__ constexpr uint32_t instr_desc(bool a_mn, bool b_mn, int n) {
    return (2u << 4) | (1u << 7) | (1u << 10) | ((a_mn ? 1u : 0u) << 15) | ((b_mn ? 1u : 0u) << 16) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}
__device__ __forceinline__ void mma_i8(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
// pattern 0: every MMA reads the same A and B tile into the same accumulator; pattern 1: 10 digit pairs over 4 A planes x 4 B planes
// into 4 accumulators (the LO sweep); pattern 2: distinct A per MMA, same B
__global__ void __launch_bounds__(128, 1) rate_kernel(int a_mn, int b_mn, int n, int iters, int pattern, long long* cycles) {
    extern __shared__ uint8_t raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bar; __shared__ uint32_t slot;
    for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) ((uint32_t*)smem)[i] = 0x01010101u * (i & 3);
    if (threadIdx.x == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar))); asm volatile("fence.mbarrier_init.release.cluster;"); }
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = slot;
    if (threadIdx.x == 0) {
        const uint32_t idesc = instr_desc(a_mn, b_mn, n);
        const uint32_t a0 = smem_u32(smem), b0 = a0 + 64 * 1024;
        const int npl = n <= 128 ? 4 : 2;                         // accumulators that fit
        long long t0 = clock64();
        for (int it = 0; it < iters; ++it) {
            if (pattern == 1) {
                for (int ta = 0; ta < 4; ++ta) for (int tb = 0; tb + ta < 4; ++tb) {
                    const uint64_t ad = a_mn ? smem_desc(a0 + ta * 8192, 128, 1024) : smem_desc(a0 + ta * 8192, 2048, 128);
                    const uint64_t bd = b_mn ? smem_desc(b0 + tb * 8192, (n / 16) * 128, 128) : smem_desc(b0 + tb * 8192, 128, 256);
                    mma_i8(tmem + (uint32_t)((ta + tb) % npl) * (n <= 128 ? 128 : 256), ad, bd, idesc, 1);
                }
            } else {
                for (int q = 0; q < 10; ++q) {
                    const uint32_t ao = pattern == 2 ? (q & 7) * 8192 : 0;
                    const uint64_t ad = a_mn ? smem_desc(a0 + ao, 128, 1024) : smem_desc(a0 + ao, 2048, 128);
                    const uint64_t bd = b_mn ? smem_desc(b0, (n / 16) * 128, 128) : smem_desc(b0, 128, 256);
                    mma_i8(tmem, ad, bd, idesc, 1);
                }
            }
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        uint32_t ok = 0;
        while (!ok) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(&bar)) : "memory");
        long long t1 = clock64();
        cycles[blockIdx.x] = t1 - t0;
    }
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}

int main() {
    int sms = 0; CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    long long* d; CK(cudaMalloc(&d, sms * 8));
    CK(cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    long long h[256];
    const int iters = 2000;
    printf("{\"sms\": %d, \"rows\": [\n", sms);
    bool first = true;
    for (int pattern = 0; pattern < 3; ++pattern)
        for (int n : {64, 112, 128, 256})
            for (int a_mn = 0; a_mn < 2; ++a_mn)
                for (int b_mn = 0; b_mn < 2; ++b_mn) {
                    for (int rep = 0; rep < 2; ++rep) {
                        rate_kernel<<<sms, 128, 200 * 1024>>>(a_mn, b_mn, n, iters, pattern, d);
                        CK(cudaDeviceSynchronize());
                    }
                    CK(cudaMemcpy(h, d, sms * 8, cudaMemcpyDeviceToHost));
                    double mx = 0, mn = 1e30;
                    for (int i = 0; i < sms; ++i) { double c = (double)h[i] / (iters * 10.0); mx = c > mx ? c : mx; mn = c < mn ? c : mn; }
                    printf("%s {\"pattern\": %d, \"N\": %d, \"a_major\": \"%s\", \"b_major\": \"%s\", \"cycles_per_mma_min\": %.1f, \"max\": %.1f, \"nominal\": %.1f}",
                           first ? "" : ",\n", pattern, n, a_mn ? "MN" : "K", b_mn ? "MN" : "K", mn, mx, n / 2.0);
                    first = false;
                }
    printf("\n]}\n");
    return 0;
}
