// Live roof measurement used by bench.py: the denominators of the roofline are taken on the same box in the same run.
#include "context.cuh"
#include "gemm.cuh"
#include "ptx.cuh"

namespace rnla {
namespace {

__global__ void __launch_bounds__(256) dmma_peak_kernel(double* out, int iters, double x, double y) {
    double c0[16], c1[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) { c0[i] = i; c1[i] = -i; }
    const double a = x + threadIdx.x * 1e-9, b = y;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1},{%2},{%3},{%0,%1};"
                         : "+d"(c0[i]), "+d"(c1[i]) : "d"(a), "d"(b));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += c0[i] + c1[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void __launch_bounds__(512) hbm_read_kernel(const double2* __restrict__ in, size_t n2, double* out) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    double s = 0;
    for (; i + 3 * stride < n2; i += 4 * stride) {
        const double2 a = __ldg(in + i), b = __ldg(in + i + stride), c = __ldg(in + i + 2 * stride), d = __ldg(in + i + 3 * stride);
        s += a.x + a.y + b.x + b.y + c.x + c.y + d.x + d.y;
    }
    for (; i < n2; i += stride) { const double2 a = __ldg(in + i); s += a.x + a.y; }
    if (s == 123.456) out[0] = s;
}

// int8 tensor-core roof: tcgen05.mma kind::i8, M = N = 128, K = 32, both operands resident in shared memory, one issuing thread per
// SM, all SMs at once (tools/i8_mma_rate.cu measured 64 cycles per MMA = the nominal 8192 MAC per clock per SM for every operand major)
__global__ void __launch_bounds__(128, 1) i8_peak_kernel(int iters) {
    extern __shared__ uint8_t raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    for (int i = threadIdx.x; i < 64 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x01fe037fu * (uint32_t)(i | 1);
    if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
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
        auto desc = [](uint32_t addr, uint32_t lbo, uint32_t sbo) {
            return (uint64_t)((addr >> 4) & 0x3fffu) | ((uint64_t)((lbo >> 4) & 0x3fffu) << 16) | ((uint64_t)((sbo >> 4) & 0x3fffu) << 32) | (1ull << 46);
        };
        const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | (1u << 16) | ((128u >> 3) << 17) | ((128u >> 4) << 24);
        const uint32_t a0 = smem_u32(smem), b0 = a0 + 32 * 1024;
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const uint64_t ad = desc(a0 + (q & 3) * 4096, 2048, 128), bd = desc(b0 + (q & 3) * 4096, 1024, 128);
                asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
                             ::"r"(tmem + (uint32_t)(q & 3) * 128), "l"(ad), "l"(bd), "r"(idesc), "r"(1u) : "memory");
            }
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        mbar_wait(&bar, 0);
    }
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}

}  // namespace
}  // namespace rnla

using namespace rnla;

// int8 tensor-core rate of this GPU in TOP/s (2 ops per multiply-accumulate): `burst` is the best of the first short launches,
// `sustained` the rate of a ~0.25 s run of back-to-back launches (power cap and clocks as they are on this box)
extern "C" rnla_status rnla_measure_int8_roof(double* burst_tops, double* sustained_tops) {
    RNLA_TRY(ensure_ctx());
    Ctx& c = ctx();
    cudaEvent_t e0, e1;
    RNLA_CUDA(cudaEventCreate(&e0)); RNLA_CUDA(cudaEventCreate(&e1));
    RNLA_CUDA(cudaFuncSetAttribute(i8_peak_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 66 * 1024));
    const int iters = 20000;                                     // 160 000 MMAs per SM per launch: about 5 ms
    const double ops = 2.0 * 128 * 128 * 32 * 8.0 * iters * c.sms;
    float best = 1e30f;
    for (int r = 0; r < 4; ++r) {
        RNLA_CUDA(cudaEventRecord(e0, c.stream));
        i8_peak_kernel<<<c.sms, 128, 66 * 1024, c.stream>>>(iters);
        RNLA_CUDA(cudaEventRecord(e1, c.stream));
        RNLA_CUDA(cudaEventSynchronize(e1));
        float ms; RNLA_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        if (r > 0 && ms < best) best = ms;
    }
    if (burst_tops) *burst_tops = ops / best * 1e-9;
    if (sustained_tops) {
        const int reps = 50;
        RNLA_CUDA(cudaEventRecord(e0, c.stream));
        for (int r = 0; r < reps; ++r) i8_peak_kernel<<<c.sms, 128, 66 * 1024, c.stream>>>(iters);
        RNLA_CUDA(cudaEventRecord(e1, c.stream));
        RNLA_CUDA(cudaEventSynchronize(e1));
        float ms; RNLA_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        *sustained_tops = ops * reps / ms * 1e-9;
    }
    g_kernel_launches += 54;
    RNLA_CUDA(cudaGetLastError());
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    return RNLA_OK;
}

extern "C" rnla_status rnla_measure_roofs(double* fp64_dmma_tflops, double* hbm_read_gbs, size_t hbm_bytes) {
    RNLA_TRY(ensure_ctx());
    Ctx& c = ctx();
    cudaEvent_t e0, e1;
    RNLA_CUDA(cudaEventCreate(&e0)); RNLA_CUDA(cudaEventCreate(&e1));
    DevBuf out;
    RNLA_CUDA(out.alloc((size_t)c.sms * 4 * 512 * 8));
    if (fp64_dmma_tflops) {
        const int iters = 20000;
        float best = 1e30f;
        for (int r = 0; r < 6; ++r) {
            RNLA_CUDA(cudaEventRecord(e0, c.stream));
            dmma_peak_kernel<<<c.sms, 256, 0, c.stream>>>(out.d(), iters, 1.0, 0.999);
            RNLA_CUDA(cudaEventRecord(e1, c.stream));
            RNLA_CUDA(cudaEventSynchronize(e1));
            float ms; RNLA_CUDA(cudaEventElapsedTime(&ms, e0, e1));
            if (r > 0 && ms < best) best = ms;
        }
        *fp64_dmma_tflops = 512.0 * 16 * iters * 8.0 * c.sms / best * 1e-9;
    }
    if (hbm_read_gbs) {
        if (hbm_bytes < ((size_t)1 << 28)) hbm_bytes = (size_t)1 << 28;
        DevBuf buf;
        RNLA_CUDA(buf.alloc(hbm_bytes));
        RNLA_CUDA(cudaMemsetAsync(buf.p, 0, hbm_bytes, c.stream));
        float best = 1e30f;
        for (int r = 0; r < 6; ++r) {
            RNLA_CUDA(cudaEventRecord(e0, c.stream));
            hbm_read_kernel<<<c.sms * 4, 512, 0, c.stream>>>(buf.as<double2>(), hbm_bytes / 16, out.d());
            RNLA_CUDA(cudaEventRecord(e1, c.stream));
            RNLA_CUDA(cudaEventSynchronize(e1));
            float ms; RNLA_CUDA(cudaEventElapsedTime(&ms, e0, e1));
            if (r > 0 && ms < best) best = ms;
        }
        *hbm_read_gbs = (double)hbm_bytes / best * 1e-6;
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    return RNLA_OK;
}
