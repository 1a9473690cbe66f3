// Round-1 roof measurements for the B200 (SURVEY.md §7 step 0 / BASELINE.md §4).
// Standalone: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench tools/microbench.cu
// Prints one JSON object: FP64 DFMA peak, FP64 DMMA (m8n8k4) peak, DMMA with concurrent
// INT32/FP32 work (pipe independence, decides whether in-kernel Philox is free), HBM read BW.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("CUDA error %s at %d\n",cudaGetErrorString(e),__LINE__); return 1;}}while(0)

__global__ void __launch_bounds__(256) k_dfma(double* out, int iters, double x, double y) {
    double a[16];
#pragma unroll
    for (int i = 0; i < 16; i++) a[i] = x + i + threadIdx.x;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 16; i++) a[i] = fma(a[i], y, x);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1},{%2},{%3},{%0,%1};"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int NACC>
__global__ void __launch_bounds__(256) k_dmma(double* out, int iters, double x, double y) {
    double c0[NACC], c1[NACC];
#pragma unroll
    for (int i = 0; i < NACC; i++) { c0[i] = i; c1[i] = -i; }
    double a = x + threadIdx.x * 1e-9, b = y;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < NACC; i++) dmma(c0[i], c1[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; i++) s += c0[i] + c1[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// DMMA in half of the warps, integer multiply-xor (Philox-like) work in the other half.
__global__ void __launch_bounds__(384) k_dmma_mixed(double* out, int iters, double x, double y, int int_per_iter) {
    int warp = threadIdx.x >> 5;
    if (warp < 8) {
        double c0[16], c1[16];
#pragma unroll
        for (int i = 0; i < 16; i++) { c0[i] = i; c1[i] = -i; }
        double a = x + threadIdx.x * 1e-9, b = y;
        for (int it = 0; it < iters; it++) {
#pragma unroll
            for (int i = 0; i < 16; i++) dmma(c0[i], c1[i], a, b);
        }
        double s = 0;
#pragma unroll
        for (int i = 0; i < 16; i++) s += c0[i] + c1[i];
        out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    } else {
        uint32_t c[4] = {threadIdx.x, blockIdx.x, 1u, 2u};
        float f = x;
        for (int it = 0; it < iters; it++) {
            for (int r = 0; r < int_per_iter; r++) {
                uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
                uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
                c[0] = hi1 ^ c[1] ^ r; c[1] = lo1; c[2] = hi0 ^ c[3] ^ it; c[3] = lo0;
                f = fmaf(f, 1.0001f, (float)(c[0] & 0xff));
            }
        }
        out[blockIdx.x * blockDim.x + threadIdx.x] = (double)(c[0] ^ c[1] ^ c[2] ^ c[3]) + f;
    }
}

__global__ void __launch_bounds__(512) k_read(const double2* __restrict__ in, size_t n2, double* out) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    double s = 0;
    for (; i + 3 * stride < n2; i += 4 * stride) {
        double2 a = __ldg(in + i), b = __ldg(in + i + stride), c = __ldg(in + i + 2 * stride), d = __ldg(in + i + 3 * stride);
        s += a.x + a.y + b.x + b.y + c.x + c.y + d.x + d.y;
    }
    for (; i < n2; i += stride) { double2 a = __ldg(in + i); s += a.x + a.y; }
    if (s == 123.456) out[0] = s;
}

template <class F>
static float time_ms(F f, int reps) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    f(); f(); cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < reps; r++) {
        cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    return best;
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    int sms = p.multiProcessorCount;
    double* out; CK(cudaMalloc(&out, sizeof(double) * sms * 8 * 512));
    int iters = 20000;
    printf("{\"gpu\": \"%s\", \"sms\": %d", p.name, sms);
    {   // DFMA: 16 FMA/thread/iter
        for (int bps = 1; bps <= 4; bps *= 2) {
            float ms = time_ms([&] { k_dfma<<<sms * bps, 256>>>(out, iters, 1.0, 0.999); }, 5);
            double fl = 2.0 * 16 * iters * 256.0 * sms * bps;
            printf(", \"dfma_tflops_bps%d\": %.3f", bps, fl / ms * 1e-9);
        }
    }
    {   // DMMA m8n8k4: 512 flop per warp instruction
        for (int bps = 1; bps <= 4; bps *= 2) {
            float ms = time_ms([&] { k_dmma<16><<<sms * bps, 256>>>(out, iters, 1.0, 0.999); }, 5);
            double fl = 512.0 * 16 * iters * 8.0 * sms * bps;
            printf(", \"dmma_tflops_acc16_bps%d\": %.3f", bps, fl / ms * 1e-9);
        }
        float ms = time_ms([&] { k_dmma<4><<<sms, 256>>>(out, iters, 1.0, 0.999); }, 5);
        printf(", \"dmma_tflops_acc4_bps1\": %.3f", 512.0 * 4 * iters * 8.0 * sms / ms * 1e-9);
        ms = time_ms([&] { k_dmma<28><<<sms, 256>>>(out, iters, 1.0, 0.999); }, 5);
        printf(", \"dmma_tflops_acc28_bps1\": %.3f", 512.0 * 28 * iters * 8.0 * sms / ms * 1e-9);
    }
    {   // mixed: does INT/FP32 work in 4 extra warps slow the 8 DMMA warps down?
        for (int ipi = 0; ipi <= 16; ipi = ipi ? ipi * 2 : 4) {
            float ms = time_ms([&] { k_dmma_mixed<<<sms, 384>>>(out, iters, 1.0, 0.999, ipi); }, 5);
            double fl = 512.0 * 16 * iters * 8.0 * sms;
            printf(", \"dmma_mixed_tflops_int%d\": %.3f", ipi, fl / ms * 1e-9);
        }
    }
    {   // HBM read-only stream over 8 GiB
        size_t bytes = (size_t)8 << 30; double2* buf; CK(cudaMalloc(&buf, bytes)); CK(cudaMemset(buf, 0, bytes));
        float ms = time_ms([&] { k_read<<<sms * 4, 512>>>(buf, bytes / 16, out); }, 5);
        printf(", \"hbm_read_gbs\": %.1f", bytes / ms * 1e-6);
        cudaFree(buf);
    }
    printf("}\n");
    return 0;
}
