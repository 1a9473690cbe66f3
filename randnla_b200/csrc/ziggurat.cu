// The reference's Gaussian sketch entries: rand_distr 0.4.3 `Normal::new(0, 1).sample` = StandardNormal = the 256-layer ziggurat
// (utils.rs `ziggurat`, normal.rs `zero_case`) drawing from ONE sequential ThreeFry2x64Rng stream (reference: src/sketch.rs:112-117,
// DMatrix::from_fn fills column-major).  A sample consumes 1 word of the stream (97.8 %), 2 (a rejected or accepted wedge test) or
// more (rejections, the tail), so entry t starts where entry t - 1 stopped.  ThreeFry is counter-based, which makes the sequential
// definition parallel in four steps:
//   1. every word position p of the stream is treated as if a sample STARTED there: value v[p] and words consumed len[p];
//   2. per block of 2048 positions and per entry offset e < 32: where the chain entered at e leaves the block, and how many
//      samples it holds (a walk through shared memory);
//   3. one thread follows the chain over the blocks (a few thousand steps): entry offset and first sample index of every block;
//   4. every block walks its own piece of the chain and writes its samples.
// The tables are regenerated on the host by the recipe that produced rand_distr's ziggurat_tables.rs (rand utils/ziggurat_tables.py:
// Doornik's zigNorInit with NORM_R, NORM_V, every entry through its "%.18f" text); the CPU checker of the tests regenerates them independently.
#include "context.cuh"
#include "drivers.cuh"
#include "gemm.cuh"
#include "panel.cuh"
#include "rng.cuh"
#include <cmath>
#include <cstdio>
#include <cstdlib>

namespace rnla {

namespace {

constexpr int ZB = 2048;      // stream positions per block
constexpr int ZW = 32;        // entry offsets resolved per block (a sample would have to span 32 words to miss: never)
constexpr double ZIG_R = 3.654152885361008796;

__device__ double d_zig_x[257], d_zig_f[257];

__device__ __forceinline__ uint64_t stream_word(uint64_t k0, uint64_t k1, uint64_t t) {
    // BlockRng64<ThreeFry2x64>: block b = threefry(ctr = (b, 0)), handed out x[0], x[1]
    uint64_t x0, x1;
    threefry2x64_20(t >> 1, 0ull, k0, k1, x0, x1);
    return (t & 1) ? x1 : x0;
}
__device__ __forceinline__ double f64_with_exponent(uint64_t frac52, int e) {     // rand 0.8.5 IntoFloat::into_float_with_exponent
    return __longlong_as_double((long long)(frac52 | ((uint64_t)(1023 + e) << 52)));
}

// step 1: the sample that would start at word p
__global__ void __launch_bounds__(256)
zig_eval_kernel(uint64_t k0, uint64_t k1, int64_t P, double* __restrict__ v, uint8_t* __restrict__ len) {
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= P) return;
    uint64_t t = (uint64_t)p;
    double x;
    for (;;) {
        const uint64_t bits = stream_word(k0, k1, t++);
        const int i = (int)(bits & 0xff);
        const double u = __dadd_rn(f64_with_exponent(bits >> 12, 1), -3.0);              // [2, 4) - 3
        x = __dmul_rn(u, d_zig_x[i]);
        if (fabs(x) < d_zig_x[i + 1]) break;
        if (i == 0) {                                                                    // the tail beyond R (normal.rs zero_case)
            double xt = 1.0, yt = 0.0;
            while (__dmul_rn(-2.0, yt) < __dmul_rn(xt, xt)) {
                const double x_ = __dadd_rn(f64_with_exponent(stream_word(k0, k1, t++) >> 12, 0), -(1.0 - 2.220446049250313e-16 / 2.0));   // Open01
                const double y_ = __dadd_rn(f64_with_exponent(stream_word(k0, k1, t++) >> 12, 0), -(1.0 - 2.220446049250313e-16 / 2.0));
                xt = __ddiv_rn(log(x_), ZIG_R);
                yt = log(y_);
            }
            x = u < 0.0 ? __dadd_rn(xt, -ZIG_R) : __dadd_rn(ZIG_R, -xt);
            break;
        }
        const double g = __dmul_rn((double)(stream_word(k0, k1, t++) >> 11), 1.0 / 9007199254740992.0);    // rng.gen::<f64>()
        const double lhs = __dadd_rn(d_zig_f[i + 1], __dmul_rn(__dadd_rn(d_zig_f[i], -d_zig_f[i + 1]), g));
        if (lhs < exp(__ddiv_rn(__dmul_rn(-x, x), 2.0))) break;
    }
    v[p] = __dadd_rn(0.0, __dmul_rn(1.0, x));                                            // Normal { mean: 0, std_dev: 1 }: mean + std_dev * z
    const uint64_t used = t - (uint64_t)p;
    len[p] = used > 255 ? 255 : (uint8_t)used;
}

// step 2: per block and entry offset, where the chain leaves the block and how many samples start inside it
__global__ void __launch_bounds__(ZW)
zig_scan_kernel(const uint8_t* __restrict__ len, int64_t P, int32_t* __restrict__ exit_off, int32_t* __restrict__ count) {
    __shared__ uint8_t sl[ZB];
    const int64_t b = blockIdx.x, base = b * ZB;
    for (int q = threadIdx.x; q < ZB; q += ZW) sl[q] = base + q < P ? len[base + q] : 1;
    __syncwarp();
    int pos = threadIdx.x, cnt = 0;
    while (pos < ZB) { pos += sl[pos]; ++cnt; }
    exit_off[b * ZW + threadIdx.x] = pos - ZB;
    count[b * ZW + threadIdx.x] = cnt;
}
// step 3: the chain over the blocks.  res[0] = samples that start inside [0, P), res[1] = 1 if an entry offset fell outside the window
__global__ void zig_chain_kernel(int64_t nblocks, const int32_t* __restrict__ exit_off, const int32_t* __restrict__ count,
                                 int32_t* __restrict__ entry, int64_t* __restrict__ first, int64_t* __restrict__ res) {
    int e = 0;
    int64_t n = 0;
    int bad = 0;
    for (int64_t b = 0; b < nblocks; ++b) {
        if (e >= ZW) { bad = 1; break; }
        entry[b] = e; first[b] = n;
        n += count[b * ZW + e];
        e = exit_off[b * ZW + e];
    }
    res[0] = n; res[1] = bad;
}
// step 4: every block writes the samples of its piece of the chain: sample t -> entry (t % rows, t / rows)
__global__ void __launch_bounds__(ZW)
zig_write_kernel(const double* __restrict__ v, const uint8_t* __restrict__ len, int64_t P, const int32_t* __restrict__ entry,
                 const int64_t* __restrict__ first, int64_t total, int64_t rows, double* __restrict__ out, int64_t ld, int64_t* __restrict__ words) {
    __shared__ uint8_t sl[ZB];
    __shared__ uint16_t sp[ZB];
    __shared__ int scount;
    const int64_t b = blockIdx.x, base = b * ZB;
    for (int q = threadIdx.x; q < ZB; q += ZW) sl[q] = base + q < P ? len[base + q] : 1;
    __syncwarp();
    if (threadIdx.x == 0) {
        int pos = entry[b], cnt = 0;
        while (pos < ZB) { sp[cnt++] = (uint16_t)pos; pos += sl[pos]; }
        scount = cnt;
        // the word after the last sample of the matrix: what the call consumed of the stream
        const int64_t f = first[b];
        if (f < total && f + cnt >= total) { const int last = sp[total - 1 - f]; *words = base + last + sl[last]; }
    }
    __syncwarp();
    const int64_t f = first[b];
    for (int i = threadIdx.x; i < scount; i += ZW) {
        const int64_t t = f + i;
        if (t < total) {
            const int64_t c = t / rows, r = t - c * rows;
            out[r + c * ld] = v[base + sp[i]];
        }
    }
}

double through_text(double v) { char buf[64]; snprintf(buf, sizeof buf, "%.18f", v); return strtod(buf, nullptr); }

}  // namespace

// the tables of rand_distr 0.4.3 (ziggurat_tables.rs), regenerated by the recipe of rand's utils/ziggurat_tables.py
void ziggurat_tables_host(double* x_out, double* f_out) {
    const double R = 3.6541528853610088, V = 0.00492867323399;
    double x[257];
    x[0] = V / std::exp(-R * R / 2.0);
    x[1] = R;
    for (int i = 2; i < 256; ++i) { const double last = x[i - 1]; x[i] = std::sqrt(-2.0 * std::log(V / last + std::exp(-last * last / 2.0))); }
    x[256] = 0.0;
    for (int i = 0; i < 257; ++i) { x_out[i] = through_text(x[i]); f_out[i] = through_text(std::exp(-x[i] * x[i] / 2.0)); }
}

// rand_core 0.6.4 SeedableRng::seed_from_u64 (PCG32 expansion) -> ThreeFry2x64 key (rust-random123/src/threefry.rs:23-27)
void threefry_key_from_u64(uint64_t state, uint64_t key[2]) {
    const uint64_t MUL = 6364136223846793005ull, INC = 11634580027462260723ull;
    uint32_t w[4];
    for (int i = 0; i < 4; ++i) {
        state = state * MUL + INC;
        const uint32_t xs = (uint32_t)(((state >> 18) ^ state) >> 27);
        const uint32_t rot = (uint32_t)(state >> 59);
        w[i] = (xs >> rot) | (xs << ((32 - rot) & 31));
    }
    key[0] = (uint64_t)w[0] | ((uint64_t)w[1] << 32);
    key[1] = (uint64_t)w[2] | ((uint64_t)w[3] << 32);
}

// One sketching operator (src/sketch.rs:102-130) into device memory.  Philox: entry (row_off + r, c) of stream `stream`.  ThreeFry: the
// reference's own operator -- a fresh sequential stream per call, column-major fill (row_off must be 0).
rnla_status fill_operator(int generator, int dist, uint64_t seed, uint32_t stream, int64_t rows, int64_t cols, int64_t row_off, double* out,
                          int64_t ld) {
    Ctx& c = ctx();
    if (generator == RNLA_GEN_THREEFRY) {
        if (row_off != 0 || c.nranks > 1)
            return fail(RNLA_ERR_INVALID_PARAMETERS, "the ThreeFry stream is sequential: row_offset must be 0 (single GPU, unsharded operator)");
        uint64_t key[2];
        threefry_key_from_u64(seed, key);
        if (dist == RNLA_GAUSSIAN) return fill_threefry_gaussian(key[0], key[1], rows, cols, out, ld, nullptr);
        RNLA_CUDA(fill_threefry(dist, key[0], key[1], rows, cols, out, ld, c.stream));
        return RNLA_OK;
    }
    RNLA_CUDA(fill_philox(dist, seed, stream, rows, cols, row_off, out, ld, c.stream));
    return RNLA_OK;
}

rnla_status fill_threefry_gaussian(uint64_t key0, uint64_t key1, int64_t rows, int64_t cols, double* out, int64_t ld, int64_t* words_consumed) {
    Ctx& c = ctx();
    const int64_t total = rows * cols;
    if (total <= 0) return RNLA_OK;
    static bool tables = false;
    if (!tables) {
        double hx[257], hf[257];
        ziggurat_tables_host(hx, hf);
        RNLA_CUDA(cudaMemcpyToSymbol(d_zig_x, hx, sizeof hx));
        RNLA_CUDA(cudaMemcpyToSymbol(d_zig_f, hf, sizeof hf));
        tables = true;
    }
    // words of the stream to evaluate: 2.2 % more than samples on average; grown if the chain needs more
    for (int64_t slack = total / 16 + 8192;; slack *= 4) {
        const int64_t P = total + slack, nblocks = (P + ZB - 1) / ZB;
        DevBuf v, len, exit_off, count, entry, first, res;
        RNLA_CUDA(v.alloc((size_t)P * 8)); RNLA_CUDA(len.alloc((size_t)P));
        RNLA_CUDA(exit_off.alloc((size_t)nblocks * ZW * 4)); RNLA_CUDA(count.alloc((size_t)nblocks * ZW * 4));
        RNLA_CUDA(entry.alloc((size_t)nblocks * 4)); RNLA_CUDA(first.alloc((size_t)nblocks * 8)); RNLA_CUDA(res.alloc(3 * 8));
        RNLA_CUDA(cudaMemsetAsync(res.p, 0, 3 * 8, c.stream));
        zig_eval_kernel<<<(unsigned)((P + 255) / 256), 256, 0, c.stream>>>(key0, key1, P, v.d(), len.as<uint8_t>());
        zig_scan_kernel<<<(unsigned)nblocks, ZW, 0, c.stream>>>(len.as<uint8_t>(), P, exit_off.as<int32_t>(), count.as<int32_t>());
        zig_chain_kernel<<<1, 1, 0, c.stream>>>(nblocks, exit_off.as<int32_t>(), count.as<int32_t>(), entry.as<int32_t>(), first.as<int64_t>(),
                                                res.as<int64_t>());
        g_kernel_launches += 3;
        int64_t h[3] = {0, 0, 0};
        RNLA_CUDA(cudaMemcpyAsync(h, res.p, 2 * 8, cudaMemcpyDeviceToHost, c.stream));
        RNLA_CUDA(cudaStreamSynchronize(c.stream));
        if (h[1]) return fail(RNLA_ERR_COMPUTATION, "ThreeFry Gaussian stream: a sample spans more than 32 words");
        if (h[0] < total) continue;                     // the evaluated words do not hold all samples: evaluate more
        zig_write_kernel<<<(unsigned)nblocks, ZW, 0, c.stream>>>(v.d(), len.as<uint8_t>(), P, entry.as<int32_t>(), first.as<int64_t>(), total, rows,
                                                                 out, ld, res.as<int64_t>() + 2);
        ++g_kernel_launches;
        RNLA_CUDA(cudaGetLastError());
        if (words_consumed) {
            RNLA_CUDA(cudaMemcpyAsync(h + 2, res.as<int64_t>() + 2, 8, cudaMemcpyDeviceToHost, c.stream));
            RNLA_CUDA(cudaStreamSynchronize(c.stream));
            *words_consumed = h[2];
        }
        return RNLA_OK;
    }
}

}  // namespace rnla
