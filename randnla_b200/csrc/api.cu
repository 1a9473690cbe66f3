// extern "C" surface (include/rnla.h): argument validation with the reference's own messages,
// host<->device staging for the host-buffer entry points, and the *_dev pass-throughs.
#include "drivers.cuh"
#include "gemm.cuh"
#include "panel.cuh"
#include <algorithm>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

using namespace rnla;

namespace rnla {
rnla_status dev_sketch_apply(int kind, int dist, uint64_t seed, int64_t d, int zeta, const double* dA, int64_t lda,
                             int64_t m_local, int64_t n, int64_t row_offset, double* dAsk, int64_t ldk);
rnla_status dev_generate_lowrank(double* dA, int64_t lda, int64_t m_local, int64_t n, int64_t row_offset, int64_t m_global,
                                 int64_t r0, const double* sigma_host, double eta, uint64_t seed);
}

static std::string fmt(const char* f, long long v) { char b[256]; snprintf(b, sizeof b, f, v); return b; }
// Rust's `{}` for f64 prints the shortest representation that round-trips; %g is close enough for messages
static std::string rust_f64(double v) {
    char b[64];
    if (v == (long long)v && std::fabs(v) < 1e15) snprintf(b, sizeof b, "%lld", (long long)v);
    else snprintf(b, sizeof b, "%.17g", v);
    return b;
}

static rnla_status h2d(DevBuf& buf, const double* host, size_t count) {
    RNLA_CUDA(buf.alloc(std::max<size_t>(count, 1) * 8));
    if (count) RNLA_CUDA(cudaMemcpyAsync(buf.p, host, count * 8, cudaMemcpyHostToDevice, ctx().stream));
    return RNLA_OK;
}
static rnla_status d2h(double* host, const double* dev, size_t count) {
    if (count) RNLA_CUDA(cudaMemcpyAsync(host, dev, count * 8, cudaMemcpyDeviceToHost, ctx().stream));
    RNLA_CUDA(cudaStreamSynchronize(ctx().stream));
    return RNLA_OK;
}

static rnla_status validate_svd_like(int64_t k, double epsilon, int64_t s) {
    // reference src/lora_drivers.rs:31-45 / :89-103
    if (k <= 0) return fail(RNLA_ERR_INVALID_PARAMETERS, fmt("Rank k must be positive, current input is %lld", (long long)k));
    if (!(epsilon > 0.0)) return fail(RNLA_ERR_INVALID_PARAMETERS, "Epsilon must be positive, current input is " + rust_f64(epsilon));
    if (s <= 0) return fail(RNLA_ERR_INVALID_PARAMETERS, fmt("Oversampling parameter s must be positive, current input is %lld", (long long)s));
    return RNLA_OK;
}

extern "C" {

// ------------------------------------------------------------------ RNG hooks
rnla_status rnla_philox4x32_10(int64_t nblocks, const uint32_t* hctr, const uint32_t* hkey, uint32_t* hout) {
    RNLA_API_GUARD;
    RNLA_TRY(ensure_ctx());
    if (nblocks <= 0) return RNLA_OK;
    Ctx& c = ctx();
    DevBuf dc, dk, dout;
    RNLA_CUDA(dc.alloc((size_t)nblocks * 16)); RNLA_CUDA(dk.alloc((size_t)nblocks * 8)); RNLA_CUDA(dout.alloc((size_t)nblocks * 16));
    RNLA_CUDA(cudaMemcpyAsync(dc.p, hctr, (size_t)nblocks * 16, cudaMemcpyHostToDevice, c.stream));
    RNLA_CUDA(cudaMemcpyAsync(dk.p, hkey, (size_t)nblocks * 8, cudaMemcpyHostToDevice, c.stream));
    RNLA_CUDA(philox_blocks(nblocks, dc.as<uint32_t>(), dk.as<uint32_t>(), dout.as<uint32_t>(), c.stream));
    RNLA_CUDA(cudaMemcpyAsync(hout, dout.p, (size_t)nblocks * 16, cudaMemcpyDeviceToHost, c.stream));
    RNLA_CUDA(cudaStreamSynchronize(c.stream));
    return RNLA_OK;
}
rnla_status rnla_threefry2x64_20(int64_t nblocks, const uint64_t* hctr, const uint64_t* hkey, uint64_t* hout) {
    RNLA_API_GUARD;
    RNLA_TRY(ensure_ctx());
    if (nblocks <= 0) return RNLA_OK;
    Ctx& c = ctx();
    DevBuf dc, dk, dout;
    RNLA_CUDA(dc.alloc((size_t)nblocks * 16)); RNLA_CUDA(dk.alloc((size_t)nblocks * 16)); RNLA_CUDA(dout.alloc((size_t)nblocks * 16));
    RNLA_CUDA(cudaMemcpyAsync(dc.p, hctr, (size_t)nblocks * 16, cudaMemcpyHostToDevice, c.stream));
    RNLA_CUDA(cudaMemcpyAsync(dk.p, hkey, (size_t)nblocks * 16, cudaMemcpyHostToDevice, c.stream));
    RNLA_CUDA(threefry_blocks(nblocks, dc.as<uint64_t>(), dk.as<uint64_t>(), dout.as<uint64_t>(), c.stream));
    RNLA_CUDA(cudaMemcpyAsync(hout, dout.p, (size_t)nblocks * 16, cudaMemcpyDeviceToHost, c.stream));
    RNLA_CUDA(cudaStreamSynchronize(c.stream));
    return RNLA_OK;
}

// ------------------------------------------------------------------ sketch operators
rnla_status rnla_sketch_fill_dev(int32_t generator, int32_t dist, uint64_t seed, uint32_t stream, int64_t rows, int64_t cols,
                                 int64_t row_offset, double* d_out, int64_t ld) {
    RNLA_API_GUARD;
    if (rows <= 0 || cols <= 0)   // src/sketch.rs:107-111
        return fail(RNLA_ERR_INVALID_DIMENSIONS, "Rows and columns must be greater than 0");
    if (dist < RNLA_GAUSSIAN || dist > RNLA_RADEMACHER) return fail(RNLA_ERR_INVALID_PARAMETERS, "unknown distribution");
    if (ld < rows) return fail(RNLA_ERR_INVALID_DIMENSIONS, "leading dimension smaller than rows");
    RNLA_TRY(ensure_ctx());
    if (generator != RNLA_GEN_PHILOX && generator != RNLA_GEN_THREEFRY) return fail(RNLA_ERR_INVALID_PARAMETERS, "unknown generator");
    return fill_operator(generator, dist, seed, stream, rows, cols, row_offset, d_out, ld);
}
rnla_status rnla_ziggurat_tables(double* x257, double* f257) {
    RNLA_API_GUARD;
    if (!x257 || !f257) return fail(RNLA_ERR_INVALID_PARAMETERS, "ziggurat tables: null output");
    ziggurat_tables_host(x257, f257);
    return RNLA_OK;
}
rnla_status rnla_sketch_fill(int32_t generator, int32_t dist, uint64_t seed, uint32_t stream, int64_t rows, int64_t cols,
                             int64_t row_offset, double* out, int64_t ld) {
    RNLA_API_GUARD;
    if (rows <= 0 || cols <= 0) return fail(RNLA_ERR_INVALID_DIMENSIONS, "Rows and columns must be greater than 0");
    if (ld < rows) return fail(RNLA_ERR_INVALID_DIMENSIONS, "leading dimension smaller than rows");
    RNLA_TRY(ensure_ctx());
    DevBuf d;
    RNLA_CUDA(d.alloc((size_t)rows * cols * 8));
    RNLA_TRY(rnla_sketch_fill_dev(generator, dist, seed, stream, rows, cols, row_offset, d.d(), rows));
    RNLA_CUDA(cudaMemcpy2DAsync(out, (size_t)ld * 8, d.p, (size_t)rows * 8, (size_t)rows * 8, (size_t)cols, cudaMemcpyDeviceToHost, ctx().stream));
    RNLA_CUDA(cudaStreamSynchronize(ctx().stream));
    return RNLA_OK;
}
rnla_status rnla_sketching_operator(int32_t dist, int64_t rows, int64_t cols, double* out) {
    RNLA_API_GUARD;
    if (rows <= 0 || cols <= 0) return fail(RNLA_ERR_INVALID_DIMENSIONS, "Rows and columns must be greater than 0");
    RNLA_TRY(ensure_ctx());
    return rnla_sketch_fill(RNLA_GEN_PHILOX, dist, ctx().opts.seed, 0, rows, cols, 0, out, rows);
}

rnla_status rnla_haar_sample(int64_t rows, int64_t cols, int32_t attr, double* out) {
    RNLA_API_GUARD;
    // src/sketch.rs:45-85
    int64_t m, n;
    if (attr == RNLA_ROW) {
        if (rows > cols) {
            char b[200]; snprintf(b, sizeof b, "Cannot have more rows (%lld) than columns (%lld) for row-orthonormal matrix", (long long)rows, (long long)cols);
            return fail(RNLA_ERR_INVALID_DIMENSIONS, b);
        }
        m = cols; n = rows;
    } else if (attr == RNLA_COLUMN) {
        if (cols > rows) {
            char b[200]; snprintf(b, sizeof b, "Cannot have more columns (%lld) than rows (%lld) for column-orthonormal matrix", (long long)cols, (long long)rows);
            return fail(RNLA_ERR_INVALID_DIMENSIONS, b);
        }
        m = rows; n = cols;
    } else return fail(RNLA_ERR_INVALID_PARAMETERS, "unknown matrix attribute");
    if (m <= 0 || n <= 0) return fail(RNLA_ERR_INVALID_DIMENSIONS, "Rows and columns must be greater than 0");
    RNLA_TRY(ensure_ctx());
    Ctx& c = ctx();
    DevBuf G, T;
    RNLA_CUDA(G.alloc((size_t)m * n * 8));
    RNLA_TRY(fill_operator(c.opts.generator, RNLA_GAUSSIAN, c.opts.seed, 0, m, n, 0, G.d(), m));   // :68-72 (generator = THREEFRY: the reference's own m n samples)
    ShardInfo sh{m, 0, m};
    RNLA_TRY(orth_inplace(G.d(), m, sh, (int)n, false, nullptr, nullptr));                  // :73-80 (R_ii >= 0, sign fix is a no-op)
    if (attr == RNLA_ROW) {
        RNLA_CUDA(T.alloc((size_t)m * n * 8));
        RNLA_CUDA(transpose_matrix(G.d(), m, T.d(), n, m, n, c.stream));
        return d2h(out, T.d(), (size_t)m * n);
    }
    return d2h(out, G.d(), (size_t)m * n);
}

// ------------------------------------------------------------------ helpers (host buffers)
rnla_status rnla_orth(const double* X, int64_t rows, int64_t cols, double* Q, double* R, int64_t* qcols) {
    RNLA_API_GUARD;
    if (rows <= 0 || cols <= 0) return fail(RNLA_ERR_INVALID_DIMENSIONS, "Rows and columns must be greater than 0");
    RNLA_TRY(ensure_ctx());
    Ctx& c = ctx();
    const int p = (int)std::min(rows, cols);
    if (qcols) *qcols = p;
    DevBuf dX, dR;
    RNLA_TRY(h2d(dX, X, (size_t)rows * cols));
    RNLA_CUDA(dR.alloc((size_t)p * std::max<int64_t>(cols, p) * 8));
    ShardInfo sh{rows, 0, rows};
    if (cols <= rows) {
        RNLA_TRY(orth_inplace(dX.d(), rows, sh, p, false, R ? dR.d() : nullptr, nullptr));
        RNLA_TRY(d2h(Q, dX.d(), (size_t)rows * p));
        if (R) RNLA_TRY(d2h(R, dR.d(), (size_t)p * p));
    } else {
        // wide input: thin Q of the leading rows x rows block, R = Q^T X (rows x cols)
        DevBuf dQ, dRt;
        RNLA_CUDA(dQ.alloc((size_t)rows * p * 8));
        RNLA_CUDA(copy_matrix(dX.d(), rows, dQ.d(), rows, rows, p, c.stream));
        RNLA_TRY(orth_inplace(dQ.d(), rows, sh, p, false, nullptr, nullptr));
        RNLA_TRY(d2h(Q, dQ.d(), (size_t)rows * p));
        if (R) {
            RNLA_CUDA(dRt.alloc((size_t)cols * p * 8));
            RNLA_TRY(dev_gemm_tn(dX.d(), rows, rows, cols, dQ.d(), rows, p, dRt.d(), cols, false));   // X^T Q  (cols x p)
            RNLA_CUDA(transpose_matrix(dRt.d(), cols, dR.d(), p, cols, p, c.stream));
            RNLA_TRY(d2h(R, dR.d(), (size_t)p * cols));
        }
    }
    return RNLA_OK;
}

rnla_status rnla_stabilizer(const double* X, int64_t rows, int64_t cols, double* L, int64_t* lcols) {
    RNLA_API_GUARD;
    if (rows <= 0 || cols <= 0) return fail(RNLA_ERR_INVALID_DIMENSIONS, "Rows and columns must be greater than 0");
    RNLA_TRY(ensure_ctx());
    const int64_t mn = std::min(rows, cols);
    if (lcols) *lcols = mn;
    DevBuf dX, dL;
    RNLA_TRY(h2d(dX, X, (size_t)rows * cols));
    RNLA_CUDA(dL.alloc((size_t)rows * mn * 8));
    RNLA_TRY(dev_stabilizer(dX.d(), rows, rows, cols, dL.d(), rows));
    return d2h(L, dL.d(), (size_t)rows * mn);
}

static rnla_status check_range_args(int64_t m, int64_t n, int64_t k) {
    if (m <= 0 || n <= 0) return fail(RNLA_ERR_INVALID_DIMENSIONS, "Rows and columns must be greater than 0");
    if (k <= 0) return fail(RNLA_ERR_INVALID_PARAMETERS, fmt("Rank k must be positive, current input is %lld", (long long)k));
    return RNLA_OK;
}

rnla_status rnla_tsog1(const double* A, int64_t m, int64_t n, int64_t k, int32_t num_passes, int32_t passes_per_stab, double* S) {
    RNLA_API_GUARD;
    RNLA_TRY(check_range_args(m, n, k));
    if (num_passes < 0 || passes_per_stab <= 0) return fail(RNLA_ERR_INVALID_PARAMETERS, "num_passes must be >= 0 and passes_per_stab > 0");
    if (k > std::min(m, n)) return fail(RNLA_ERR_INVALID_DIMENSIONS, "tsog1: k must not exceed min(m, n)");
    RNLA_TRY(ensure_ctx());
    phases_reset();
    DevBuf dA, dS;
    RNLA_TRY(h2d(dA, A, (size_t)m * n));
    RNLA_CUDA(dS.alloc((size_t)n * k * 8));
    ShardInfo sh{m, 0, m};
    RNLA_TRY(dev_tsog1(dA.d(), m, sh, n, (int)k, num_passes, passes_per_stab, ctx().opts, dS.d()));
    return d2h(S, dS.d(), (size_t)n * k);
}

rnla_status rnla_rf1(const double* A, int64_t m, int64_t n, int64_t k, double* Q, int64_t* qcols) {
    RNLA_API_GUARD;
    RNLA_TRY(check_range_args(m, n, k));
    RNLA_TRY(ensure_ctx());
    phases_reset();
    const rnla_options o = ctx().opts;
    const int l = (int)std::min<int64_t>(k, std::min(m, n));
    if (qcols) *qcols = l;
    DevBuf dA, dQ;
    RNLA_TRY(h2d(dA, A, (size_t)m * n));
    RNLA_CUDA(dQ.alloc((size_t)m * l * 8));
    ShardInfo sh{m, 0, m};
    RNLA_TRY(dev_rf1(dA.d(), m, sh, n, l, o.num_passes > 0 ? o.num_passes : 2, o.passes_per_stab > 0 ? o.passes_per_stab : 1, o, dQ.d(), m));
    return d2h(Q, dQ.d(), (size_t)m * l);
}

rnla_status rnla_qb1(const double* A, int64_t m, int64_t n, int64_t k, double epsilon, double* Q, double* B, int64_t* qcols) {
    RNLA_API_GUARD;
    (void)epsilon;   // ignored, as in the reference (src/lora_helpers.rs:18-19)
    RNLA_TRY(check_range_args(m, n, k));
    RNLA_TRY(ensure_ctx());
    phases_reset();
    Ctx& c = ctx();
    const rnla_options o = c.opts;
    const int l = (int)std::min<int64_t>(k, std::min(m, n));
    if (qcols) *qcols = l;
    DevBuf dA, dQ, dBt, dB;
    RNLA_TRY(h2d(dA, A, (size_t)m * n));
    RNLA_CUDA(dQ.alloc((size_t)m * l * 8)); RNLA_CUDA(dBt.alloc((size_t)n * l * 8)); RNLA_CUDA(dB.alloc((size_t)n * l * 8));
    ShardInfo sh{m, 0, m};
    RNLA_TRY(dev_qb1(dA.d(), m, sh, n, l, o.num_passes > 0 ? o.num_passes : 2, o.passes_per_stab > 0 ? o.passes_per_stab : 1, o, dQ.d(), m, dBt.d()));
    RNLA_CUDA(transpose_matrix(dBt.d(), n, dB.d(), l, n, l, c.stream));
    RNLA_TRY(d2h(Q, dQ.d(), (size_t)m * l));
    return d2h(B, dB.d(), (size_t)l * n);
}

// ------------------------------------------------------------------ drivers (host buffers)
rnla_status rnla_rand_svd(const double* A, int64_t m, int64_t n, int64_t k, double epsilon, int64_t s,
                          double* U, double* S, double* Vt, int64_t* r_out) {
    RNLA_API_GUARD;
    RNLA_TRY(validate_svd_like(k, epsilon, s));
    if (m <= 0 || n <= 0) return fail(RNLA_ERR_INVALID_DIMENSIONS, "Rows and columns must be greater than 0");
    RNLA_TRY(ensure_ctx());
    Ctx& c = ctx();
    const int64_t l = std::min<int64_t>(k + s, std::min(m, n));
    const int64_t r = std::min(k, l);
    DevBuf dA, dU, dS, dVt;
    const rnla_options& o = c.opts;
    const int q = o.num_passes > 0 ? o.num_passes : 2;
    const bool streamed = o.mode == RNLA_MODE_INTENDED && q >= 2 && q % 2 == 0 && m >= 8192;
    // outputs first: nothing below may return between arming the one-shot upload hook and the driver call that consumes it
    RNLA_CUDA(dU.alloc((size_t)m * r * 8)); RNLA_CUDA(dS.alloc((size_t)r * 8)); RNLA_CUDA(dVt.alloc((size_t)r * n * 8));
    struct HookGuard { Ctx& c; ~HookGuard() { c.first_pass_hook = nullptr; c.block_landed_hook = nullptr; } } hook_guard{c};
    if (!streamed) {
        RNLA_TRY(h2d(dA, A, (size_t)m * n));
    } else {
        // The first product of the power iteration, Y = A * Omega, only needs the rows of A it multiplies: upload A in row
        // blocks on a copy stream and multiply each block as soon as it has landed.  The whole pass hides behind the PCIe
        // transfer (32 GB at ~55 GB/s against 26 ms of DMMA at the headline size); only the last block's GEMM is exposed.
        RNLA_CUDA(dA.alloc((size_t)m * n * 8));
        double* dAp = dA.d();
        c.first_pass_hook = [&c, A, dAp, m, n, l, o](double* S, double* Y, int64_t ldy) -> rnla_status {
            struct Ev {                                         // destroyed on every return path
                cudaEvent_t e = nullptr;
                ~Ev() { if (e) cudaEventDestroy(e); }
            } ev_ready, ev_landed;
            RNLA_CUDA(cudaEventCreateWithFlags(&ev_ready.e, cudaEventDisableTiming));
            RNLA_CUDA(cudaEventCreateWithFlags(&ev_landed.e, cudaEventDisableTiming));
            cudaEvent_t ready = ev_ready.e, landed = ev_landed.e;
            const bool fused = o.generator == RNLA_GEN_PHILOX && (o.fused_sketch == 1 || (o.fused_sketch == 2 && (double)n * l * 8.0 > 48.0 * 1024 * 1024));
            if (!fused) RNLA_TRY(fill_operator(o.generator, o.dist, o.seed, 1 /* STREAM_RANGE_N */, n, l, 0, S, n));
            RNLA_CUDA(cudaEventRecord(ready, c.stream));                    // dA is allocated stream-ordered on c.stream
            RNLA_CUDA(cudaStreamWaitEvent(c.copy_stream, ready, 0));
            const int64_t nblk = std::min<int64_t>(16, (m + 8191) / 8192);
            const int64_t mb = (((m + nblk - 1) / nblk + 127) / 128) * 128;
            rnla_status rc = RNLA_OK;
            for (int64_t r0 = 0; r0 < m && rc == RNLA_OK; r0 += mb) {
                const int64_t rows = std::min(mb, m - r0);
                cudaError_t e = cudaMemcpy2DAsync(dAp + r0, (size_t)m * 8, A + r0, (size_t)m * 8, (size_t)rows * 8, (size_t)n,
                                                  cudaMemcpyHostToDevice, c.copy_stream);
                if (e == cudaSuccess) e = cudaEventRecord(landed, c.copy_stream);
                if (e == cudaSuccess) e = cudaStreamWaitEvent(c.stream, landed, 0);
                if (e != cudaSuccess) { rc = cuda_fail(e, "streamed upload of A", __FILE__, __LINE__); break; }
                rc = fused ? dev_sketch_gemm(dAp + r0, m, rows, n, o.dist, o.seed, 1 /* STREAM_RANGE_N */, l, Y + r0, ldy)
                           : dev_gemm_nn(dAp + r0, m, rows, n, S, n, l, Y + r0, ldy);
                if (rc == RNLA_OK && c.block_landed_hook) rc = c.block_landed_hook(r0, rows);     // int8 passes: split this block now
            }
            return rc;
        };
    }
    int64_t rr = 0;
    {
        const rnla_status st = dev_rand_svd(dA.d(), m, m, n, k, s, c.opts, dU.d(), m, dS.d(), dVt.d(), r, &rr);
        c.first_pass_hook = nullptr;
        if (st != RNLA_OK) { cudaStreamSynchronize(c.copy_stream); return st; }
    }
    std::vector<double> sig((size_t)r);
    RNLA_TRY(d2h(U, dU.d(), (size_t)m * r));
    RNLA_TRY(d2h(Vt, dVt.d(), (size_t)r * n));
    RNLA_TRY(d2h(sig.data(), dS.d(), (size_t)r));
    // S is returned as a dense r x r diagonal matrix (src/lora_drivers.rs:64)
    memset(S, 0, (size_t)r * r * 8);
    for (int64_t i = 0; i < r; ++i) S[i + i * r] = sig[(size_t)i];
    if (r_out) *r_out = r;
    return RNLA_OK;
}

rnla_status rnla_rand_evd1(const double* A, int64_t n, int64_t k, double epsilon, int64_t s, double* V, double* lambda, int64_t* r_out) {
    RNLA_API_GUARD;
    RNLA_TRY(validate_svd_like(k, epsilon, s));
    if (n <= 0) return fail(RNLA_ERR_INVALID_DIMENSIONS, "Rows and columns must be greater than 0");
    RNLA_TRY(ensure_ctx());
    Ctx& c = ctx();
    const int64_t l = std::min<int64_t>(k + s, n);
    const int64_t r = std::min(k, l);
    DevBuf dA, dV, dL;
    RNLA_TRY(h2d(dA, A, (size_t)n * n));
    RNLA_CUDA(dV.alloc((size_t)n * r * 8)); RNLA_CUDA(dL.alloc((size_t)r * 8));
    int64_t rr = 0;
    RNLA_TRY(dev_rand_evd1(dA.d(), n, n, n, k, s, c.opts, dV.d(), n, dL.d(), &rr));
    RNLA_TRY(d2h(V, dV.d(), (size_t)n * r));
    RNLA_TRY(d2h(lambda, dL.d(), (size_t)r));
    if (r_out) *r_out = rr;
    return RNLA_OK;
}

rnla_status rnla_rand_evd2(const double* A, int64_t n, int64_t k, int64_t s, double* V, double* lambda, int64_t* r_out) {
    RNLA_API_GUARD;
    // src/lora_drivers.rs:169-173: only k is validated (s may be 0)
    if (k <= 0) return fail(RNLA_ERR_INVALID_PARAMETERS, fmt("Rank k must be positive, current input is %lld", (long long)k));
    if (s < 0) return fail(RNLA_ERR_INVALID_PARAMETERS, "Oversampling parameter s must be non-negative");
    if (n <= 0) return fail(RNLA_ERR_INVALID_DIMENSIONS, "Rows and columns must be greater than 0");
    RNLA_TRY(ensure_ctx());
    Ctx& c = ctx();
    const int64_t kk = std::min(k, n);
    DevBuf dA, dV, dL;
    RNLA_TRY(h2d(dA, A, (size_t)n * n));
    RNLA_CUDA(dV.alloc((size_t)n * kk * 8)); RNLA_CUDA(dL.alloc((size_t)kk * 8));
    int64_t rr = 0;
    RNLA_TRY(dev_rand_evd2(dA.d(), n, n, n, k, s, c.opts, dV.d(), n, dL.d(), &rr));
    RNLA_TRY(d2h(V, dV.d(), (size_t)n * rr));
    RNLA_TRY(d2h(lambda, dL.d(), (size_t)rr));
    if (r_out) *r_out = rr;
    return RNLA_OK;
}

// ------------------------------------------------------------------ drivers (device buffers)
static const rnla_options& pick(const rnla_options* o) { return o ? *o : ctx().opts; }

rnla_status rnla_rand_svd_dev(const double* dA, int64_t lda, int64_t m_local, int64_t n, int64_t k, int64_t s,
                              const rnla_options* opt, double* dU, int64_t ldu, double* dSigma, double* dVt, int64_t ldvt, int64_t* r) {
    RNLA_API_GUARD;
    RNLA_TRY(validate_svd_like(k, 1.0, s));
    RNLA_TRY(ensure_ctx());
    const rnla_options o = pick(opt);
    host_trace_mark("enter rnla_rand_svd_dev");
    const rnla_status st = dev_rand_svd(dA, lda, m_local, n, k, s, o, dU, ldu, dSigma, dVt, ldvt, r);
    host_trace_mark("leave rnla_rand_svd_dev");
    return st;
}
rnla_status rnla_rand_evd1_dev(const double* dA, int64_t lda, int64_t n, int64_t k, int64_t s, const rnla_options* opt,
                               double* dV, int64_t ldv, double* dLambda, int64_t* r) {
    RNLA_API_GUARD;
    RNLA_TRY(validate_svd_like(k, 1.0, s));
    RNLA_TRY(ensure_ctx());
    const rnla_options o = pick(opt);
    return dev_rand_evd1(dA, lda, n, n, k, s, o, dV, ldv, dLambda, r);
}
rnla_status rnla_rand_evd2_dev(const double* dA, int64_t lda, int64_t n, int64_t k, int64_t s, const rnla_options* opt,
                               double* dV, int64_t ldv, double* dLambda, int64_t* r) {
    RNLA_API_GUARD;
    if (k <= 0) return fail(RNLA_ERR_INVALID_PARAMETERS, fmt("Rank k must be positive, current input is %lld", (long long)k));
    RNLA_TRY(ensure_ctx());
    const rnla_options o = pick(opt);
    return dev_rand_evd2(dA, lda, n, n, k, s, o, dV, ldv, dLambda, r);
}

// ------------------------------------------------------------------ sketch step of sketch_and_precondition
int64_t rnla_sketch_dim(int64_t m, int64_t n, double sampling_factor, int32_t rule) {
    RNLA_API_GUARD;
    if (rule == 0) {
        // src/sketch_and_precondition.rs:49,105
        if (sampling_factor * (double)n > (double)m) return m;
        return (int64_t)std::floor(sampling_factor * (double)n);
    }
    // :172
    int64_t d = (int64_t)std::floor(sampling_factor * (double)n);
    if (d < 1) d = 1;
    if (d > m) d = m;
    return d;
}

rnla_status rnla_sketch_apply_dev(int32_t kind, int32_t dist, uint64_t seed, int64_t d, int32_t zeta, const double* dA, int64_t lda,
                                  int64_t m_local, int64_t n, int64_t row_offset, double* dA_sk, int64_t ld_sk) {
    RNLA_API_GUARD;
    RNLA_TRY(ensure_ctx());
    return dev_sketch_apply(kind, dist, seed, d, zeta, dA, lda, m_local, n, row_offset, dA_sk, ld_sk);
}

rnla_status rnla_sketch_apply(int32_t kind, int32_t dist, uint64_t seed, int64_t d, int32_t zeta, const double* A, int64_t m, int64_t n,
                              const double* b, int64_t nrhs, double* A_sk, double* b_sk) {
    RNLA_API_GUARD;
    if (m <= 0 || n <= 0 || d <= 0) return fail(RNLA_ERR_INVALID_DIMENSIONS, "Rows and columns must be greater than 0");
    RNLA_TRY(ensure_ctx());
    DevBuf dA, dAsk, db, dbsk;
    RNLA_TRY(h2d(dA, A, (size_t)m * n));
    RNLA_CUDA(dAsk.alloc((size_t)d * n * 8));
    RNLA_TRY(dev_sketch_apply(kind, dist, seed, d, zeta, dA.d(), m, m, n, 0, dAsk.d(), d));
    RNLA_TRY(d2h(A_sk, dAsk.d(), (size_t)d * n));
    if (b && nrhs > 0) {
        RNLA_TRY(h2d(db, b, (size_t)m * nrhs));
        RNLA_CUDA(dbsk.alloc((size_t)d * nrhs * 8));
        RNLA_TRY(dev_sketch_apply(kind, dist, seed, d, zeta, db.d(), m, m, nrhs, 0, dbsk.d(), d));
        RNLA_TRY(d2h(b_sk, dbsk.d(), (size_t)d * nrhs));
    }
    return RNLA_OK;
}

// ------------------------------------------------------------------ building blocks
rnla_status rnla_gemm_nn_dev(const double* dA, int64_t lda, int64_t m, int64_t K, const double* dB, int64_t ldb, int64_t N, double* dC, int64_t ldc) {
    RNLA_API_GUARD;
    RNLA_TRY(ensure_ctx());
    return dev_gemm_nn(dA, lda, m, K, dB, ldb, N, dC, ldc);
}
// ---- blendenpik_overdetermined, end to end (reference src/sketch_and_precondition.rs:26-59)
static rnla_status validate_lsq(int64_t m, int64_t n, double epsilon, int64_t l, double sampling_factor) {
    char buf[160];
    if (m < n) {                                                                                   // :29-33
        snprintf(buf, sizeof buf, "Need more columns than rows, found %lld rows and %lld columns", (long long)m, (long long)n);
        return fail(RNLA_ERR_NOT_OVERDETERMINED, buf);
    }
    if (sampling_factor < 1.0) {                                                                   // :34-38
        snprintf(buf, sizeof buf, "Sampling factor must be greater than 1, current input is %g", sampling_factor);
        return fail(RNLA_ERR_INVALID_PARAMETERS, buf);
    }
    if (epsilon <= 0.0) {                                                                          // :39-43
        snprintf(buf, sizeof buf, "Epsilon must be positive, current input is %g", epsilon);
        return fail(RNLA_ERR_INVALID_PARAMETERS, buf);
    }
    if (l <= 0) {                                                                                  // :44-48
        snprintf(buf, sizeof buf, "Number of iterations must be positive, current input is %lld", (long long)l);
        return fail(RNLA_ERR_INVALID_PARAMETERS, buf);
    }
    return RNLA_OK;
}
rnla_status rnla_blendenpik_overdetermined_dev(const double* dA, int64_t lda, int64_t m_local, int64_t n, const double* db,
                                               double epsilon, int64_t l, double sampling_factor, int32_t kind, int32_t dist,
                                               int32_t zeta, double* dx, int64_t* iterations, int32_t* converged) {
    RNLA_API_GUARD;
    RNLA_TRY(ensure_ctx());
    ShardInfo sh;
    RNLA_TRY(shard_layout(m_local, &sh));
    RNLA_TRY(validate_lsq(sh.rows_global, n, epsilon, l, sampling_factor));
    return dev_blendenpik(dA, lda, m_local, n, db, epsilon, l, sampling_factor, kind, dist, zeta, ctx().opts.seed, dx, iterations, converged);
}
rnla_status rnla_blendenpik_overdetermined(const double* A, int64_t m, int64_t n, const double* b, double epsilon, int64_t l,
                                           double sampling_factor, int32_t kind, int32_t dist, int32_t zeta, double* x,
                                           int64_t* iterations, int32_t* converged) {
    RNLA_API_GUARD;
    RNLA_TRY(validate_lsq(m, n, epsilon, l, sampling_factor));
    RNLA_TRY(ensure_ctx());
    DevBuf dA, db, dx;
    RNLA_TRY(h2d(dA, A, (size_t)m * n));
    RNLA_TRY(h2d(db, b, (size_t)m));
    RNLA_CUDA(dx.alloc((size_t)n * 8));
    RNLA_TRY(dev_blendenpik(dA.d(), m, m, n, db.d(), epsilon, l, sampling_factor, kind, dist, zeta, ctx().opts.seed, dx.d(), iterations, converged));
    return d2h(x, dx.d(), (size_t)n);
}

// ---- planning decisions, callable without a GPU (tests/test_host_logic.py) ---------------------------------------------
void rnla_plan_gemm(int64_t m, int64_t n, int64_t N, int32_t sms, int32_t* out /* 4 */) {
    RNLA_API_GUARD;
    int chunks = 0, tiles = 0; int64_t chunk_rows = 0;
    gemm_tn_plan_info(m, n, N, sms, &chunks, &chunk_rows, &tiles);
    out[0] = gemm_nn_ksplit(m, n, N, sms); out[1] = chunks; out[2] = (int32_t)chunk_rows; out[3] = tiles;
}
int32_t rnla_plan_saso_block(int64_t d, int32_t zeta, int32_t width, int64_t n, int64_t nchunks, int32_t sms, int32_t* shape /* 4 */,
                             int32_t* desc /* 4 per CTA */, int32_t cap, int32_t* nslots) {
    RNLA_API_GUARD;
    int bpt = 0, cb = 0, parts = 0;
    if (!saso_block_shape(d, zeta, width, n, &bpt, &cb, &parts)) return -1;
    const int ncg = (int)((n + cb - 1) / cb);
    std::vector<SbFrag> work; std::vector<int> fix_cg, fix_off, fix_cnt, slots;
    saso_block_worklist(ncg, nchunks, sms, work, fix_cg, fix_off, fix_cnt, slots);
    shape[0] = bpt; shape[1] = cb; shape[2] = parts; shape[3] = ncg;
    for (int i = 0; i < (int)work.size() && i < cap; ++i) { desc[4 * i] = work[i].cg; desc[4 * i + 1] = work[i].lo; desc[4 * i + 2] = work[i].hi; desc[4 * i + 3] = work[i].slot; }
    if (nslots) *nslots = (int)slots.size();
    return (int32_t)work.size();
}

rnla_status rnla_lsrn_overdetermined_dev(const double* dA, int64_t lda, int64_t m_local, int64_t n, const double* db, double epsilon,
                                         int64_t l, double sampling_factor, int32_t kind, int32_t dist, int32_t zeta, double* dx,
                                         int64_t* iterations, int32_t* converged) {
    RNLA_API_GUARD;
    RNLA_TRY(ensure_ctx());
    ShardInfo sh;
    RNLA_TRY(shard_layout(m_local, &sh));
    RNLA_TRY(validate_lsq(sh.rows_global, n, epsilon, l, sampling_factor));                          // :85-104
    return dev_lsrn(dA, lda, m_local, n, db, epsilon, l, sampling_factor, kind, dist, zeta, ctx().opts.seed, dx, iterations, converged);
}
rnla_status rnla_lsrn_overdetermined(const double* A, int64_t m, int64_t n, const double* b, double epsilon, int64_t l,
                                     double sampling_factor, int32_t kind, int32_t dist, int32_t zeta, double* x,
                                     int64_t* iterations, int32_t* converged) {
    RNLA_API_GUARD;
    RNLA_TRY(validate_lsq(m, n, epsilon, l, sampling_factor));
    RNLA_TRY(ensure_ctx());
    DevBuf dA, db, dx;
    RNLA_TRY(h2d(dA, A, (size_t)m * n));
    RNLA_TRY(h2d(db, b, (size_t)m));
    RNLA_CUDA(dx.alloc((size_t)n * 8));
    RNLA_TRY(dev_lsrn(dA.d(), m, m, n, db.d(), epsilon, l, sampling_factor, kind, dist, zeta, ctx().opts.seed, dx.d(), iterations, converged));
    return d2h(x, dx.d(), (size_t)n);
}

// ---- reference src/cg.rs: cgls, conjugate_grad, verify_solution
rnla_status rnla_cgls_dev(const double* dA, int64_t lda, int64_t m_local, int64_t n, const double* db, double tolerance,
                          int64_t num_iterations, double* dx, int64_t* iterations, int32_t* converged) {
    RNLA_API_GUARD;
    if (!dA || !db || !dx) return fail(RNLA_ERR_INVALID_PARAMETERS, "cgls: null argument");
    if (m_local < 0 || n < 1) return fail(RNLA_ERR_INVALID_DIMENSIONS, "cgls: a must have at least one column");
    RNLA_TRY(ensure_ctx());
    phases_reset();
    int64_t it = 0; int32_t conv = 0;
    RNLA_TRY(dev_cgls_operator(dA, lda, m_local, n, db, nullptr, dx, tolerance, num_iterations, &it, &conv));
    RNLA_CUDA(cudaStreamSynchronize(ctx().stream));
    if (iterations) *iterations = it;
    if (converged) *converged = conv;
    return RNLA_OK;
}
rnla_status rnla_cgls(const double* A, int64_t m, int64_t n, const double* b, double tolerance, int64_t num_iterations,
                      const double* x0, double* x, int64_t* iterations, int32_t* converged) {
    RNLA_API_GUARD;
    if (!A || !b || !x) return fail(RNLA_ERR_INVALID_PARAMETERS, "cgls: null argument");
    if (m < 1 || n < 1) return fail(RNLA_ERR_INVALID_DIMENSIONS, "cgls: a must have at least one row and one column");
    RNLA_TRY(ensure_ctx());
    DevBuf dA, db, dx;
    RNLA_TRY(h2d(dA, A, (size_t)m * n));
    RNLA_TRY(h2d(db, b, (size_t)m));
    if (x0) RNLA_TRY(h2d(dx, x0, (size_t)n));
    else { RNLA_CUDA(dx.alloc((size_t)n * 8)); RNLA_CUDA(cudaMemsetAsync(dx.p, 0, (size_t)n * 8, ctx().stream)); }        // :29
    RNLA_TRY(rnla_cgls_dev(dA.d(), m, m, n, db.d(), tolerance, num_iterations, dx.d(), iterations, converged));
    return d2h(x, dx.d(), (size_t)n);
}
rnla_status rnla_conjugate_grad_dev(const double* dA, int64_t lda, int64_t n, const double* db, double* dx, int64_t* iterations,
                                    int32_t* converged) {
    RNLA_API_GUARD;
    if (!dA || !db || !dx) return fail(RNLA_ERR_INVALID_PARAMETERS, "conjugate_grad: null argument");
    RNLA_TRY(ensure_ctx());
    return dev_conjugate_grad(dA, lda, n, db, dx, iterations, converged);
}
rnla_status rnla_conjugate_grad(const double* A, int64_t n, const double* b, const double* x0, double* x, int64_t* iterations,
                                int32_t* converged) {
    RNLA_API_GUARD;
    if (!A || !b || !x) return fail(RNLA_ERR_INVALID_PARAMETERS, "conjugate_grad: null argument");
    if (n < 1) return fail(RNLA_ERR_INVALID_DIMENSIONS, "conjugate_grad: empty system");
    RNLA_TRY(ensure_ctx());
    DevBuf dA, db, dx;
    RNLA_TRY(h2d(dA, A, (size_t)n * n));
    RNLA_TRY(h2d(db, b, (size_t)n));
    if (x0) RNLA_TRY(h2d(dx, x0, (size_t)n));
    else {                                                                                         // DVector::from_element(n, 1.0)  :88
        std::vector<double> ones((size_t)n, 1.0);
        RNLA_TRY(h2d(dx, ones.data(), (size_t)n));
    }
    RNLA_TRY(dev_conjugate_grad(dA.d(), n, n, db.d(), dx.d(), iterations, converged));
    return d2h(x, dx.d(), (size_t)n);
}
rnla_status rnla_verify_solution(const double* A, int64_t m, int64_t n, const double* b, const double* x, double* residual_norm) {
    RNLA_API_GUARD;
    if (!A || !b || !x || !residual_norm) return fail(RNLA_ERR_INVALID_PARAMETERS, "verify_solution: null argument");
    if (m < 1 || n < 1) return fail(RNLA_ERR_INVALID_DIMENSIONS, "verify_solution: empty system");
    RNLA_TRY(ensure_ctx());
    DevBuf dA, db, dx;
    RNLA_TRY(h2d(dA, A, (size_t)m * n));
    RNLA_TRY(h2d(db, b, (size_t)m));
    RNLA_TRY(h2d(dx, x, (size_t)n));
    return dev_verify_solution(dA.d(), m, m, n, db.d(), dx.d(), residual_norm);
}

// ---- lsqr (reference src/solvers.rs:115-278)
rnla_status rnla_lsqr_dev(const double* dA, int64_t lda, int64_t m_local, int64_t n, const double* db, double damp, double atol,
                          double btol, double conlim, int64_t iter_lim, int32_t calc_var, const double* dx0, double* dx,
                          rnla_lsqr_result* result, double* arnorms, int64_t arnorms_cap, double* dvar) {
    RNLA_API_GUARD;
    if (!result || !dx || !dA || !db) return fail(RNLA_ERR_INVALID_PARAMETERS, "lsqr: null argument");
    if (m_local < 0 || n < 1) return fail(RNLA_ERR_INVALID_DIMENSIONS, "lsqr: a must have at least one row and one column");
    RNLA_TRY(ensure_ctx());
    return dev_lsqr(dA, lda, m_local, n, db, damp, atol, btol, conlim, iter_lim, calc_var, dx0, dx, result, arnorms, arnorms_cap, dvar);
}
rnla_status rnla_lsqr(const double* A, int64_t m, int64_t n, const double* b, double damp, double atol, double btol, double conlim,
                      int64_t iter_lim, int32_t calc_var, const double* x0, double* x, rnla_lsqr_result* result, double* arnorms,
                      int64_t arnorms_cap, double* var) {
    RNLA_API_GUARD;
    if (!result || !x || !A || !b) return fail(RNLA_ERR_INVALID_PARAMETERS, "lsqr: null argument");
    if (calc_var && !var) return fail(RNLA_ERR_INVALID_PARAMETERS, "lsqr: calc_var needs a var buffer");
    if (m < 1 || n < 1) return fail(RNLA_ERR_INVALID_DIMENSIONS, "lsqr: a must have at least one row and one column");
    RNLA_TRY(ensure_ctx());
    DevBuf dA, db, dx0, dx, dvar;
    RNLA_TRY(h2d(dA, A, (size_t)m * n));
    RNLA_TRY(h2d(db, b, (size_t)m));
    if (x0) RNLA_TRY(h2d(dx0, x0, (size_t)n));
    RNLA_CUDA(dx.alloc((size_t)n * 8));
    if (var) RNLA_CUDA(dvar.alloc((size_t)n * 8));
    RNLA_TRY(dev_lsqr(dA.d(), m, m, n, db.d(), damp, atol, btol, conlim, iter_lim, calc_var, x0 ? dx0.d() : nullptr, dx.d(), result,
                      arnorms, arnorms_cap, var ? dvar.d() : nullptr));
    if (var) RNLA_TRY(d2h(var, dvar.d(), (size_t)n));
    return d2h(x, dx.d(), (size_t)n);
}

rnla_status rnla_gemv_dev(const double* dA, int64_t lda, int64_t m, int64_t n, int32_t trans, const double* dx, double* dy) {
    RNLA_API_GUARD;
    RNLA_TRY(ensure_ctx());
    return trans ? dev_gemv_t(dA, lda, m, n, dx, dy) : dev_gemv_n(dA, lda, m, n, dx, dy);
}

int32_t rnla_plan_normal_pass(uint64_t base_address, int64_t lda, int64_t n, int32_t* out /* 7 */) {
    if (n < 1 || n > 2048 || lda < 1 || (base_address & 7) != 0) return 0;
    const NormalPassPlan p = normal_pass_plan(base_address, lda, n);
    out[0] = p.cluster; out[1] = p.ncb; out[2] = p.ne; out[3] = p.shift_e; out[4] = p.shift_o; out[5] = p.pitch; out[6] = p.stage_bytes;
    return 1;
}
int32_t rnla_normal_pass_supported(const double* dA, int64_t lda, int64_t m_local, int64_t n) {
    RNLA_API_GUARD;
    if (ensure_ctx() != RNLA_OK) return 0;
    return normal_pass_supported(dA, lda, m_local, n) ? 1 : 0;
}
rnla_status rnla_normal_pass_dev(const double* dA, int64_t lda, int64_t m_local, int64_t n, const double* dx, double cq, const double* dy,
                                 double cy, double* du, double* dt) {
    RNLA_API_GUARD;
    RNLA_TRY(ensure_ctx());
    return dev_normal_pass(dA, lda, m_local, n, dx, cq, dy, cy, du, dt);
}

rnla_status rnla_sketch_gemm_dev(const double* dA, int64_t lda, int64_t m, int64_t K, int32_t dist, uint64_t seed, uint32_t stream,
                                 int64_t N, double* dC, int64_t ldc) {
    RNLA_API_GUARD;
    RNLA_TRY(ensure_ctx());
    return dev_sketch_gemm(dA, lda, m, K, dist, seed, stream, N, dC, ldc);
}
rnla_status rnla_gemm_tn_dev(const double* dA, int64_t lda, int64_t m, int64_t n, const double* dQ, int64_t ldq, int64_t N,
                             double* dZ, int64_t ldz, int32_t allreduce) {
    RNLA_API_GUARD;
    RNLA_TRY(ensure_ctx());
    return dev_gemm_tn(dA, lda, m, n, dQ, ldq, N, dZ, ldz, allreduce != 0);
}
rnla_status rnla_orth_dev(double* dX, int64_t ldx, int64_t rows_local, int64_t cols, int32_t sharded, double* dR, int64_t* deficient) {
    RNLA_API_GUARD;
    RNLA_TRY(ensure_ctx());
    ShardInfo sh{rows_local, 0, rows_local};
    if (sharded) RNLA_TRY(shard_layout(rows_local, &sh));
    return orth_inplace(dX, ldx, sh, (int)cols, sharded != 0, dR, deficient);
}
rnla_status rnla_small_svd_dev(const double* dM, int64_t ldm, int64_t p, double* dU, double* dSigma, double* dV) {
    RNLA_API_GUARD;
    RNLA_TRY(ensure_ctx());
    Ctx& c = ctx();
    if (p <= 0 || p > 1024) return fail(RNLA_ERR_INVALID_DIMENSIONS, "small_svd: 1 <= p <= 1024");
    // the same route the drivers take for B: M = Q R (CholeskyQR2 with deficiency handling), one-sided Jacobi on R^T
    // (R = Ur diag(sigma) Vr^T), U = Q Ur, V = Vr
    DevBuf Q, R, Ur, work, info;
    RNLA_CUDA(Q.alloc((size_t)p * p * 8)); RNLA_CUDA(R.alloc((size_t)p * p * 8)); RNLA_CUDA(Ur.alloc((size_t)p * p * 8));
    RNLA_CUDA(work.alloc(jacobi_svd_work_doubles((int)p) * 8)); RNLA_CUDA(info.alloc(8));
    RNLA_CUDA(copy_matrix(dM, ldm, Q.d(), p, p, p, c.stream));
    ShardInfo sh{p, 0, p};
    RNLA_TRY(orth_inplace(Q.d(), p, sh, (int)p, false, R.d(), nullptr));
    RNLA_CUDA(jacobi_svd(R.d(), p, (int)p, Ur.d(), p, dSigma, dV, p, work.d(), info.as<int>(), c.stream, 1));
    int h[2];
    RNLA_CUDA(cudaMemcpyAsync(h, info.p, 8, cudaMemcpyDeviceToHost, c.stream));
    RNLA_CUDA(cudaStreamSynchronize(c.stream));
    g_last_jacobi_sweeps = h[0];
    if (h[1]) return fail(RNLA_ERR_MATRIX_DECOMPOSITION, "SVD decomposition failed");
    RNLA_TRY(dev_gemm_nn(Q.d(), p, p, p, Ur.d(), p, p, dU, p));
    RNLA_CUDA(cudaStreamSynchronize(c.stream));
    return RNLA_OK;
}
int32_t rnla_last_jacobi_sweeps(void) { return g_last_jacobi_sweeps; }

rnla_status rnla_small_eigh_dev(const double* dC, int64_t ldc, int64_t p, double* dW, double* dLambda) {
    RNLA_API_GUARD;
    RNLA_TRY(ensure_ctx());
    Ctx& c = ctx();
    if (p <= 0 || p > 1024) return fail(RNLA_ERR_INVALID_DIMENSIONS, "small_eigh: 1 <= p <= 1024");
    DevBuf work, info;
    RNLA_CUDA(work.alloc(jacobi_svd_work_doubles((int)p) * 8)); RNLA_CUDA(info.alloc(8));
    RNLA_CUDA(jacobi_eigh(dC, ldc, (int)p, dW, p, dLambda, 0, work.d(), info.as<int>(), c.stream));
    int h[2];
    RNLA_CUDA(cudaMemcpyAsync(h, info.p, 8, cudaMemcpyDeviceToHost, c.stream));
    RNLA_CUDA(cudaStreamSynchronize(c.stream));
    if (h[1]) return fail(RNLA_ERR_COMPUTATION, "symmetric eigen-decomposition did not converge");
    return RNLA_OK;
}

rnla_status rnla_generate_lowrank_dev(double* dA, int64_t lda, int64_t m_local, int64_t n, int64_t row_offset, int64_t m_global,
                                      int64_t r0, const double* sigma_host, double eta, uint64_t seed) {
    RNLA_API_GUARD;
    RNLA_TRY(ensure_ctx());
    return dev_generate_lowrank(dA, lda, m_local, n, row_offset, m_global, r0, sigma_host, eta, seed);
}

}  // extern "C"

// ================================================================================================================
// SURVEY.md section 8(f) rows 2-4 and the saddle-point driver: pivoted QR, CQRRPT, sketch-and-solve, ID / CUR
// ================================================================================================================
static rnla_status d2h_idx(int64_t* host, const int64_t* dev, size_t count) {
    if (count) RNLA_CUDA(cudaMemcpyAsync(host, dev, count * 8, cudaMemcpyDeviceToHost, ctx().stream));
    RNLA_CUDA(cudaStreamSynchronize(ctx().stream));
    return RNLA_OK;
}

rnla_status rnla_qrcp_dev(double* dR, int64_t ldr, int64_t m, int64_t n, int64_t steps, int64_t* dperm, double* dQ, int64_t ldq,
                          int64_t qcols) {
    RNLA_API_GUARD;
    RNLA_TRY(ensure_ctx());
    if (m <= 0 || n <= 0) return fail(RNLA_ERR_INVALID_DIMENSIONS, "Rows and columns must be greater than 0");
    if (steps <= 0) return fail(RNLA_ERR_INVALID_PARAMETERS, "k must be positive");                          // pivot_decompositions.rs:201
    if (steps > std::min(m, n)) return fail(RNLA_ERR_INVALID_PARAMETERS, "k must be <= min(m,n)");            // :200
    if (qcols < 0 || qcols > m) return fail(RNLA_ERR_INVALID_DIMENSIONS, "qrcp: 0 <= qcols <= m");
    phases_reset();
    RNLA_TRY(dev_qrcp(dR, ldr, m, n, steps, dperm, qcols ? dQ : nullptr, ldq, qcols));
    RNLA_CUDA(cudaStreamSynchronize(ctx().stream));
    return RNLA_OK;
}
rnla_status rnla_qrcp(const double* A, int64_t m, int64_t n, int64_t steps, int64_t qcols, double* Q, double* R, int64_t* perm) {
    RNLA_API_GUARD;
    RNLA_TRY(ensure_ctx());
    if (m <= 0 || n <= 0) return fail(RNLA_ERR_INVALID_DIMENSIONS, "Rows and columns must be greater than 0");
    DevBuf dR, dQ, dp;
    RNLA_TRY(h2d(dR, A, (size_t)m * n));
    RNLA_CUDA(dp.alloc((size_t)n * 8));
    const bool wantq = Q != nullptr && qcols > 0;
    if (wantq) RNLA_CUDA(dQ.alloc((size_t)m * qcols * 8));
    RNLA_TRY(rnla_qrcp_dev(dR.d(), m, m, n, steps, dp.as<int64_t>(), wantq ? dQ.d() : nullptr, m, wantq ? qcols : 0));
    if (wantq) RNLA_TRY(d2h(Q, dQ.d(), (size_t)m * qcols));
    RNLA_TRY(d2h(R, dR.d(), (size_t)m * n));
    return d2h_idx(perm, dp.as<int64_t>(), (size_t)n);
}

// ---- lupp (reference src/pivot_decompositions.rs:21-86)
rnla_status rnla_lupp_dev(double* dW, int64_t ldw, int64_t n, double* dL, int64_t ldl, double* dU, int64_t ldu, int64_t* dperm) {
    RNLA_API_GUARD;
    if (!dW || !dL || !dU || !dperm) return fail(RNLA_ERR_INVALID_PARAMETERS, "lupp: null argument");
    if (n < 1) return fail(RNLA_ERR_INVALID_DIMENSIONS, "lupp: empty matrix");          // the reference underflows `n - 1` (:32)
    RNLA_TRY(ensure_ctx());
    phases_reset();
    int64_t sing = -1;
    RNLA_TRY(dev_lupp(dW, ldw, n, dL, ldl, dU, ldu, dperm, &sing));
    if (sing >= 0) return fail(RNLA_ERR_SINGULAR_MATRIX, "Matrix must be nonsingular for an LU decomposition");   // :44-48
    return RNLA_OK;
}
rnla_status rnla_lupp(const double* A, int64_t rows, int64_t cols, double* L, double* U, int64_t* perm) {
    RNLA_API_GUARD;
    if (rows != cols) {                                                                            // :23-27
        char buf[160];
        snprintf(buf, sizeof buf, "Matrix must be square, found matrix with %lld rows and %lld columns", (long long)rows, (long long)cols);
        return fail(RNLA_ERR_NOT_SQUARE, buf);
    }
    if (!A || !L || !U || !perm) return fail(RNLA_ERR_INVALID_PARAMETERS, "lupp: null argument");
    if (rows < 1) return fail(RNLA_ERR_INVALID_DIMENSIONS, "lupp: empty matrix");
    RNLA_TRY(ensure_ctx());
    const int64_t n = rows;
    DevBuf dW, dL, dU, dp;
    RNLA_TRY(h2d(dW, A, (size_t)n * n));
    RNLA_CUDA(dL.alloc((size_t)n * n * 8)); RNLA_CUDA(dU.alloc((size_t)n * n * 8)); RNLA_CUDA(dp.alloc((size_t)n * 8));
    RNLA_TRY(rnla_lupp_dev(dW.d(), n, n, dL.d(), n, dU.d(), n, dp.as<int64_t>()));
    RNLA_TRY(d2h(L, dL.d(), (size_t)n * n));
    RNLA_TRY(d2h(U, dU.d(), (size_t)n * n));
    return d2h_idx(perm, dp.as<int64_t>(), (size_t)n);
}

rnla_status rnla_sap_chol_qrcp_dev(const double* dA, int64_t lda, int64_t m, int64_t n, int64_t d, int32_t kind, int32_t dist,
                                   int32_t zeta, double* dQ, int64_t ldq, double* dR, int64_t ldr, int64_t* dJ, int64_t* k) {
    RNLA_API_GUARD;
    RNLA_TRY(ensure_ctx());
    int64_t kk = 0;
    RNLA_TRY(dev_sap_chol_qrcp(dA, lda, m, n, d, kind, dist, zeta, ctx().opts.seed, dQ, ldq, dR, ldr, dJ, &kk));
    if (k) *k = kk;
    return RNLA_OK;
}
rnla_status rnla_sap_chol_qrcp(const double* A, int64_t m, int64_t n, int64_t d, int32_t kind, int32_t dist, int32_t zeta,
                               double* Q, double* R, int64_t* J, int64_t* k) {
    RNLA_API_GUARD;
    if (!(n <= d && d <= m) || n <= 0) return fail(RNLA_ERR_INVALID_PARAMETERS, "d must satisfy n \xe2\x89\xa4 d \xe2\x89\xaa m");   // cqrrpt.rs:29
    RNLA_TRY(ensure_ctx());
    DevBuf dA, dQ, dR, dJ;
    RNLA_TRY(h2d(dA, A, (size_t)m * n));
    RNLA_CUDA(dQ.alloc((size_t)m * n * 8)); RNLA_CUDA(dR.alloc((size_t)n * n * 8)); RNLA_CUDA(dJ.alloc((size_t)n * 8));
    int64_t kk = 0;
    RNLA_TRY(dev_sap_chol_qrcp(dA.d(), m, m, n, d, kind, dist, zeta, ctx().opts.seed, dQ.d(), m, dR.d(), n, dJ.as<int64_t>(), &kk));
    if (k) *k = kk;
    RNLA_TRY(d2h(Q, dQ.d(), (size_t)m * kk));
    if (kk) RNLA_CUDA(cudaMemcpy2DAsync(R, (size_t)kk * 8, dR.p, (size_t)n * 8, (size_t)kk * 8, (size_t)n, cudaMemcpyDeviceToHost, ctx().stream));
    return d2h_idx(J, dJ.as<int64_t>(), (size_t)n);
}

rnla_status rnla_sketched_least_squares_dev(int32_t which, const double* dA, int64_t lda, int64_t m, int64_t n, const double* db,
                                            int32_t kind, int32_t dist, int32_t zeta, double* dx) {
    RNLA_API_GUARD;
    RNLA_TRY(ensure_ctx());
    if (which != 0 && which != 1) return fail(RNLA_ERR_INVALID_PARAMETERS, "sketched_least_squares: which = 0 (QR) or 1 (SVD)");
    return dev_sketched_least_squares(which, dA, lda, m, n, db, kind, dist, zeta, ctx().opts.seed, dx);
}
static rnla_status sketched_ls_host(int which, const double* A, int64_t m, int64_t n, const double* b, int32_t kind, int32_t dist,
                                    int32_t zeta, double* x) {
    RNLA_TRY(ensure_ctx());
    if (m <= 0 || n <= 0) return fail(RNLA_ERR_INVALID_DIMENSIONS, "Rows and columns must be greater than 0");
    DevBuf dA, db, dx;
    RNLA_TRY(h2d(dA, A, (size_t)m * n));
    RNLA_TRY(h2d(db, b, (size_t)m));
    RNLA_CUDA(dx.alloc((size_t)n * 8));
    RNLA_TRY(dev_sketched_least_squares(which, dA.d(), m, m, n, db.d(), kind, dist, zeta, ctx().opts.seed, dx.d()));
    return d2h(x, dx.d(), (size_t)n);
}
rnla_status rnla_sketched_least_squares_qr(const double* A, int64_t m, int64_t n, const double* b, int32_t kind, int32_t dist,
                                           int32_t zeta, double* x) {
    RNLA_API_GUARD;
    return sketched_ls_host(0, A, m, n, b, kind, dist, zeta, x);
}
rnla_status rnla_sketched_least_squares_svd(const double* A, int64_t m, int64_t n, const double* b, int32_t kind, int32_t dist,
                                            int32_t zeta, double* x) {
    RNLA_API_GUARD;
    return sketched_ls_host(1, A, m, n, b, kind, dist, zeta, x);
}

rnla_status rnla_osid_qrcp(const double* Y, int64_t l, int64_t w, int64_t k, int32_t attr, double* X, int64_t* J) {
    RNLA_API_GUARD;
    if (k <= 0) return fail(RNLA_ERR_INVALID_PARAMETERS, "k must be positive)");                              // id.rs:278
    if (k > std::min(l, w)) return fail(RNLA_ERR_INVALID_PARAMETERS, "k must be <= min(l,w)");                // id.rs:279
    RNLA_TRY(ensure_ctx());
    phases_reset();
    DevBuf dY, dX, dJ;
    RNLA_TRY(h2d(dY, Y, (size_t)l * w));
    const bool col = attr == RNLA_COLUMN;
    const int64_t xr = col ? k : l, xc = col ? w : k;
    RNLA_CUDA(dX.alloc((size_t)xr * xc * 8)); RNLA_CUDA(dJ.alloc((size_t)k * 8));
    RNLA_TRY(dev_osid_qrcp(dY.d(), l, l, w, k, attr, dX.d(), xr, dJ.as<int64_t>()));
    RNLA_TRY(d2h(X, dX.d(), (size_t)xr * xc));
    return d2h_idx(J, dJ.as<int64_t>(), (size_t)k);
}
rnla_status rnla_osid_randomised_dev(const double* dA, int64_t lda, int64_t m, int64_t n, int64_t k, int32_t attr,
                                     const rnla_options* opt, double* dX, int64_t ldx, int64_t* dJ) {
    RNLA_API_GUARD;
    RNLA_TRY(ensure_ctx());
    phases_reset();
    RNLA_TRY(dev_osid_randomised(dA, lda, m, n, k, attr, opt ? *opt : ctx().opts, dX, ldx, dJ));
    RNLA_CUDA(cudaStreamSynchronize(ctx().stream));
    return RNLA_OK;
}
rnla_status rnla_osid_randomised(const double* A, int64_t m, int64_t n, int64_t k, int32_t attr, double* X, int64_t* J) {
    RNLA_API_GUARD;
    if (k <= 0) return fail(RNLA_ERR_INVALID_PARAMETERS, "k must be positive)");                              // id.rs:223
    if (k > std::min(m, n)) return fail(RNLA_ERR_INVALID_PARAMETERS, "k must be <= min(l,w)");                // id.rs:224
    RNLA_TRY(ensure_ctx());
    DevBuf dA, dX, dJ;
    RNLA_TRY(h2d(dA, A, (size_t)m * n));
    const bool col = attr == RNLA_COLUMN;
    const int64_t xr = col ? k : m, xc = col ? n : k;
    RNLA_CUDA(dX.alloc((size_t)xr * xc * 8)); RNLA_CUDA(dJ.alloc((size_t)k * 8));
    RNLA_TRY(rnla_osid_randomised_dev(dA.d(), m, m, n, k, attr, nullptr, dX.d(), xr, dJ.as<int64_t>()));
    RNLA_TRY(d2h(X, dX.d(), (size_t)xr * xc));
    return d2h_idx(J, dJ.as<int64_t>(), (size_t)k);
}
rnla_status rnla_two_sided_id(const double* A, int64_t m, int64_t n, int64_t k, int32_t randomised, double* Z, int64_t* I, int64_t* J,
                              double* X) {
    RNLA_API_GUARD;
    if (k <= 0) return fail(RNLA_ERR_INVALID_PARAMETERS, "k must be positive)");
    if (k > std::min(m, n)) return fail(RNLA_ERR_INVALID_PARAMETERS, "k must be <= min(l,w)");
    RNLA_TRY(ensure_ctx());
    DevBuf dA, dZ, dX, dI, dJ;
    RNLA_TRY(h2d(dA, A, (size_t)m * n));
    RNLA_CUDA(dZ.alloc((size_t)m * k * 8)); RNLA_CUDA(dX.alloc((size_t)k * n * 8));
    RNLA_CUDA(dI.alloc((size_t)k * 8)); RNLA_CUDA(dJ.alloc((size_t)k * 8));
    RNLA_TRY(dev_two_sided_id(randomised, dA.d(), m, m, n, k, ctx().opts, dZ.d(), m, dI.as<int64_t>(), dJ.as<int64_t>(), dX.d(), k));
    RNLA_TRY(d2h(Z, dZ.d(), (size_t)m * k));
    RNLA_TRY(d2h(X, dX.d(), (size_t)k * n));
    RNLA_TRY(d2h_idx(I, dI.as<int64_t>(), (size_t)k));
    return d2h_idx(J, dJ.as<int64_t>(), (size_t)k);
}
rnla_status rnla_cur_dev(const double* dA, int64_t lda, int64_t m, int64_t n, int64_t k, int32_t randomised, const rnla_options* opt,
                         int64_t* dJ, double* dU, int64_t ldu, int64_t* dI) {
    RNLA_API_GUARD;
    RNLA_TRY(ensure_ctx());
    return dev_cur(randomised, dA, lda, m, n, k, opt ? *opt : ctx().opts, dJ, dU, ldu, dI);
}
rnla_status rnla_cur(const double* A, int64_t m, int64_t n, int64_t k, int32_t randomised, int64_t* J, double* U, int64_t* I) {
    RNLA_API_GUARD;
    if (k <= 0) return fail(RNLA_ERR_INVALID_PARAMETERS, "k must be positive)");
    if (k > std::min(m, n)) return fail(RNLA_ERR_INVALID_PARAMETERS, "k must be <= min(l,w)");
    RNLA_TRY(ensure_ctx());
    DevBuf dA, dU, dI, dJ;
    RNLA_TRY(h2d(dA, A, (size_t)m * n));
    RNLA_CUDA(dU.alloc((size_t)k * k * 8)); RNLA_CUDA(dI.alloc((size_t)k * 8)); RNLA_CUDA(dJ.alloc((size_t)k * 8));
    RNLA_TRY(dev_cur(randomised, dA.d(), m, m, n, k, ctx().opts, dJ.as<int64_t>(), dU.d(), k, dI.as<int64_t>()));
    RNLA_TRY(d2h(U, dU.d(), (size_t)k * k));
    RNLA_TRY(d2h_idx(I, dI.as<int64_t>(), (size_t)k));
    return d2h_idx(J, dJ.as<int64_t>(), (size_t)k);
}

rnla_status rnla_sketch_saddle_point_precondition_dev(const double* dA, int64_t lda, int64_t m, int64_t n, const double* db,
                                                      const double* dc, double mu, double epsilon, int64_t l, double sampling_factor,
                                                      double* dx, double* dy, int64_t* iterations, int32_t* converged) {
    RNLA_API_GUARD;
    RNLA_TRY(ensure_ctx());
    RNLA_TRY(validate_lsq(m, n, epsilon, l, sampling_factor));                                             // sketch_and_precondition.rs:152-171
    return dev_saddle_point(dA, lda, m, n, db, dc, mu, epsilon, l, sampling_factor, RNLA_GAUSSIAN, ctx().opts.seed, dx, dy, iterations, converged);
}
rnla_status rnla_sketch_saddle_point_precondition(const double* A, int64_t m, int64_t n, const double* b, const double* c, double mu,
                                                  double epsilon, int64_t l, double sampling_factor, double* x, double* y,
                                                  int64_t* iterations, int32_t* converged) {
    RNLA_API_GUARD;
    RNLA_TRY(validate_lsq(m, n, epsilon, l, sampling_factor));
    RNLA_TRY(ensure_ctx());
    DevBuf dA, db, dc, dx, dy;
    RNLA_TRY(h2d(dA, A, (size_t)m * n));
    RNLA_TRY(h2d(db, b, (size_t)m));
    if (c) RNLA_TRY(h2d(dc, c, (size_t)n));
    RNLA_CUDA(dx.alloc((size_t)n * 8)); RNLA_CUDA(dy.alloc((size_t)m * 8));
    RNLA_TRY(dev_saddle_point(dA.d(), m, m, n, db.d(), c ? dc.d() : nullptr, mu, epsilon, l, sampling_factor, RNLA_GAUSSIAN,
                              ctx().opts.seed, dx.d(), dy.d(), iterations, converged));
    RNLA_TRY(d2h(x, dx.d(), (size_t)n));
    return d2h(y, dy.d(), (size_t)m);
}

// ---- INT8 tensor-core products (i8gemm.cu), exposed for tests and benches -------------------------------------------------------
// trans = 0: C (m x N) = A (m x n) * B (n x N);  trans != 0: C (n x N) = A^T * B (m x N).  `reps` products with one split of A into
// `planes` digit planes (4: 31-bit operands, all_pairs != 0 adds the digit pairs of groups 4..6; 6: 47-bit; 7: 55-bit, FP64-grade).
rnla_status rnla_i8_gemm_dev(int32_t trans, int32_t planes, int32_t all_pairs, const double* dA, int64_t lda, int64_t m, int64_t n,
                             const double* dB, int64_t ldb, int64_t N, double* dC, int64_t ldc, int32_t reps) {
    RNLA_API_GUARD;
    RNLA_TRY(ensure_ctx());
    if (N < 1 || N > 256 || m < 1 || n < 1) return fail(RNLA_ERR_INVALID_DIMENSIONS, "i8 gemm: 1 <= N <= 256");
    if (planes != 4 && planes != 6 && planes != 7) return fail(RNLA_ERR_INVALID_PARAMETERS, "i8 gemm: planes must be 4, 6 or 7");
    phases_reset();
    bool usable = false;
    RNLA_TRY(i8_prepare(dA, lda, m, n, planes, &usable));
    if (!usable) { i8_release(); return fail(RNLA_ERR_COMPUTATION, "i8 gemm: A holds non-finite values or rows too small to scale"); }
    rnla_status st = RNLA_OK;
    for (int r = 0; r < std::max(reps, 1) && st == RNLA_OK; ++r) {
        PhaseScope ph(trans ? "i8:At*B" : "i8:A*B");
        i8_set_precision(planes, all_pairs != 0);
        st = trans ? i8_gemm_tn(dB, ldb, N, dC, ldc) : i8_gemm_nn(dB, ldb, N, dC, ldc);
    }
    i8_deactivate();
    cudaError_t e = cudaStreamSynchronize(ctx().stream);
    i8_release();
    RNLA_TRY(st);
    RNLA_CUDA(e);
    return RNLA_OK;
}
// tests: drain the int32 accumulators every `stages` stages of 64 contraction indices (0 restores the exactness bound)
rnla_status rnla_debug_i8_flush(int32_t stages) {
    RNLA_API_GUARD;
    i8_debug_flush(stages);
    return RNLA_OK;
}
// round-1 entry point, kept: reps > 0: the ten leading digit pairs of the 31-bit split; reps < 0: A B with all 16 pairs, A^T B on
// the 55-bit split
rnla_status rnla_i8_range_gemm_dev(int32_t trans, const double* dA, int64_t lda, int64_t m, int64_t n, const double* dB, int64_t ldb,
                                   int64_t N, double* dC, int64_t ldc, int32_t reps) {
    if (reps >= 0) return rnla_i8_gemm_dev(trans, 4, 0, dA, lda, m, n, dB, ldb, N, dC, ldc, reps);
    return rnla_i8_gemm_dev(trans, trans ? 7 : 4, 1, dA, lda, m, n, dB, ldb, N, dC, ldc, -reps);
}

// the digit-plane workspace of the int8 passes persists across calls (tens of GB at the headline size); give it back
rnla_status rnla_release_workspace(void) {
    RNLA_API_GUARD;
    if (ensure_ctx() != RNLA_OK) return RNLA_OK;
    cudaStreamSynchronize(ctx().stream);
    i8_free_workspace();
    return RNLA_OK;
}
