// Process-global context: device, stream, options, NCCL communicator (dlopen'ed), timings, error text.
#pragma once
#include <cstdint>
#include <functional>
#include <mutex>
#include <string>
#include <vector>
#include <cuda_runtime.h>
#include "../../include/rnla.h"

namespace rnla {

struct PhaseTiming { std::string name; cudaEvent_t e0, e1; };

struct Ctx {
    bool ready = false;
    int device = 0;
    int sms = 148;
    cudaStream_t stream = nullptr;
    cudaStream_t own_stream = nullptr;
    rnla_options opts;
    // communicator (one process per GPU)
    void* nccl_lib = nullptr;
    void* comm = nullptr;
    int nranks = 1, rank = 0;
    // timings of the last driver call
    std::vector<PhaseTiming> phases;
    std::vector<int> open_phases;            // indices of the phases begun and not yet ended (phases nest)
    bool kernel_timing = false;              // also record kernel-level entries ("k:..."), rnla_set_kernel_timing
    std::vector<cudaEvent_t> event_pool;
    std::vector<std::string> timing_names;   // storage handed out by rnla_get_timings
    // host-buffer entry points: copy stream for the upload of A, and a one-shot hook that replaces the first product
    // Y = A * Omega of the power iteration by "upload a row block, multiply it" (the pass hides behind the PCIe copy)
    cudaStream_t copy_stream = nullptr;
    std::function<rnla_status(double* S, double* Y, int64_t ldy)> first_pass_hook;
    // called by that hook after each row block [r0, r0 + rows) of A has landed and been multiplied (the int8 split of the block)
    std::function<rnla_status(int64_t r0, int64_t rows)> block_landed_hook;
};

Ctx& ctx();
// Every extern "C" entry point holds this lock for its whole duration: the context (stream, phase timings, workspace pool,
// one-shot hooks) is process-global, so calls from several host threads are serialised.  Recursive: some entry points are
// implemented on top of others.
std::recursive_mutex& api_mutex();
#define RNLA_API_GUARD std::lock_guard<std::recursive_mutex> rnla_api_guard_(::rnla::api_mutex())
void set_error(const std::string& msg);
rnla_status fail(rnla_status code, const std::string& msg);
rnla_status cuda_fail(cudaError_t e, const char* what, const char* file, int line);
rnla_status ensure_ctx();

#define RNLA_CUDA(expr)                                                             \
    do {                                                                            \
        cudaError_t _e = (expr);                                                    \
        if (_e != cudaSuccess) return ::rnla::cuda_fail(_e, #expr, __FILE__, __LINE__); \
    } while (0)
#define RNLA_TRY(expr)                          \
    do {                                        \
        rnla_status _s = (expr);                \
        if (_s != RNLA_OK) return _s;           \
    } while (0)

// stream-ordered device buffer
struct DevBuf {
    void* p = nullptr;
    size_t bytes = 0;
    DevBuf() = default;
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    ~DevBuf() { release(); }
    cudaError_t alloc(size_t n) {
        release();
        if (n == 0) n = 8;
        bytes = n;
        return cudaMallocAsync(&p, n, ctx().stream);
    }
    void release() {
        if (p) { cudaFreeAsync(p, ctx().stream); p = nullptr; }
    }
    double* d() const { return static_cast<double*>(p); }
    template <class T> T* as() const { return static_cast<T*>(p); }
};

// collectives over the communicator; no-ops when nranks == 1
rnla_status allreduce_sum_f64(double* buf, size_t count);
rnla_status allgather_i64(const int64_t* send_dev, int64_t* recv_dev, size_t count_per_rank);

void phase_begin(const char* name);
void host_trace_mark(const char* name);     // RNLA_TRACE_HOST=1 diagnostics
void phase_end();
void kernel_phase_begin(const char* name);   // no-ops unless kernel timing is on
void kernel_phase_end();
void phases_reset();

struct PhaseScope {
    explicit PhaseScope(const char* n) { phase_begin(n); }
    ~PhaseScope() { phase_end(); }
};

}  // namespace rnla
