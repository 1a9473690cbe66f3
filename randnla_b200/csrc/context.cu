#include "context.cuh"
#include "drivers.cuh"
#include "gemm.cuh"
#include <dlfcn.h>
#include <nccl.h>      // types only: the library is dlopen'ed so librnla.so has no link-time NCCL dependency
#include <cstdio>
#include <chrono>
#include <cstring>
#include <mutex>

namespace rnla {

static thread_local std::string t_err;
static Ctx g_ctx;
static std::mutex g_mu;
static std::recursive_mutex g_api_mu;
std::recursive_mutex& api_mutex() { return g_api_mu; }

Ctx& ctx() { return g_ctx; }
void set_error(const std::string& msg) { t_err = msg; }
rnla_status fail(rnla_status code, const std::string& msg) { t_err = msg; return code; }
rnla_status cuda_fail(cudaError_t e, const char* what, const char* file, int line) {
    char buf[512];
    snprintf(buf, sizeof buf, "CUDA error '%s' in %s (%s:%d)", cudaGetErrorString(e), what, file, line);
    t_err = buf;
    cudaGetLastError();
    return RNLA_ERR_COMPUTATION;
}

static void default_opts(rnla_options* o) {
    memset(o, 0, sizeof *o);
    o->mode = RNLA_MODE_INTENDED;
    o->dist = RNLA_GAUSSIAN;
    o->seed = 0;
    o->num_passes = 0;
    o->passes_per_stab = 0;
    o->fused_sketch = 2;
    o->range_passes_int8 = -1;          // auto: FP64-grade passes on the integer tensor cores where the shape is supported
    const char* r8 = getenv("RNLA_RANGE_INT8");
    if (r8 && r8[0] >= '0' && r8[0] <= '3' && !r8[1]) o->range_passes_int8 = r8[0] - '0';
    if (r8 && !strcmp(r8, "auto")) o->range_passes_int8 = -1;
    o->generator = RNLA_GEN_PHILOX;
    const char* gen = getenv("RNLA_GENERATOR");
    if (gen && (!strcmp(gen, "threefry") || !strcmp(gen, "THREEFRY") || !strcmp(gen, "1"))) o->generator = RNLA_GEN_THREEFRY;
    const char* m = getenv("RNLA_MODE");
    if (m && (!strcmp(m, "literal") || !strcmp(m, "LITERAL") || !strcmp(m, "1"))) o->mode = RNLA_MODE_LITERAL;
}

static rnla_status init_locked(int device) {
    Ctx& c = g_ctx;
    if (c.ready) {
        if (device >= 0 && device != c.device)
            return fail(RNLA_ERR_COMPUTATION, "rnla_init: context already bound to another device");
        return RNLA_OK;
    }
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return fail(RNLA_ERR_COMPUTATION,
                    "no usable CUDA device: randnla_b200 has no CPU fallback (cudaGetDeviceCount failed or returned 0)");
    }
    if (device < 0) device = 0;
    if (device >= ndev) return fail(RNLA_ERR_COMPUTATION, "rnla_init: device index out of range");
    RNLA_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    RNLA_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10)
        return fail(RNLA_ERR_COMPUTATION, std::string("device '") + prop.name + "' is not sm_100-class; this library is built for sm_100a only");
    c.device = device;
    c.sms = prop.multiProcessorCount;
    RNLA_CUDA(cudaStreamCreateWithFlags(&c.own_stream, cudaStreamNonBlocking));
    RNLA_CUDA(cudaStreamCreateWithFlags(&c.copy_stream, cudaStreamNonBlocking));
    c.stream = c.own_stream;
    cudaMemPool_t pool;
    RNLA_CUDA(cudaDeviceGetDefaultMemPool(&pool, device));
    uint64_t thresh = UINT64_MAX;
    RNLA_CUDA(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thresh));
    default_opts(&c.opts);
    c.ready = true;
    return RNLA_OK;
}

rnla_status ensure_ctx() {
    std::lock_guard<std::mutex> lk(g_mu);
    if (g_ctx.ready) { cudaSetDevice(g_ctx.device); return RNLA_OK; }
    int dev = 0;
    if (const char* lr = getenv("LOCAL_RANK")) dev = atoi(lr);
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) == cudaSuccess && ndev > 0) dev %= ndev; else cudaGetLastError();
    return init_locked(dev);
}

// ---------------------------------------------------------------- NCCL (dlopen)
struct NcclApi {
    ncclResult_t (*GetUniqueId)(ncclUniqueId*);
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int);
    ncclResult_t (*CommDestroy)(ncclComm_t);
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t);
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t);
    const char* (*GetErrorString)(ncclResult_t);
};
static NcclApi g_nccl;

static rnla_status load_nccl() {
    Ctx& c = g_ctx;
    if (c.nccl_lib) return RNLA_OK;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    void* h = nullptr;
    for (const char* n : names) { h = dlopen(n, RTLD_NOW | RTLD_GLOBAL); if (h) break; }
    if (!h) return fail(RNLA_ERR_COMPUTATION, std::string("cannot dlopen libnccl.so.2: ") + dlerror());
#define LOAD(sym)                                                                           \
    *(void**)(&g_nccl.sym) = dlsym(h, "nccl" #sym);                                          \
    if (!g_nccl.sym) return fail(RNLA_ERR_COMPUTATION, "libnccl is missing symbol nccl" #sym);
    LOAD(GetUniqueId) LOAD(CommInitRank) LOAD(CommDestroy) LOAD(AllReduce) LOAD(AllGather) LOAD(GetErrorString)
#undef LOAD
    c.nccl_lib = h;
    return RNLA_OK;
}
static rnla_status nccl_fail(ncclResult_t r, const char* what) {
    return fail(RNLA_ERR_COMPUTATION, std::string("NCCL error in ") + what + ": " + g_nccl.GetErrorString(r));
}

rnla_status allreduce_sum_f64(double* buf, size_t count) {
    Ctx& c = g_ctx;
    if (c.nranks <= 1 || count == 0) return RNLA_OK;
    ncclResult_t r = g_nccl.AllReduce(buf, buf, count, ncclFloat64, ncclSum, (ncclComm_t)c.comm, c.stream);
    if (r != ncclSuccess) return nccl_fail(r, "ncclAllReduce");
    ++g_kernel_launches;
    return RNLA_OK;
}
rnla_status allgather_i64(const int64_t* send_dev, int64_t* recv_dev, size_t count_per_rank) {
    Ctx& c = g_ctx;
    if (c.nranks <= 1) {
        RNLA_CUDA(cudaMemcpyAsync(recv_dev, send_dev, count_per_rank * 8, cudaMemcpyDeviceToDevice, c.stream));
        return RNLA_OK;
    }
    ncclResult_t r = g_nccl.AllGather(send_dev, recv_dev, count_per_rank, ncclInt64, (ncclComm_t)c.comm, c.stream);
    if (r != ncclSuccess) return nccl_fail(r, "ncclAllGather");
    ++g_kernel_launches;
    return RNLA_OK;
}

// ---------------------------------------------------------------- phase timings
static cudaEvent_t get_event() {
    Ctx& c = g_ctx;
    if (!c.event_pool.empty()) { cudaEvent_t e = c.event_pool.back(); c.event_pool.pop_back(); return e; }
    cudaEvent_t e; cudaEventCreate(&e); return e;
}
void phases_reset() {
    Ctx& c = g_ctx;
    for (auto& p : c.phases) { c.event_pool.push_back(p.e0); c.event_pool.push_back(p.e1); }
    c.phases.clear();
    c.open_phases.clear();
}
// Phases nest: a kernel-level entry ("k:..." names, recorded only while kernel timing is on, rnla_set_kernel_timing) may sit inside
// a driver-level phase; phase_end closes the innermost open one.
// RNLA_TRACE_HOST=1: report on stderr every stretch of more than 20 ms of HOST time between two consecutive phase marks (where does a
// call spend time that no phase accounts for?)
static void host_mark(const char* what, const char* name) {
    static const bool on = [] { const char* e = getenv("RNLA_TRACE_HOST"); return e && e[0] == '1'; }();
    if (!on) return;
    static std::chrono::steady_clock::time_point last = std::chrono::steady_clock::now();
    static std::string last_what = "start";
    const auto now = std::chrono::steady_clock::now();
    const double ms = std::chrono::duration<double, std::milli>(now - last).count();
    if (ms > 20.0) fprintf(stderr, "[rnla host gap] %.1f ms between <%s> and <%s %s>\n", ms, last_what.c_str(), what, name);
    last = now; last_what = std::string(what) + " " + name;
}
void host_trace_mark(const char* name) { host_mark("mark", name); }
void phase_begin(const char* name) {
    host_mark("begin", name);
    Ctx& c = g_ctx;
    PhaseTiming p; p.name = name; p.e0 = get_event(); p.e1 = get_event();
    cudaEventRecord(p.e0, c.stream);
    c.phases.push_back(p);
    c.open_phases.push_back((int)c.phases.size() - 1);
}
void phase_end() {
    Ctx& c = g_ctx;
    if (c.open_phases.empty()) return;
    const int i = c.open_phases.back();
    c.open_phases.pop_back();
    if (i < (int)c.phases.size()) { host_mark("end", c.phases[(size_t)i].name.c_str()); cudaEventRecord(c.phases[(size_t)i].e1, c.stream); }
}
void kernel_phase_begin(const char* name) { if (g_ctx.kernel_timing) phase_begin(name); }
void kernel_phase_end() { if (g_ctx.kernel_timing) phase_end(); }

}  // namespace rnla

using namespace rnla;

extern "C" {

int32_t rnla_version(void) { return 100; }
const char* rnla_last_error_message(void) { return t_err.c_str(); }

rnla_status rnla_init(int32_t device) {
    RNLA_API_GUARD;
    std::lock_guard<std::mutex> lk(g_mu);
    return init_locked(device);
}
void rnla_shutdown(void) {
    RNLA_API_GUARD;
    std::lock_guard<std::mutex> lk(g_mu);
    Ctx& c = g_ctx;
    if (!c.ready) return;
    cudaStreamSynchronize(c.stream);
    i8_free_workspace();
    if (c.comm) { g_nccl.CommDestroy((ncclComm_t)c.comm); c.comm = nullptr; c.nranks = 1; c.rank = 0; }
    phases_reset();
    for (auto e : c.event_pool) cudaEventDestroy(e);
    c.event_pool.clear();
    cudaStreamDestroy(c.own_stream);
    if (c.copy_stream) { cudaStreamDestroy(c.copy_stream); c.copy_stream = nullptr; }
    c.first_pass_hook = nullptr;
    c.stream = c.own_stream = nullptr;
    c.ready = false;
}
void* rnla_stream(void) { return ensure_ctx() == RNLA_OK ? (void*)g_ctx.stream : nullptr; }
rnla_status rnla_set_stream(void* cuda_stream) {
    RNLA_API_GUARD;
    RNLA_TRY(ensure_ctx());
    RNLA_CUDA(cudaStreamSynchronize(g_ctx.stream));
    g_ctx.stream = cuda_stream ? (cudaStream_t)cuda_stream : g_ctx.own_stream;
    return RNLA_OK;
}
rnla_status rnla_synchronize(void) {
    RNLA_API_GUARD;
    RNLA_TRY(ensure_ctx());
    RNLA_CUDA(cudaStreamSynchronize(g_ctx.stream));
    return RNLA_OK;
}
void rnla_default_options(rnla_options* opt) { if (opt) default_opts(opt); }
rnla_status rnla_set_options(const rnla_options* opt) {
    RNLA_API_GUARD;
    if (!opt) return fail(RNLA_ERR_INVALID_PARAMETERS, "rnla_set_options: null options");
    if (opt->mode != RNLA_MODE_INTENDED && opt->mode != RNLA_MODE_LITERAL)
        return fail(RNLA_ERR_INVALID_PARAMETERS, "rnla_set_options: unknown mode");
    if (opt->dist < RNLA_GAUSSIAN || opt->dist > RNLA_RADEMACHER)
        return fail(RNLA_ERR_INVALID_PARAMETERS, "rnla_set_options: unknown distribution");
    if (opt->generator != RNLA_GEN_PHILOX && opt->generator != RNLA_GEN_THREEFRY)
        return fail(RNLA_ERR_INVALID_PARAMETERS, "rnla_set_options: unknown generator");
    RNLA_TRY(ensure_ctx());
    g_ctx.opts = *opt;
    return RNLA_OK;
}
void rnla_get_options(rnla_options* opt) {
    RNLA_API_GUARD;
    if (!opt) return;
    if (g_ctx.ready) *opt = g_ctx.opts; else default_opts(opt);
}
uint64_t rnla_kernel_launches(void) { return g_kernel_launches; }

rnla_status rnla_set_kernel_timing(int32_t on) {
    RNLA_API_GUARD;
    g_ctx.kernel_timing = on != 0;
    return RNLA_OK;
}
int32_t rnla_get_timings(const char** names, double* ms, int32_t cap) {
    RNLA_API_GUARD;
    Ctx& c = g_ctx;
    if (!c.ready) return 0;
    cudaStreamSynchronize(c.stream);
    c.timing_names.clear();
    for (auto& p : c.phases) c.timing_names.push_back(p.name);
    int32_t n = 0;
    for (size_t i = 0; i < c.phases.size() && n < cap; ++i, ++n) {
        float t = 0.f;
        cudaEventElapsedTime(&t, c.phases[i].e0, c.phases[i].e1);
        if (names) names[n] = c.timing_names[i].c_str();
        if (ms) ms[n] = t;
    }
    return (int32_t)c.phases.size();
}

rnla_status rnla_comm_unique_id(uint8_t id[128]) {
    RNLA_API_GUARD;
    RNLA_TRY(ensure_ctx());
    RNLA_TRY(load_nccl());
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    ncclUniqueId uid;
    ncclResult_t r = g_nccl.GetUniqueId(&uid);
    if (r != ncclSuccess) return nccl_fail(r, "ncclGetUniqueId");
    memcpy(id, &uid, 128);
    return RNLA_OK;
}
rnla_status rnla_comm_init(int32_t nranks, int32_t rank, const uint8_t id[128]) {
    RNLA_API_GUARD;
    RNLA_TRY(ensure_ctx());
    if (nranks < 1 || rank < 0 || rank >= nranks) return fail(RNLA_ERR_INVALID_PARAMETERS, "rnla_comm_init: bad rank/nranks");
    Ctx& c = g_ctx;
    if (c.comm) return fail(RNLA_ERR_COMPUTATION, "rnla_comm_init: communicator already initialised");
    if (nranks == 1) { c.nranks = 1; c.rank = 0; return RNLA_OK; }
    RNLA_TRY(load_nccl());
    ncclUniqueId uid;
    memcpy(&uid, id, 128);
    ncclComm_t comm;
    ncclResult_t r = g_nccl.CommInitRank(&comm, nranks, uid, rank);
    if (r != ncclSuccess) return nccl_fail(r, "ncclCommInitRank");
    c.comm = comm; c.nranks = nranks; c.rank = rank;
    return RNLA_OK;
}
rnla_status rnla_comm_destroy(void) {
    RNLA_API_GUARD;
    Ctx& c = g_ctx;
    if (c.comm) {
        cudaStreamSynchronize(c.stream);
        g_nccl.CommDestroy((ncclComm_t)c.comm);
        c.comm = nullptr;
    }
    c.nranks = 1; c.rank = 0;
    return RNLA_OK;
}
int32_t rnla_comm_size(void) { return g_ctx.nranks; }
int32_t rnla_comm_rank(void) { return g_ctx.rank; }

rnla_status rnla_malloc(void** dptr, size_t bytes) {
    RNLA_API_GUARD;
    RNLA_TRY(ensure_ctx());
    RNLA_CUDA(cudaMalloc(dptr, bytes ? bytes : 8));
    return RNLA_OK;
}
rnla_status rnla_free(void* dptr) {
    RNLA_API_GUARD;
    RNLA_TRY(ensure_ctx());
    RNLA_CUDA(cudaStreamSynchronize(g_ctx.stream));
    RNLA_CUDA(cudaFree(dptr));
    return RNLA_OK;
}
rnla_status rnla_memcpy_h2d(void* dst, const void* src, size_t bytes) {
    RNLA_API_GUARD;
    RNLA_TRY(ensure_ctx());
    RNLA_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, g_ctx.stream));
    RNLA_CUDA(cudaStreamSynchronize(g_ctx.stream));
    return RNLA_OK;
}
rnla_status rnla_memcpy_d2h(void* dst, const void* src, size_t bytes) {
    RNLA_API_GUARD;
    RNLA_TRY(ensure_ctx());
    RNLA_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, g_ctx.stream));
    RNLA_CUDA(cudaStreamSynchronize(g_ctx.stream));
    return RNLA_OK;
}

}  // extern "C"
