// capi.cu — context, log residency and the host-buffer entry points of include/siesta_gpu.h.
#include <cstdlib>
#include <cstring>
#include <vector>

#include "common.cuh"

namespace siesta {

static thread_local std::string t_last_error;
std::atomic<long long> g_kernel_launches{0};

void set_error(const std::string& msg) { t_last_error = msg; }

}  // namespace siesta

using namespace siesta;

extern "C" const char* siesta_last_error(void) { return t_last_error.c_str(); }
extern "C" int64_t siesta_kernel_launches(void) { return (int64_t)g_kernel_launches.load(); }

extern "C" int siesta_init(int32_t device_id, siesta_ctx** out) {
    if (!out) return SIESTA_E_INVALID;
    int n_dev = 0;
    cudaError_t e = cudaGetDeviceCount(&n_dev);
    if (e != cudaSuccess || n_dev == 0) {
        set_error(std::string("no CUDA device: ") + (e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0") +
                  " (libsiesta_gpu has no CPU fallback)");
        return SIESTA_E_CUDA;
    }
    if (device_id < 0 || device_id >= n_dev) {
        set_error("device id out of range");
        return SIESTA_E_INVALID;
    }
    cudaDeviceProp prop;
    SIESTA_CUDA_OK(cudaGetDeviceProperties(&prop, device_id));
    if (prop.major != 10) {
        set_error(std::string("device ") + prop.name + " is sm_" + std::to_string(prop.major * 10 + prop.minor) +
                  "; libsiesta_gpu is built for sm_100a (B200) only");
        return SIESTA_E_CUDA;
    }
    SIESTA_CUDA_OK(cudaSetDevice(device_id));
    Ctx* c = new Ctx();
    c->device = device_id;
    c->sm_count = prop.multiProcessorCount;
    SIESTA_CUDA_OK(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    // keep the stream-ordered pool warm between requests (workspaces are re-used, not returned to the driver)
    cudaMemPool_t pool;
    SIESTA_CUDA_OK(cudaDeviceGetDefaultMemPool(&pool, device_id));
    unsigned long long keep = ~0ull;
    SIESTA_CUDA_OK(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep));
    *out = reinterpret_cast<siesta_ctx*>(c);
    return SIESTA_OK;
}

extern "C" void siesta_shutdown(siesta_ctx* ctx) {
    if (!ctx) return;
    Ctx* c = reinterpret_cast<Ctx*>(ctx);
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
}

// One pass over the activity column when a log becomes resident: are all ids inside [0, n_activities)?
// (kernel K1's filter then tests ids with two shift instructions per event instead of a checked table lookup)
__global__ void act_range_kernel(const int32_t* act, int64_t n, int32_t n_act, int* bad) {
    int any = 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        any |= (unsigned)act[i] >= (unsigned)n_act;
    if (__any_sync(0xffffffffu, any) && (threadIdx.x & 31) == 0) atomicOr(bad, 1);
}

static int validate_act(Log* L) {
    L->act_valid = false;
    if (L->n_events == 0) {
        L->act_valid = true;
        return SIESTA_OK;
    }
    Ctx* c = L->ctx;
    int* d_bad = nullptr;
    SIESTA_CUDA_OK(cudaMallocAsync((void**)&d_bad, sizeof(int), c->stream));
    SIESTA_CUDA_OK(cudaMemsetAsync(d_bad, 0, sizeof(int), c->stream));
    const int64_t want = (L->n_events + 255) / 256;
    const int grid = (int)(want < (int64_t)c->sm_count * 16 ? want : (int64_t)c->sm_count * 16);
    act_range_kernel<<<grid, 256, 0, c->stream>>>(L->d_act, L->n_events, L->n_activities, d_bad);
    SIESTA_LAUNCHED();
    int bad = 1;
    SIESTA_CUDA_OK(cudaMemcpyAsync(&bad, d_bad, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    SIESTA_CUDA_OK(cudaStreamSynchronize(c->stream));
    cudaFreeAsync(d_bad, c->stream);
    L->act_valid = bad == 0;
    return SIESTA_OK;
}

static int check_csr(const int64_t* trace_off, int64_t n_traces, int64_t n_events, int32_t* max_len) {
    if (n_traces < 0 || n_events < 0 || trace_off[0] != 0 || trace_off[n_traces] != n_events) {
        set_error("CSR log: trace_off must start at 0 and end at n_events");
        return SIESTA_E_INVALID;
    }
    int64_t mx = 0;
    for (int64_t t = 0; t < n_traces; ++t) {
        const int64_t len = trace_off[t + 1] - trace_off[t];
        if (len < 0) {
            set_error("CSR log: trace_off must be non-decreasing");
            return SIESTA_E_INVALID;
        }
        mx = len > mx ? len : mx;
    }
    *max_len = (int32_t)(mx > 0x7fffffff ? 0x7fffffff : mx);
    return SIESTA_OK;
}

extern "C" int siesta_log_load(siesta_ctx* ctx, const int64_t* trace_off, const int32_t* act, const int64_t* ts_ms,
                               int64_t n_traces, int64_t n_events, int32_t n_activities, siesta_log** out) {
    if (!ctx || !trace_off || (!act && n_events) || (!ts_ms && n_events) || !out || n_activities < 0) {
        set_error("siesta_log_load: null argument");
        return SIESTA_E_INVALID;
    }
    Ctx* c = reinterpret_cast<Ctx*>(ctx);
    int32_t max_len = 0;
    int rc = check_csr(trace_off, n_traces, n_events, &max_len);
    if (rc) return rc;
    SIESTA_CUDA_OK(cudaSetDevice(c->device));
    Log* L = new Log();
    L->ctx = c;
    L->n_traces = n_traces;
    L->n_events = n_events;
    L->n_activities = n_activities;
    L->max_trace_len = max_len;
    L->owns = true;
    void *d_off = nullptr, *d_act = nullptr, *d_ts = nullptr;
    const size_t ne = (size_t)(n_events ? n_events : 1);
    cudaError_t e;
    if ((e = cudaMalloc(&d_off, (size_t)(n_traces + 1) * 8)) != cudaSuccess || (e = cudaMalloc(&d_act, ne * 4 + 16)) != cudaSuccess ||
        (e = cudaMalloc(&d_ts, ne * 8 + 16)) != cudaSuccess) {
        set_error(std::string("siesta_log_load: cudaMalloc: ") + cudaGetErrorString(e));
        cudaFree(d_off);
        cudaFree(d_act);
        cudaFree(d_ts);
        delete L;
        return SIESTA_E_NOMEM;
    }
    L->d_trace_off = (const int64_t*)d_off;
    L->d_act = (const int32_t*)d_act;
    L->d_ts_ms = (const int64_t*)d_ts;
    SIESTA_CUDA_OK(cudaMemcpyAsync(d_off, trace_off, (size_t)(n_traces + 1) * 8, cudaMemcpyHostToDevice, c->stream));
    if (n_events) {
        SIESTA_CUDA_OK(cudaMemcpyAsync(d_act, act, (size_t)n_events * 4, cudaMemcpyHostToDevice, c->stream));
        SIESTA_CUDA_OK(cudaMemcpyAsync(d_ts, ts_ms, (size_t)n_events * 8, cudaMemcpyHostToDevice, c->stream));
    }
    SIESTA_CUDA_OK(cudaStreamSynchronize(c->stream));
    if ((rc = validate_act(L))) {
        siesta_log_free(reinterpret_cast<siesta_log*>(L));
        return rc;
    }
    *out = reinterpret_cast<siesta_log*>(L);
    return SIESTA_OK;
}

extern "C" int siesta_log_wrap_device(siesta_ctx* ctx, const int64_t* d_trace_off, const int32_t* d_act,
                                      const int64_t* d_ts_ms, int64_t n_traces, int64_t n_events, int32_t n_activities,
                                      int32_t max_trace_len, siesta_log** out) {
    if (!ctx || !d_trace_off || !out || n_traces < 0 || n_events < 0) {
        set_error("siesta_log_wrap_device: bad argument");
        return SIESTA_E_INVALID;
    }
    Log* L = new Log();
    L->ctx = reinterpret_cast<Ctx*>(ctx);
    L->d_trace_off = d_trace_off;
    L->d_act = d_act;
    L->d_ts_ms = d_ts_ms;
    L->n_traces = n_traces;
    L->n_events = n_events;
    L->n_activities = n_activities;
    L->max_trace_len = max_trace_len;
    L->owns = false;
    SIESTA_CUDA_OK(cudaSetDevice(L->ctx->device));
    int rc = validate_act(L);
    if (rc) {
        delete L;
        return rc;
    }
    *out = reinterpret_cast<siesta_log*>(L);
    return SIESTA_OK;
}

extern "C" void siesta_log_free(siesta_log* log) {
    if (!log) return;
    Log* L = reinterpret_cast<Log*>(log);
    if (L->owns) {
        cudaSetDevice(L->ctx->device);
        cudaFree((void*)L->d_trace_off);
        cudaFree((void*)L->d_act);
        cudaFree((void*)L->d_ts_ms);
    }
    delete L;
}

extern "C" int64_t siesta_log_n_traces(const siesta_log* log) { return log ? reinterpret_cast<const Log*>(log)->n_traces : 0; }
extern "C" int64_t siesta_log_n_events(const siesta_log* log) { return log ? reinterpret_cast<const Log*>(log)->n_events : 0; }

template <class T>
static int fetch(T** dst, const void* d_src, int64_t n, cudaStream_t s) {
    *dst = (T*)std::malloc(sizeof(T) * (size_t)(n > 0 ? n : 1));
    if (!*dst) return SIESTA_E_NOMEM;
    if (n > 0 && d_src) SIESTA_CUDA_OK(cudaMemcpyAsync(*dst, d_src, sizeof(T) * (size_t)n, cudaMemcpyDeviceToHost, s));
    return SIESTA_OK;
}

extern "C" void siesta_matches_free(siesta_matches* m) {
    if (!m) return;
    std::free(m->trace_idx);
    std::free(m->occ_off);
    std::free(m->ev_off);
    std::free(m->ev_pos);
    std::free(m->ev_rank);
    std::free(m->ev_act);
    std::free(m->ev_ts_ms);
    std::free(m->err_trace_idx);
    std::free(m);
}

extern "C" int siesta_detect(siesta_log* log, const siesta_nfa* nfa, const int64_t* cand, int64_t n_cand, uint32_t flags,
                             siesta_matches** out) {
    if (!log || !nfa || !out || (cand && n_cand < 0)) {
        set_error("siesta_detect: null argument");
        return SIESTA_E_INVALID;
    }
    Log* L = reinterpret_cast<Log*>(log);
    SIESTA_CUDA_OK(cudaSetDevice(L->ctx->device));
    // per-call stream: concurrent requests on one log do not serialise on the ctx stream
    cudaStream_t stream;
    SIESTA_CUDA_OK(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
    int64_t* d_cand = nullptr;
    int rc = SIESTA_OK;
    siesta_dev_matches dm;
    std::memset(&dm, 0, sizeof(dm));
    siesta_matches* m = nullptr;
    do {
        if (cand) {
            for (int64_t i = 0; i < n_cand; ++i)
                if (cand[i] < 0 || cand[i] >= L->n_traces) {
                    set_error("siesta_detect: candidate trace index out of range");
                    rc = SIESTA_E_INVALID;
                    break;
                }
            if (rc) break;
            if (cudaMalloc((void**)&d_cand, (size_t)(n_cand > 0 ? n_cand : 1) * 8) != cudaSuccess) {
                set_error("siesta_detect: cudaMalloc(candidates)");
                rc = SIESTA_E_NOMEM;
                break;
            }
            if (n_cand && cudaMemcpyAsync(d_cand, cand, (size_t)n_cand * 8, cudaMemcpyHostToDevice, stream) != cudaSuccess) {
                set_error("siesta_detect: H2D(candidates)");
                rc = SIESTA_E_CUDA;
                break;
            }
        }
        rc = siesta_detect_device(log, nfa, d_cand, n_cand, flags, stream, &dm);
        if (rc) break;
        m = (siesta_matches*)std::calloc(1, sizeof(siesta_matches));
        m->n_traces = dm.n_traces;
        m->n_occurrences = dm.n_occurrences;
        m->n_events = dm.n_events;
        m->n_matches_emitted = dm.n_matches_emitted;
        m->n_ref_errors = dm.n_ref_errors;
        m->kernel_ms = dm.kernel_ms;
        m->detect_ms = dm.detect_ms;
        if ((rc = fetch(&m->trace_idx, dm.d_trace_idx, dm.n_traces, stream)) ||
            (rc = fetch(&m->occ_off, dm.d_occ_off, dm.n_traces + 1, stream)) ||
            (rc = fetch(&m->ev_off, dm.d_ev_off, dm.n_occurrences + 1, stream)) ||
            (rc = fetch(&m->ev_pos, dm.d_ev_pos, dm.n_events, stream)) ||
            (rc = fetch(&m->err_trace_idx, dm.d_err_trace_idx, dm.n_ref_errors, stream)))
            break;
        if (!(flags & SIESTA_F_NO_EVENT_COLUMNS)) {
            if ((rc = fetch(&m->ev_rank, dm.d_ev_rank, dm.n_events, stream)) ||
                (rc = fetch(&m->ev_act, dm.d_ev_act, dm.n_events, stream)) ||
                (rc = fetch(&m->ev_ts_ms, dm.d_ev_ts_ms, dm.n_events, stream)))
                break;
        }
        if (cudaStreamSynchronize(stream) != cudaSuccess) {
            set_error("siesta_detect: D2H");
            rc = SIESTA_E_CUDA;
        }
    } while (0);
    siesta_dev_matches_free(&dm);
    if (d_cand) cudaFree(d_cand);
    cudaStreamDestroy(stream);
    if (rc) {
        siesta_matches_free(m);
        return rc;
    }
    *out = m;
    return SIESTA_OK;
}

extern "C" int siesta_evaluate_events(siesta_ctx* ctx, const int64_t* trace_off, const int32_t* act, const int64_t* ts_ms,
                                      int64_t n_traces, int64_t n_events, int32_t n_activities, const siesta_nfa* nfa,
                                      uint32_t flags, siesta_matches** out) {
    siesta_log* log = nullptr;
    int rc = siesta_log_load(ctx, trace_off, act, ts_ms, n_traces, n_events, n_activities, &log);
    if (rc) return rc;
    rc = siesta_detect(log, nfa, nullptr, 0, flags, out);
    siesta_log_free(log);
    return rc;
}
