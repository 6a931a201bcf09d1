// capi.cu — context, log residency and the host-buffer entry points of include/siesta_gpu.h.
#include <algorithm>
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
    {   // the first slab of pinned counter blocks now, not inside a request
        void* slab = nullptr;
        SIESTA_CUDA_OK(cudaHostAlloc(&slab, 32 * 128, cudaHostAllocDefault));
        c->pinned_slabs.push_back(slab);
        for (int i = 0; i < 32; ++i) c->pinned_counters.push_back(reinterpret_cast<unsigned long long*>(reinterpret_cast<char*>(slab) + 128 * i));
    }
    *out = reinterpret_cast<siesta_ctx*>(c);
    return SIESTA_OK;
}

namespace siesta {
void* dev_arena_alloc(Ctx* c, size_t bytes) {
    if (bytes < 256) bytes = 256;
    {
        std::lock_guard<std::mutex> g(c->arena_mu);
        DevBlock* best = nullptr;
        for (DevBlock& b : c->dev_arena)
            if (!b.used && b.size >= bytes && b.size <= 2 * bytes + (64u << 20) && (!best || b.size < best->size)) best = &b;
        if (best) {
            best->used = true;
            return best->p;
        }
    }
    const size_t want = ((bytes + bytes / 8) + ((size_t)2 << 20) - 1) & ~(((size_t)2 << 20) - 1);   // slack: the next request may be a little larger
    void* p = nullptr;
    cudaSetDevice(c->device);
    cudaError_t e = cudaMalloc(&p, want);
    if (e != cudaSuccess) {   // give the unused blocks back to the driver and try once more
        cudaGetLastError();
        std::vector<void*> drop;
        {
            std::lock_guard<std::mutex> g(c->arena_mu);
            for (size_t i = 0; i < c->dev_arena.size();) {
                if (!c->dev_arena[i].used) {
                    drop.push_back(c->dev_arena[i].p);
                    c->dev_arena.erase(c->dev_arena.begin() + (long)i);
                } else ++i;
            }
        }
        for (void* q : drop) cudaFree(q);
        e = cudaMalloc(&p, want);
    }
    if (e != cudaSuccess) {
        cudaGetLastError();
        set_error(std::string("cudaMalloc(") + std::to_string(want) + "): " + cudaGetErrorString(e));
        return nullptr;
    }
    std::lock_guard<std::mutex> g(c->arena_mu);
    c->dev_arena.push_back(DevBlock{p, want, true});
    return p;
}

void dev_arena_free(Ctx* c, void* p) {
    if (!p) return;
    std::lock_guard<std::mutex> g(c->arena_mu);
    for (DevBlock& b : c->dev_arena)
        if (b.p == p) {
            b.used = false;
            return;
        }
}
}  // namespace siesta

extern "C" void siesta_shutdown(siesta_ctx* ctx) {
    if (!ctx) return;
    Ctx* c = reinterpret_cast<Ctx*>(ctx);
    cudaSetDevice(c->device);
    cudaDeviceSynchronize();
    for (DevBlock& b : c->dev_arena) cudaFree(b.p);   // results must have been freed before the ctx
    if (c->stream) cudaStreamDestroy(c->stream);
    for (HostBlock& b : c->arena) cudaFreeHost(b.p);  // result objects must have been freed before the ctx
    for (void* p : c->pinned_slabs) cudaFreeHost(p);
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

namespace siesta {
// Marks L->act_valid = false if events [first_event, first_event + n_events) hold an id outside [0, n_activities).
int validate_act_range(Log* L, int64_t first_event, int64_t n_events, cudaStream_t stream) {
    if (n_events == 0) return SIESTA_OK;
    Ctx* c = L->ctx;
    int* d_bad = nullptr;
    SIESTA_CUDA_OK(cudaMallocAsync((void**)&d_bad, sizeof(int), stream));
    SIESTA_CUDA_OK(cudaMemsetAsync(d_bad, 0, sizeof(int), stream));
    const int64_t want = (n_events + 255) / 256;
    const int grid = (int)(want < (int64_t)c->sm_count * 16 ? want : (int64_t)c->sm_count * 16);
    act_range_kernel<<<grid, 256, 0, stream>>>(L->d_act + first_event, n_events, L->n_activities, d_bad);
    SIESTA_LAUNCHED();
    int bad = 1;
    SIESTA_CUDA_OK(cudaMemcpyAsync(&bad, d_bad, sizeof(int), cudaMemcpyDeviceToHost, stream));
    SIESTA_CUDA_OK(cudaStreamSynchronize(stream));
    cudaFreeAsync(d_bad, stream);
    if (bad) L->act_valid = false;
    return SIESTA_OK;
}
}  // namespace siesta

// A wrapped CSR lives in memory the caller filled: offsets must start at 0, never decrease and end at n_events, or every
// kernel would read out of bounds.  One pass over trace_off on the device.
__global__ void csr_check_kernel(const int64_t* off, int64_t n_traces, int64_t n_events, int* bad) {
    int any = 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i <= n_traces; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t v = off[i];
        if (v < 0 || v > n_events) any = 1;
        if (i == 0 && v != 0) any = 1;
        if (i == n_traces && v != n_events) any = 1;
        if (i < n_traces && off[i + 1] < v) any = 1;
    }
    if (__any_sync(0xffffffffu, any) && (threadIdx.x & 31) == 0) atomicOr(bad, 1);
}

static int validate_device_csr(Log* L) {
    Ctx* c = L->ctx;
    int* d_bad = nullptr;
    SIESTA_CUDA_OK(cudaMallocAsync((void**)&d_bad, sizeof(int), c->stream));
    struct Free {
        int* p;
        cudaStream_t s;
        ~Free() { cudaFreeAsync(p, s); }
    } guard{d_bad, c->stream};
    SIESTA_CUDA_OK(cudaMemsetAsync(d_bad, 0, sizeof(int), c->stream));
    const int64_t want = (L->n_traces + 1 + 255) / 256;
    const int grid = (int)(want < (int64_t)c->sm_count * 16 ? want : (int64_t)c->sm_count * 16);
    csr_check_kernel<<<grid, 256, 0, c->stream>>>(L->d_trace_off, L->n_traces, L->n_events, d_bad);
    SIESTA_LAUNCHED();
    int bad = 1;
    SIESTA_CUDA_OK(cudaMemcpyAsync(&bad, d_bad, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    SIESTA_CUDA_OK(cudaStreamSynchronize(c->stream));
    if (bad) {
        set_error("siesta_log_wrap_device: trace_off must start at 0, be non-decreasing and end at n_events");
        return SIESTA_E_INVALID;
    }
    return SIESTA_OK;
}

static int validate_act(Log* L) {
    L->act_valid = true;
    return validate_act_range(L, 0, L->n_events, L->ctx->stream);
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
    e = cudaMemcpyAsync(d_off, trace_off, (size_t)(n_traces + 1) * 8, cudaMemcpyHostToDevice, c->stream);
    if (e == cudaSuccess && n_events) e = cudaMemcpyAsync(d_act, act, (size_t)n_events * 4, cudaMemcpyHostToDevice, c->stream);
    if (e == cudaSuccess && n_events) e = cudaMemcpyAsync(d_ts, ts_ms, (size_t)n_events * 8, cudaMemcpyHostToDevice, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    if (e != cudaSuccess) {
        set_error(std::string("siesta_log_load: host -> device copy: ") + cudaGetErrorString(e));
        siesta_log_free(reinterpret_cast<siesta_log*>(L));   // owns its three allocations
        return SIESTA_E_CUDA;
    }
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
    if (!ctx || !d_trace_off || !out || n_traces < 0 || n_events < 0 || n_activities < 0 || (n_events > 0 && (!d_act || !d_ts_ms))) {
        set_error("siesta_log_wrap_device: bad argument (null column with n_events > 0?)");
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
    int rc = cudaSetDevice(L->ctx->device) == cudaSuccess ? SIESTA_OK : SIESTA_E_CUDA;
    if (rc == SIESTA_OK) rc = validate_device_csr(L);
    if (rc == SIESTA_OK) rc = validate_act(L);
    if (rc) {
        delete L;
        return rc;
    }
    *out = reinterpret_cast<siesta_log*>(L);
    return SIESTA_OK;
}

extern "C" void siesta_log_set_first_trace(siesta_log* log, int64_t first_trace) {
    if (log) reinterpret_cast<Log*>(log)->first_trace = first_trace;
}

extern "C" int siesta_log_set_blocks(siesta_log* log, int32_t n_blocks, const int64_t* local_first, const int64_t* global_first) {
    Log* L = reinterpret_cast<Log*>(log);
    if (!L || n_blocks < 0 || n_blocks > SIESTA_MAX_BLOCKS || (n_blocks > 0 && (!local_first || !global_first))) {
        set_error("siesta_log_set_blocks: bad argument (0 <= n_blocks <= SIESTA_MAX_BLOCKS)");
        return SIESTA_E_INVALID;
    }
    if (n_blocks > 0) {
        bool ok = local_first[0] == 0 && local_first[n_blocks] == L->n_traces;
        for (int b = 0; b < n_blocks && ok; ++b) ok = local_first[b] <= local_first[b + 1] && global_first[b] >= 0;
        if (!ok) {
            set_error("siesta_log_set_blocks: local_first must ascend from 0 to the shard's trace count, global_first >= 0");
            return SIESTA_E_INVALID;
        }
        for (int b = 0; b <= n_blocks; ++b) L->blk_local[b] = local_first[b];
        for (int b = 0; b < n_blocks; ++b) L->blk_global[b] = global_first[b];
    }
    L->n_blocks = n_blocks;
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
        if (L->d_src_event) cudaFree((void*)L->d_src_event);
    }
    delete L;
}

extern "C" int64_t siesta_log_n_traces(const siesta_log* log) { return log ? reinterpret_cast<const Log*>(log)->n_traces : 0; }
extern "C" int64_t siesta_log_n_events(const siesta_log* log) { return log ? reinterpret_cast<const Log*>(log)->n_events : 0; }

// ------------------------------------------------------------------------------------------ host result objects
// One block per result: [ResultHeader][siesta_matches][arrays ...].  Large results live in pinned memory from the
// ctx arena (device -> host copies run at PCIe speed and the block is reused by the next request).
namespace {
constexpr uint64_t RESULT_MAGIC = 0x5349455354414d31ull;  // "SIESTAM1"
constexpr size_t PIN_THRESHOLD = 256 * 1024;
constexpr size_t ARENA_KEEP_BYTES = (size_t)4 << 30;

struct ResultHeader {
    uint64_t magic;
    Ctx* ctx;       // arena owner, or nullptr for a malloc block
    void* base;
    uint64_t pad[5];
};
static_assert(sizeof(ResultHeader) == 64, "ResultHeader is one cache line");

void* arena_alloc(Ctx* c, size_t bytes) {
    size_t want = (size_t)1 << 20;
    while (want < bytes) want <<= 1;
    {
        std::lock_guard<std::mutex> g(c->arena_mu);
        for (HostBlock& b : c->arena)
            if (!b.used && b.size >= bytes && b.size <= 4 * want) {
                b.used = true;
                return b.p;
            }
    }
    void* p = nullptr;
    if (cudaHostAlloc(&p, want, cudaHostAllocDefault) != cudaSuccess) {
        cudaGetLastError();
        return nullptr;
    }
    std::lock_guard<std::mutex> g(c->arena_mu);
    c->arena.push_back(HostBlock{p, want, true});
    return p;
}

void arena_free(Ctx* c, void* p) {
    std::lock_guard<std::mutex> g(c->arena_mu);
    size_t kept = 0;
    for (HostBlock& b : c->arena) kept += b.used ? 0 : b.size;
    for (size_t i = 0; i < c->arena.size(); ++i)
        if (c->arena[i].p == p) {
            if (kept + c->arena[i].size > ARENA_KEEP_BYTES) {
                cudaFreeHost(p);
                c->arena.erase(c->arena.begin() + i);
            } else {
                c->arena[i].used = false;
            }
            return;
        }
}

inline size_t align64(size_t x) { return (x + 63) & ~(size_t)63; }

}  // namespace
namespace siesta {
// Device results of one or more consecutive chunks / shards (already rebased: see RebaseOffsets) -> one host
// siesta_matches.  part_ctx (optional): the context (device, stream) each part lives on - shards of a multi-GPU log.
int assemble_matches(Ctx* c, std::vector<siesta_dev_matches>& parts, uint32_t flags, cudaStream_t stream, siesta_matches** out,
                     const std::vector<Ctx*>* part_ctx) {
    int64_t n_tr = 0, n_occ = 0, n_ev = 0, n_err = 0, n_emit = 0, n_unsup = 0;
    double k_ms = 0, d_ms = 0;
    bool counted = true;
    for (const siesta_dev_matches& p : parts) {
        n_tr += p.n_traces;
        n_occ += p.n_occurrences;
        n_ev += p.n_events;
        n_err += p.n_ref_errors;
        n_unsup += p.n_unsupported;
        if (p.n_matches_emitted < 0) counted = false;
        else n_emit += p.n_matches_emitted;
        k_ms += p.kernel_ms;
        d_ms += p.detect_ms;
    }
    const bool all_cols = !(flags & SIESTA_F_NO_EVENT_COLUMNS);
    size_t off = sizeof(ResultHeader) + align64(sizeof(siesta_matches));
    const size_t o_trace = off; off += align64((size_t)(n_tr ? n_tr : 1) * 8);
    const size_t o_occ = off;   off += align64((size_t)(n_tr + 1) * 8);
    const size_t o_evoff = off; off += align64((size_t)(n_occ + 1) * 8);
    const size_t o_pos = off;   off += align64((size_t)(n_ev ? n_ev : 1) * 4);
    const size_t o_err = off;   off += align64((size_t)(n_err ? n_err : 1) * 8);
    const size_t o_unsup = off; off += align64((size_t)(n_unsup ? n_unsup : 1) * 8);
    size_t o_rank = 0, o_act = 0, o_ts = 0;
    if (all_cols) {
        o_rank = off; off += align64((size_t)(n_ev ? n_ev : 1) * 4);
        o_act = off;  off += align64((size_t)(n_ev ? n_ev : 1) * 4);
        o_ts = off;   off += align64((size_t)(n_ev ? n_ev : 1) * 8);
    }
    char* base = nullptr;
    Ctx* owner = nullptr;
    if (off >= PIN_THRESHOLD) {
        base = (char*)arena_alloc(c, off);
        if (base) owner = c;
    }
    if (!base) base = (char*)std::malloc(off);
    if (!base) return SIESTA_E_NOMEM;
    ResultHeader* h = reinterpret_cast<ResultHeader*>(base);
    h->magic = RESULT_MAGIC;
    h->ctx = owner;
    h->base = base;
    siesta_matches* m = reinterpret_cast<siesta_matches*>(base + sizeof(ResultHeader));
    std::memset(m, 0, sizeof(*m));
    m->n_traces = n_tr;
    m->n_occurrences = n_occ;
    m->n_events = n_ev;
    m->n_matches_emitted = counted ? n_emit : -1;
    m->n_ref_errors = n_err;
    m->n_unsupported = n_unsup;
    m->unsupported_trace_idx = reinterpret_cast<int64_t*>(base + o_unsup);
    m->kernel_ms = k_ms;
    m->detect_ms = d_ms;
    m->trace_idx = reinterpret_cast<int64_t*>(base + o_trace);
    m->occ_off = reinterpret_cast<int64_t*>(base + o_occ);
    m->ev_off = reinterpret_cast<int64_t*>(base + o_evoff);
    m->ev_pos = reinterpret_cast<int32_t*>(base + o_pos);
    m->err_trace_idx = reinterpret_cast<int64_t*>(base + o_err);
    if (all_cols) {
        m->ev_rank = reinterpret_cast<int32_t*>(base + o_rank);
        m->ev_act = reinterpret_cast<int32_t*>(base + o_act);
        m->ev_ts_ms = reinterpret_cast<int64_t*>(base + o_ts);
    }
    m->occ_off[0] = 0;
    m->ev_off[0] = 0;
    int64_t a_tr = 0, a_occ = 0, a_ev = 0, a_err = 0, a_unsup = 0;
    cudaError_t e = cudaSuccess;
    cudaStream_t cur = stream;
    auto d2h = [&](void* dst, const void* src, size_t bytes) {
        if (bytes && src && e == cudaSuccess) e = cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, cur);
    };
    for (size_t pi = 0; pi < parts.size(); ++pi) {
        const siesta_dev_matches& p = parts[pi];
        if (part_ctx) {
            if (e == cudaSuccess) e = cudaSetDevice((*part_ctx)[pi]->device);
            cur = (*part_ctx)[pi]->stream;
        }
        d2h(m->trace_idx + a_tr, p.d_trace_idx, (size_t)p.n_traces * 8);
        // each part carries its own tail entry; the next part overwrites it with the same value
        d2h(m->occ_off + a_tr, p.d_occ_off, (size_t)(p.n_traces + 1) * 8);
        d2h(m->ev_off + a_occ, p.d_ev_off, (size_t)(p.n_occurrences + 1) * 8);
        d2h(m->ev_pos + a_ev, p.d_ev_pos, (size_t)p.n_events * 4);
        d2h(m->err_trace_idx + a_err, p.d_err_trace_idx, (size_t)p.n_ref_errors * 8);
        d2h(m->unsupported_trace_idx + a_unsup, p.d_unsupported_trace_idx, (size_t)p.n_unsupported * 8);
        if (all_cols) {
            d2h(m->ev_rank + a_ev, p.d_ev_rank, (size_t)p.n_events * 4);
            d2h(m->ev_act + a_ev, p.d_ev_act, (size_t)p.n_events * 4);
            d2h(m->ev_ts_ms + a_ev, p.d_ev_ts_ms, (size_t)p.n_events * 8);
        }
        a_tr += p.n_traces;
        a_occ += p.n_occurrences;
        a_ev += p.n_events;
        a_err += p.n_ref_errors;
        a_unsup += p.n_unsupported;
    }
    if (part_ctx) {
        for (size_t pi = 0; pi < parts.size() && e == cudaSuccess; ++pi) {
            e = cudaSetDevice((*part_ctx)[pi]->device);
            if (e == cudaSuccess) e = cudaStreamSynchronize((*part_ctx)[pi]->stream);
        }
    } else if (e == cudaSuccess) {
        e = cudaStreamSynchronize(stream);
    }
    if (e != cudaSuccess) {
        set_error(std::string("device -> host copy of the occurrences: ") + cudaGetErrorString(e));
        siesta_matches_free(m);
        return SIESTA_E_CUDA;
    }
    *out = m;
    return SIESTA_OK;
}
}  // namespace siesta

extern "C" void siesta_matches_free(siesta_matches* m) {
    if (!m) return;
    ResultHeader* h = reinterpret_cast<ResultHeader*>(reinterpret_cast<char*>(m) - sizeof(ResultHeader));
    if (h->magic != RESULT_MAGIC) return;  // not ours (or freed twice)
    h->magic = 0;
    if (h->ctx) arena_free(h->ctx, h->base);
    else std::free(h->base);
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
    std::vector<siesta_dev_matches> parts(1);
    std::memset(&parts[0], 0, sizeof(siesta_dev_matches));
    do {
        if (cand) {
            for (int64_t i = 0; i < n_cand; ++i)
                if (cand[i] < 0 || cand[i] >= L->n_traces) {
                    set_error("siesta_detect: candidate trace index out of range");
                    rc = SIESTA_E_INVALID;
                    break;
                }
            if (rc) break;
            if (cudaMallocAsync((void**)&d_cand, (size_t)(n_cand > 0 ? n_cand : 1) * 8, stream) != cudaSuccess) {
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
        rc = siesta_detect_device(log, nfa, d_cand, n_cand, flags, stream, &parts[0]);
        if (rc) break;
        rc = assemble_matches(L->ctx, parts, flags, stream, out, nullptr);
    } while (0);
    siesta_dev_matches_free(&parts[0]);
    if (d_cand) cudaFreeAsync(d_cand, stream);
    cudaStreamSynchronize(stream);
    cudaStreamDestroy(stream);
    return rc;
}

// Literal SaseConnector.evaluate signature: the request's events arrive in host buffers.  The log is cut into chunks
// of whole traces; chunk c + 1 (and c + 2) travel host -> device on a copy stream while chunk c is verified, so the
// call runs at the speed of the host link.  Chunk results are rebased on the device (RebaseOffsets) and land in one
// host block.
//
// The offsets of a chunk are validated on the host right before the chunk is enqueued (the check of chunk c + 1 runs while
// chunk c is on the link) instead of in one serial host pass over all of trace_off in front of the first copy.
// act8 != nullptr: the activity column arrives as one byte per event (siesta_evaluate_events_act8); it crosses the link
// narrow and widen_act8_kernel rebuilds the int32 column the kernels read, chunk by chunk, on the run stream.
__global__ void widen_act8_kernel(const uint8_t* __restrict__ src, int32_t* __restrict__ dst, int64_t e0, int64_t e1) {
    // groups of four events at absolute indices [4 g, 4 g + 4): both columns come from cudaMalloc, so a whole group is one
    // aligned 4-byte load and one aligned 16-byte store; the ragged groups at the two ends of the chunk go event by event
    const int64_t g0 = e0 >> 2, g1 = (e1 + 3) >> 2;
    for (int64_t g = g0 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < g1; g += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i = g << 2;
        if (i >= e0 && i + 4 <= e1) {
            const uint32_t w = *reinterpret_cast<const uint32_t*>(src + i);
            *reinterpret_cast<int4*>(dst + i) = make_int4((int)(w & 255u), (int)((w >> 8) & 255u), (int)((w >> 16) & 255u), (int)(w >> 24));
        } else {
            for (int64_t j = (i > e0 ? i : e0); j < e1 && j < i + 4; ++j) dst[j] = (int32_t)src[j];
        }
    }
}

static int check_csr_range(const int64_t* trace_off, int64_t t0, int64_t t1, int64_t n_events, int32_t* max_len) {
    int64_t mx = *max_len, bad = 0;
    for (int64_t t = t0; t < t1; ++t) {
        const int64_t len = trace_off[t + 1] - trace_off[t];
        bad |= len | trace_off[t + 1];   // no offset below 0, so no difference overflows
        mx = len > mx ? len : mx;
    }
    if (bad < 0 || trace_off[t0] < 0 || trace_off[t1] > n_events) {
        set_error("CSR log: trace_off must be non-decreasing and lie in [0, n_events]");
        return SIESTA_E_INVALID;
    }
    *max_len = (int32_t)(mx > 0x7fffffff ? 0x7fffffff : mx);
    return SIESTA_OK;
}

static int evaluate_events_impl(siesta_ctx* ctx, const int64_t* trace_off, const int32_t* act, const uint8_t* act8,
                                const int64_t* ts_ms, int64_t n_traces, int64_t n_events, int32_t n_activities,
                                const siesta_nfa* nfa, uint32_t flags, siesta_matches** out) {
    if (!ctx || !trace_off || (!act && !act8 && n_events) || (!ts_ms && n_events) || !nfa || !out || n_activities < 0 ||
        (act8 && n_activities > 256)) {
        set_error(act8 && n_activities > 256 ? "siesta_evaluate_events_act8: more than 256 activities do not fit a byte column"
                                             : "siesta_evaluate_events: null argument");
        return SIESTA_E_INVALID;
    }
    Ctx* c = reinterpret_cast<Ctx*>(ctx);
    if (n_traces < 0 || n_events < 0 || trace_off[0] != 0 || trace_off[n_traces] != n_events) {
        set_error("CSR log: trace_off must start at 0 and end at n_events");
        return SIESTA_E_INVALID;
    }
    int rc = SIESTA_OK;
    int needs_ts = 1;
    {
        DevNfa dn;  // fail on a malformed NFA before anything is copied
        if ((rc = validate_nfa(nfa, flags, &dn))) return rc;
        std::vector<uint16_t> lut;
        int n_pos = 0;
        build_lut(nfa, dn, n_activities, flags, lut, &needs_ts, &n_pos);
    }
    SIESTA_CUDA_OK(cudaSetDevice(c->device));
    // Timestamps the query does not MATCH on (no time constraint: needs_ts == 0) are only read for the events it reports,
    // a few percent of the log.  If the caller's timestamp column is page-locked, mapped host memory (cudaHostAlloc,
    // cudaHostRegister, torch pin_memory), the kernels read those few values straight from it over the host link
    // instead of the whole column travelling to the device first: 4 B/event cross the link instead of 12.
    const int64_t* ts_mapped = nullptr;
    // (returnAll asks for relative seconds only in Occurrence.overlaps: when kernel K1-P takes the request, the traces with more
    //  than one engine match - the only ones that test overlaps - are a small minority and read theirs through the mapping too)
    const bool few_ts = !needs_ts || ((flags & SIESTA_F_RETURN_ALL) && detect_nkp_eligible(nfa, flags, n_activities));
    if (few_ts && n_events > 0 && std::getenv("SIESTA_NO_TS_ZERO_COPY") == nullptr) {
        cudaPointerAttributes pa;
        if (cudaPointerGetAttributes(&pa, ts_ms) == cudaSuccess && pa.type == cudaMemoryTypeHost && pa.devicePointer)
            ts_mapped = reinterpret_cast<const int64_t*>(pa.devicePointer);
        cudaGetLastError();
    }

    // chunk boundaries: whole traces, about CHUNK_EVENTS events each (about 50 - 64 MB on the link: each chunk costs two
    // host waits, so fewer, larger chunks when only the activity column travels, and fewer again when it travels as bytes)
    int64_t CHUNK_EVENTS = ts_mapped ? (act8 ? (48 << 20) : (16 << 20)) : (4 << 20);
    if (const char* env = std::getenv("SIESTA_CHUNK_EVENTS")) {  // test aid: force many small chunks
        const long long v = std::atoll(env);
        if (v > 0) CHUNK_EVENTS = v;
    }
    std::vector<int64_t> cut{0};
    while (cut.back() < n_traces) {
        const int64_t t0 = cut.back();
        // (offsets are validated chunk by chunk below; on a malformed array the search still ends inside it)
        const int64_t* lim = std::upper_bound(trace_off + t0 + 1, trace_off + n_traces + 1, trace_off[t0] + CHUNK_EVENTS);
        int64_t t1 = (int64_t)(lim - trace_off) - 1;   // last trace end <= budget
        if (t1 <= t0) t1 = t0 + 1;                      // a single trace longer than the budget
        cut.push_back(t1);
    }
    const int n_chunks = (int)cut.size() - 1;

    // three streams: s_copy carries the chunks over the host link back to back; s_prep widens a byte column and checks the
    // activity ids of chunk c as soon as it has landed (its verdict goes to a pinned flag the host reads when it gets to the
    // chunk: no dedicated wait, no allocation per chunk); s_run verifies
    cudaStream_t s_copy = nullptr, s_run = nullptr, s_prep = nullptr;
    SIESTA_CUDA_OK(cudaStreamCreateWithFlags(&s_copy, cudaStreamNonBlocking));
    if (cudaStreamCreateWithFlags(&s_run, cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithFlags(&s_prep, cudaStreamNonBlocking) != cudaSuccess) {
        cudaStreamDestroy(s_copy);
        if (s_run) cudaStreamDestroy(s_run);
        set_error("siesta_evaluate_events: cudaStreamCreate");
        return SIESTA_E_CUDA;
    }
    int* h_bad = reinterpret_cast<int*>(arena_alloc(c, (size_t)(n_chunks > 0 ? n_chunks : 1) * sizeof(int)));   // pinned, reused across requests
    if (!h_bad) {
        cudaStreamDestroy(s_copy);
        cudaStreamDestroy(s_run);
        cudaStreamDestroy(s_prep);
        set_error("siesta_evaluate_events: cudaHostAlloc");
        return SIESTA_E_NOMEM;
    }
    int64_t* d_off = nullptr;
    int32_t* d_act = nullptr;
    uint8_t* d_act8 = nullptr;
    int64_t* d_ts = nullptr;
    int* d_bad = nullptr;
    std::vector<cudaEvent_t> ready((size_t)n_chunks, nullptr), copied((size_t)n_chunks, nullptr);
    std::vector<int32_t> chunk_max_len((size_t)n_chunks, 0);
    std::vector<siesta_dev_matches> parts;
    parts.reserve((size_t)n_chunks);
    cudaError_t e = cudaSuccess;
    auto ok = [&](cudaError_t x) {
        if (e == cudaSuccess) e = x;
        return e == cudaSuccess;
    };
    const size_t ne = (size_t)(n_events ? n_events : 1);
    ok(cudaMallocAsync((void**)&d_off, (size_t)(n_traces + 1) * 8, s_copy));
    ok(cudaMallocAsync((void**)&d_act, ne * 4 + 32, s_copy));
    if (act8) ok(cudaMallocAsync((void**)&d_act8, ne + 32, s_copy));
    if (!ts_mapped) ok(cudaMallocAsync((void**)&d_ts, ne * 8 + 32, s_copy));
    ok(cudaMallocAsync((void**)&d_bad, (size_t)(n_chunks > 0 ? n_chunks : 1) * sizeof(int), s_copy));
    if (e == cudaSuccess) ok(cudaMemsetAsync(d_bad, 0, (size_t)(n_chunks > 0 ? n_chunks : 1) * sizeof(int), s_copy));
    int enq = 0;
    int32_t max_len = 0;
    auto enqueue_copy = [&](int k) {
        const int64_t t0 = cut[k], t1 = cut[k + 1];
        if (rc == SIESTA_OK) rc = check_csr_range(trace_off, t0, t1, n_events, &max_len);
        if (rc || e != cudaSuccess) return;
        chunk_max_len[(size_t)k] = max_len;   // longest trace of chunks 0 .. k: an upper bound for chunk k is all a view needs
        const int64_t e0 = trace_off[t0], e1 = trace_off[t1];
        ok(cudaMemcpyAsync(d_off + t0, trace_off + t0, (size_t)(t1 - t0 + 1) * 8, cudaMemcpyHostToDevice, s_copy));
        if (e1 > e0) {
            if (act8) ok(cudaMemcpyAsync(d_act8 + e0, act8 + e0, (size_t)(e1 - e0), cudaMemcpyHostToDevice, s_copy));
            else ok(cudaMemcpyAsync(d_act + e0, act + e0, (size_t)(e1 - e0) * 4, cudaMemcpyHostToDevice, s_copy));
            if (!ts_mapped) ok(cudaMemcpyAsync(d_ts + e0, ts_ms + e0, (size_t)(e1 - e0) * 8, cudaMemcpyHostToDevice, s_copy));
        }
        ok(cudaEventCreateWithFlags(&copied[k], cudaEventDisableTiming));
        ok(cudaEventRecord(copied[k], s_copy));
        ok(cudaStreamWaitEvent(s_prep, copied[k], 0));
        if (e1 > e0 && e == cudaSuccess) {
            const int64_t n = e1 - e0;
            if (act8) {   // chunks are widened in order on one stream: the ragged 4-event groups at a border are written by their own chunk
                const int64_t want = (n / 4 + 256) / 256;
                const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(want, (int64_t)c->sm_count * 8));
                widen_act8_kernel<<<grid, 256, 0, s_prep>>>(d_act8, d_act, e0, e1);
                SIESTA_LAUNCHED();
            }
            const int64_t want = (n + 255) / 256;
            const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(want, (int64_t)c->sm_count * 16));
            act_range_kernel<<<grid, 256, 0, s_prep>>>(d_act + e0, n, n_activities, d_bad + k);
            SIESTA_LAUNCHED();
        }
        ok(cudaMemcpyAsync(h_bad + k, d_bad + k, sizeof(int), cudaMemcpyDeviceToHost, s_prep));
        ok(cudaEventCreateWithFlags(&ready[k], cudaEventDisableTiming));
        ok(cudaEventRecord(ready[k], s_prep));
    };
    // ---- streamed result.  When every matching trace holds ONE occurrence of k events (detect_uniform_k), the request's
    // trace count bounds every column, so the host block is laid out for the worst case before the first chunk and the
    // columns of a chunk travel device -> host as soon as the chunk is verified - under the later chunks' host -> device
    // copies (the link is full duplex) instead of after the last one.  Error / outlier lists get a fixed capacity; a request
    // that outgrows it, or a block above 1 GB, is assembled at the end as before.
    struct Streamed {
        bool on = false;
        char* base = nullptr;
        siesta_matches* m = nullptr;
        cudaStream_t s = nullptr;
        std::vector<cudaEvent_t> done;
        int64_t a_tr = 0, a_occ = 0, a_ev = 0, a_err = 0, a_unsup = 0, n_emit = 0;
        bool counted = true;
        double k_ms = 0, d_ms = 0;
    } S;
    constexpr int64_t LIST_CAP = 65536;
    const bool all_cols = !(flags & SIESTA_F_NO_EVENT_COLUMNS);
    {
        const int uk = detect_uniform_k(nfa, flags);
        if (uk > 0 && n_chunks > 1 && std::getenv("SIESTA_NO_STREAMED_RESULT") == nullptr) {
            const size_t cap_tr = (size_t)n_traces, cap_ev = (size_t)n_traces * (size_t)uk;
            size_t off = sizeof(ResultHeader) + align64(sizeof(siesta_matches));
            const size_t o_trace = off; off += align64(cap_tr * 8);
            const size_t o_occ = off;   off += align64((cap_tr + 1) * 8);
            const size_t o_evoff = off; off += align64((cap_tr + 1) * 8);
            const size_t o_pos = off;   off += align64(cap_ev * 4);
            const size_t o_err = off;   off += align64((size_t)LIST_CAP * 8);
            const size_t o_unsup = off; off += align64((size_t)LIST_CAP * 8);
            size_t o_rank = 0, o_act = 0, o_ts = 0;
            if (all_cols) {
                o_rank = off; off += align64(cap_ev * 4);
                o_act = off;  off += align64(cap_ev * 4);
                o_ts = off;   off += align64(cap_ev * 8);
            }
            if (off <= ((size_t)1 << 30) && cudaStreamCreateWithFlags(&S.s, cudaStreamNonBlocking) == cudaSuccess) {
                S.base = (char*)arena_alloc(c, off);
                if (S.base) {
                    ResultHeader* h = reinterpret_cast<ResultHeader*>(S.base);
                    h->magic = RESULT_MAGIC;
                    h->ctx = c;
                    h->base = S.base;
                    siesta_matches* m = reinterpret_cast<siesta_matches*>(S.base + sizeof(ResultHeader));
                    std::memset(m, 0, sizeof(*m));
                    m->trace_idx = reinterpret_cast<int64_t*>(S.base + o_trace);
                    m->occ_off = reinterpret_cast<int64_t*>(S.base + o_occ);
                    m->ev_off = reinterpret_cast<int64_t*>(S.base + o_evoff);
                    m->ev_pos = reinterpret_cast<int32_t*>(S.base + o_pos);
                    m->err_trace_idx = reinterpret_cast<int64_t*>(S.base + o_err);
                    m->unsupported_trace_idx = reinterpret_cast<int64_t*>(S.base + o_unsup);
                    if (all_cols) {
                        m->ev_rank = reinterpret_cast<int32_t*>(S.base + o_rank);
                        m->ev_act = reinterpret_cast<int32_t*>(S.base + o_act);
                        m->ev_ts_ms = reinterpret_cast<int64_t*>(S.base + o_ts);
                    }
                    m->occ_off[0] = 0;
                    m->ev_off[0] = 0;
                    S.m = m;
                    S.on = true;
                }
            }
            cudaGetLastError();
        }
    }
    auto stream_part = [&](const siesta_dev_matches& p) {   // the chunk's columns -> their final place in the host block
        if (!S.on) return;
        if (S.a_err + p.n_ref_errors > LIST_CAP || S.a_unsup + p.n_unsupported > LIST_CAP || S.a_tr + p.n_traces > n_traces) {
            S.on = false;   // (outgrew the fixed lists: assembled at the end)
            return;
        }
        cudaEvent_t ev = nullptr;
        cudaError_t x = cudaEventCreateWithFlags(&ev, cudaEventDisableTiming);
        if (x == cudaSuccess) x = cudaEventRecord(ev, s_run);          // the placement of this chunk is the last thing on s_run
        if (x == cudaSuccess) x = cudaStreamWaitEvent(S.s, ev, 0);
        if (ev) S.done.push_back(ev);
        siesta_matches* m = S.m;
        auto d2h = [&](void* dst, const void* src, size_t bytes) {
            if (bytes && src && x == cudaSuccess) x = cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, S.s);
        };
        d2h(m->trace_idx + S.a_tr, p.d_trace_idx, (size_t)p.n_traces * 8);
        d2h(m->occ_off + S.a_tr, p.d_occ_off, (size_t)(p.n_traces + 1) * 8);     // the tail entry is overwritten by the next part
        d2h(m->ev_off + S.a_occ, p.d_ev_off, (size_t)(p.n_occurrences + 1) * 8);
        d2h(m->ev_pos + S.a_ev, p.d_ev_pos, (size_t)p.n_events * 4);
        d2h(m->err_trace_idx + S.a_err, p.d_err_trace_idx, (size_t)p.n_ref_errors * 8);
        d2h(m->unsupported_trace_idx + S.a_unsup, p.d_unsupported_trace_idx, (size_t)p.n_unsupported * 8);
        if (all_cols) {
            d2h(m->ev_rank + S.a_ev, p.d_ev_rank, (size_t)p.n_events * 4);
            d2h(m->ev_act + S.a_ev, p.d_ev_act, (size_t)p.n_events * 4);
            d2h(m->ev_ts_ms + S.a_ev, p.d_ev_ts_ms, (size_t)p.n_events * 8);
        }
        if (x != cudaSuccess) {
            cudaGetLastError();
            S.on = false;
            return;
        }
        S.a_tr += p.n_traces;
        S.a_occ += p.n_occurrences;
        S.a_ev += p.n_events;
        S.a_err += p.n_ref_errors;
        S.a_unsup += p.n_unsupported;
        if (p.n_matches_emitted < 0) S.counted = false;
        else S.n_emit += p.n_matches_emitted;
        S.k_ms += p.kernel_ms;
        S.d_ms += p.detect_ms;
    };
    RebaseOffsets base{0, 0, 0};
    bool ids_valid = true;
    for (int k = 0; k < n_chunks && e == cudaSuccess && rc == SIESTA_OK; ++k) {
        while (enq < n_chunks && enq <= k + 2 && rc == SIESTA_OK) enqueue_copy(enq++);
        if (e != cudaSuccess || rc != SIESTA_OK) break;
        ok(cudaStreamWaitEvent(s_run, ready[k], 0));
        ok(cudaEventSynchronize(ready[k]));   // the chunk has landed (usually long ago: the copies run two chunks ahead) and h_bad[k] is final
        if (e != cudaSuccess) break;
        Log view;
        view.ctx = c;
        view.d_trace_off = d_off + cut[k];  // offsets are global event indices: act / ts_ms stay whole
        view.d_act = d_act;
        view.d_ts_ms = ts_mapped ? ts_mapped : d_ts;
        view.n_traces = cut[k + 1] - cut[k];
        // the view ends where the chunk ends: the 32-byte sector that holds the chunk's last event also holds the first
        // events of chunk k + 1, which may not have arrived yet - K1-P indexes its table with whatever a sector holds, so a
        // trace whose last sector crosses the end of the view must take the staged kernel (it does: detect_nkp.cu `fits`)
        view.n_events = trace_off[cut[k + 1]];
        view.n_activities = n_activities;
        view.max_trace_len = chunk_max_len[(size_t)k];
        view.owns = false;
        if (h_bad[k]) ids_valid = false;   // sticky: the first sector of chunk k + 1 holds the last events of chunk k
        view.act_valid = ids_valid;
        siesta_dev_matches dm;
        base.trace = cut[k];
        rc = detect_device_impl(&view, nfa, nullptr, 0, flags, s_run, base, &dm);
        if (rc) break;
        parts.push_back(dm);
        stream_part(dm);
        base.occ += dm.n_occurrences;
        base.ev += dm.n_events;
    }
    if (e != cudaSuccess && rc == SIESTA_OK) {
        set_error(std::string("siesta_evaluate_events: ") + cudaGetErrorString(e));
        rc = SIESTA_E_CUDA;
    }
    if (S.s) {
        if (cudaStreamSynchronize(S.s) != cudaSuccess) {
            cudaGetLastError();
            S.on = false;
        }
        for (cudaEvent_t ev : S.done) cudaEventDestroy(ev);
        cudaStreamDestroy(S.s);
    }
    if (rc == SIESTA_OK && S.on && S.a_tr == base.occ) {   // (one occurrence per matching trace: the two counts agree)
        siesta_matches* m = S.m;
        m->n_traces = S.a_tr;
        m->n_occurrences = S.a_occ;
        m->n_events = S.a_ev;
        m->n_ref_errors = S.a_err;
        m->n_unsupported = S.a_unsup;
        m->n_matches_emitted = S.counted ? S.n_emit : -1;
        m->kernel_ms = S.k_ms;
        m->detect_ms = S.d_ms;
        m->occ_off[S.a_tr] = S.a_occ;
        m->ev_off[S.a_occ] = S.a_ev;
        *out = m;
    } else {
        if (S.base) {   // not streamed after all: the block goes back, the parts are assembled as a whole
            reinterpret_cast<ResultHeader*>(S.base)->magic = 0;
            arena_free(c, S.base);
        }
        if (rc == SIESTA_OK) rc = assemble_matches(c, parts, flags, s_run, out, nullptr);
    }
    for (siesta_dev_matches& p : parts) siesta_dev_matches_free(&p);
    cudaStreamSynchronize(s_copy);
    cudaStreamSynchronize(s_prep);
    cudaStreamSynchronize(s_run);
    if (d_off) cudaFreeAsync(d_off, s_run);
    if (d_act) cudaFreeAsync(d_act, s_run);
    if (d_act8) cudaFreeAsync(d_act8, s_run);
    if (d_ts) cudaFreeAsync(d_ts, s_run);
    if (d_bad) cudaFreeAsync(d_bad, s_run);
    for (cudaEvent_t ev : ready)
        if (ev) cudaEventDestroy(ev);
    for (cudaEvent_t ev : copied)
        if (ev) cudaEventDestroy(ev);
    cudaStreamSynchronize(s_run);
    arena_free(c, h_bad);
    cudaStreamDestroy(s_copy);
    cudaStreamDestroy(s_prep);
    cudaStreamDestroy(s_run);
    return rc;
}

extern "C" int siesta_evaluate_events(siesta_ctx* ctx, const int64_t* trace_off, const int32_t* act, const int64_t* ts_ms,
                                      int64_t n_traces, int64_t n_events, int32_t n_activities, const siesta_nfa* nfa,
                                      uint32_t flags, siesta_matches** out) {
    return evaluate_events_impl(ctx, trace_off, act, nullptr, ts_ms, n_traces, n_events, n_activities, nfa, flags, out);
}

// The same request with the activity column as ONE BYTE per event (alphabets of at most 256 activities - every log the
// reference's papers use): the call runs at the speed of the host link, and the activity column is what crosses it, so the
// caller's serialiser (the JNI shim filling a direct buffer from List<Event>) writes the dictionary id into a byte.
extern "C" int siesta_evaluate_events_act8(siesta_ctx* ctx, const int64_t* trace_off, const uint8_t* act8, const int64_t* ts_ms,
                                           int64_t n_traces, int64_t n_events, int32_t n_activities, const siesta_nfa* nfa,
                                           uint32_t flags, siesta_matches** out) {
    if (!act8 && n_events) {
        set_error("siesta_evaluate_events_act8: null argument");
        return SIESTA_E_INVALID;
    }
    return evaluate_events_impl(ctx, trace_off, nullptr, act8 ? act8 : reinterpret_cast<const uint8_t*>(""), ts_ms, n_traces,
                                n_events, n_activities, nfa, flags, out);
}
