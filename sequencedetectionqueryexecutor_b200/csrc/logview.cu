// logview.cu — derived logs built ON THE DEVICE from a resident CSR log: the from / till window and the group streams.
//
//   siesta_log_filter_time  Trace.filter(from, till) (J/model/DBModel/Trace.java:25-29; the same test in
//                           SparkDatabaseRepository.addFilterIds :204-218 and Utils.evaluateEvent, Utils.java:67-79):
//                           an event stays iff from <= timestamp <= till, either bound optional.  The reference filters
//                           the event list BEFORE Utils.transformToSaseEvents numbers it, so in the derived log the
//                           index of an event in its trace is its rank among the surviving events; `src_event` keeps the
//                           event's index in the source log.
//   siesta_log_group        the streams of /detection over groups of traces (SparkDatabaseRepository.querySingleTableGroups
//                           :307-336): a trace belongs to the FIRST group that lists it; a group's stream is the events of
//                           its traces merged by timestamp; groups whose events do not cover every event type of the query
//                           are dropped.  The reference sorts a Spark iterable of undefined order with a stable sort, so
//                           ties between equal timestamps are undefined there; here ties go by (order of the trace in the
//                           group's list, position in the trace).  Traces must be sorted by timestamp (the log's invariant).
// Both return a new resident log that owns its memory; every kernel of the library runs on it unchanged
// (SaseConnector.evaluateGroups, SaseConnector.java:85-110, is then siesta_detect on the group log).
#include <algorithm>
#include <cstring>
#include <vector>

#include "common.cuh"

namespace siesta {

constexpr int LV_T = 256;

// exclusive scan of int64 counts in three small kernels (chunks of LV_T, chunk totals, add back)
__global__ void __launch_bounds__(LV_T) lv_scan_chunks_kernel(int64_t* v, int64_t n, int64_t* totals) {
    __shared__ int64_t ws[LV_T / 32];
    const int64_t i = (int64_t)blockIdx.x * LV_T + threadIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t x = i < n ? v[i] : 0;
    int64_t inc = x;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int64_t y = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= d) inc += y;
    }
    if (lane == 31) ws[warp] = inc;
    __syncthreads();
    int64_t before = 0, all = 0;
    for (int k = 0; k < LV_T / 32; ++k) {
        if (k < warp) before += ws[k];
        all += ws[k];
    }
    if (i < n) v[i] = before + inc - x;
    if (threadIdx.x == 0) totals[blockIdx.x] = all;
}
__global__ void lv_scan_totals_kernel(int64_t* totals, int64_t n_chunks, int64_t* grand) {
    int64_t run = 0;   // one thread: n_chunks = n / 256 values, and this is a load-time step
    for (int64_t c = 0; c < n_chunks; ++c) {
        const int64_t t = totals[c];
        totals[c] = run;
        run += t;
    }
    *grand = run;
}
__global__ void __launch_bounds__(LV_T) lv_scan_add_kernel(int64_t* v, int64_t n, const int64_t* totals, const int64_t* grand) {
    const int64_t i = (int64_t)blockIdx.x * LV_T + threadIdx.x;
    if (i < n) v[i] += totals[blockIdx.x];
    if (i == n) v[n] = *grand;   // the CSR's closing offset
}

static int lv_exclusive_scan(int64_t* d_v, int64_t n, cudaStream_t stream, int64_t* h_total) {
    const int64_t n_chunks = (n + 1 + LV_T - 1) / LV_T;   // one extra element: the closing offset
    int64_t* d_tot = nullptr;
    SIESTA_CUDA_OK(cudaMallocAsync((void**)&d_tot, (size_t)(n_chunks + 1) * 8, stream));
    lv_scan_chunks_kernel<<<(unsigned)n_chunks, LV_T, 0, stream>>>(d_v, n, d_tot);
    SIESTA_LAUNCHED();
    lv_scan_totals_kernel<<<1, 1, 0, stream>>>(d_tot, n_chunks, d_tot + n_chunks);
    SIESTA_LAUNCHED();
    lv_scan_add_kernel<<<(unsigned)n_chunks, LV_T, 0, stream>>>(d_v, n, d_tot, d_tot + n_chunks);
    SIESTA_LAUNCHED();
    SIESTA_CUDA_OK(cudaMemcpyAsync(h_total, d_tot + n_chunks, 8, cudaMemcpyDeviceToHost, stream));
    SIESTA_CUDA_OK(cudaStreamSynchronize(stream));
    cudaFreeAsync(d_tot, stream);
    return SIESTA_OK;
}

// ------------------------------------------------------------------------------------------------ from / till
__device__ __forceinline__ bool lv_keep(long long ts, long long from, long long till, int has_from, int has_till) {
    return (!has_from || ts >= from) && (!has_till || ts <= till);
}
// one warp per trace: lanes stride over the events (coalesced)
__global__ void __launch_bounds__(LV_T) lv_count_kept_kernel(const int64_t* off, const int64_t* ts, int64_t n_traces, long long from,
                                                             long long till, int has_from, int has_till, int64_t* cnt) {
    const int lane = threadIdx.x & 31;
    for (int64_t t = ((int64_t)blockIdx.x * LV_T + threadIdx.x) >> 5; t < n_traces; t += ((int64_t)gridDim.x * LV_T) >> 5) {
        const int64_t o0 = off[t], o1 = off[t + 1];
        int c = 0;
        for (int64_t e = o0 + lane; e < o1; e += 32) c += lv_keep(ts[e], from, till, has_from, has_till);
#pragma unroll
        for (int d = 16; d; d >>= 1) c += __shfl_xor_sync(0xffffffffu, c, d);
        if (lane == 0) cnt[t] = c;
    }
}
__global__ void __launch_bounds__(LV_T) lv_copy_kept_kernel(const int64_t* off, const int32_t* act, const int64_t* ts, int64_t n_traces,
                                                            long long from, long long till, int has_from, int has_till,
                                                            const int64_t* new_off, int32_t* o_act, int64_t* o_ts, int64_t* o_src) {
    const int lane = threadIdx.x & 31;
    for (int64_t t = ((int64_t)blockIdx.x * LV_T + threadIdx.x) >> 5; t < n_traces; t += ((int64_t)gridDim.x * LV_T) >> 5) {
        const int64_t o0 = off[t], o1 = off[t + 1];
        int64_t at = new_off[t];
        for (int64_t e0 = o0; e0 < o1; e0 += 32) {
            const int64_t e = e0 + lane;
            const long long v = e < o1 ? ts[e] : 0;
            const bool k = e < o1 && lv_keep(v, from, till, has_from, has_till);
            const unsigned m = __ballot_sync(0xffffffffu, k);
            if (k) {
                const int64_t to = at + __popc(m & ((1u << lane) - 1u));
                o_act[to] = act[e];
                o_ts[to] = v;
                o_src[to] = e;
            }
            at += __popc(m);
        }
    }
}

// ------------------------------------------------------------------------------------------------ groups
// first group that lists the trace (SparkDatabaseRepository.java:315-318), -1 = none; one thread per (group, member)
__global__ void lv_first_group_kernel(const int64_t* g_off, const int64_t* g_tr, int32_t n_groups, int64_t n_traces, int* owner, int* bad) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < g_off[n_groups]; i += (int64_t)gridDim.x * blockDim.x) {
        int lo = 0, hi = n_groups;   // group of member slot i
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (g_off[mid] <= i) lo = mid; else hi = mid;
        }
        const int64_t t = g_tr[i];
        if (t < 0 || t >= n_traces) { *bad = 1; continue; }
        atomicMin(owner + t, lo);
    }
}
// per group: events of the members it owns, and does it cover every event type of the query?  one warp per group
__global__ void __launch_bounds__(LV_T) lv_group_sizes_kernel(const int64_t* g_off, const int64_t* g_tr, int32_t n_groups, const int* owner,
                                                              const int64_t* off, const int32_t* act, const int32_t* types, int32_t n_types,
                                                              int64_t* cnt /* [n_groups]: events, 0 if dropped */, int* kept) {
    const int lane = threadIdx.x & 31;
    for (int g = (int)(((int64_t)blockIdx.x * LV_T + threadIdx.x) >> 5); g < n_groups; g += (int)(((int64_t)gridDim.x * LV_T) >> 5)) {
        unsigned long long seen = 0;   // bit i <=> types[i] occurs (n_types <= 64)
        int64_t n = 0;
        for (int64_t i = g_off[g]; i < g_off[g + 1]; ++i) {
            const int64_t t = g_tr[i];
            if (owner[t] != g) continue;
            bool dup = false;          // a trace listed twice in the same group counts once
            for (int64_t k = g_off[g]; k < i && !dup; ++k) dup = g_tr[k] == t;
            if (dup) continue;
            const int64_t o0 = off[t], o1 = off[t + 1];
            n += o1 - o0;
            for (int64_t e = o0 + lane; e < o1; e += 32) {
                const int a = act[e];
                for (int q = 0; q < n_types; ++q)
                    if (types[q] == a) seen |= 1ull << q;
            }
        }
#pragma unroll
        for (int d = 16; d; d >>= 1) seen |= __shfl_xor_sync(0xffffffffu, seen, d);
        const bool ok = n > 0 && seen == (n_types == 64 ? ~0ull : ((1ull << n_types) - 1ull));
        if (lane == 0) {
            cnt[g] = ok ? n : 0;
            kept[g] = ok ? 1 : 0;
        }
    }
}
// merge: every event finds its rank in the group's stream by counting, in every member trace, the events that sort before
// it - (timestamp, order of the trace in the group, position) - with one binary search per member (traces are sorted by
// timestamp).  One warp per group, lanes stride over the group's events.
__global__ void __launch_bounds__(LV_T) lv_group_merge_kernel(const int64_t* g_off, const int64_t* g_tr, int32_t n_groups, const int* owner,
                                                              const int* kept, const int64_t* slot /* [n_groups]: output trace of the group */,
                                                              const int64_t* off, const int32_t* act, const int64_t* ts,
                                                              const int64_t* new_off, int32_t* o_act, int64_t* o_ts, int64_t* o_src) {
    const int lane = threadIdx.x & 31;
    for (int g = (int)(((int64_t)blockIdx.x * LV_T + threadIdx.x) >> 5); g < n_groups; g += (int)(((int64_t)gridDim.x * LV_T) >> 5)) {
        if (!kept[g]) continue;
        const int64_t base = new_off[slot[g]];
        const int64_t m0 = g_off[g], m1 = g_off[g + 1];
        for (int64_t i = m0; i < m1; ++i) {
            const int64_t t = g_tr[i];
            if (owner[t] != g) continue;
            bool dup = false;
            for (int64_t k = m0; k < i && !dup; ++k) dup = g_tr[k] == t;
            if (dup) continue;
            const int64_t o0 = off[t], o1 = off[t + 1];
            for (int64_t e = o0 + lane; e < o1; e += 32) {
                const long long v = ts[e];
                int64_t rank = e - o0;                      // the events of its own trace before it
                for (int64_t j = m0; j < m1; ++j) {
                    const int64_t u = g_tr[j];
                    if (j == i || owner[u] != g) continue;
                    bool dup2 = false;
                    for (int64_t k = m0; k < j && !dup2; ++k) dup2 = g_tr[k] == u;
                    if (dup2) continue;
                    // events of trace u that sort before e: ts < v, or ts == v when u comes earlier in the group's list
                    int64_t lo = off[u], hi = off[u + 1];
                    const bool earlier = j < i;
                    while (lo < hi) {
                        const int64_t mid = (lo + hi) >> 1;
                        const long long w = ts[mid];
                        if (w < v || (earlier && w == v)) lo = mid + 1; else hi = mid;
                    }
                    rank += lo - off[u];
                }
                o_act[base + rank] = act[e];
                o_ts[base + rank] = v;
                o_src[base + rank] = e;
            }
        }
    }
}
__global__ void lv_pick_kept_kernel(const int* kept, const int64_t* slot, int32_t n_groups, const int64_t* cnt, int64_t* out_cnt, int32_t* out_gid) {
    for (int g = blockIdx.x * blockDim.x + threadIdx.x; g < n_groups; g += gridDim.x * blockDim.x)
        if (kept[g]) {
            out_cnt[slot[g]] = cnt[g];
            out_gid[slot[g]] = g + 1;   // the reference numbers groups from 1 (SparkDatabaseRepository.java:317)
        }
}

static Log* new_owned_log(Ctx* c, int64_t n_traces, int64_t n_events, int32_t n_act, int64_t* d_off, int32_t* d_act, int64_t* d_ts, int64_t* d_src) {
    Log* L = new Log();
    L->ctx = c;
    L->d_trace_off = d_off;
    L->d_act = d_act;
    L->d_ts_ms = d_ts;
    L->d_src_event = d_src;
    L->n_traces = n_traces;
    L->n_events = n_events;
    L->n_activities = n_act;
    L->max_trace_len = 0;
    L->owns = true;
    return L;
}

}  // namespace siesta

using namespace siesta;

extern "C" int siesta_log_filter_time(siesta_log* log, int64_t from_ms, int32_t has_from, int64_t till_ms, int32_t has_till, siesta_log** out) {
    Log* S = reinterpret_cast<Log*>(log);
    if (!S || !out) {
        set_error("siesta_log_filter_time: null argument");
        return SIESTA_E_INVALID;
    }
    Ctx* c = S->ctx;
    SIESTA_CUDA_OK(cudaSetDevice(c->device));
    cudaStream_t stream = c->stream;
    const int64_t T = S->n_traces;
    int64_t* d_off = nullptr;
    SIESTA_CUDA_OK(cudaMalloc((void**)&d_off, (size_t)(T + 1) * 8 + 16));
    SIESTA_CUDA_OK(cudaMemsetAsync(d_off, 0, (size_t)(T + 1) * 8, stream));
    const int grid = (int)std::min<int64_t>((T * 32 + LV_T - 1) / LV_T + 1, (int64_t)c->sm_count * 16);
    if (T > 0) {
        lv_count_kept_kernel<<<grid, LV_T, 0, stream>>>(S->d_trace_off, S->d_ts_ms, T, from_ms, till_ms, has_from, has_till, d_off);
        SIESTA_LAUNCHED();
    }
    int64_t E2 = 0;
    int rc = lv_exclusive_scan(d_off, T, stream, &E2);
    if (rc) {
        cudaFree(d_off);
        return rc;
    }
    int32_t* d_act = nullptr;
    int64_t *d_ts = nullptr, *d_src = nullptr;
    const size_t ne = (size_t)(E2 ? E2 : 1);
    cudaError_t e;
    if ((e = cudaMalloc((void**)&d_act, ne * 4 + 32)) != cudaSuccess || (e = cudaMalloc((void**)&d_ts, ne * 8 + 32)) != cudaSuccess ||
        (e = cudaMalloc((void**)&d_src, ne * 8 + 32)) != cudaSuccess) {
        set_error(std::string("siesta_log_filter_time: cudaMalloc: ") + cudaGetErrorString(e));
        cudaFree(d_off); cudaFree(d_act); cudaFree(d_ts); cudaFree(d_src);
        return SIESTA_E_NOMEM;
    }
    if (T > 0) {
        lv_copy_kept_kernel<<<grid, LV_T, 0, stream>>>(S->d_trace_off, S->d_act, S->d_ts_ms, T, from_ms, till_ms, has_from, has_till, d_off,
                                                      d_act, d_ts, d_src);
        SIESTA_LAUNCHED();
    }
    SIESTA_CUDA_OK(cudaStreamSynchronize(stream));
    Log* L = new_owned_log(c, T, E2, S->n_activities, d_off, d_act, d_ts, d_src);
    L->first_trace = S->first_trace;
    L->act_valid = S->act_valid;     // a subset of validated ids
    *out = reinterpret_cast<siesta_log*>(L);
    return SIESTA_OK;
}

extern "C" int siesta_log_group(siesta_log* log, const int64_t* group_off, const int64_t* group_traces, int32_t n_groups,
                                const int32_t* types, int32_t n_types, siesta_log** out, int32_t* group_ids, int32_t* n_kept) {
    Log* S = reinterpret_cast<Log*>(log);
    if (!S || !out || !group_off || (!group_traces && n_groups && group_off[n_groups]) || n_groups < 0 || !group_ids || !n_kept ||
        n_types < 0 || n_types > 64 || (n_types && !types)) {
        set_error("siesta_log_group: bad argument (at most 64 event types)");
        return SIESTA_E_INVALID;
    }
    for (int g = 0; g < n_groups; ++g)
        if (group_off[g + 1] < group_off[g]) {
            set_error("siesta_log_group: group_off must be non-decreasing");
            return SIESTA_E_INVALID;
        }
    Ctx* c = S->ctx;
    SIESTA_CUDA_OK(cudaSetDevice(c->device));
    cudaStream_t stream = c->stream;
    const int64_t n_members = n_groups ? group_off[n_groups] : 0;
    const size_t ng = (size_t)std::max(n_groups, 1);
    // scratch: group lists, owner per trace, per-group counts / kept / slot, bad flag
    int64_t *d_goff = nullptr, *d_gtr = nullptr, *d_cnt = nullptr, *d_slot = nullptr;
    int *d_owner = nullptr, *d_kept = nullptr, *d_bad = nullptr;
    int32_t* d_types = nullptr;
    struct Free {
        std::vector<void*> p;
        cudaStream_t s;
        ~Free() { for (void* q : p) if (q) cudaFreeAsync(q, s); }
    } scratch{{}, stream};
    auto alloc = [&](void** p, size_t bytes) {
        const cudaError_t e = cudaMallocAsync(p, bytes ? bytes : 16, stream);
        if (e == cudaSuccess) scratch.p.push_back(*p);
        return e;
    };
    SIESTA_CUDA_OK(alloc((void**)&d_goff, (ng + 1) * 8));
    SIESTA_CUDA_OK(alloc((void**)&d_gtr, (size_t)std::max<int64_t>(n_members, 1) * 8));
    SIESTA_CUDA_OK(alloc((void**)&d_cnt, (ng + 1) * 8));
    SIESTA_CUDA_OK(alloc((void**)&d_slot, (ng + 1) * 8));
    SIESTA_CUDA_OK(alloc((void**)&d_owner, (size_t)std::max<int64_t>(S->n_traces, 1) * 4));
    SIESTA_CUDA_OK(alloc((void**)&d_kept, ng * 4));
    SIESTA_CUDA_OK(alloc((void**)&d_bad, 4));
    SIESTA_CUDA_OK(alloc((void**)&d_types, (size_t)std::max(n_types, 1) * 4));
    SIESTA_CUDA_OK(cudaMemcpyAsync(d_goff, group_off, (size_t)(n_groups + 1) * 8, cudaMemcpyHostToDevice, stream));
    if (n_members) SIESTA_CUDA_OK(cudaMemcpyAsync(d_gtr, group_traces, (size_t)n_members * 8, cudaMemcpyHostToDevice, stream));
    if (n_types) SIESTA_CUDA_OK(cudaMemcpyAsync(d_types, types, (size_t)n_types * 4, cudaMemcpyHostToDevice, stream));
    SIESTA_CUDA_OK(cudaMemsetAsync(d_owner, 0x7f, (size_t)std::max<int64_t>(S->n_traces, 1) * 4, stream));   // 0x7f7f7f7f = no group
    SIESTA_CUDA_OK(cudaMemsetAsync(d_bad, 0, 4, stream));
    const int gw = (int)std::min<int64_t>(((int64_t)n_groups * 32 + LV_T - 1) / LV_T + 1, (int64_t)c->sm_count * 16);
    std::vector<int> h_kept((size_t)ng, 0);
    if (n_groups > 0) {
        lv_first_group_kernel<<<(int)std::min<int64_t>((n_members + 255) / 256 + 1, 4096), 256, 0, stream>>>(d_goff, d_gtr, n_groups, S->n_traces, d_owner, d_bad);
        SIESTA_LAUNCHED();
        lv_group_sizes_kernel<<<gw, LV_T, 0, stream>>>(d_goff, d_gtr, n_groups, d_owner, S->d_trace_off, S->d_act, d_types, n_types, d_cnt, d_kept);
        SIESTA_LAUNCHED();
    }
    int bad = 0;
    SIESTA_CUDA_OK(cudaMemcpyAsync(&bad, d_bad, 4, cudaMemcpyDeviceToHost, stream));
    SIESTA_CUDA_OK(cudaMemcpyAsync(h_kept.data(), d_kept, (size_t)n_groups * 4, cudaMemcpyDeviceToHost, stream));
    SIESTA_CUDA_OK(cudaStreamSynchronize(stream));
    if (bad) {
        set_error("siesta_log_group: trace index out of range");
        return SIESTA_E_INVALID;
    }
    // the kept groups become the traces of the new log, in group order
    std::vector<int64_t> h_slot((size_t)ng, 0);
    int32_t K = 0;
    for (int g = 0; g < n_groups; ++g) {
        h_slot[(size_t)g] = K;
        if (h_kept[(size_t)g]) ++K;
    }
    SIESTA_CUDA_OK(cudaMemcpyAsync(d_slot, h_slot.data(), (size_t)ng * 8, cudaMemcpyHostToDevice, stream));
    int64_t* d_off = nullptr;
    int32_t* d_gid = nullptr;
    SIESTA_CUDA_OK(cudaMalloc((void**)&d_off, (size_t)(K + 1) * 8 + 16));
    SIESTA_CUDA_OK(alloc((void**)&d_gid, (size_t)std::max(K, 1) * 4));
    SIESTA_CUDA_OK(cudaMemsetAsync(d_off, 0, (size_t)(K + 1) * 8, stream));
    if (n_groups > 0) {
        lv_pick_kept_kernel<<<(n_groups + 255) / 256, 256, 0, stream>>>(d_kept, d_slot, n_groups, d_cnt, d_off, d_gid);
        SIESTA_LAUNCHED();
    }
    int64_t E2 = 0;
    int rc = lv_exclusive_scan(d_off, K, stream, &E2);
    if (rc) {
        cudaFree(d_off);
        return rc;
    }
    int32_t* d_act = nullptr;
    int64_t *d_ts = nullptr, *d_src = nullptr;
    const size_t ne = (size_t)(E2 ? E2 : 1);
    cudaError_t e;
    if ((e = cudaMalloc((void**)&d_act, ne * 4 + 32)) != cudaSuccess || (e = cudaMalloc((void**)&d_ts, ne * 8 + 32)) != cudaSuccess ||
        (e = cudaMalloc((void**)&d_src, ne * 8 + 32)) != cudaSuccess) {
        set_error(std::string("siesta_log_group: cudaMalloc: ") + cudaGetErrorString(e));
        cudaFree(d_off); cudaFree(d_act); cudaFree(d_ts); cudaFree(d_src);
        return SIESTA_E_NOMEM;
    }
    if (K > 0) {
        lv_group_merge_kernel<<<gw, LV_T, 0, stream>>>(d_goff, d_gtr, n_groups, d_owner, d_kept, d_slot, S->d_trace_off, S->d_act, S->d_ts_ms,
                                                      d_off, d_act, d_ts, d_src);
        SIESTA_LAUNCHED();
        SIESTA_CUDA_OK(cudaMemcpyAsync(group_ids, d_gid, (size_t)K * 4, cudaMemcpyDeviceToHost, stream));
    }
    SIESTA_CUDA_OK(cudaStreamSynchronize(stream));
    *n_kept = K;
    Log* L = new_owned_log(c, K, E2, S->n_activities, d_off, d_act, d_ts, d_src);
    L->act_valid = S->act_valid;
    *out = reinterpret_cast<siesta_log*>(L);
    return SIESTA_OK;
}

extern "C" int siesta_log_source_events(siesta_log* log, int64_t* out /* [n_events] */) {
    Log* L = reinterpret_cast<Log*>(log);
    if (!L || !out) {
        set_error("siesta_log_source_events: null argument");
        return SIESTA_E_INVALID;
    }
    if (!L->d_src_event) {
        set_error("siesta_log_source_events: not a derived log (siesta_log_filter_time / siesta_log_group)");
        return SIESTA_E_INVALID;
    }
    SIESTA_CUDA_OK(cudaSetDevice(L->ctx->device));
    if (L->n_events) SIESTA_CUDA_OK(cudaMemcpy(out, L->d_src_event, (size_t)L->n_events * 8, cudaMemcpyDeviceToHost));
    return SIESTA_OK;
}
