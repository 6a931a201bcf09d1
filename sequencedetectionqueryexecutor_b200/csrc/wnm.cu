// wnm.cu — kernel W: why-not-match (sm_100a).  Algorithm and its derivation from the reference's engine: wnm.cuh.
//
// Replaces WhyNotMatchSASE.evaluate (J/model/WhyNotMatch/UsingSase/WhyNotMatchSASE.java:37-55) for the traces
// QueryPlanWhyNotMatch.execute (J/model/Queries/QueryPlans/Detection/QueryPlanWhyNotMatch.java:78-89) hands it.
//
// A WARP owns a trace:
//   1. lanes = events: the events of the pattern's activities are compacted by ballot into the warp's list
//      {lo, variants, primary - lo, position in the trace, state mask} (global scratch, a few KB, L1-resident);
//   2. lanes = listed events: every variant computes its index in the stably sorted uncertain stream with clamped
//      divisions over the other events (wnm_rank) and writes {value, change, position, state mask} there - the stream
//      lives in shared memory (13 B per uncertain event);
//   3. lanes = START events (uncertain events of the first state's activity with change <= k): each lane sweeps the stream
//      once per start (wnm_sweep: per state the family's latest event and the cheapest partial run, back pointers in a
//      lane-interleaved global scratch so the lanes' stores coalesce) and keeps its best (total change, reversed tuple);
//   4. the warp reduces the lanes' bests by shuffles and lane 0 writes the trace's almost-match.
// Work is handed out by an atomic counter.  Cost per trace: O(starts x stream x states) - the reference's engine
// materialises every combination of uncertain events as a run object.
#include <algorithm>
#include <cstring>
#include <vector>

#include "common.cuh"
#include "wnm.cuh"

namespace siesta {

struct WnmParams {
    const int64_t* trace_off;
    const int32_t* act;
    const int64_t* ts_ms;
    const int64_t* cand;
    int64_t n;
    int n_cap, r_cap;                // capacity of the stream / of the list of relevant events (per warp)
    long long* r_lo;                 // [warps][r_cap]
    int *r_nv, *r_d, *r_src;         // [warps][r_cap]
    unsigned char* r_mask;           // [warps][r_cap]
    unsigned short* prev;            // [warps][(m - 1) * n_cap * 32]
    int *o_status, *o_total, *o_pos, *o_val, *o_chg, *o_spos;   // per candidate (x m)
    unsigned long long* counter;     // [0] next trace, [1] longest stream seen, [2] most relevant events seen
};

// pass 0: sizes.  One thread per candidate: relevant events and uncertain events of its trace -> maxima, so the scratch
// of the evaluation is sized by the request, not by the worst case
__global__ void wnm_size_kernel(const __grid_constant__ WnmParams P, const __grid_constant__ WnmProgram W) {
    unsigned long long mx_n = 0, mx_r = 0;
    for (int64_t ci = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; ci < P.n; ci += (int64_t)gridDim.x * blockDim.x) {
        const int64_t t = P.cand ? P.cand[ci] : ci;
        const int64_t o0 = P.trace_off[t], o1 = P.trace_off[t + 1];
        unsigned long long n = 0, r = 0;
        for (int64_t i = o0; i < o1; ++i) {
            if (!wnm_state_mask(W, P.act[i])) continue;
            ++r;
            n += (unsigned long long)wnm_variants(W, wnm_primary(W, P.ts_ms ? P.ts_ms[i] : 0, i - o0));
        }
        mx_n = n > mx_n ? n : mx_n;
        mx_r = r > mx_r ? r : mx_r;
    }
    for (int d = 16; d; d >>= 1) {
        const unsigned long long a = __shfl_xor_sync(0xffffffffu, mx_n, d), b = __shfl_xor_sync(0xffffffffu, mx_r, d);
        mx_n = a > mx_n ? a : mx_n;
        mx_r = b > mx_r ? b : mx_r;
    }
    if ((threadIdx.x & 31) == 0) {
        atomicMax(P.counter + 1, mx_n);
        atomicMax(P.counter + 2, mx_r);
    }
}

constexpr int WNM_WARPS = 4;

__global__ void __launch_bounds__(WNM_WARPS * 32) wnm_kernel(const __grid_constant__ WnmParams P, const __grid_constant__ WnmProgram W) {
    extern __shared__ int s_wnm[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int64_t warp = (int64_t)blockIdx.x * WNM_WARPS + wib;
    const int m = W.m, n_cap = P.n_cap, r_cap = P.r_cap;
    // the warp's slice of shared memory: value[n_cap] change[n_cap] position[n_cap] (int) + state mask[n_cap] (byte)
    int* s_val = s_wnm + (size_t)wib * (3 * n_cap + n_cap / 4);
    int* s_chg = s_val + n_cap;
    int* s_src = s_chg + n_cap;
    unsigned char* s_mask = reinterpret_cast<unsigned char*>(s_src + n_cap);
    long long* r_lo = P.r_lo + warp * r_cap;
    int* r_nv = P.r_nv + warp * r_cap;
    int* r_d = P.r_d + warp * r_cap;
    int* r_src = P.r_src + warp * r_cap;
    unsigned char* r_mask = P.r_mask + warp * r_cap;
    unsigned short* prev = P.prev + (size_t)warp * (size_t)(m > 1 ? m - 1 : 1) * n_cap * 32 + lane;

    for (;;) {
        unsigned long long cu = 0;
        if (lane == 0) cu = atomicAdd(P.counter, 1ull);
        const long long ci = (long long)__shfl_sync(0xffffffffu, cu, 0);
        if (ci >= P.n) break;
        const int64_t t = P.cand ? P.cand[ci] : ci;
        const int64_t o0 = P.trace_off[t], o1 = P.trace_off[t + 1];
        // ---- 1. the events of the pattern's activities
        int n_rel = 0;
        long long n_str = 0;
        for (int64_t b = o0; b < o1; b += 32) {
            const int64_t i = b + lane;
            unsigned msk = 0;
            long long prim = 0;
            if (i < o1) {
                msk = wnm_state_mask(W, P.act[i]);
                if (msk) prim = wnm_primary(W, P.ts_ms ? P.ts_ms[i] : 0, i - o0);
            }
            const unsigned bal = __ballot_sync(0xffffffffu, msk != 0);
            const int at = n_rel + __popc(bal & ((1u << lane) - 1u));
            int nv = 0;
            if (msk) {
                nv = wnm_variants(W, prim);
                if (at < r_cap) {
                    const long long lo = wnm_lo(W, prim);
                    r_lo[at] = lo;
                    r_nv[at] = nv;
                    r_d[at] = (int)(prim - lo);
                    r_src[at] = (int)(i - o0);
                    r_mask[at] = (unsigned char)msk;
                }
            }
            long long tot = nv;
            for (int d = 16; d; d >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, d);
            n_str += tot;
            n_rel += __popc(bal);
        }
        if (n_rel == 0) {
            if (lane == 0) P.o_status[ci] = 0;
            continue;
        }
        if (n_rel > r_cap || n_str > n_cap) {   // beyond SIESTA_WNM_MAX_STREAM: listed for the caller
            if (lane == 0) P.o_status[ci] = 2;
            continue;
        }
        const int n = (int)n_str;
        __syncwarp();
        // ---- 2. the uncertain stream, in place of the stable sort
        for (int q = lane; q < n_rel; q += 32) {
            const long long lo = r_lo[q];
            const int nv = r_nv[q], d = r_d[q], src = r_src[q];
            const unsigned char msk = r_mask[q];
            for (int v = 0; v < nv; ++v) {
                const int idx = wnm_rank(W, r_lo, r_nv, n_rel, q, v);
                const long long sh = (long long)v * W.step;
                s_val[idx] = (int)(lo + sh);                             // (int) i, WhyNotMatchSASE.java:72
                s_chg[idx] = (int)(sh > d ? sh - d : d - sh);            // (int) Math.abs(original - i)
                s_src[idx] = src;
                s_mask[idx] = msk;
            }
        }
        __syncwarp();
        // ---- 3. one sweep per start event, starts dealt to the lanes round-robin
        int best_f = WNM_INF, best_t[SIESTA_MAX_STATES], tup[SIESTA_MAX_STATES];
        for (int j = 0; j < SIESTA_MAX_STATES; ++j) best_t[j] = tup[j] = -1;
        int n_start = 0;
        for (int s = 0; s < n; ++s) {
            if (!(s_mask[s] & 1u) || s_chg[s] > W.k) continue;   // state 0: the right activity and `change <= k`
            if ((n_start++ & 31) != lane) continue;
            const int f = wnm_sweep<unsigned short>(W, s_val, s_chg, s_mask, n, s, prev, n_cap, 32, tup);
            if (f != WNM_INF && (best_f == WNM_INF || wnm_better(m, f, tup, best_f, best_t))) {
                best_f = f;
                for (int j = 0; j < SIESTA_MAX_STATES; ++j) best_t[j] = tup[j];
            }
        }
        // ---- 4. the warp's best
        for (int d = 16; d; d >>= 1) {
            const int of = __shfl_xor_sync(0xffffffffu, best_f, d);
            int ot[SIESTA_MAX_STATES];
            for (int j = 0; j < SIESTA_MAX_STATES; ++j) ot[j] = __shfl_xor_sync(0xffffffffu, best_t[j], d);
            if (of != WNM_INF && (best_f == WNM_INF || wnm_better(m, of, ot, best_f, best_t))) {
                best_f = of;
                for (int j = 0; j < SIESTA_MAX_STATES; ++j) best_t[j] = ot[j];
            }
        }
        if (lane == 0) {
            P.o_status[ci] = best_f != WNM_INF ? 1 : 0;
            if (best_f != WNM_INF) {
                P.o_total[ci] = best_f;
                for (int j = 0; j < m; ++j) {
                    const int e = best_t[j];
                    P.o_pos[ci * m + j] = s_src[e];
                    P.o_val[ci * m + j] = s_val[e];
                    P.o_chg[ci * m + j] = s_chg[e];
                    P.o_spos[ci * m + j] = e;
                }
            }
        }
        __syncwarp();
    }
}

// The request as the reference would build it, checked: which shapes the closed form covers (wnm.cuh).
int wnm_build(const int32_t* pattern, int32_t m, const siesta_wnm_constraint* cons, int32_t n_cons, int32_t u, int32_t step, int32_t k,
              uint32_t flags, WnmProgram* W) {
    if (!pattern || m < 1 || m > SIESTA_MAX_STATES || n_cons < 0 || (n_cons && !cons) || u < 0 || step < 1 || k < 0) {
        set_error("siesta_why_not_match: pattern of 1 .. 8 events, uncertainty >= 0, step >= 1, k >= 0 expected");
        return SIESTA_E_INVALID;
    }
    if (n_cons > WNM_MAX_CONS) {
        set_error("siesta_why_not_match: more than 16 constraints");
        return SIESTA_E_UNSUPPORTED;
    }
    if (flags & ~(uint32_t)SIESTA_F_EVT_POS) {
        set_error("siesta_why_not_match: flags other than SIESTA_F_EVT_POS");
        return SIESTA_E_INVALID;
    }
    std::memset(W, 0, sizeof(*W));
    W->m = m;
    for (int j = 0; j < m; ++j) W->type[j] = pattern[j];
    W->n_cons = n_cons;
    for (int q = 0; q < n_cons; ++q) {
        const siesta_wnm_constraint& c = cons[q];
        if (c.pos_a < 0 || c.pos_a >= c.pos_b || c.pos_b >= m || c.value < 0 || (c.kind != SIESTA_WNM_GAP && c.kind != SIESTA_WNM_TIME) ||
            (c.method != SIESTA_WNM_WITHIN && c.method != SIESTA_WNM_ATLEAST)) {
            set_error("siesta_why_not_match: a constraint must name an earlier and a later event of the pattern and a value >= 0");
            return SIESTA_E_INVALID;
        }
        W->c_a[q] = c.pos_a;
        W->c_b[q] = c.pos_b;
        W->c_kind[q] = c.kind;
        W->c_method[q] = c.method;
        W->c_value[q] = c.value;
        if (c.kind == SIESTA_WNM_TIME) W->time_at[c.pos_b] = 1;
    }
    W->u = u;
    W->step = step;
    W->k = k;
    W->evt_pos = (flags & SIESTA_F_EVT_POS) ? 1 : 0;
    return SIESTA_OK;
}

}  // namespace siesta

using namespace siesta;

extern "C" int siesta_why_not_match(siesta_log* log, const int32_t* pattern, int32_t m, const siesta_wnm_constraint* cons, int32_t n_cons,
                                    int32_t uncertainty, int32_t step, int32_t k, const int64_t* cand, int64_t n_cand, uint32_t flags,
                                    siesta_almost_matches** out) {
    Log* L = reinterpret_cast<Log*>(log);
    if (!L || !out || (cand == nullptr && n_cand != 0) || n_cand < 0) {
        set_error("siesta_why_not_match: null argument");
        return SIESTA_E_INVALID;
    }
    WnmProgram W;
    int rc = wnm_build(pattern, m, cons, n_cons, uncertainty, step, k, flags, &W);
    if (rc) return rc;
    Ctx* c = L->ctx;
    SIESTA_CUDA_OK(cudaSetDevice(c->device));
    cudaStream_t stream = nullptr;
    SIESTA_CUDA_OK(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
    const int64_t n = cand ? n_cand : L->n_traces;
    for (int64_t i = 0; cand && i < n; ++i)
        if (cand[i] < 0 || cand[i] >= L->n_traces) {
            cudaStreamDestroy(stream);
            set_error("siesta_why_not_match: candidate trace index out of range");
            return SIESTA_E_INVALID;
        }
    struct Bufs {   // everything the request allocates, released on every path
        std::vector<void*> dev;
        cudaStream_t s;
        cudaEvent_t e0 = nullptr, e1 = nullptr;
        ~Bufs() {
            for (void* p : dev) cudaFreeAsync(p, s);
            if (e0) cudaEventDestroy(e0);
            if (e1) cudaEventDestroy(e1);
            cudaStreamSynchronize(s);
            cudaStreamDestroy(s);
        }
    } B{{}, stream};
    cudaError_t e = cudaSuccess;
    auto dalloc = [&](size_t bytes) -> void* {
        void* p = nullptr;
        if (e == cudaSuccess) e = cudaMallocAsync(&p, bytes ? bytes : 16, stream);
        if (p) B.dev.push_back(p);
        return p;
    };
    const size_t nn = (size_t)std::max<int64_t>(n, 1);
    WnmParams P;
    std::memset(&P, 0, sizeof(P));
    P.trace_off = L->d_trace_off;
    P.act = L->d_act;
    P.ts_ms = L->d_ts_ms;
    P.n = n;
    int64_t* d_cand = nullptr;
    if (cand) {
        d_cand = (int64_t*)dalloc(nn * 8);
        if (e == cudaSuccess && n) e = cudaMemcpyAsync(d_cand, cand, (size_t)n * 8, cudaMemcpyHostToDevice, stream);
        P.cand = d_cand;
    }
    P.counter = (unsigned long long*)dalloc(64);
    P.o_status = (int*)dalloc(nn * 4);
    P.o_total = (int*)dalloc(nn * 4);
    P.o_pos = (int*)dalloc(nn * m * 4);
    P.o_val = (int*)dalloc(nn * m * 4);
    P.o_chg = (int*)dalloc(nn * m * 4);
    P.o_spos = (int*)dalloc(nn * m * 4);
    if (e == cudaSuccess) e = cudaMemsetAsync(P.counter, 0, 64, stream);
    if (e == cudaSuccess) e = cudaEventCreate(&B.e0);
    if (e == cudaSuccess) e = cudaEventCreate(&B.e1);
    if (e != cudaSuccess) {
        set_error(std::string("siesta_why_not_match: ") + cudaGetErrorString(e));
        return SIESTA_E_CUDA;
    }
    std::vector<int> h_status(nn, 0), h_total, h_pos, h_val, h_chg, h_spos;
    float ms = 0.f;
    if (n > 0) {
        SIESTA_CUDA_OK(cudaEventRecord(B.e0, stream));
        const int g0 = (int)std::min<int64_t>((n + 127) / 128, (int64_t)c->sm_count * 8);
        wnm_size_kernel<<<g0, 128, 0, stream>>>(P, W);
        SIESTA_LAUNCHED();
        unsigned long long h_cnt[3] = {0, 0, 0};
        SIESTA_CUDA_OK(cudaMemcpyAsync(h_cnt, P.counter, 24, cudaMemcpyDeviceToHost, stream));
        SIESTA_CUDA_OK(cudaStreamSynchronize(stream));
        // capacities of this request (multiples of 32, at most SIESTA_WNM_MAX_STREAM: longer streams are listed)
        P.n_cap = (int)std::min<unsigned long long>(SIESTA_WNM_MAX_STREAM, std::max<unsigned long long>(32, (h_cnt[1] + 31) & ~31ull));
        P.r_cap = (int)std::min<unsigned long long>(SIESTA_WNM_MAX_STREAM, std::max<unsigned long long>(32, (h_cnt[2] + 31) & ~31ull));
        const size_t smem = (size_t)WNM_WARPS * (3 * (size_t)P.n_cap + P.n_cap / 4) * sizeof(int);
        SIESTA_CUDA_OK(cudaFuncSetAttribute(wnm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int per_sm = 0;
        SIESTA_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, wnm_kernel, WNM_WARPS * 32, smem));
        // resident CTAs: as many as fit, as long as the back-pointer scratch of all warps stays below 1 GB
        const size_t prev_per_warp = (size_t)std::max(1, m - 1) * P.n_cap * 32 * sizeof(unsigned short);
        const int by_scratch = (int)std::max<size_t>(1, ((size_t)1 << 30) / (prev_per_warp * WNM_WARPS * (size_t)c->sm_count));
        per_sm = std::max(1, std::min(per_sm, std::min(12, by_scratch)));
        const int grid = (int)std::min<int64_t>((n + WNM_WARPS - 1) / WNM_WARPS, (int64_t)c->sm_count * per_sm);
        const size_t warps = (size_t)grid * WNM_WARPS;
        P.r_lo = (long long*)dalloc(warps * P.r_cap * 8);
        P.r_nv = (int*)dalloc(warps * P.r_cap * 4);
        P.r_d = (int*)dalloc(warps * P.r_cap * 4);
        P.r_src = (int*)dalloc(warps * P.r_cap * 4);
        P.r_mask = (unsigned char*)dalloc(warps * P.r_cap);
        P.prev = (unsigned short*)dalloc(warps * (size_t)std::max(1, m - 1) * P.n_cap * 32 * sizeof(unsigned short));
        if (e != cudaSuccess) {
            set_error(std::string("siesta_why_not_match: ") + cudaGetErrorString(e));
            return e == cudaErrorMemoryAllocation ? SIESTA_E_NOMEM : SIESTA_E_CUDA;
        }
        wnm_kernel<<<grid, WNM_WARPS * 32, smem, stream>>>(P, W);
        SIESTA_LAUNCHED();
        SIESTA_CUDA_OK(cudaGetLastError());
        SIESTA_CUDA_OK(cudaEventRecord(B.e1, stream));
        h_total.resize(nn);
        h_pos.resize(nn * m);
        h_val.resize(nn * m);
        h_chg.resize(nn * m);
        h_spos.resize(nn * m);
        SIESTA_CUDA_OK(cudaMemcpyAsync(h_status.data(), P.o_status, (size_t)n * 4, cudaMemcpyDeviceToHost, stream));
        SIESTA_CUDA_OK(cudaMemcpyAsync(h_total.data(), P.o_total, (size_t)n * 4, cudaMemcpyDeviceToHost, stream));
        SIESTA_CUDA_OK(cudaMemcpyAsync(h_pos.data(), P.o_pos, (size_t)n * m * 4, cudaMemcpyDeviceToHost, stream));
        SIESTA_CUDA_OK(cudaMemcpyAsync(h_val.data(), P.o_val, (size_t)n * m * 4, cudaMemcpyDeviceToHost, stream));
        SIESTA_CUDA_OK(cudaMemcpyAsync(h_chg.data(), P.o_chg, (size_t)n * m * 4, cudaMemcpyDeviceToHost, stream));
        SIESTA_CUDA_OK(cudaMemcpyAsync(h_spos.data(), P.o_spos, (size_t)n * m * 4, cudaMemcpyDeviceToHost, stream));
        SIESTA_CUDA_OK(cudaStreamSynchronize(stream));
        SIESTA_CUDA_OK(cudaEventElapsedTime(&ms, B.e0, B.e1));
    }
    // ---- the response: traces with an almost-match in candidate order (AlmostMatch objects, createResponse :160-173)
    int64_t n_hit = 0, n_unsup = 0;
    for (int64_t i = 0; i < n; ++i) {
        n_hit += h_status[(size_t)i] == 1;
        n_unsup += h_status[(size_t)i] == 2;
    }
    const size_t ev = (size_t)std::max<int64_t>(n_hit, 1) * m;
    const size_t bytes = sizeof(siesta_almost_matches) + 64 + (size_t)std::max<int64_t>(n_hit, 1) * 12 + ev * 16 + (size_t)std::max<int64_t>(n_unsup, 1) * 8 + 64;
    char* base = (char*)std::malloc(bytes);
    if (!base) return SIESTA_E_NOMEM;
    siesta_almost_matches* r = reinterpret_cast<siesta_almost_matches*>(base);
    std::memset(r, 0, sizeof(*r));
    char* p = base + ((sizeof(siesta_almost_matches) + 63) & ~(size_t)63);
    r->trace_idx = reinterpret_cast<int64_t*>(p); p += (size_t)std::max<int64_t>(n_hit, 1) * 8;
    r->unsupported_trace_idx = reinterpret_cast<int64_t*>(p); p += (size_t)std::max<int64_t>(n_unsup, 1) * 8;
    r->total_change = reinterpret_cast<int32_t*>(p); p += (size_t)std::max<int64_t>(n_hit, 1) * 4;
    r->ev_pos = reinterpret_cast<int32_t*>(p); p += ev * 4;
    r->ev_value = reinterpret_cast<int32_t*>(p); p += ev * 4;
    r->ev_change = reinterpret_cast<int32_t*>(p); p += ev * 4;
    r->ev_stream_pos = reinterpret_cast<int32_t*>(p);
    r->n_traces = n_hit;
    r->n_states = m;
    r->n_unsupported = n_unsup;
    r->kernel_ms = ms;
    int64_t a = 0, b = 0;
    for (int64_t i = 0; i < n; ++i) {
        const int64_t t = (cand ? cand[i] : i) + L->first_trace;
        if (h_status[(size_t)i] == 2) r->unsupported_trace_idx[b++] = t;
        if (h_status[(size_t)i] != 1) continue;
        r->trace_idx[a] = t;
        r->total_change[a] = h_total[(size_t)i];
        for (int j = 0; j < m; ++j) {
            r->ev_pos[a * m + j] = h_pos[(size_t)i * m + j];
            r->ev_value[a * m + j] = h_val[(size_t)i * m + j];
            r->ev_change[a * m + j] = h_chg[(size_t)i * m + j];
            r->ev_stream_pos[a * m + j] = h_spos[(size_t)i * m + j];
        }
        ++a;
    }
    *out = r;
    return SIESTA_OK;
}

extern "C" void siesta_almost_matches_free(siesta_almost_matches* m) { std::free(m); }
