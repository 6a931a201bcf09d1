// intersect.cu — kernel K2: the pair index over a resident CSR log and the sorted trace-id intersection.
//
// Replaces SparkDatabaseRepository.getCommonIds (J/storage/repositories/SparkDatabaseRepository.java:160-178):
// "the traces that contain ALL true pairs" = intersection of the per-pair posting lists, each an ascending,
// duplicate-free list of dense trace indices.  Posting lists either come from the caller (index.parquet, loaded
// with siesta_index_load) or are derived on the GPU from the CSR log under the SeqTable view (siesta_index_build):
// a trace is listed under (A,B) iff it holds an A before a B (A == B: at least two occurrences).
//
// Kernels
//   pair_flags_kernel   one warp per trace: first / last position and count of the (<= 32) activities the requested
//                       pairs mention, then one lane per pair writes flag[pair][trace]
//   list_hits_kernel    one thread per element of every list: hit[trace] += 1 (byte counters, word atomics); a trace is in
//                       the intersection of n duplicate-free lists iff its counter reaches n: every list is read
//                       exactly once, coalesced (8 B x sum |list_i|), no search
//   compaction          block counts -> one-block scan -> ordered write (ascending output, deterministic)
#include <algorithm>
#include <cstring>
#include <vector>

#include "common.cuh"

namespace siesta {

constexpr int IT = 256;
constexpr int MAX_BUILD_PAIRS = 32;

struct Index {
    Log* log = nullptr;
    int32_t n_pairs = 0;
    std::vector<int32_t> pair_a, pair_b;
    std::vector<int64_t> off;       // [n_pairs + 1] into d_lists
    int64_t* d_lists = nullptr;     // all posting lists back to back (device)
};

// ------------------------------------------------------------------------------------------------ compaction
// flags: u8 [n_rows][n]; out rows start at row_off[row]; counts per row returned in row_cnt.
// a flag counts if it is non-zero (want == 0) or equal to `want`
__device__ __forceinline__ bool flag_set(uint8_t f, int want) { return want ? f == want : f != 0; }

__global__ void __launch_bounds__(IT) flag_count_kernel(const uint8_t* flags, int64_t n, int64_t n_blk, unsigned long long* blk, int want) {
    const int row = blockIdx.y;
    const int64_t i = (int64_t)blockIdx.x * IT + threadIdx.x;
    const unsigned f = (i < n && flag_set(flags[(int64_t)row * n + i], want)) ? 1u : 0u;
    const unsigned ball = __ballot_sync(0xffffffffu, f);
    __shared__ unsigned s[IT / 32];
    if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = __popc(ball);
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned tot = 0;
        for (int k = 0; k < IT / 32; ++k) tot += s[k];
        blk[(int64_t)row * n_blk + blockIdx.x] = tot;
    }
}

// one block per row: exclusive scan of the row's block counts in place; total into row_cnt[row]
__global__ void __launch_bounds__(1024) row_scan_kernel(unsigned long long* blk, int64_t n_blk, unsigned long long* row_cnt) {
    __shared__ unsigned long long ws[32];
    __shared__ unsigned long long carry_s;
    unsigned long long* row = blk + (int64_t)blockIdx.x * n_blk;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    for (int64_t base = 0; base < n_blk; base += 1024) {
        const int64_t i = base + threadIdx.x;
        const unsigned long long v = i < n_blk ? row[i] : 0ull;
        unsigned long long inc = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            unsigned long long y = __shfl_up_sync(0xffffffffu, inc, d);
            if (lane >= d) inc += y;
        }
        if (lane == 31) ws[warp] = inc;
        __syncthreads();
        unsigned long long before = 0, all = 0;
        for (int k = 0; k < 32; ++k) {
            if (k < warp) before += ws[k];
            all += ws[k];
        }
        const unsigned long long carry = carry_s;
        if (i < n_blk) row[i] = carry + before + inc - v;
        __syncthreads();
        if (threadIdx.x == 0) carry_s = carry + all;
        __syncthreads();
    }
    if (threadIdx.x == 0) row_cnt[blockIdx.x] = carry_s;
}

// values[i] (or i itself when values == nullptr) of the flagged elements, in order, to out + row_off[row]
__global__ void __launch_bounds__(IT) flag_write_kernel(const uint8_t* flags, const int64_t* values, int64_t n, int64_t n_blk,
                                                       const unsigned long long* blk, const int64_t* row_off, int64_t* out, int want) {
    const int row = blockIdx.y;
    const int64_t i = (int64_t)blockIdx.x * IT + threadIdx.x;
    const unsigned f = (i < n && flag_set(flags[(int64_t)row * n + i], want)) ? 1u : 0u;
    const unsigned ball = __ballot_sync(0xffffffffu, f);
    __shared__ unsigned s[IT / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) s[warp] = __popc(ball);
    __syncthreads();
    if (!f) return;
    unsigned before = 0;
    for (int k = 0; k < warp; ++k) before += s[k];
    const int64_t pos = (int64_t)blk[(int64_t)row * n_blk + blockIdx.x] + before + __popc(ball & ((1u << lane) - 1u));
    out[(row_off ? row_off[row] : 0) + pos] = values ? values[i] : i;
}

// ------------------------------------------------------------------------------------------------ index build
struct PairFlagParams {
    const int64_t* trace_off;
    const int32_t* act;
    int64_t n_traces;
    const int8_t* slot_of;  // [n_act] activity -> scratch slot (0..31) or -1
    int32_t n_act;
    int32_t n_pairs;
    int8_t slot_a[MAX_BUILD_PAIRS], slot_b[MAX_BUILD_PAIRS];
    uint8_t* flags;         // [n_pairs][n_traces]
};

__global__ void __launch_bounds__(IT) pair_flags_kernel(const __grid_constant__ PairFlagParams P) {
    __shared__ uint32_t s_cnt[IT / 32][32], s_first[IT / 32][32], s_last[IT / 32][32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long warps_total = (long long)gridDim.x * (IT / 32);
    for (long long t = (long long)blockIdx.x * (IT / 32) + warp; t < P.n_traces; t += warps_total) {
        const long long lo = P.trace_off[t], hi = P.trace_off[t + 1];
        s_cnt[warp][lane] = 0;
        s_first[warp][lane] = 0xffffffffu;
        s_last[warp][lane] = 0;
        __syncwarp();
        for (long long i = lo + lane; i < hi; i += 32) {
            const int x = __ldg(P.act + i);
            const int sl = (x >= 0 && x < P.n_act) ? P.slot_of[x] : -1;
            if (sl >= 0) {
                atomicAdd(&s_cnt[warp][sl], 1u);
                atomicMin(&s_first[warp][sl], (uint32_t)(i - lo));
                atomicMax(&s_last[warp][sl], (uint32_t)(i - lo));
            }
        }
        __syncwarp();
        if (lane < P.n_pairs) {
            const int sa = P.slot_a[lane], sb = P.slot_b[lane];
            bool f = false;
            if (sa >= 0 && sb >= 0) {
                if (sa == sb) f = s_cnt[warp][sa] >= 2;
                else f = s_cnt[warp][sa] && s_cnt[warp][sb] && s_first[warp][sa] < s_last[warp][sb];
            }
            P.flags[(long long)lane * P.n_traces + t] = f ? 1 : 0;
        }
        __syncwarp();
    }
}

// ------------------------------------------------------------------------------------------------ intersection
struct HitParams {
    int32_t n_lists;
    const int64_t* lists[MAX_BUILD_PAIRS];
    int64_t start[MAX_BUILD_PAIRS + 1];  // list k owns the flat element range [start[k], start[k + 1])
    int64_t n_traces;
    uint32_t* hit;  // [ceil(n_traces / 4)] words = one byte counter per trace
};

// One thread per element of every list.  Lists are duplicate-free and n_lists <= 32, so a byte never overflows into
// its neighbour; ascending ids make the atomics of a warp fall into a few adjacent words.
__global__ void __launch_bounds__(IT) list_hits_kernel(const __grid_constant__ HitParams P) {
    const int64_t i = (int64_t)blockIdx.x * IT + threadIdx.x;
    if (i >= P.start[P.n_lists]) return;
    int k = 0;
    while (i >= P.start[k + 1]) ++k;
    const int64_t t = __ldg(P.lists[k] + (i - P.start[k]));
    if (t >= 0 && t < P.n_traces) atomicAdd(P.hit + (t >> 2), 1u << (8 * (int)(t & 3)));
}

// union over the OR-expansions: mark the traces whose counter reached `want`, and clear the counters for the next one
__global__ void __launch_bounds__(IT) mark_hits_kernel(uint8_t* hit, uint8_t* mark, int64_t n, int want) {
    const int64_t i = (int64_t)blockIdx.x * IT + threadIdx.x;
    if (i >= n) return;
    if (hit[i] == want) mark[i] = 1;
    hit[i] = 0;
}

// hit[t] = number of the given lists that hold trace t (hit must be zero on entry)
static int count_list_hits(Index* ix, const int32_t* ids, int n, uint32_t* d_hit, cudaStream_t stream) {
    HitParams P;
    std::memset(&P, 0, sizeof(P));
    P.n_lists = n;
    for (int k = 0; k < n; ++k) {
        P.lists[k] = ix->d_lists + ix->off[ids[k]];
        P.start[k + 1] = P.start[k] + (ix->off[ids[k] + 1] - ix->off[ids[k]]);
    }
    P.n_traces = ix->log->n_traces;
    P.hit = d_hit;
    if (P.start[n] > 0) {
        list_hits_kernel<<<(unsigned)((P.start[n] + IT - 1) / IT), IT, 0, stream>>>(P);
        SIESTA_LAUNCHED();
        SIESTA_CUDA_OK(cudaGetLastError());
    }
    return SIESTA_OK;
}

static int compact_rows(cudaStream_t stream, const uint8_t* d_flags, const int64_t* d_values, int64_t n, int n_rows,
                        std::vector<int64_t>& counts, int64_t** d_out_all, std::vector<int64_t>& row_off, int want = 0) {
    const int64_t n_blk = std::max<int64_t>((n + IT - 1) / IT, 1);
    unsigned long long *d_blk = nullptr, *d_cnt = nullptr;
    int64_t* d_off = nullptr;
    SIESTA_CUDA_OK(cudaMallocAsync((void**)&d_blk, sizeof(unsigned long long) * (size_t)n_blk * n_rows, stream));
    SIESTA_CUDA_OK(cudaMallocAsync((void**)&d_cnt, sizeof(unsigned long long) * (size_t)n_rows, stream));
    dim3 grid((unsigned)n_blk, (unsigned)n_rows);
    flag_count_kernel<<<grid, IT, 0, stream>>>(d_flags, n, n_blk, d_blk, want);
    SIESTA_LAUNCHED();
    row_scan_kernel<<<n_rows, 1024, 0, stream>>>(d_blk, n_blk, d_cnt);
    SIESTA_LAUNCHED();
    std::vector<unsigned long long> h_cnt((size_t)n_rows);
    SIESTA_CUDA_OK(cudaMemcpyAsync(h_cnt.data(), d_cnt, sizeof(unsigned long long) * (size_t)n_rows, cudaMemcpyDeviceToHost, stream));
    SIESTA_CUDA_OK(cudaStreamSynchronize(stream));
    counts.assign((size_t)n_rows, 0);
    row_off.assign((size_t)n_rows + 1, 0);
    for (int r = 0; r < n_rows; ++r) {
        counts[r] = (int64_t)h_cnt[r];
        row_off[r + 1] = row_off[r] + counts[r];
    }
    SIESTA_CUDA_OK(cudaMallocAsync((void**)d_out_all, sizeof(int64_t) * (size_t)std::max<int64_t>(row_off[n_rows], 1), stream));
    SIESTA_CUDA_OK(cudaMallocAsync((void**)&d_off, sizeof(int64_t) * (size_t)(n_rows + 1), stream));
    SIESTA_CUDA_OK(cudaMemcpyAsync(d_off, row_off.data(), sizeof(int64_t) * (size_t)(n_rows + 1), cudaMemcpyHostToDevice, stream));
    flag_write_kernel<<<grid, IT, 0, stream>>>(d_flags, d_values, n, n_blk, d_blk, d_off, *d_out_all, want);
    SIESTA_LAUNCHED();
    SIESTA_CUDA_OK(cudaGetLastError());
    SIESTA_CUDA_OK(cudaStreamSynchronize(stream));  // row_off (host vector) was the source of an async copy
    cudaFreeAsync(d_blk, stream);
    cudaFreeAsync(d_cnt, stream);
    cudaFreeAsync(d_off, stream);
    return SIESTA_OK;
}

}  // namespace siesta

using namespace siesta;

extern "C" int siesta_index_build(siesta_log* log, const int32_t* pair_a, const int32_t* pair_b, int32_t n_pairs,
                                  siesta_index** out) {
    if (!log || !pair_a || !pair_b || !out || n_pairs < 1) {
        set_error("siesta_index_build: bad argument");
        return SIESTA_E_INVALID;
    }
    if (n_pairs > MAX_BUILD_PAIRS) {
        set_error("siesta_index_build: at most " + std::to_string(MAX_BUILD_PAIRS) + " pairs per call");
        return SIESTA_E_UNSUPPORTED;
    }
    Log* L = reinterpret_cast<Log*>(log);
    SIESTA_CUDA_OK(cudaSetDevice(L->ctx->device));
    cudaStream_t stream = L->ctx->stream;
    // activities mentioned by the pairs -> scratch slots
    std::vector<int8_t> slot_of((size_t)std::max(1, L->n_activities), -1);
    PairFlagParams P;
    std::memset(&P, 0, sizeof(P));
    int n_slots = 0;
    auto slot = [&](int a) -> int {
        if (a < 0 || a >= L->n_activities) return -1;  // activity unknown to the log: empty list
        if (slot_of[a] < 0) slot_of[a] = (int8_t)n_slots++;
        return slot_of[a];
    };
    for (int p = 0; p < n_pairs; ++p) {
        P.slot_a[p] = (int8_t)slot(pair_a[p]);
        P.slot_b[p] = (int8_t)slot(pair_b[p]);
        if (n_slots > 32) {
            set_error("siesta_index_build: the pairs of one call may mention at most 32 distinct activities");
            return SIESTA_E_UNSUPPORTED;
        }
    }
    const int64_t T = L->n_traces;
    int8_t* d_slot = nullptr;
    uint8_t* d_flags = nullptr;
    SIESTA_CUDA_OK(cudaMallocAsync((void**)&d_slot, slot_of.size(), stream));
    SIESTA_CUDA_OK(cudaMallocAsync((void**)&d_flags, (size_t)std::max<int64_t>(T, 1) * n_pairs, stream));
    SIESTA_CUDA_OK(cudaMemcpyAsync(d_slot, slot_of.data(), slot_of.size(), cudaMemcpyHostToDevice, stream));
    P.trace_off = L->d_trace_off;
    P.act = L->d_act;
    P.n_traces = T;
    P.slot_of = d_slot;
    P.n_act = L->n_activities;
    P.n_pairs = n_pairs;
    P.flags = d_flags;
    if (T > 0) {
        const int64_t ctas = (T + IT / 32 - 1) / (IT / 32);
        const int grid = (int)std::min<int64_t>(ctas, (int64_t)L->ctx->sm_count * 8);
        pair_flags_kernel<<<grid, IT, 0, stream>>>(P);
        SIESTA_LAUNCHED();
        SIESTA_CUDA_OK(cudaGetLastError());
    }
    Index* ix = new Index();
    ix->log = L;
    ix->n_pairs = n_pairs;
    ix->pair_a.assign(pair_a, pair_a + n_pairs);
    ix->pair_b.assign(pair_b, pair_b + n_pairs);
    std::vector<int64_t> counts;
    int rc = compact_rows(stream, d_flags, nullptr, T, n_pairs, counts, &ix->d_lists, ix->off);
    cudaFreeAsync(d_slot, stream);
    cudaFreeAsync(d_flags, stream);
    if (rc) {
        delete ix;
        return rc;
    }
    *out = reinterpret_cast<siesta_index*>(ix);
    return SIESTA_OK;
}

extern "C" int siesta_index_load(siesta_log* log, int32_t n_pairs, const int32_t* pair_a, const int32_t* pair_b,
                                 const int64_t* post_off, const int64_t* trace_idx, siesta_index** out) {
    if (!log || n_pairs < 1 || !pair_a || !pair_b || !post_off || !out || post_off[0] != 0) {
        set_error("siesta_index_load: bad argument");
        return SIESTA_E_INVALID;
    }
    Log* L = reinterpret_cast<Log*>(log);
    for (int p = 0; p < n_pairs; ++p) {
        if (post_off[p + 1] < post_off[p]) {
            set_error("siesta_index_load: post_off must be non-decreasing");
            return SIESTA_E_INVALID;
        }
        for (int64_t i = post_off[p]; i < post_off[p + 1]; ++i)
            if (trace_idx[i] < 0 || trace_idx[i] >= L->n_traces || (i > post_off[p] && trace_idx[i] <= trace_idx[i - 1])) {
                set_error("siesta_index_load: posting lists must be strictly ascending trace indices of this log");
                return SIESTA_E_INVALID;
            }
    }
    SIESTA_CUDA_OK(cudaSetDevice(L->ctx->device));
    Index* ix = new Index();
    ix->log = L;
    ix->n_pairs = n_pairs;
    ix->pair_a.assign(pair_a, pair_a + n_pairs);
    ix->pair_b.assign(pair_b, pair_b + n_pairs);
    ix->off.assign(post_off, post_off + n_pairs + 1);
    const int64_t total = post_off[n_pairs];
    SIESTA_CUDA_OK(cudaMallocAsync((void**)&ix->d_lists, sizeof(int64_t) * (size_t)std::max<int64_t>(total, 1), L->ctx->stream));
    if (total) SIESTA_CUDA_OK(cudaMemcpyAsync(ix->d_lists, trace_idx, sizeof(int64_t) * (size_t)total, cudaMemcpyHostToDevice, L->ctx->stream));
    SIESTA_CUDA_OK(cudaStreamSynchronize(L->ctx->stream));
    *out = reinterpret_cast<siesta_index*>(ix);
    return SIESTA_OK;
}

extern "C" void siesta_index_free(siesta_index* index) {
    if (!index) return;
    Index* ix = reinterpret_cast<Index*>(index);
    cudaSetDevice(ix->log->ctx->device);
    if (ix->d_lists) cudaFreeAsync(ix->d_lists, ix->log->ctx->stream);
    delete ix;
}

extern "C" int64_t siesta_index_list_len(const siesta_index* index, int32_t pair) {
    const Index* ix = reinterpret_cast<const Index*>(index);
    if (!ix || pair < 0 || pair >= ix->n_pairs) return -1;
    return ix->off[pair + 1] - ix->off[pair];
}

extern "C" int siesta_index_get_list(const siesta_index* index, int32_t pair, int64_t* out, int64_t cap) {
    const Index* ix = reinterpret_cast<const Index*>(index);
    if (!ix || pair < 0 || pair >= ix->n_pairs || !out) {
        set_error("siesta_index_get_list: bad argument");
        return SIESTA_E_INVALID;
    }
    const int64_t n = ix->off[pair + 1] - ix->off[pair];
    if (cap < n) {
        set_error("siesta_index_get_list: buffer too small");
        return SIESTA_E_INVALID;
    }
    SIESTA_CUDA_OK(cudaSetDevice(ix->log->ctx->device));
    if (n) SIESTA_CUDA_OK(cudaMemcpyAsync(out, ix->d_lists + ix->off[pair], sizeof(int64_t) * (size_t)n, cudaMemcpyDeviceToHost, ix->log->ctx->stream));
    SIESTA_CUDA_OK(cudaStreamSynchronize(ix->log->ctx->stream));
    return SIESTA_OK;
}

// result stays on the device (ascending trace indices); free with siesta_device_free
namespace {
// events and stream-ordered scratch of one call, released on every return path
struct CallScratch {
    cudaStream_t s;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    std::vector<void*> bufs;
    explicit CallScratch(cudaStream_t stream) : s(stream) {}
    ~CallScratch() {
        if (e0) cudaEventDestroy(e0);
        if (e1) cudaEventDestroy(e1);
        for (void* p : bufs) cudaFreeAsync(p, s);
    }
    cudaError_t alloc(void** p, size_t bytes) {
        const cudaError_t e = cudaMallocAsync(p, bytes, s);
        if (e == cudaSuccess) bufs.push_back(*p);
        return e;
    }
};
}  // namespace

extern "C" int siesta_intersect_device(siesta_index* index, const int32_t* pair_ids, int32_t n, int64_t** d_out, int64_t* out_n,
                                       double* kernel_ms) {
    Index* ix = reinterpret_cast<Index*>(index);
    if (!ix || !pair_ids || n < 1 || n > MAX_BUILD_PAIRS || !d_out || !out_n) {
        set_error("siesta_intersect: bad argument (1.." + std::to_string(MAX_BUILD_PAIRS) + " lists)");
        return SIESTA_E_INVALID;
    }
    for (int k = 0; k < n; ++k)
        if (pair_ids[k] < 0 || pair_ids[k] >= ix->n_pairs) {
            set_error("siesta_intersect: pair id out of range");
            return SIESTA_E_INVALID;
        }
    SIESTA_CUDA_OK(cudaSetDevice(ix->log->ctx->device));
    cudaStream_t stream = ix->log->ctx->stream;
    // a list named twice would be counted twice: count each distinct list once
    std::vector<int32_t> ids(pair_ids, pair_ids + n);
    std::sort(ids.begin(), ids.end());
    ids.erase(std::unique(ids.begin(), ids.end()), ids.end());
    const int nd = (int)ids.size();
    const int64_t T = ix->log->n_traces;
    const size_t hit_bytes = (size_t)((std::max<int64_t>(T, 1) + 3) / 4) * 4;
    CallScratch cs(stream);
    cudaEvent_t& e0 = cs.e0;
    cudaEvent_t& e1 = cs.e1;
    SIESTA_CUDA_OK(cudaEventCreate(&e0));
    SIESTA_CUDA_OK(cudaEventCreate(&e1));
    SIESTA_CUDA_OK(cudaEventRecord(e0, stream));
    uint32_t* d_hit = nullptr;
    SIESTA_CUDA_OK(cs.alloc((void**)&d_hit, hit_bytes));
    SIESTA_CUDA_OK(cudaMemsetAsync(d_hit, 0, hit_bytes, stream));
    int rc = count_list_hits(ix, ids.data(), nd, d_hit, stream);
    std::vector<int64_t> counts, row_off;
    if (rc == SIESTA_OK) rc = compact_rows(stream, reinterpret_cast<const uint8_t*>(d_hit), nullptr, T, 1, counts, d_out, row_off, nd);
    SIESTA_CUDA_OK(cudaEventRecord(e1, stream));
    SIESTA_CUDA_OK(cudaStreamSynchronize(stream));
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    if (rc) return rc;
    *out_n = counts[0];
    if (kernel_ms) *kernel_ms = ms;
    return SIESTA_OK;
}

extern "C" int siesta_candidates_device(siesta_index* index, const int32_t* exp_off, const int32_t* pair_ids, int32_t n_exp,
                                        int64_t** d_out, int64_t* out_n, double* kernel_ms) {
    Index* ix = reinterpret_cast<Index*>(index);
    if (!ix || !exp_off || !pair_ids || n_exp < 1 || !d_out || !out_n) {
        set_error("siesta_candidates: bad argument");
        return SIESTA_E_INVALID;
    }
    for (int x = 0; x < n_exp; ++x) {
        const int n = exp_off[x + 1] - exp_off[x];
        if (n < 1 || n > MAX_BUILD_PAIRS) {
            set_error("siesta_candidates: every expansion needs 1.." + std::to_string(MAX_BUILD_PAIRS) + " true pairs "
                      "(patterns without a true pair take the Single plan: verify all traces)");
            return SIESTA_E_INVALID;
        }
        for (int k = exp_off[x]; k < exp_off[x + 1]; ++k)
            if (pair_ids[k] < 0 || pair_ids[k] >= ix->n_pairs) {
                set_error("siesta_candidates: pair id out of range");
                return SIESTA_E_INVALID;
            }
    }
    SIESTA_CUDA_OK(cudaSetDevice(ix->log->ctx->device));
    cudaStream_t stream = ix->log->ctx->stream;
    const int64_t T = ix->log->n_traces;
    CallScratch cs(stream);
    cudaEvent_t& e0 = cs.e0;
    cudaEvent_t& e1 = cs.e1;
    SIESTA_CUDA_OK(cudaEventCreate(&e0));
    SIESTA_CUDA_OK(cudaEventCreate(&e1));
    SIESTA_CUDA_OK(cudaEventRecord(e0, stream));
    uint8_t* d_mark = nullptr;
    uint32_t* d_hit = nullptr;
    const size_t hit_bytes = (size_t)((std::max<int64_t>(T, 1) + 3) / 4) * 4;
    SIESTA_CUDA_OK(cs.alloc((void**)&d_mark, (size_t)std::max<int64_t>(T, 1)));
    SIESTA_CUDA_OK(cudaMemsetAsync(d_mark, 0, (size_t)std::max<int64_t>(T, 1), stream));
    SIESTA_CUDA_OK(cs.alloc((void**)&d_hit, hit_bytes));
    SIESTA_CUDA_OK(cudaMemsetAsync(d_hit, 0, hit_bytes, stream));
    for (int x = 0; x < n_exp; ++x) {
        std::vector<int32_t> ids(pair_ids + exp_off[x], pair_ids + exp_off[x + 1]);
        std::sort(ids.begin(), ids.end());
        ids.erase(std::unique(ids.begin(), ids.end()), ids.end());
        int rcx = count_list_hits(ix, ids.data(), (int)ids.size(), d_hit, stream);
        if (rcx) return rcx;
        if (T > 0) {
            mark_hits_kernel<<<(unsigned)((T + IT - 1) / IT), IT, 0, stream>>>(reinterpret_cast<uint8_t*>(d_hit), d_mark, T, (int)ids.size());
            SIESTA_LAUNCHED();
            SIESTA_CUDA_OK(cudaGetLastError());
        }
    }
    std::vector<int64_t> counts, row_off;
    int rc = compact_rows(stream, d_mark, nullptr, T, 1, counts, d_out, row_off);
    SIESTA_CUDA_OK(cudaEventRecord(e1, stream));
    SIESTA_CUDA_OK(cudaStreamSynchronize(stream));
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    if (rc) return rc;
    *out_n = counts[0];
    if (kernel_ms) *kernel_ms = ms;
    return SIESTA_OK;
}

extern "C" int siesta_candidates(siesta_index* index, const int32_t* exp_off, const int32_t* pair_ids, int32_t n_exp, int64_t* out,
                                 int64_t cap, int64_t* out_n) {
    if (!out_n || (cap && !out)) {
        set_error("siesta_candidates: null output");
        return SIESTA_E_INVALID;
    }
    int64_t* d = nullptr;
    int rc = siesta_candidates_device(index, exp_off, pair_ids, n_exp, &d, out_n, nullptr);
    if (rc) return rc;
    Index* ix = reinterpret_cast<Index*>(index);
    if (*out_n > cap) {
        set_error("siesta_candidates: output buffer too small");
        rc = SIESTA_E_INVALID;
    } else if (*out_n) {
        cudaError_t e = cudaMemcpyAsync(out, d, sizeof(int64_t) * (size_t)*out_n, cudaMemcpyDeviceToHost, ix->log->ctx->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ix->log->ctx->stream);
        if (e != cudaSuccess) {
            set_error(std::string("siesta_candidates: D2H: ") + cudaGetErrorString(e));
            rc = SIESTA_E_CUDA;
        }
    }
    cudaFreeAsync(d, ix->log->ctx->stream);
    return rc;
}

extern "C" void siesta_device_free(siesta_log* log, void* d_ptr) {
    if (!log || !d_ptr) return;
    Log* L = reinterpret_cast<Log*>(log);
    cudaSetDevice(L->ctx->device);
    cudaFreeAsync(d_ptr, L->ctx->stream);
}

extern "C" int siesta_intersect(siesta_index* index, const int32_t* pair_ids, int32_t n, int64_t* out, int64_t cap, int64_t* out_n) {
    if (!out || !out_n) {
        set_error("siesta_intersect: null output");
        return SIESTA_E_INVALID;
    }
    int64_t* d = nullptr;
    int rc = siesta_intersect_device(index, pair_ids, n, &d, out_n, nullptr);
    if (rc) return rc;
    Index* ix = reinterpret_cast<Index*>(index);
    if (*out_n > cap) {
        set_error("siesta_intersect: output buffer too small");
        rc = SIESTA_E_INVALID;
    } else if (*out_n) {
        cudaError_t e = cudaMemcpyAsync(out, d, sizeof(int64_t) * (size_t)*out_n, cudaMemcpyDeviceToHost, ix->log->ctx->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ix->log->ctx->stream);
        if (e != cudaSuccess) {
            set_error(std::string("siesta_intersect: D2H: ") + cudaGetErrorString(e));
            rc = SIESTA_E_CUDA;
        }
    }
    cudaFreeAsync(d, ix->log->ctx->stream);
    return rc;
}
