// stats.cu — kernel K4: pair statistics behind /stats (count, sum / min / max of durations, sum of squares).
//
// In the reference /stats is a lookup into count.parquet (QueryPlanStats.execute, J/model/Queries/QueryPlans/
// QueryPlanStats.java:43-48 -> S3Connector.getCounts, J/storage/repositories/S3/S3Connector.java:125-169; record layout
// J/model/DBModel/Count.java:12-24).  The table is written by the SIESTA preprocess component, which is NOT part of the
// reference repository: its pairing policy is unpinned here.  K4 computes the same record
// from the resident log under the policy SIESTA publishes for its index (non-overlapping skip-till-next-match pairs):
// per trace and pair (A,B), take the first A, then the first B after it, emit the pair, continue after that B
// (A == B: consecutive occurrences are paired (1,2), (3,4), ...).  duration = ts_ms(B) - ts_ms(A).
//
// One warp per trace, lanes = pairs (<= 32 per call).  The warp loads 32 events at a time (coalesced), keeps only those
// whose activity appears in some pair (ballot), and broadcasts them one by one; lane p advances pair p's two-state
// automaton.  Accumulators live in registers for the whole launch (persistent warps); sum of squares is exact
// (128-bit).  A second kernel folds the per-warp partials, so the result is deterministic.
#include <algorithm>
#include <cstring>
#include <vector>

#include "common.cuh"

namespace siesta {

constexpr int ST = 256;
constexpr int MAX_STAT_PAIRS = 32;

struct StatParams {
    const int64_t* trace_off;
    const int32_t* act;
    const int64_t* ts_ms;
    int64_t n_traces;
    int32_t n_act;
    int32_t n_pairs;
    int32_t pa[MAX_STAT_PAIRS], pb[MAX_STAT_PAIRS];
    const uint8_t* rel;        // [n_act] activity appears in some pair
    unsigned long long* part;  // [n_warps][MAX_STAT_PAIRS][6]: count, sum, min, max, sq_lo, sq_hi
};

__device__ __forceinline__ long long shfl_ll(long long v, int src) {
    int lo = __shfl_sync(0xffffffffu, (int)(v & 0xffffffffll), src);
    int hi = __shfl_sync(0xffffffffu, (int)(v >> 32), src);
    return ((long long)hi << 32) | (unsigned int)lo;
}

__global__ void __launch_bounds__(ST) pair_stats_kernel(const __grid_constant__ StatParams P) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long gw = (long long)blockIdx.x * (ST / 32) + warp;
    const long long warps_total = (long long)gridDim.x * (ST / 32);
    const int A = lane < P.n_pairs ? P.pa[lane] : -1, B = lane < P.n_pairs ? P.pb[lane] : -1;
    long long cnt = 0, sum = 0, mn = 0x7fffffffffffffffll, mx = -0x7fffffffffffffffll - 1;
    unsigned long long sq_lo = 0, sq_hi = 0;
    for (long long t = gw; t < P.n_traces; t += warps_total) {
        long long lo = 0, hi = 0;
        if (lane == 0) { lo = P.trace_off[t]; hi = P.trace_off[t + 1]; }
        lo = shfl_ll(lo, 0);
        hi = shfl_ll(hi, 0);
        bool waiting_b = false;
        long long ta = 0;
        for (long long base = lo; base < hi; base += 32) {
            const long long idx = base + lane;
            int x = -1;
            long long ts = 0;
            if (idx < hi) x = __ldg(P.act + idx);
            const bool rel = x >= 0 && x < P.n_act && P.rel[x];
            if (rel) ts = __ldg(reinterpret_cast<const long long*>(P.ts_ms) + idx);
            for (unsigned m = __ballot_sync(0xffffffffu, rel); m; m &= m - 1) {
                const int k = __ffs(m) - 1;
                const int ex = __shfl_sync(0xffffffffu, x, k);
                const long long et = shfl_ll(ts, k);
                if (waiting_b) {
                    if (ex == B) {
                        const long long d = et - ta;
                        ++cnt;
                        sum += d;
                        mn = d < mn ? d : mn;
                        mx = d > mx ? d : mx;
                        const unsigned long long ad = d < 0 ? (unsigned long long)(-d) : (unsigned long long)d;
                        const unsigned long long p_lo = ad * ad, p_hi = __umul64hi(ad, ad);
                        sq_lo += p_lo;
                        sq_hi += p_hi + (sq_lo < p_lo ? 1ull : 0ull);
                        waiting_b = false;
                    }
                } else if (ex == A) {
                    waiting_b = true;
                    ta = et;
                }
            }
        }
    }
    if (lane < P.n_pairs) {
        unsigned long long* o = P.part + ((size_t)gw * MAX_STAT_PAIRS + lane) * 6;
        o[0] = (unsigned long long)cnt;
        o[1] = (unsigned long long)sum;
        o[2] = (unsigned long long)mn;
        o[3] = (unsigned long long)mx;
        o[4] = sq_lo;
        o[5] = sq_hi;
    }
}

// one warp per pair: fold the per-warp partials; out[p] = count, sum, min, max, four 32-bit limbs of the sum of squares
__global__ void __launch_bounds__(32) pair_stats_fold_kernel(const unsigned long long* part, long long n_warps, int n_pairs, long long* out) {
    const int p = blockIdx.x, lane = threadIdx.x;
    long long cnt = 0, sum = 0, mn = 0x7fffffffffffffffll, mx = -0x7fffffffffffffffll - 1;
    unsigned long long l0 = 0, l1 = 0, l2 = 0, l3 = 0;  // 32-bit limbs accumulated in 64 bits (n_warps < 2^32)
    for (long long w = lane; w < n_warps; w += 32) {
        const unsigned long long* o = part + ((size_t)w * MAX_STAT_PAIRS + p) * 6;
        cnt += (long long)o[0];
        sum += (long long)o[1];
        mn = (long long)o[2] < mn ? (long long)o[2] : mn;
        mx = (long long)o[3] > mx ? (long long)o[3] : mx;
        l0 += o[4] & 0xffffffffull;
        l1 += o[4] >> 32;
        l2 += o[5] & 0xffffffffull;
        l3 += o[5] >> 32;
    }
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) {
        cnt += shfl_ll(cnt, lane ^ d);
        sum += shfl_ll(sum, lane ^ d);
        const long long omn = shfl_ll(mn, lane ^ d), omx = shfl_ll(mx, lane ^ d);
        mn = omn < mn ? omn : mn;
        mx = omx > mx ? omx : mx;
        l0 += (unsigned long long)shfl_ll((long long)l0, lane ^ d);
        l1 += (unsigned long long)shfl_ll((long long)l1, lane ^ d);
        l2 += (unsigned long long)shfl_ll((long long)l2, lane ^ d);
        l3 += (unsigned long long)shfl_ll((long long)l3, lane ^ d);
    }
    if (lane == 0) {
        // carry-normalise to true 32-bit limbs
        l1 += l0 >> 32; l0 &= 0xffffffffull;
        l2 += l1 >> 32; l1 &= 0xffffffffull;
        l3 += l2 >> 32; l2 &= 0xffffffffull;
        long long* o = out + (size_t)p * 8;
        o[0] = cnt; o[1] = sum; o[2] = mn; o[3] = mx;
        o[4] = (long long)l0; o[5] = (long long)l1; o[6] = (long long)l2; o[7] = (long long)l3;
    }
}

}  // namespace siesta

using namespace siesta;

static int pair_stats_chunk(siesta_log* log, const int32_t* pair_a, const int32_t* pair_b, int32_t n_pairs, int64_t* d_out, void* stream_,
                            double* kernel_ms);

// Any number of pairs: the kernel keeps one pair per lane, so a request is served in passes of MAX_STAT_PAIRS pairs.
extern "C" int siesta_pair_stats_device(siesta_log* log, const int32_t* pair_a, const int32_t* pair_b, int32_t n_pairs,
                                        int64_t* d_out, void* stream_, double* kernel_ms) {
    if (!log || !pair_a || !pair_b || n_pairs < 1 || !d_out) {
        set_error("siesta_pair_stats: null argument or no pair");
        return SIESTA_E_INVALID;
    }
    double total = 0;
    for (int32_t at = 0; at < n_pairs; at += MAX_STAT_PAIRS) {
        double ms = 0;
        const int rc = pair_stats_chunk(log, pair_a + at, pair_b + at, std::min<int32_t>(MAX_STAT_PAIRS, n_pairs - at), d_out + 8 * (int64_t)at,
                                        stream_, &ms);
        if (rc) return rc;
        total += ms;
    }
    if (kernel_ms) *kernel_ms = total;
    return SIESTA_OK;
}

static int pair_stats_chunk(siesta_log* log, const int32_t* pair_a, const int32_t* pair_b, int32_t n_pairs, int64_t* d_out, void* stream_,
                            double* kernel_ms) {
    Log* L = reinterpret_cast<Log*>(log);
    SIESTA_CUDA_OK(cudaSetDevice(L->ctx->device));
    cudaStream_t stream = stream_ ? reinterpret_cast<cudaStream_t>(stream_) : L->ctx->stream;
    StatParams P;
    std::memset(&P, 0, sizeof(P));
    P.trace_off = L->d_trace_off;
    P.act = L->d_act;
    P.ts_ms = L->d_ts_ms;
    P.n_traces = L->n_traces;
    P.n_act = L->n_activities;
    P.n_pairs = n_pairs;
    std::vector<uint8_t> rel((size_t)std::max(1, L->n_activities), 0);
    for (int p = 0; p < n_pairs; ++p) {
        P.pa[p] = pair_a[p];
        P.pb[p] = pair_b[p];
        if (pair_a[p] >= 0 && pair_a[p] < L->n_activities) rel[pair_a[p]] = 1;
        if (pair_b[p] >= 0 && pair_b[p] < L->n_activities) rel[pair_b[p]] = 1;
    }
    int per_sm = 0;
    SIESTA_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, pair_stats_kernel, ST, 0));
    if (per_sm < 1) per_sm = 1;
    const int64_t ctas_needed = std::max<int64_t>((L->n_traces + ST / 32 - 1) / (ST / 32), 1);
    const int grid = (int)std::min<int64_t>(ctas_needed, (int64_t)L->ctx->sm_count * per_sm);
    const long long n_warps = (long long)grid * (ST / 32);
    uint8_t* d_rel = nullptr;
    unsigned long long* d_part = nullptr;
    SIESTA_CUDA_OK(cudaMallocAsync((void**)&d_rel, rel.size(), stream));
    SIESTA_CUDA_OK(cudaMallocAsync((void**)&d_part, sizeof(unsigned long long) * (size_t)n_warps * MAX_STAT_PAIRS * 6, stream));
    SIESTA_CUDA_OK(cudaMemcpyAsync(d_rel, rel.data(), rel.size(), cudaMemcpyHostToDevice, stream));
    P.rel = d_rel;
    P.part = d_part;
    cudaEvent_t e0, e1;
    SIESTA_CUDA_OK(cudaEventCreate(&e0));
    SIESTA_CUDA_OK(cudaEventCreate(&e1));
    SIESTA_CUDA_OK(cudaEventRecord(e0, stream));
    pair_stats_kernel<<<grid, ST, 0, stream>>>(P);
    SIESTA_LAUNCHED();
    pair_stats_fold_kernel<<<n_pairs, 32, 0, stream>>>(d_part, n_warps, n_pairs, reinterpret_cast<long long*>(d_out));
    SIESTA_LAUNCHED();
    SIESTA_CUDA_OK(cudaGetLastError());
    SIESTA_CUDA_OK(cudaEventRecord(e1, stream));
    SIESTA_CUDA_OK(cudaStreamSynchronize(stream));  // `rel` (host vector) was the source of an async copy
    float ms = 0.f;
    SIESTA_CUDA_OK(cudaEventElapsedTime(&ms, e0, e1));
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFreeAsync(d_rel, stream);
    cudaFreeAsync(d_part, stream);
    if (kernel_ms) *kernel_ms = ms;
    return SIESTA_OK;
}

extern "C" int siesta_pair_stats(siesta_log* log, const int32_t* pair_a, const int32_t* pair_b, int32_t n_pairs,
                                 siesta_pair_count* out, double* kernel_ms) {
    Log* L = reinterpret_cast<Log*>(log);
    if (!L || !out || n_pairs < 1) {
        set_error("siesta_pair_stats: bad argument");
        return SIESTA_E_INVALID;
    }
    SIESTA_CUDA_OK(cudaSetDevice(L->ctx->device));
    int64_t* d = nullptr;
    SIESTA_CUDA_OK(cudaMallocAsync((void**)&d, sizeof(int64_t) * 8 * (size_t)n_pairs, L->ctx->stream));
    int rc = siesta_pair_stats_device(log, pair_a, pair_b, n_pairs, d, L->ctx->stream, kernel_ms);
    std::vector<int64_t> h((size_t)n_pairs * 8);
    if (rc == SIESTA_OK) {
        cudaError_t e = cudaMemcpyAsync(h.data(), d, sizeof(int64_t) * h.size(), cudaMemcpyDeviceToHost, L->ctx->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(L->ctx->stream);
        if (e != cudaSuccess) {
            set_error(std::string("siesta_pair_stats: D2H: ") + cudaGetErrorString(e));
            rc = SIESTA_E_CUDA;
        }
    }
    cudaFreeAsync(d, L->ctx->stream);
    if (rc) return rc;
    for (int p = 0; p < n_pairs; ++p) {
        const int64_t* o = h.data() + (size_t)p * 8;
        out[p].count = o[0];
        out[p].sum_duration_ms = o[1];
        out[p].min_duration_ms = o[0] ? o[2] : 0;
        out[p].max_duration_ms = o[0] ? o[3] : 0;
        out[p].sum_squares_lo = (uint64_t)o[4] | ((uint64_t)o[5] << 32);
        out[p].sum_squares_hi = (uint64_t)o[6] | ((uint64_t)o[7] << 32);
    }
    return SIESTA_OK;
}
