// explore.cu — /explore in accurate mode (row X): for every candidate continuation, the exact number of occurrences
// of pattern + candidate and the summed duration of those occurrences.
//
// Replaces QueryPlanExplorationAccurate.patternDetection (J/model/Queries/QueryPlans/Exploration/
// QueryPlanExplorationAccurate.java:82-102): SimplePattern + [next] -> SaseConnector.evaluate(..., false) ->
// clearOccurrences(true) per trace -> completions = number of occurrences, average duration =
// sum(Occurrence.getDuration) / completions with getDuration = (last.ms - first.ms) / 1000.0 (J/model/Occurrence.java:
// 55-61).  The reference runs one full detection per candidate; so does this version (each a K1 launch on the resident
// log, no host round trip for the events); sharing one pass between the candidates is the next step.
// The library returns exact integers (completions, summed milliseconds); the double division stays with the caller.
#include <cstring>
#include <vector>

#include "common.cuh"

namespace siesta {

__global__ void __launch_bounds__(256) occurrence_duration_kernel(const int64_t* ev_off, const int64_t* ev_ts, int64_t n_occ,
                                                                  unsigned long long* sum_ms) {
    long long acc = 0;
    for (int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; o < n_occ; o += (int64_t)gridDim.x * blockDim.x) {
        const int64_t a = ev_off[o], b = ev_off[o + 1];
        if (b > a) acc += ev_ts[b - 1] - ev_ts[a];
    }
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) {
        const int lo = __shfl_xor_sync(0xffffffffu, (int)(acc & 0xffffffffll), d);
        const int hi = __shfl_xor_sync(0xffffffffu, (int)(acc >> 32), d);
        acc += ((long long)hi << 32) | (unsigned int)lo;
    }
    if ((threadIdx.x & 31) == 0 && acc) atomicAdd(sum_ms, (unsigned long long)acc);
}

}  // namespace siesta

using namespace siesta;

extern "C" int siesta_explore_accurate(siesta_log* log, const int32_t* pattern_activities, int32_t n_pattern,
                                       const int32_t* candidates, int32_t n_candidates, uint32_t flags,
                                       int64_t* completions, int64_t* sum_duration_ms, double* kernel_ms) {
    Log* L = reinterpret_cast<Log*>(log);
    if (!L || !pattern_activities || n_pattern < 1 || n_pattern + 1 > SIESTA_MAX_STATES || (n_candidates && !candidates) ||
        n_candidates < 0 || !completions || !sum_duration_ms) {
        set_error("siesta_explore_accurate: bad argument (pattern of 1.." + std::to_string(SIESTA_MAX_STATES - 1) + " events)");
        return SIESTA_E_INVALID;
    }
    if (flags & ~(uint32_t)SIESTA_F_EVT_POS) {
        set_error("siesta_explore_accurate: only SIESTA_F_EVT_POS may be set (returnAll is implied)");
        return SIESTA_E_INVALID;
    }
    SIESTA_CUDA_OK(cudaSetDevice(L->ctx->device));
    cudaStream_t stream = L->ctx->stream;
    unsigned long long* d_sum = nullptr;
    SIESTA_CUDA_OK(cudaMallocAsync((void**)&d_sum, sizeof(unsigned long long) * (size_t)(n_candidates ? n_candidates : 1), stream));
    SIESTA_CUDA_OK(cudaMemsetAsync(d_sum, 0, sizeof(unsigned long long) * (size_t)(n_candidates ? n_candidates : 1), stream));
    double ms = 0;
    int rc = SIESTA_OK;
    for (int c = 0; c < n_candidates && rc == SIESTA_OK; ++c) {
        siesta_nfa nfa;  // SimplePattern.getNfa: all states "normal" (J/model/Patterns/SimplePattern.java:96-104)
        std::memset(&nfa, 0, sizeof(nfa));
        nfa.n_states = n_pattern + 1;
        for (int s = 0; s <= n_pattern; ++s) {
            nfa.states[s].kind = SIESTA_STATE_NORMAL;
            nfa.states[s].n_types = 1;
            nfa.states[s].types[0] = s < n_pattern ? pattern_activities[s] : candidates[c];
        }
        siesta_dev_matches dm;
        rc = detect_device_impl(L, &nfa, nullptr, 0, flags | SIESTA_F_RETURN_ALL, stream, RebaseOffsets{0, 0, 0}, &dm);
        if (rc) break;
        completions[c] = dm.n_occurrences;
        ms += dm.kernel_ms;
        if (dm.n_ref_errors) {
            set_error("siesta_explore_accurate: the reference engine throws on this pattern");
            rc = SIESTA_E_REFERENCE_THROWS;
        } else if (dm.n_occurrences > 0) {
            const int grid = (int)std::min<int64_t>((dm.n_occurrences + 255) / 256, (int64_t)L->ctx->sm_count * 8);
            occurrence_duration_kernel<<<grid, 256, 0, stream>>>(dm.d_ev_off, dm.d_ev_ts_ms, dm.n_occurrences, d_sum + c);
            SIESTA_LAUNCHED();
            if (cudaGetLastError() != cudaSuccess) rc = SIESTA_E_CUDA;
        }
        // the duration kernel reads dm's buffers: they are freed on the ctx stream, in order
        siesta_dev_matches_free(&dm);
    }
    if (rc == SIESTA_OK && n_candidates) {
        cudaError_t e = cudaMemcpyAsync(sum_duration_ms, d_sum, sizeof(int64_t) * (size_t)n_candidates, cudaMemcpyDeviceToHost, stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
        if (e != cudaSuccess) {
            set_error(std::string("siesta_explore_accurate: D2H: ") + cudaGetErrorString(e));
            rc = SIESTA_E_CUDA;
        }
    }
    cudaFreeAsync(d_sum, stream);
    if (kernel_ms) *kernel_ms = ms;
    return rc;
}
