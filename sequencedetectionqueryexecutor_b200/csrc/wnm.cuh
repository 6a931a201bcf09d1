// wnm.cuh — kernel W (why-not-match): the per-trace evaluation, shared by csrc/wnm.cu (device) and tests/host_harness
// (the same code compiled for the host and compared with the literal oracle, oracle/wnm_oracle.cpp).
//
// Replaces WhyNotMatchSASE.evaluate (J/model/WhyNotMatch/UsingSase/WhyNotMatchSASE.java:37-55, getUnCertainStream :63-83,
// getNFA :91-152, createResponse :160-173) on the SASE engine's skip-till-any-match path (S/engine/Engine.java:159-178,
// 341-350, 593-645).  The reference enumerates EVERY combination of uncertain events as a run (exponential); what it
// reports has a closed form that one sweep per start event computes:
//
//  * the uncertain stream.  Event q of the trace (activities of the pattern only) becomes V_q = (hi_q - lo_q) / step + 1
//    events at lo_q + v * step, lo_q = max(primary_q - u, 0), hi_q = primary_q + u, change = |shift|; the stream is the
//    STABLE sort by shifted value, so the index of (q, v) is v + #{(q', v') : value' < value} + #{q' < q : value' == value}
//    - a sum of clamped divisions over the other events, no sort (wnm_rank).
//  * all states are "normal" and the strategy is skip-till-any-match: a run is any index-increasing tuple whose events
//    pass the predicates when they are taken.  Matches leave the engine ordered by their last event, then by the
//    position of the run they extend in the run list, which is the order of ITS last event, and so on down to the start:
//    the emission order is the lexicographic order of the REVERSED tuples.  createResponse keeps the later of two matches
//    of equal total change, so the answer is: least total change, then the greatest reversed tuple.
//  * predicates as the engine evaluates them:
//      - `change <= k - $2.change - ..`: on state 0 it reads `change <= k`; on every later state its last operand names the
//        state itself and PredicateOptimized.evaluate returns true (PredicateOptimized.java:348-350) after reading value
//        vectors that exist by then: vacuous.
//      - `position|timestamp > $previous.x` (only on a state a constraint of that kind ends at) compares with the run's own
//        previous event; positions increase anyway, timestamps must increase strictly.
//      - `x <=|>= $A.x + c`: $A is read from the value vector of state A, and Run.clone is shallow (Run.java:319-327): the
//        array is shared by all runs that descend from one START event and is overwritten whenever any of them takes an
//        event at state A (initializeValueVector :332-355).  For A = 0 that is the start itself; for A >= 1 it is the LAST
//        event any run of the family took at state A before this event - not the run's own.
//    So, for a fixed start, whether an event e is taken at state j depends on e, on the family's latest events, and on
//    the run only through "some run at state j - 1 whose last event lies before e (and strictly earlier in time, if a time
//    constraint ends at j)".  One forward sweep per start keeps, per state, the latest taken event and the least cost of a
//    partial run ending before e (ties to the later event: the greater reversed tuple), plus a back pointer per taken event.
//  * a constraint (A >= 1, B) between two states of the SAME activity: the event e that a run at state B is offered may be
//    taken at state A in the same step, and the family's slot then holds e itself when the run at B is evaluated - always,
//    because the earliest run at state A takes whatever any run at A takes and was created before every run at a later
//    state, so it sits earlier in the run list.  The sweep decides the states of an event in ascending order for that.
#pragma once
#include <cstdint>

#include "../../include/siesta_gpu.h"

#if defined(__CUDACC__)
#define WNM_HD __host__ __device__ __forceinline__
#else
#define WNM_HD inline
#endif

namespace siesta {

constexpr int WNM_MAX_CONS = 16;
constexpr int WNM_INF = 0x7fffffff;

struct WnmProgram {
    int m;                            // states (all "normal")
    int type[SIESTA_MAX_STATES];      // activity of state j
    int n_cons;
    int c_a[WNM_MAX_CONS], c_b[WNM_MAX_CONS], c_kind[WNM_MAX_CONS], c_method[WNM_MAX_CONS];
    long long c_value[WNM_MAX_CONS];
    int time_at[SIESTA_MAX_STATES];   // a time constraint ends at state j: `timestamp > $previous.timestamp` is asked there
    int u, step, k;
    int evt_pos;                      // primary metric = position (SIESTA_F_EVT_POS) instead of epoch seconds
};

// states whose activity is `a`, as a bit mask (0: not an event of the pattern)
WNM_HD unsigned wnm_state_mask(const WnmProgram& W, int a) {
    unsigned msk = 0;
    for (int j = 0; j < W.m; ++j) msk |= (W.type[j] == a ? 1u : 0u) << j;
    return msk;
}
// Event.getPrimaryMetric: EventTs.java:87-89 (epoch ms / 1000, Java's truncating division), EventPos.java:84-86
WNM_HD long long wnm_primary(const WnmProgram& W, long long ts_ms, long long pos) { return W.evt_pos ? pos : ts_ms / 1000; }
WNM_HD long long wnm_lo(const WnmProgram& W, long long primary) { return primary - W.u > 0 ? primary - W.u : 0; }
WNM_HD int wnm_variants(const WnmProgram& W, long long primary) {   // (clamped far above SIESTA_WNM_MAX_STREAM: such a trace is listed)
    const long long v = (primary + W.u - wnm_lo(W, primary)) / W.step + 1;
    return v > (1 << 20) ? (1 << 20) : (int)v;
}

// index of variant v of relevant event q in the stable sort by shifted value (getUnCertainStream :76-80)
WNM_HD int wnm_rank(const WnmProgram& W, const long long* lo, const int* nv, int n_rel, int q, int v) {
    const long long x = lo[q] + (long long)v * W.step;
    long long idx = v;
    for (int p = 0; p < n_rel; ++p) {
        if (p == q) continue;
        const long long d = x - lo[p] - (p < q ? 0 : 1);   // variants of p at or below x (p earlier: ties go first), below x otherwise
        if (d < 0) continue;
        if (d >= (long long)(nv[p] - 1) * W.step) idx += nv[p];   // all of them: the usual case, events lie further apart than 2 u
        else idx += (int)d / W.step + 1;                           // (d < 2 u: 32-bit division)
    }
    return (int)idx;
}

// The sweep of one start event s over the stream [0, n): value / change / state mask per stream index.
// prev: back pointers, entry (j, e) at prev[((j - 1) * n_cap + e) * prev_stride] for states j = 1 .. m - 1.
// Returns the least total change of a match that starts at s (WNM_INF: none) and its stream indices in tup[0 .. m).
// M = number of states, a template parameter: every per-state array is indexed by unrolled loops only, so the sweep's
// state lives in registers (with a run-time state count the arrays sat in local memory).
template <int M, typename PrevT>
WNM_HD int wnm_sweep_m(const WnmProgram& W, const int* __restrict__ val, const int* __restrict__ chg, const unsigned char* __restrict__ smask,
                       int n, int s, PrevT* prev, int n_cap, int prev_stride, int* tup) {
    int latest[M];                // latest event the family took at state j (-1: none yet)
    int all_f[M], all_i[M];       // least cost of a run at state j over every taken event so far (ties: later)
    int old_f[M], old_i[M];       // ... over the events strictly earlier in time than the current one
#pragma unroll
    for (int j = 0; j < M; ++j) {
        latest[j] = -1;
        all_f[j] = old_f[j] = WNM_INF;
        all_i[j] = old_i[j] = -1;
    }
    latest[0] = s;
    all_f[0] = chg[s];
    all_i[0] = s;
    int best_f = WNM_INF, best_e = -1;
    int cur_val = val[s];
    for (int e = s + 1; e < n; ++e) {
        const int ve = val[e];
        if (ve != cur_val) {   // time moved on: everything taken so far is strictly earlier
            cur_val = ve;
#pragma unroll
            for (int j = 0; j < M - 1; ++j) {
                old_f[j] = all_f[j];
                old_i[j] = all_i[j];
            }
        }
        const unsigned sm = smask[e] >> 1;
        if (!sm) continue;
        // Pass 1, states ascending: which states take e.  A run is offered e once, at its own state, and the partial runs it
        // extends are those that existed before e (all_* / old_* are only updated in pass 2).  A constraint (A, j) reads the
        // family's slot of state A as it is when the run at state j is evaluated: if e itself is taken at state A (same
        // activity at both states), the run that takes it - the earliest run at state A takes whatever any run at A takes, and
        // it was created before every run at state j > A - sits earlier in the run list, so the slot already holds e.
        unsigned acc = 0;
        int fj[M];
#pragma unroll
        for (int j = 1; j < M; ++j) {
            fj[j] = WNM_INF;
            if (!((sm >> (j - 1)) & 1u)) continue;
            const int pf = W.time_at[j] ? old_f[j - 1] : all_f[j - 1];
            const int pi = W.time_at[j] ? old_i[j - 1] : all_i[j - 1];
            if (pf == WNM_INF) continue;
            bool ok = true;
            for (int q = 0; q < W.n_cons; ++q) {
                if (W.c_b[q] != j) continue;
                const int a = W.c_a[q];
                int r = -1;
#pragma unroll
                for (int x = 0; x < M - 1; ++x) r = x == a ? latest[x] : r;
                if ((acc >> a) & 1u) r = e;
                if (r < 0) {   // (cannot happen with c_a < c_b: a run at state j went through c_a)
                    ok = false;
                    break;
                }
                const long long lhs = W.c_kind[q] == SIESTA_WNM_GAP ? e : ve;
                const long long rhs = (W.c_kind[q] == SIESTA_WNM_GAP ? r : val[r]) + W.c_value[q];
                if (W.c_method[q] == SIESTA_WNM_WITHIN ? !(lhs <= rhs) : !(lhs >= rhs)) {
                    ok = false;
                    break;
                }
            }
            if (!ok) continue;
            acc |= 1u << j;
            fj[j] = pf + chg[e];
            prev[((size_t)(j - 1) * n_cap + e) * prev_stride] = (PrevT)pi;
        }
        // Pass 2: the taken states' bookkeeping
#pragma unroll
        for (int j = 1; j < M; ++j) {
            if (!((acc >> j) & 1u)) continue;
            const int f = fj[j];
            if (j == M - 1) {
                if (f <= best_f) {
                    best_f = f;
                    best_e = e;
                }
            } else {
                latest[j] = e;
                if (f <= all_f[j]) {
                    all_f[j] = f;
                    all_i[j] = e;
                }
            }
        }
    }
    if (best_f == WNM_INF) return WNM_INF;
    int e = best_e;
#pragma unroll
    for (int j = M - 1; j >= 1; --j) {
        tup[j] = e;
        e = (int)prev[((size_t)(j - 1) * n_cap + e) * prev_stride];
    }
    tup[0] = e;
    return best_f;
}

template <typename PrevT>
WNM_HD int wnm_sweep(const WnmProgram& W, const int* __restrict__ val, const int* __restrict__ chg, const unsigned char* __restrict__ smask,
                     int n, int s, PrevT* prev, int n_cap, int prev_stride, int* tup) {
    switch (W.m) {
        case 1: tup[0] = s; return chg[s];
        case 2: return wnm_sweep_m<2>(W, val, chg, smask, n, s, prev, n_cap, prev_stride, tup);
        case 3: return wnm_sweep_m<3>(W, val, chg, smask, n, s, prev, n_cap, prev_stride, tup);
        case 4: return wnm_sweep_m<4>(W, val, chg, smask, n, s, prev, n_cap, prev_stride, tup);
        case 5: return wnm_sweep_m<5>(W, val, chg, smask, n, s, prev, n_cap, prev_stride, tup);
        case 6: return wnm_sweep_m<6>(W, val, chg, smask, n, s, prev, n_cap, prev_stride, tup);
        case 7: return wnm_sweep_m<7>(W, val, chg, smask, n, s, prev, n_cap, prev_stride, tup);
        default: return wnm_sweep_m<8>(W, val, chg, smask, n, s, prev, n_cap, prev_stride, tup);
    }
}

// (cost, tuple) a better than b: less total change, then the later match = the greater reversed tuple
WNM_HD bool wnm_better(int m, int fa, const int* ta, int fb, const int* tb) {
    if (fa != fb) return fa < fb;
    for (int j = m - 1; j >= 0; --j)
        if (ta[j] != tb[j]) return ta[j] > tb[j];
    return false;
}

}  // namespace siesta
