// Test harness: kernel W's per-trace evaluation (csrc/wnm.cuh) compiled for the host and driven serially, so the closed
// form can be compared with the literal oracle (oracle/wnm_oracle.cpp) on the CPU.  Not part of the product.
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../sequencedetectionqueryexecutor_b200/csrc/wnm.cuh"

using namespace siesta;

extern "C" int wnm_host_eval(const int64_t* trace_off, const int32_t* act, const int64_t* ts_ms, int64_t n_traces,
                             const int32_t* pattern, int32_t m, const siesta_wnm_constraint* cons, int32_t n_cons, int32_t u,
                             int32_t step, int32_t k, uint32_t flags, int32_t* status /* [n_traces] */, int32_t* total,
                             int32_t* ev_pos, int32_t* ev_value, int32_t* ev_change, int32_t* ev_spos /* [n_traces * m] */) {
    WnmProgram W;
    std::memset(&W, 0, sizeof(W));
    W.m = m;
    for (int j = 0; j < m; ++j) W.type[j] = pattern[j];
    W.n_cons = n_cons;
    for (int q = 0; q < n_cons; ++q) {
        W.c_a[q] = cons[q].pos_a;
        W.c_b[q] = cons[q].pos_b;
        W.c_kind[q] = cons[q].kind;
        W.c_method[q] = cons[q].method;
        W.c_value[q] = cons[q].value;
        if (cons[q].kind == SIESTA_WNM_TIME) W.time_at[cons[q].pos_b] = 1;
    }
    W.u = u;
    W.step = step;
    W.k = k;
    W.evt_pos = (flags & SIESTA_F_EVT_POS) ? 1 : 0;
    for (int64_t t = 0; t < n_traces; ++t) {
        std::vector<long long> lo;
        std::vector<int> nv, d, src;
        std::vector<unsigned char> msk;
        for (int64_t i = trace_off[t]; i < trace_off[t + 1]; ++i) {
            const unsigned sm = wnm_state_mask(W, act[i]);
            if (!sm) continue;
            const long long prim = wnm_primary(W, ts_ms[i], i - trace_off[t]);
            lo.push_back(wnm_lo(W, prim));
            nv.push_back(wnm_variants(W, prim));
            d.push_back((int)(prim - lo.back()));
            src.push_back((int)(i - trace_off[t]));
            msk.push_back((unsigned char)sm);
        }
        status[t] = 0;
        int n = 0;
        for (int x : nv) n += x;
        if (n == 0) continue;
        std::vector<int> val((size_t)n, -1), chg((size_t)n, -1), ssrc((size_t)n, -1);
        std::vector<unsigned char> smask((size_t)n, 0);
        for (int q = 0; q < (int)lo.size(); ++q)
            for (int v = 0; v < nv[(size_t)q]; ++v) {
                const int idx = wnm_rank(W, lo.data(), nv.data(), (int)lo.size(), q, v);
                if (idx < 0 || idx >= n || val[(size_t)idx] != -1) return 1;   // the ranks must be a permutation
                const long long sh = (long long)v * step;
                val[(size_t)idx] = (int)(lo[(size_t)q] + sh);
                chg[(size_t)idx] = (int)(sh > d[(size_t)q] ? sh - d[(size_t)q] : d[(size_t)q] - sh);
                ssrc[(size_t)idx] = src[(size_t)q];
                smask[(size_t)idx] = msk[(size_t)q];
            }
        std::vector<unsigned short> prev((size_t)(m > 1 ? m - 1 : 1) * n);
        int best_f = WNM_INF, best_t[SIESTA_MAX_STATES], tup[SIESTA_MAX_STATES];
        for (int s = 0; s < n; ++s) {
            if (!(smask[(size_t)s] & 1u) || chg[(size_t)s] > k) continue;
            const int f = wnm_sweep<unsigned short>(W, val.data(), chg.data(), smask.data(), n, s, prev.data(), n, 1, tup);
            if (f != WNM_INF && (best_f == WNM_INF || wnm_better(m, f, tup, best_f, best_t))) {
                best_f = f;
                for (int j = 0; j < m; ++j) best_t[j] = tup[j];
            }
        }
        if (best_f == WNM_INF) continue;
        status[t] = 1;
        total[t] = best_f;
        for (int j = 0; j < m; ++j) {
            const int e = best_t[j];
            ev_pos[t * m + j] = ssrc[(size_t)e];
            ev_value[t * m + j] = val[(size_t)e];
            ev_change[t * m + j] = chg[(size_t)e];
            ev_spos[t * m + j] = e;
        }
    }
    return 0;
}
