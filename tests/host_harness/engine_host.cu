// engine_host.cu — TEST HARNESS (not the product): runs the device engine (detect_engine.cuh), compiled for the
// host, over a CSR log on the CPU so tests can compare it with the oracle without a GPU.  The filter and the
// output assembly here are simple serial restatements of kernel K1's phases A and C.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../sequencedetectionqueryexecutor_b200/csrc/detect_fast.cuh"

namespace siesta {
static thread_local std::string t_err;
void set_error(const std::string& m) { t_err = m; }
std::atomic<long long> g_kernel_launches{0};
}  // namespace siesta

using namespace siesta;

template <class T>
static T* dup(const std::vector<T>& v) {
    T* p = (T*)std::malloc(sizeof(T) * (v.size() ? v.size() : 1));
    if (!v.empty()) std::memcpy(p, v.data(), sizeof(T) * v.size());
    return p;
}

struct Out {
    std::vector<int64_t> trace_idx, occ_off{0}, ev_off{0}, ev_ts, err;
    std::vector<int32_t> ev_pos, ev_rank, ev_act;
    int64_t emitted = 0;
};

template <int W, int R, int NF>
static int run_one(const DevNfa& dn, const std::vector<uint32_t>& meta, const std::vector<int32_t>& ts, int needs_ts, uint32_t flags,
                   std::vector<typename MaskOps<W>::T>& sel, unsigned* n_emitted) {
    constexpr int NE = 32 * W;
    const bool evt_pos = (flags & SIESTA_F_EVT_POS) != 0;
    if ((int)meta.size() > NE) return 3;
    TraceEvents ev{meta.data(), needs_ts ? ts.data() : nullptr, 1, (int)meta.size(), evt_pos, 1};
    // same dispatch as kernel K1: closed-form evaluators for the NK / FK2 classes (detect_fast.cuh)
    if (dn.fast_class == FAST_FK2) {
        typename MaskOps<W>::T m = 0;
        *n_emitted = 0;
        if (!fk2_eval<W>(dn, ev, m)) return 0;
        sel.assign(1, m);
        return 1;
    }
    if (dn.fast_class == FAST_NP1) {
        typename MaskOps<W>::T m = 0;
        NkMasks<W> nm;
        nm.init();
        for (int j = 0; j < ev.n; ++j) nm.on_event(j, ev.word(j));
        if (!(dn.need_vv ? np1p_eval<W>(dn, ev, nm.T, m, *n_emitted)
                         : np1_eval<W>(dn, nm.T, (flags & (SIESTA_F_RETURN_ALL | SIESTA_F_COUNT_MATCHES)) != 0, m, *n_emitted))) return 0;
        sel.assign(1, m);
        return 1;
    }
    if (dn.fast_class == FAST_NK) {
        std::vector<typename MaskOps<W>::T> aux(NE), s(NE);
        int nsel = 0;
        NkMasks<W> nm;
        nm.init();
        for (int j = 0; j < ev.n; ++j) nm.on_event(j, ev.word(j));
        // same rule as kernel K1: stop at the first completed start when the walks are monotone and nothing is counted
        const bool first_only = !(flags & (SIESTA_F_RETURN_ALL | SIESTA_F_COUNT_MATCHES)) && !needs_ts;
        if (!nk_eval<W>(dn, ev, nm.T, (flags & SIESTA_F_RETURN_ALL) != 0, evt_pos, aux.data(), 1, s.data(), nsel, *n_emitted, first_only)) return 0;
        sel.assign(s.begin(), s.begin() + nsel);
        return 1;
    }
    auto* arrays = new RunArrays<W, R, NF>();
    auto* eng = new RunEngine<RunStore<W, R, NF, 1>>(dn, ev, arrays->store());
    const bool prune = (flags & SIESTA_F_LITERAL_RUNS) == 0;
    const bool dedup = prune && !(flags & SIESTA_F_RETURN_ALL) && !(flags & SIESTA_F_COUNT_MATCHES);
    BestEmit<W> be;
    eng->run(be, prune, dedup);
    int status = 0;
    if (eng->ovf) status = 3;
    else if (eng->err) status = 2;
    else if (be.n > 0) {
        status = 1;
        *n_emitted = be.n;
        sel.assign(1, be.best);
        if ((flags & SIESTA_F_RETURN_ALL) && be.n > 1) {
            GreedyEmit<W, NE> ge(ev, be.best, evt_pos);
            eng->run(ge, prune, false);
            if (ge.ovf || eng->ovf) status = 3;
            else sel.assign(ge.sel, ge.sel + ge.nsel);
        }
    }
    delete eng;
    delete arrays;
    return status;
}

extern "C" const char* engine_host_last_error() { return siesta::t_err.c_str(); }

// which evaluator validate_nfa (csrc/nfa.cpp, the product's own host code) picks: FAST_* of detect_fast.cuh, or < 0 = the error
extern "C" int engine_host_fast_class(const siesta_nfa* nfa, uint32_t flags) {
    DevNfa dn;
    const int rc = validate_nfa(nfa, flags, &dn);
    return rc ? (rc < 0 ? rc : -rc) : (int)dn.fast_class;
}

extern "C" int engine_host_detect(const int64_t* trace_off, const int32_t* act, const int64_t* ts_ms, int64_t n_traces,
                                  int32_t n_act, const siesta_nfa* nfa, const int64_t* cand, int64_t n_cand, uint32_t flags,
                                  int64_t* n_wide, siesta_matches** out) {
    DevNfa dn;
    int rc = validate_nfa(nfa, flags, &dn);
    if (rc) return rc;
    std::vector<uint16_t> lut;
    int needs_ts = 0, n_pos = 0;
    build_lut(nfa, dn, n_act, flags, lut, &needs_ts, &n_pos);
    const bool evt_pos = (flags & SIESTA_F_EVT_POS) != 0;
    Out o;
    *n_wide = 0;
    const int64_t n = cand ? n_cand : n_traces;
    for (int64_t i = 0; i < n; ++i) {
        const int64_t t = cand ? cand[i] : i;
        const int64_t b0 = trace_off[t], b1 = trace_off[t + 1];
        std::vector<uint32_t> meta;
        std::vector<int32_t> ts;
        long long t0 = 0;
        for (int64_t p = b0; p < b1; ++p) {
            const int a = act[p];
            const uint32_t m = (a >= 0 && a < n_act) ? lut[a] : 0;
            if (!m) continue;
            if (meta.empty()) t0 = ts_ms[p];
            meta.push_back(m | ((uint32_t)(p - b0) << 16));
            ts.push_back((int)((ts_ms[p] - t0) / 1000));
        }
        if (meta.empty()) continue;
        if (b1 - b0 > 65536) return SIESTA_E_UNSUPPORTED;
        unsigned emitted = 0;
        // kernel K1-P (detect_nkp_kernel): class NK as window walks in rank space (32 filtered events) or over the 64
        // raw position slots that start at the 32-byte sector of the trace's first event (detect_fast.cuh)
        const int lead = (int)(b0 & 7);
        NkwProgram prog;
        const int space = nkw_build(dn, flags, &prog);
        // returnAll: K1-P answers a trace with ONE engine match (that occurrence is the selection); with more, the overlap test
        // decides and the trace goes to the staged kernel (below), as a trace that does not fit
        const bool return_all_h = (flags & SIESTA_F_RETURN_ALL) != 0;
        const bool first_only = !(flags & (SIESTA_F_COUNT_MATCHES | SIESTA_F_RETURN_ALL));
        if (space != NKW_NONE && (!needs_ts || return_all_h) && (b1 - b0) + lead <= 64 && (space == NKW_RAW || meta.size() <= 32)) {
            std::vector<int> sel_idx;   // the selected occurrence as indices into the filtered list
            bool hit;
            if (space == NKW_RANK) {
                uint32_t T[SIESTA_MAX_STATES + 1] = {0, 0, 0, 0, 0, 0, 0, 0, 0}, best = 0, sel_g[32], aux[32];
                for (size_t j = 0; j < meta.size(); ++j)
                    for (int k = 0; k < SIESTA_MAX_STATES; ++k)
                        if (meta[j] & (1u << k)) T[k] |= 1u << j;
                hit = nkw_eval<uint32_t>(prog, T, best, emitted, first_only);
                if (prog.markov) {   // the backward all-starts evaluation must agree with the per-start walks
                    uint32_t best_m = 0;
                    unsigned emitted_m = 0;
                    const bool hit_m = nkw_eval_markov<uint32_t>(prog, T, best_m, emitted_m);
                    if (hit_m != hit || (hit && best_m != best) || (!first_only && emitted_m != emitted)) {
                        siesta::set_error("nkw_eval_markov<rank> disagrees with nkw_eval");
                        return -98;
                    }
                }
                // cross-check against the generic greedy walk on the filtered list (the staged kernel's evaluator)
                TraceEvents ev{meta.data(), nullptr, 1, (int)meta.size(), evt_pos, 1};
                int nsel = 0;
                unsigned emitted_g = 0;
                const bool hit_g = nk_eval<1>(dn, ev, T, false, evt_pos, aux, 1, sel_g, nsel, emitted_g, first_only);
                if (hit != hit_g || (hit && (best != sel_g[0] || emitted != emitted_g))) {
                    siesta::set_error("nkw_eval<rank> disagrees with nk_eval<TraceEvents>");
                    return -99;
                }
                for (uint32_t m = best; m; m &= m - 1) sel_idx.push_back(__builtin_ffs((int)m) - 1);
            } else {
                unsigned long long T[SIESTA_MAX_STATES + 1] = {0, 0, 0, 0, 0, 0, 0, 0, 0}, R = 0, best = 0, sel_g[1];
                for (uint32_t m : meta) {
                    const int slot = (int)(m >> 16) + lead;
                    R |= 1ull << slot;
                    for (int k = 0; k < SIESTA_MAX_STATES; ++k)
                        if (m & (1u << k)) T[k] |= 1ull << slot;
                }
                hit = nkw_eval<unsigned long long>(prog, T, best, emitted, first_only);
                if (prog.markov) {
                    unsigned long long best_m = 0;
                    unsigned emitted_m = 0;
                    const bool hit_m = nkw_eval_markov<unsigned long long>(prog, T, best_m, emitted_m);
                    if (hit_m != hit || (hit && best_m != best) || (!first_only && emitted_m != emitted)) {
                        siesta::set_error("nkw_eval_markov<raw> disagrees with nkw_eval");
                        return -98;
                    }
                }
                PosEvents pe{R, lead, evt_pos};
                int nsel = 0;
                unsigned emitted_g = 0;
                const bool hit_g = nk_eval<2>(dn, pe, T, false, evt_pos, (unsigned long long*)nullptr, 0, sel_g, nsel, emitted_g, first_only);
                if (hit != hit_g || (hit && (best != sel_g[0] || emitted != emitted_g))) {
                    siesta::set_error("nkw_eval<raw> disagrees with nk_eval<PosEvents>");
                    return -99;
                }
                for (unsigned long long m = best; m; m &= m - 1) sel_idx.push_back(pe.rank(__builtin_ffsll((long long)m) - 1));
            }
            if (!hit) continue;
            if (!(return_all_h && emitted > 1)) {
                o.emitted += emitted;
                o.trace_idx.push_back(t);
                for (int j : sel_idx) {
                    const int src = (int)(meta[j] >> 16);
                    o.ev_pos.push_back(src);
                    o.ev_rank.push_back(j);
                    o.ev_act.push_back(act[b0 + src]);
                    const long long raw = ts_ms[b0 + src];
                    o.ev_ts.push_back(evt_pos ? raw : (long long)((int)((raw - t0) / 1000)) * 1000 + t0);
                }
                o.ev_off.push_back((int64_t)o.ev_pos.size());
                o.occ_off.push_back((int64_t)o.ev_off.size() - 1);
                continue;
            }
            emitted = 0;   // more than one engine match: evaluated again below, with the overlap test
        }
        // kernel K1-L: class NK on a trace beyond the mask kernels' limits - the forward walk over the filtered list.  The
        // harness takes it from 25 relevant events on, so that the usual soak traces exercise it too.
        if (dn.fast_class == FAST_NK && meta.size() > 24 && std::getenv("SIESTA_HARNESS_NO_LONG") == nullptr) {
            std::vector<uint16_t> word(meta.size());
            std::vector<int32_t> posv(meta.size());
            for (size_t j = 0; j < meta.size(); ++j) {
                word[j] = (uint16_t)(meta[j] & 0xFFFFu);
                posv[j] = (int32_t)(meta[j] >> 16);
            }
            LongEvents le{(int)meta.size(), word.data(), posv.data(), needs_ts ? ts.data() : nullptr, evt_pos};
            const bool return_all = (flags & SIESTA_F_RETURN_ALL) != 0;
            const int np = n_pos > 0 ? n_pos : 1;
            // every start's run; emission order = (completion, start)
            struct Run { int c, s; std::vector<int> e; };
            std::vector<Run> runs;
            for (int s0 = 0; s0 < le.n; ++s0) {
                if (!(word[s0] & 1u)) continue;
                int o[SIESTA_MAX_STATES];
                const int k = nk_long_walk(dn, le, s0, o);
                if (k) runs.push_back(Run{o[k - 1], s0, std::vector<int>(o, o + k)});
            }
            if (runs.empty()) continue;
            (void)np;
            std::stable_sort(runs.begin(), runs.end(), [](const Run& a, const Run& b) { return a.c != b.c ? a.c < b.c : a.s < b.s; });
            std::vector<const Run*> sel{&runs[0]};   // all runs have the same size: the first emitted one is the first-largest
            if (return_all)
                for (size_t r = 1; r < runs.size(); ++r) {
                    bool ov = false;
                    for (const Run* q : sel) ov = ov || nk_long_overlaps(le, evt_pos, runs[r].e.front(), runs[r].e.back(), q->e.front(), q->e.back());
                    if (!ov) sel.push_back(&runs[r]);
                }
            o.emitted += (int64_t)runs.size();
            o.trace_idx.push_back(t);
            for (const Run* q : sel) {
                for (int j : q->e) {
                    const int src = posv[j];
                    o.ev_pos.push_back(src);
                    o.ev_rank.push_back(j);
                    o.ev_act.push_back(act[b0 + src]);
                    const long long raw = ts_ms[b0 + src];
                    o.ev_ts.push_back(evt_pos ? raw : (long long)((int)((raw - t0) / 1000)) * 1000 + t0);
                }
                o.ev_off.push_back((int64_t)o.ev_pos.size());
            }
            o.occ_off.push_back((int64_t)o.ev_off.size() - 1);
            continue;
        }
        int status;
        std::vector<uint32_t> sel1;
        std::vector<unsigned long long> sel2;
        bool wide = false;
        status = run_one<1, 16, 16>(dn, meta, ts, needs_ts, flags, sel1, &emitted);
        if (status == 3) {
            wide = true;
            ++*n_wide;
            status = run_one<2, 1024, 128>(dn, meta, ts, needs_ts, flags, sel2, &emitted);
            if (status == 3) return SIESTA_E_UNSUPPORTED;
        }
        if (status == 2) {
            o.err.push_back(t);
            continue;
        }
        if (status != 1) continue;
        o.emitted += emitted;
        o.trace_idx.push_back(t);
        const size_t ns = wide ? sel2.size() : sel1.size();
        for (size_t k = 0; k < ns; ++k) {
            unsigned long long m = wide ? sel2[k] : sel1[k];
            while (m) {
                const int j = __builtin_ffsll((long long)m) - 1;
                m &= m - 1;
                const int src = (int)(meta[j] >> 16);
                o.ev_pos.push_back(src);
                o.ev_rank.push_back(j);
                o.ev_act.push_back(act[b0 + src]);
                const long long raw = ts_ms[b0 + src];
                o.ev_ts.push_back(evt_pos ? raw : (long long)((int)((raw - t0) / 1000)) * 1000 + t0);
            }
            o.ev_off.push_back((int64_t)o.ev_pos.size());
        }
        o.occ_off.push_back((int64_t)o.ev_off.size() - 1);
    }
    siesta_matches* m = (siesta_matches*)std::calloc(1, sizeof(siesta_matches));
    m->n_traces = (int64_t)o.trace_idx.size();
    m->n_occurrences = (int64_t)o.ev_off.size() - 1;
    m->n_events = (int64_t)o.ev_pos.size();
    m->n_matches_emitted = (flags & (SIESTA_F_COUNT_MATCHES | SIESTA_F_RETURN_ALL | SIESTA_F_LITERAL_RUNS)) ? o.emitted : -1;
    m->n_ref_errors = (int64_t)o.err.size();
    m->trace_idx = dup(o.trace_idx);
    m->occ_off = dup(o.occ_off);
    m->ev_off = dup(o.ev_off);
    m->ev_pos = dup(o.ev_pos);
    m->ev_rank = dup(o.ev_rank);
    m->ev_act = dup(o.ev_act);
    m->ev_ts_ms = dup(o.ev_ts);
    m->err_trace_idx = dup(o.err);
    *out = m;
    return 0;
}
