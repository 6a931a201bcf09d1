/*
 * wnm_oracle.cpp — CPU ORACLE (test infrastructure, NOT the product).
 *
 * A literal restatement of the reference's why-not-match path: the uncertain stream, the NFA with its predicates,
 * the SASE engine under skip-till-any-match exactly as it executes (value vectors shared by the clones of a run,
 * the operand loop of PredicateOptimized.evaluate with its early returns), and the choice of the reported match.
 * Exponential like the original (every combination of events becomes a run): only for small cases.
 * Only tests/ may load this.
 *
 * Parity pins: the reference's WhyNotMatchSASETest (src/test/java/.../model/WhyNotMatch/UsingSase/
 * WhyNotMatchSASETest.java:31-38 stream sizes 7 x 3 and 6 x 3; :70-81 one almost-match for the trace A@101 B@104
 * C@109 with (0,1) within 2 s, (1,2) at least 7 s, u = 3, step = 1, k = 3) are replayed by tests/test_wnm.py.
 * The reference itself cannot run here (no JVM).
 *
 * Paths cited below are relative to the reference's src/main/java/:
 *   S/ = edu/umass/cs/sase/      J/ = com/datalab/siesta/queryprocessor/
 */
#include "../include/siesta_gpu.h"

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <unordered_map>
#include <vector>

namespace {

enum { A_POSITION = 0, A_TIMESTAMP = 1, A_CHANGE = 2 };

/* UncertainTimeEvent (J/model/WhyNotMatch/UsingSase/UncertainTimeEvent.java): getId() == position */
struct UEvent {
    int type;
    int position;
    int timestamp;
    int change;
    int src; /* index of the original event inside its trace (not part of the Java object) */
};
int attr_of(const UEvent& e, int a) { /* getAttributeByName :33-41 */
    return a == A_POSITION ? e.position : a == A_TIMESTAMP ? e.timestamp : e.change;
}

struct NullPointer {}; /* the Java NullPointerException that PredicateOptimized.evaluate catches (:353-357) */
struct TooLarge {};

/* One predicate string of WhyNotMatchSASE.getNFA (:91-152), kept as its operand list. */
struct Pred {
    int kind; /* 0: "<attr> > $previous.<attr>"   1: "<attr> <=|>= $<ref+1>.<attr> + c"   2: "change <= k - $2.change .. - $<state_num+1>.change" */
    int attr;
    int op; /* 0 <=, 1 >= */
    int ref;
    long long c;
    int state_num;
    int k;
    /* PredicateOptimized.checkSingle (:71-81): single iff no operand names another state by number */
    bool isSingleState() const { return kind == 0 || (kind == 2 && state_num == 0); }
    /* linkAggregationOperand (:57-69): the LAST operand with an aggregation gives relatedState / attributeName */
    int relatedState() const { return kind == 1 ? ref : state_num; }
    int attributeName() const { return kind == 1 ? attr : A_CHANGE; }
};

struct WNfa {
    int size = 0;
    std::vector<int> type;
    std::vector<std::vector<Pred>> preds;           /* State.edges[0].predicates */
    std::vector<std::vector<int>> templates;        /* NFA.valueVectors[state]: attribute of each ValueVectorTemplate */
    std::vector<bool> hasValueVector;
    bool needValueVector = false;

    /* WhyNotMatchSASE.getNFA (:91-108) + generatePredicatesFromConstraints (:136-158) + generateConstraintOfK (:118-127) */
    void build(const int32_t* pattern, int m, const siesta_wnm_constraint* cons, int n_cons, int k) {
        size = m;
        type.assign(pattern, pattern + m);
        preds.assign(m, {});
        for (int i = 0; i < m; ++i) {
            for (int q = 0; q < n_cons; ++q) {
                const siesta_wnm_constraint& c = cons[q];
                if (c.pos_b != i) continue;
                const int a = c.kind == SIESTA_WNM_GAP ? A_POSITION : A_TIMESTAMP;
                if (i > 0) preds[i].push_back(Pred{0, a, 0, 0, 0, 0, 0});
                preds[i].push_back(Pred{1, a, c.method == SIESTA_WNM_WITHIN ? 0 : 1, c.pos_a, c.value, 0, 0});
            }
            preds[i].push_back(Pred{2, A_CHANGE, 0, 0, 0, i, k});
        }
        /* NFA.compileValueVectorOptimized, S/query/NFA.java:415-474 */
        templates.assign(m, {});
        for (int i = 0; i < m; ++i)
            for (const Pred& p : preds[i])
                if (!p.isSingleState()) {
                    const int stateNumber = p.relatedState();
                    if (stateNumber < 0 || stateNumber >= m) throw TooLarge{}; /* ArrayIndexOutOfBounds in Java: callers validate */
                    templates[stateNumber].push_back(p.attributeName());
                    needValueVector = true;
                }
        hasValueVector.assign(m, false);
        for (int i = 0; i < m; ++i) hasValueVector[i] = !templates[i].empty();
    }
};

struct VVElem { /* ValueVectorElementSet: attribute + the value of the event it was initialised with */
    int attr;
    int value;
};
typedef std::shared_ptr<std::vector<VVElem>> VVState; /* valueVector[state], null until initialised */
struct VVArray {                                        /* the outer ValueVectorElement[][]: SHARED by every clone */
    std::vector<VVState> slot;
};

struct Run {
    std::vector<int> eventIds;
    std::vector<int> state;
    int size = 0, count = 0, cur = 0;
    bool isFull = false;
    std::shared_ptr<VVArray> vv;
    const WNfa* nfa = nullptr;

    void initializeRun(const WNfa* n) { /* S/engine/Run.java:130-158 */
        nfa = n;
        size = n->size;
        state.assign(size, 0);
        eventIds.clear();
        cur = 0;
        isFull = false;
        count = 0;
        vv = nullptr;
        if (n->needValueVector) {
            vv = std::make_shared<VVArray>();
            vv->slot.assign(size, nullptr);
        }
    }
    bool checkMatch() const { /* :181-191 */
        if (!isFull) return false;
        for (int v : state)
            if (v != 2) return false;
        return true;
    }
    void initializeValueVector(const UEvent& e) { /* :332-355 */
        auto v = std::make_shared<std::vector<VVElem>>();
        for (int a : nfa->templates[cur]) v->push_back(VVElem{a, attr_of(e, a)});
        vv->slot[cur] = v; /* written into the array every clone of this run's family shares */
    }
    int getNeededValueVector(int stateNumber, int attribute) const { /* :375-385 */
        if (!vv || !vv->slot[stateNumber]) throw NullPointer{};
        for (const VVElem& x : *vv->slot[stateNumber])
            if (x.attr == attribute) return x.value;
        return 0;
    }
    void addEvent(const UEvent& e) { /* :196-225 -> addEventToNormalorOr :247-262 (every state is "normal") */
        eventIds.push_back(e.position);
        state[cur] = 2;
        count++;
        int c = 0;
        for (int v : state) c += v == 2;
        if (cur == nfa->size - 1 || c == size) {
            isFull = true;
        } else {
            if (nfa->needValueVector && nfa->hasValueVector[cur]) initializeValueVector(e);
            cur++;
        }
    }
    int getPreviousEventId() const { return eventIds.at((size_t)count - 1); } /* :284-286 */
};
typedef std::shared_ptr<Run> RunP;

struct Engine {
    const WNfa& nfa;
    std::vector<RunP> activeRuns;
    std::unordered_map<int, const UEvent*> buffer;
    std::vector<std::vector<int>> matches; /* event ids, in emission order */
    size_t run_limit;

    Engine(const WNfa& n, size_t lim) : nfa(n), run_limit(lim) {}

    /* PredicateOptimized.evaluate(Event, Event), S/query/PredicateOptimized.java:302-323 (State.canStartWithEvent) */
    bool evaluate2(const Pred& p, const UEvent& cur, const UEvent& prev) const {
        (void)prev;
        switch (p.kind) {
            case 0: return attr_of(cur, p.attr) > attr_of(prev, p.attr);
            case 1: { /* the numbered operand is filled from the CURRENT event */
                const long long rhs = (long long)attr_of(cur, p.attr) + p.c;
                return p.op == 0 ? attr_of(cur, p.attr) <= rhs : attr_of(cur, p.attr) >= rhs;
            }
            default: {
                long long rhs = p.k;
                for (int t = 1; t <= p.state_num; ++t) rhs -= cur.change;
                return cur.change <= rhs;
            }
        }
    }
    /* PredicateOptimized.evaluate(Event, Run, EventBuffer), :331-368: the operands in order, with the early returns */
    bool evaluate3(const Pred& p, const UEvent& e, const Run& r) const {
        switch (p.kind) {
            case 0: {
                const UEvent* prev = buffer.at(r.getPreviousEventId());
                return attr_of(e, p.attr) > attr_of(*prev, p.attr);
            }
            case 1: {
                if (p.ref == r.cur) return true; /* stateNumber - 1 == r.getCurrentState() */
                int v;
                try {
                    v = r.getNeededValueVector(p.ref, p.attr);
                } catch (NullPointer&) {
                    return false;
                }
                const long long rhs = (long long)v + p.c;
                return p.op == 0 ? attr_of(e, p.attr) <= rhs : attr_of(e, p.attr) >= rhs;
            }
            default: {
                long long rhs = p.k;
                for (int t = 1; t <= p.state_num; ++t) { /* operands $2.change .. $(state_num + 1).change */
                    if (t == r.cur) return true;
                    try {
                        rhs -= r.getNeededValueVector(t, A_CHANGE);
                    } catch (NullPointer&) {
                        return false;
                    }
                }
                return e.change <= rhs;
            }
        }
    }
    void bufferEvent(const UEvent& e) {
        if (!buffer.count(e.position)) buffer[e.position] = &e;
    }
    /* Engine.checkPredicate, S/engine/Engine.java:1102-1163 for a "normal" state */
    bool checkPredicate(const UEvent& e, const Run& r) const {
        const int s = r.cur;
        if (nfa.type[s] != e.type) return false;
        for (const Pred& p : nfa.preds[s])
            if (!evaluate3(p, e, r)) return false;
        return true;
    }
    void outputMatch(const Run& r) { matches.push_back(r.eventIds); }
    /* Engine.evaluateEventForSkipTillAny, :593-645 (checkTimeWindow: the window is Integer.MAX_VALUE, NFAWrapper.java:21-25) */
    void evaluateEventForSkipTillAny(const UEvent& e, const RunP& r) {
        if (!checkPredicate(e, *r)) return;
        bufferEvent(e);
        RunP newRun = std::make_shared<Run>(*r); /* Run.clone :319-327: eventIds and state copied, valueVector shared */
        const int oldState = newRun->cur;
        newRun->addEvent(e);
        const int newState = newRun->cur;
        if (oldState != newState) {
            activeRuns.push_back(newRun);
            if (activeRuns.size() > run_limit) throw TooLarge{};
        } else if (newRun->isFull) {
            if (newRun->checkMatch()) outputMatch(*newRun);
        }
    }
    /* Engine.createNewRun, :933-998 with a "normal" first state (the trailing block :983-996 as the tests pin it: off) */
    void createNewRun(const UEvent& e) {
        if (nfa.type[0] != e.type) return; /* State.canStartWithEvent, S/query/State.java:302-314 */
        for (const Pred& p : nfa.preds[0])
            if (!evaluate2(p, e, e)) return;
        bufferEvent(e);
        RunP newRun = std::make_shared<Run>();
        newRun->initializeRun(&nfa);
        newRun->addEvent(e);
        if (newRun->checkMatch()) outputMatch(*newRun);
        else activeRuns.push_back(newRun);
    }
    /* Engine.runSkipTillAnyEngine, :159-178 */
    void run(const std::vector<UEvent>& stream) {
        for (const UEvent& e : stream) {
            const size_t size = activeRuns.size(); /* evaluateRunsForSkipTillAny :341-350 */
            for (size_t i = 0; i < size; ++i) {
                RunP r = activeRuns[i];
                if (r->isFull) continue;
                evaluateEventForSkipTillAny(e, r);
            }
            createNewRun(e);
        }
    }
};

/* WhyNotMatchSASE.getUnCertainStream, :63-83 */
std::vector<UEvent> uncertain_stream(const std::vector<int>& type, const std::vector<long long>& primary, const std::vector<int>& src,
                                     int uncertainty, int step) {
    std::vector<UEvent> buf;
    for (size_t q = 0; q < type.size(); ++q) {
        const long long original = primary[q];
        const long long minimum = std::max(original - uncertainty, 0LL);
        const long long maximum = original + uncertainty;
        for (long long i = minimum; i <= maximum; i += step)
            buf.push_back(UEvent{type[q], 0, (int)i, (int)std::llabs(original - i), src[q]});
    }
    std::stable_sort(buf.begin(), buf.end(), [](const UEvent& a, const UEvent& b) { return a.timestamp < b.timestamp; });
    for (size_t i = 0; i < buf.size(); ++i) buf[i].position = (int)i;
    return buf;
}

template <typename T>
T* dup(const std::vector<T>& v) {
    T* p = (T*)std::malloc(sizeof(T) * (v.empty() ? 1 : v.size()));
    if (!v.empty()) std::memcpy(p, v.data(), sizeof(T) * v.size());
    return p;
}

}  // namespace

extern "C" {

/* Size of the uncertain stream of one list of primary metrics (WhyNotMatchSASETest.testGetUncertainStream). */
int64_t oracle_wnm_stream_size(const int64_t* primary, int32_t n, int32_t uncertainty, int32_t step) {
    std::vector<int> type((size_t)n, 0), src((size_t)n, 0);
    std::vector<long long> prim(primary, primary + n);
    return (int64_t)uncertain_stream(type, prim, src, uncertainty, step).size();
}

/* WhyNotMatchSASE.evaluate (:37-55) + createResponse (:160-173) over the candidate traces of a CSR log.
 * Returns 0, or 1 when a trace needs more than run_limit runs (the caller asked for a case that is too large). */
int oracle_why_not_match(const int64_t* trace_off, const int32_t* act, const int64_t* ts_ms, int64_t n_traces,
                         const int32_t* pattern, int32_t m, const siesta_wnm_constraint* cons, int32_t n_cons,
                         int32_t uncertainty, int32_t step, int32_t k, const int64_t* cand, int64_t n_cand, uint32_t flags,
                         int64_t run_limit, siesta_almost_matches** out) {
    WNfa nfa;
    try {
        nfa.build(pattern, m, cons, n_cons, k);
    } catch (TooLarge&) {
        return 2;
    }
    std::vector<int64_t> tr;
    std::vector<int32_t> total, ev_pos, ev_value, ev_change, ev_spos;
    const int64_t n = cand ? n_cand : n_traces;
    for (int64_t ci = 0; ci < n; ++ci) {
        const int64_t t = cand ? cand[ci] : ci;
        std::vector<int> type, src;
        std::vector<long long> primary;
        for (int64_t i = trace_off[t]; i < trace_off[t + 1]; ++i) {
            bool in = false;
            for (int s = 0; s < m; ++s) in |= pattern[s] == act[i];
            if (!in) continue;
            type.push_back(act[i]);
            src.push_back((int)(i - trace_off[t]));
            /* Event.getPrimaryMetric: EventTs.java:87-89 (epoch ms / 1000), EventPos.java:84-86 (position) */
            primary.push_back((flags & SIESTA_F_EVT_POS) ? (long long)(i - trace_off[t]) : (long long)(ts_ms[i] / 1000));
        }
        if (type.empty()) continue;
        const std::vector<UEvent> stream = uncertain_stream(type, primary, src, uncertainty, step);
        Engine eng(nfa, (size_t)run_limit);
        try {
            eng.run(stream);
        } catch (TooLarge&) {
            return 1;
        }
        if (eng.matches.empty()) continue; /* .filter(x -> !x._2.isEmpty()) */
        /* createResponse: reduce((x, y) -> x._2 < y._2 ? x : y) - of equal sums the later match stays */
        size_t best = 0;
        long long best_sum = 0;
        for (size_t q = 0; q < eng.matches.size(); ++q) {
            long long sum = 0;
            for (int id : eng.matches[q]) sum += stream[(size_t)id].change;
            if (q == 0 || !(best_sum < sum)) {
                best = q;
                best_sum = sum;
            }
        }
        tr.push_back(t);
        total.push_back((int32_t)best_sum);
        for (int id : eng.matches[best]) {
            const UEvent& e = stream[(size_t)id];
            ev_pos.push_back(e.src);
            ev_value.push_back(e.timestamp);
            ev_change.push_back(e.change);
            ev_spos.push_back(e.position);
        }
    }
    siesta_almost_matches* r = (siesta_almost_matches*)std::calloc(1, sizeof(siesta_almost_matches));
    r->n_traces = (int64_t)tr.size();
    r->n_states = m;
    r->trace_idx = dup(tr);
    r->total_change = dup(total);
    r->ev_pos = dup(ev_pos);
    r->ev_value = dup(ev_value);
    r->ev_change = dup(ev_change);
    r->ev_stream_pos = dup(ev_spos);
    r->unsupported_trace_idx = (int64_t*)std::malloc(8);
    *out = r;
    return 0;
}

void oracle_almost_matches_free(siesta_almost_matches* r) {
    if (!r) return;
    std::free(r->trace_idx);
    std::free(r->total_change);
    std::free(r->ev_pos);
    std::free(r->ev_value);
    std::free(r->ev_change);
    std::free(r->ev_stream_pos);
    std::free(r->unsupported_trace_idx);
    std::free(r);
}

}  // extern "C"
