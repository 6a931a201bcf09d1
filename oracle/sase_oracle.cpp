/*
 * sase_oracle.cpp — CPU ORACLE (test infrastructure, NOT the product).
 *
 * A literal restatement of the reference's per-trace verification path so the
 * CUDA path can be checked bit-for-bit.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load this.
 *
 * Parity pins: the reference's own engine known-answer tests
 * (src/test/java/.../SaseConnection/EvaluateNewQueries.java,
 * EvaluateComplexQueries.java — 31 methods, 21 scenarios on the stream
 * "A B A C D A B E") are replayed by tests/test_oracle_kat.py.  The reference
 * itself cannot run here (no JVM), see DESIGN.md.
 *
 * Paths cited below are relative to the reference's src/main/java/:
 *   S/ = edu/umass/cs/sase/      J/ = com/datalab/siesta/queryprocessor/
 *
 * Every place where the Java code would throw (ArrayIndexOutOfBounds,
 * IndexOutOfBounds, NullPointerException) raises RefThrow here; the trace is
 * then reported in err_trace_idx (SaseConnector.java:60-62 turns it into a
 * RuntimeException that fails the whole request).
 */
#include "../include/siesta_gpu.h"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <thread>
#include <unordered_map>
#include <vector>

namespace {

struct RefThrow {
    const char* what;
};
thread_local const char* g_last_throw = "";

/* SaseEvent (J/SaseConnection/SaseEvent.java:18-153): id == position. */
struct OEvent {
    int type;       /* dense activity id (names already case-folded) */
    int position;   /* == getId()                                     */
    int timestamp;  /* int seconds, or list index for EventPos        */
    int src;        /* index of the event inside its CSR trace        */
    int rank;       /* index in the filtered list                     */
};

inline bool is_kleene(int kind) { /* State.isKleeneClosure: State.java:116, AdditionalState.java:35 */
    return kind == SIESTA_STATE_KLEENE_PLUS || kind == SIESTA_STATE_KLEENE_STAR;
}

/* ValueVectorElementSet per state (S/engine/ValueVectorElementSet.java); the
 * outer array is shared by every clone of a run (Run.clone is shallow:
 * S/engine/Run.java:319-327). */
struct ValueVector {
    const OEvent* slot[SIESTA_MAX_STATES];
    ValueVector() {
        for (auto& s : slot) s = nullptr;
    }
};

struct Nfa {
    siesta_nfa d;
    bool ignore_preds;
    bool need_vv;                     /* NFA.needValueVector, NFA.java:445-447 */
    bool has_vv[SIESTA_MAX_STATES];   /* NFA.hasValueVector,  NFA.java:462-469 */
    int size() const { return d.n_states; }
    int n_preds(int s) const { return ignore_preds ? 0 : d.states[s].n_preds; }
    /* NFA.compileValueVectorOptimized (NFA.java:415-474): a state gets a
     * value vector iff some predicate on any edge references it. */
    void compile() {
        need_vv = false;
        for (int i = 0; i < SIESTA_MAX_STATES; ++i) has_vv[i] = false;
        for (int i = 0; i < size(); ++i)
            for (int k = 0; k < n_preds(i); ++k) {
                int ref = d.states[i].preds[k].ref_state;
                if (ref < 0 || ref >= size()) throw RefThrow{"value vector template index out of bounds (NFA.java:436)"};
                has_vv[ref] = true;
                need_vv = true;
            }
    }
    /* State.checkEventType (State.java:135-137) / AdditionalState.checkEventType (:44-52) */
    bool check_type(int s, const OEvent& e) const {
        const siesta_state& st = d.states[s];
        if (st.kind == SIESTA_STATE_NORMAL || st.kind == SIESTA_STATE_KLEENE_PLUS) return st.types[0] == e.type;
        for (int i = 0; i < st.n_types; ++i)
            if (st.types[i] == e.type) return true;
        return false;
    }
    int n_edges(int s) const { return is_kleene(d.states[s].kind) ? 3 : 1; }
    /* nfa.getStates(i): a run whose cursor moved past a trailing negative state indexes states[size]
     * (Run.java:203 increments currentState unconditionally) -> ArrayIndexOutOfBounds */
    int kind_at(int s) const {
        if (s < 0 || s >= size()) throw RefThrow{"getStates(currentState) out of bounds (Engine.java:656,691,1104)"};
        return d.states[s].kind;
    }
};

struct Run {
    std::vector<int> eventIds;
    std::vector<int> state;
    int size = 0, count = 0, cur = 0;
    bool isFull = false, kinit = false, containsNegative = false;
    bool was_reset = false; /* resetRun nulls nfa: a second resetRun NPEs (Run.java:162-166) */
    std::shared_ptr<ValueVector> vv;
    const Nfa* nfa = nullptr;

    /* Run.initializeRun, Run.java:130-158 */
    void initialize(const Nfa* n) {
        nfa = n;
        size = n->size();
        state.assign(size, 0);
        eventIds.clear();
        cur = 0;
        isFull = false;
        count = 0;
        kinit = false;
        containsNegative = false;
        was_reset = false;
        vv = n->need_vv ? std::make_shared<ValueVector>() : nullptr;
        for (int i = 0; i < size; ++i)
            if (n->d.states[i].kind == SIESTA_STATE_NEGATIVE) state[i] = 2;
        for (int i = 0; i < size; ++i)
            if (n->d.states[i].kind == SIESTA_STATE_KLEENE_STAR) state[i] = 3;
    }
    /* Run.checkMatch, Run.java:181-191 */
    bool checkMatch() const {
        if (!isFull) return false;
        for (int v : state)
            if (v != 2) return false;
        return true;
    }
    /* Run.getPreviousEventId, Run.java:284-286: ArrayList.get(count-1) */
    int getPreviousEventId() const {
        if (count - 1 < 0 || count - 1 >= (int)eventIds.size()) throw RefThrow{"eventIds.get(count-1) out of bounds (Run.java:285)"};
        return eventIds[count - 1];
    }
    void initializeValueVector(const OEvent& e) { vv->slot[cur] = &e; } /* Run.java:332-355 */
    void updateValueVector(const OEvent& e) {                            /* Run.java:360-364 */
        if (vv->slot[cur] == nullptr) throw RefThrow{"updateValueVector on uninitialised state (Run.java:361)"};
        vv->slot[cur] = &e;
    }
    /* Run.addEventToNormalorOr, Run.java:247-262 */
    void addEventToNormalorOr(const OEvent& e) {
        eventIds.push_back(e.position);
        state[cur] = 2;
        count++;
        int c = 0;
        for (int v : state) c += (v == 2);
        if (cur == nfa->size() - 1 || c == size) {
            isFull = true;
        } else {
            if (nfa->need_vv && nfa->has_vv[cur]) initializeValueVector(e);
            cur++;
        }
    }
    /* Run.addEventToKleene, Run.java:264-278 */
    void addEventToKleene(const OEvent& e) {
        eventIds.push_back(e.position);
        if (nfa->need_vv && nfa->has_vv[cur]) {
            if (kinit) updateValueVector(e);
            else initializeValueVector(e);
        }
        kinit = true;
        state[cur] = 3;
        count++;
    }
    /* Run.addEventToNextState, Run.java:228-245 */
    void addEventToNextState(const OEvent& e) {
        if (cur + 1 >= nfa->size()) throw RefThrow{"getStates(currentState+1) out of bounds (Run.java:229)"};
        int nk = nfa->d.states[cur + 1].kind;
        if (nk == SIESTA_STATE_NORMAL || nk == SIESTA_STATE_OR) {
            state[cur] = 2;
            cur++;
            addEventToNormalorOr(e);
        } else if (is_kleene(nk)) {
            cur++;
            addEventToKleene(e);
        } else if (nk == SIESTA_STATE_NEGATIVE) {
            containsNegative = true;
            cur++;
        }
    }
    /* Run.addEvent, Run.java:196-225 (startTimeStamp is only used by the time
     * window, which SIESTA sets to Integer.MAX_VALUE: NFAWrapper.java:21-25) */
    void addEvent(const OEvent& e) {
        int k = nfa->d.states[cur].kind;
        if (k == SIESTA_STATE_NORMAL || k == SIESTA_STATE_OR) {
            addEventToNormalorOr(e);
        } else if (k == SIESTA_STATE_NEGATIVE) {
            if (nfa->check_type(cur, e)) {
                containsNegative = true;
                cur++;
            } else {
                addEventToNextState(e);
            }
        } else if (k == SIESTA_STATE_KLEENE_STAR) {
            if (nfa->check_type(cur, e)) addEventToKleene(e);
            else addEventToNextState(e);
        } else if (k == SIESTA_STATE_KLEENE_PLUS) {
            addEventToKleene(e);
        }
    }
    /* Run.proceed, Run.java:307-315 */
    void proceed() {
        state[cur] = 2;
        if (cur == size - 1) isFull = true;
        else cur++;
    }
};

using RunP = std::shared_ptr<Run>;

struct Match {
    std::vector<const OEvent*> events; /* nullptr = id never buffered (Match.java:62-71) */
};

int attr_of(const OEvent& e, int attr) { /* SaseEvent.getAttributeByName, SaseEvent.java:80-89 */
    return attr == SIESTA_ATTR_POSITION ? e.position : e.timestamp;
}

bool compare(int64_t lhs, int op, int64_t rhs) { /* jeval double compare of int32 values: exact */
    return op == SIESTA_OP_LE ? lhs <= rhs : lhs >= rhs;
}

struct Engine {
    const Nfa& nfa;
    bool head_mode;
    std::vector<RunP> activeRuns;
    std::vector<RunP> toDeleteRuns;
    std::unordered_map<int, const OEvent*> buffer; /* EventBuffer.java:51-62 */
    std::vector<Match> matches;

    Engine(const Nfa& n, bool head) : nfa(n), head_mode(head) {}

    void bufferEvent(const OEvent& e) {
        if (buffer.find(e.position) == buffer.end()) buffer[e.position] = &e;
    }
    /* PredicateOptimized.evaluate(Event, Event), PredicateOptimized.java:302-323:
     * the related-state operand is filled from the current event itself. */
    bool evalEdgeSelf(int s, int edge, const OEvent& e) const {
        if (edge >= nfa.n_edges(s)) throw RefThrow{"getEdges(i) out of bounds"};
        if (edge == 2) return true;
        for (int k = 0; k < nfa.n_preds(s); ++k) {
            const siesta_pred& p = nfa.d.states[s].preds[k];
            int64_t v = attr_of(e, p.attr);
            if (!compare(v, p.op, v + p.constant)) return false;
        }
        return true;
    }
    /* Edge.evaluatePredicate(Event, Run, EventBuffer) (Edge.java:122-134) over
     * PredicateOptimized.evaluate(Event, Run, EventBuffer) (:331-368). */
    bool evalEdge(int s, int edge, const OEvent* e, const Run& r) const {
        if (edge >= nfa.n_edges(s)) throw RefThrow{"getEdges(i) out of bounds (Engine.java:1215)"};
        if (edge == 2) return true; /* SIESTA never adds proceed-edge predicates */
        for (int k = 0; k < nfa.n_preds(s); ++k) {
            const siesta_pred& p = nfa.d.states[s].preds[k];
            if (e == nullptr) throw RefThrow{"predicate on null event"};
            int64_t lhs = attr_of(*e, p.attr);
            if (p.ref_state == r.cur) continue;       /* "return true" for this predicate (:348-350) */
            const OEvent* ref = r.vv ? r.vv->slot[p.ref_state] : nullptr;
            if (ref == nullptr) return false;          /* NPE caught -> false (:353-357)              */
            if (!compare(lhs, p.op, (int64_t)attr_of(*ref, p.attr) + p.constant)) return false;
        }
        return true;
    }
    /* State.canStartWithEvent (State.java:302-314) / AdditionalState (:54-62) */
    bool canStartWithEvent(int s, const OEvent& e) const {
        if (!nfa.check_type(s, e)) return false;
        return evalEdgeSelf(s, 0, e);
    }
    /* Engine.checkProceed, Engine.java:1205-1224 */
    bool checkProceed(const Run& r) const {
        int cur = r.cur;
        int prevId = r.getPreviousEventId();
        auto it = buffer.find(prevId);
        const OEvent* prev = it == buffer.end() ? nullptr : it->second;
        int kind = nfa.kind_at(cur); /* State s = nfa.getStates(currentState), Engine.java:1210 */
        if (r.state[cur] == 0 && kind == SIESTA_STATE_KLEENE_PLUS) return false;
        return evalEdge(cur, 2, prev, r);
    }
    /* Engine.checkPredicatesForNextState, Engine.java:1165-1180 */
    bool checkPredicatesForNextState(int cur, const OEvent& e, const Run& r) const {
        int s2 = cur + 1;
        if (nfa.check_type(s2, e)) {
            if (!is_kleene(nfa.d.states[cur].kind)) return evalEdge(s2, 0, &e, r);
            return r.kinit ? evalEdge(s2, 1, &e, r) : evalEdge(s2, 0, &e, r);
        }
        return false;
    }
    /* Engine.checkPredicate, Engine.java:1102-1163 */
    bool checkPredicate(const OEvent& e, const Run& r) const {
        int cur = r.cur;
        int kind = nfa.kind_at(cur);
        if (kind == SIESTA_STATE_NEGATIVE) {
            if (nfa.check_type(cur, e)) return evalEdge(cur, 0, &e, r);
            else if (nfa.size() > cur + 1) return checkPredicatesForNextState(cur, e, r);
        }
        if (kind == SIESTA_STATE_KLEENE_STAR) {
            if (nfa.check_type(cur, e)) return r.kinit ? evalEdge(cur, 1, &e, r) : evalEdge(cur, 0, &e, r);
        }
        if (!nfa.check_type(cur, e)) return false;
        if (!is_kleene(kind)) return evalEdge(cur, 0, &e, r);
        return r.kinit ? evalEdge(cur, 1, &e, r) : evalEdge(cur, 0, &e, r);
    }
    void outputMatch(const Run& r) { /* Match(Run, NFA, EventBuffer), Match.java:62-71 */
        Match m;
        for (int i = 0; i < r.count; ++i) {
            if (i >= (int)r.eventIds.size()) throw RefThrow{"Match: eventIds.get(i) out of bounds"};
            auto it = buffer.find(r.eventIds[i]);
            m.events.push_back(it == buffer.end() ? nullptr : it->second);
        }
        matches.push_back(std::move(m));
    }
    RunP cloneRun(const Run& r) { /* Engine.cloneRun :1283-1288 -> Run.clone :319-327 (vv shared) */
        return std::make_shared<Run>(r);
    }
    /* Engine.evaluateEventForSkipTillNext, Engine.java:654-725 */
    void evaluateEventForSkipTillNext(const OEvent& e, RunP rp) {
        Run& r = *rp;
        int ts = r.cur;
        if (nfa.kind_at(ts) == SIESTA_STATE_KLEENE_STAR && !r.kinit && checkProceed(r) &&
            nfa.d.states[ts].types[0] == e.type) {
            RunP nr = cloneRun(r);
            nr->proceed();
            if (nr->checkMatch()) outputMatch(r); /* emits r's ids, not the clone's (:665) */
            else activeRuns.push_back(nr);
        }
        if (checkPredicate(e, r)) {
            /* checkTimeWindow: timeWindow == Integer.MAX_VALUE -> always true */
            bufferEvent(e);
            r.addEvent(e);
            if (r.containsNegative) toDeleteRuns.push_back(rp);
            if (r.isFull) {
                if (r.checkMatch()) {
                    outputMatch(r);
                    toDeleteRuns.push_back(rp);
                }
            }
            ts = r.cur;
            if (is_kleene(nfa.kind_at(ts))) {
                if (checkProceed(r)) {
                    RunP nr = cloneRun(r);
                    nr->kinit = true;
                    activeRuns.push_back(nr);
                    r.proceed();
                    if (r.checkMatch()) {
                        outputMatch(r);
                        toDeleteRuns.push_back(rp);
                    }
                }
            }
        }
    }
    /* Engine.createNewRun, Engine.java:933-998 */
    void createNewRun(const OEvent& e) {
        int k0 = nfa.d.states[0].kind;
        if (canStartWithEvent(0, e)) {
            if (is_kleene(k0)) {
                bufferEvent(e);
                RunP nr = std::make_shared<Run>();
                nr->initialize(&nfa);
                nr->addEvent(e);
                if (checkProceed(*nr)) {
                    nr->proceed();
                    activeRuns.push_back(nr);
                }
            }
            bufferEvent(e);
            RunP nr = std::make_shared<Run>();
            nr->initialize(&nfa);
            nr->addEvent(e);
            if (nr->checkMatch()) outputMatch(*nr);
            else activeRuns.push_back(nr);
        } else if (k0 == SIESTA_STATE_KLEENE_STAR || k0 == SIESTA_STATE_NEGATIVE) {
            if (nfa.size() > 1) {
                if (canStartWithEvent(1, e)) {
                    bufferEvent(e);
                    RunP nr = std::make_shared<Run>();
                    nr->initialize(&nfa);
                    nr->addEvent(e);
                    if (nfa.kind_at(nr->cur) == SIESTA_STATE_KLEENE_STAR && checkProceed(*nr)) nr->proceed();
                    if (nr->checkMatch()) outputMatch(*nr);
                    else activeRuns.push_back(nr);
                }
            }
        }
        if (head_mode) { /* trailing block, Engine.java:983-996; off in the mode the reference's tests pin */
            if (nfa.size() < 2) throw RefThrow{"getStates(1) out of bounds (Engine.java:985)"};
            if (nfa.d.states[1].kind == SIESTA_STATE_KLEENE_STAR) {
                RunP nr = std::make_shared<Run>();
                nr->initialize(&nfa);
                nr->addEvent(e); /* not buffered */
                checkProceed(*nr);
                nr->proceed();
                if (nr->checkMatch()) outputMatch(*nr);
                else activeRuns.push_back(nr);
            }
        }
    }
    /* Engine.cleanRuns, Engine.java:1404-1418 */
    void cleanRuns() {
        for (auto& rp : toDeleteRuns) {
            if (rp->was_reset) throw RefThrow{"resetRun on a reset run: nfa is null (Run.java:163)"};
            rp->was_reset = true;
            auto it = std::find(activeRuns.begin(), activeRuns.end(), rp);
            if (it != activeRuns.end()) activeRuns.erase(it);
        }
        toDeleteRuns.clear();
    }
    /* Engine.runSkipTillNextEngine, Engine.java:207-224 */
    void run(const std::vector<OEvent>& stream) {
        for (const OEvent& e : stream) {
            size_t size = activeRuns.size(); /* evaluateRunsForSkipTillNext :360-369 */
            for (size_t i = 0; i < size; ++i) {
                RunP r = activeRuns[i];
                if (r->isFull) continue;
                evaluateEventForSkipTillNext(e, r);
            }
            if (!toDeleteRuns.empty()) cleanRuns();
            createNewRun(e);
        }
    }
};

/* Occurrence.overlaps, J/model/Occurrence.java:36-49.  `a` is this, `b` the argument. */
bool overlaps(const Match& a, const Match& b, bool by_position) {
    if (b.events.empty() || a.events.empty()) throw RefThrow{"Occurrence.overlaps on empty occurrence"};
    const OEvent* f = b.events.front();
    const OEvent* l = b.events.back();
    const OEvent* af = a.events.front();
    const OEvent* al = a.events.back();
    bool notOverlaps;
    if (by_position) notOverlaps = al->position < f->position || af->position > l->position;
    else notOverlaps = al->timestamp < f->timestamp || af->timestamp > l->timestamp; /* Timestamp.before/after on ts*1000+minTs */
    return !notOverlaps;
}

/* Occurrences.clearOccurrences, J/model/Occurrences.java:58-89 */
std::vector<int> clearOccurrences(const std::vector<Match>& occ, bool returnAll, bool by_position) {
    std::vector<int> response;
    int e = 0;
    for (int i = 1; i < (int)occ.size(); ++i)
        if (occ[i].events.size() > occ[e].events.size()) e = i;
    response.push_back(e);
    if (!returnAll) return response;
    for (int i = 1; i < (int)occ.size(); ++i) {
        bool ov = false;
        for (int o : response)
            if (overlaps(occ[i], occ[o], by_position)) {
                ov = true;
                break;
            }
        if (!ov) response.push_back(i);
    }
    return response;
}

struct TraceResult {
    int status = 0; /* 0 none, 1 match, 2 reference throws */
    int64_t n_emitted = 0;
    std::vector<std::vector<const OEvent*>> selected;
    int64_t t0_ms = 0;
};

/* Utils.transformToSaseEvents (J/model/Utils/Utils.java:48-65) over the trace
 * filtered to the pattern's event types (Trace.clearTrace, J/model/DBModel/
 * Trace.java:53-59; SparkDatabaseRepository.querySeqTable :94-107). */
void build_stream(const Nfa& nfa, const int32_t* act, const int64_t* ts_ms, int64_t lo, int64_t hi, bool evt_pos,
                  std::vector<OEvent>& out, int64_t& t0) {
    out.clear();
    t0 = 0;
    int rank = 0;
    for (int64_t i = lo; i < hi; ++i) {
        bool rel = false;
        for (int s = 0; s < nfa.size() && !rel; ++s)
            for (int k = 0; k < nfa.d.states[s].n_types && !rel; ++k) rel = nfa.d.states[s].types[k] == act[i];
        if (!rel) continue;
        OEvent e;
        e.type = act[i];
        e.src = (int)(i - lo);
        e.rank = rank;
        if (evt_pos) { /* EventPos.transformSaseEvent, EventPos.java:51-55 */
            e.position = e.src;
            e.timestamp = rank;
        } else {       /* EventTs.transformSaseEvent, EventTs.java:52-58 */
            if (rank == 0) t0 = ts_ms[i];
            e.position = rank;
            e.timestamp = rank == 0 ? 0 : (int)((ts_ms[i] - t0) / 1000);
        }
        out.push_back(e);
        rank++;
    }
}

void run_trace(const Nfa& nfa, const std::vector<OEvent>& stream, uint32_t flags, TraceResult& res) {
    res = TraceResult();
    if (stream.empty()) return; /* SaseConnector.java:53-55 */
    try {
        Engine eng(nfa, (flags & SIESTA_F_MODE_HEAD) != 0);
        eng.run(stream);
        res.n_emitted = (int64_t)eng.matches.size();
        if (eng.matches.empty()) return;
        for (const Match& m : eng.matches)
            for (const OEvent* ev : m.events)
                if (ev == nullptr) throw RefThrow{"SaseEvent::getEventBoth on null (SaseConnector.java:67-70)"};
        std::vector<int> sel = clearOccurrences(eng.matches, (flags & SIESTA_F_RETURN_ALL) != 0, (flags & SIESTA_F_EVT_POS) != 0);
        for (int i : sel) res.selected.push_back(eng.matches[i].events);
        res.status = 1;
    } catch (const RefThrow& rt) {
        res = TraceResult();
        res.status = 2;
        g_last_throw = rt.what;
    }
}

template <class T>
T* dup(const std::vector<T>& v) {
    T* p = (T*)std::malloc(sizeof(T) * (v.size() ? v.size() : 1));
    if (!v.empty()) std::memcpy(p, v.data(), sizeof(T) * v.size());
    return p;
}

}  // namespace

extern "C" {

/* Engine-level entry used by the KAT pins: one explicit stream, every match in
 * emission order.  status: 0 ok, 2 reference throws.  Caller frees with
 * oracle_free. match_ids holds event ids (-1 = null event). */
int oracle_run_stream(const siesta_nfa* nfa_desc, const int32_t* type, const int32_t* id, const int32_t* ts, int32_t n,
                      uint32_t flags, int32_t* status, int64_t* n_matches, int64_t** match_off, int32_t** match_ids) {
    std::vector<int64_t> off{0};
    std::vector<int32_t> ids;
    *status = 0;
    try {
        Nfa nfa;
        nfa.d = *nfa_desc;
        nfa.ignore_preds = (flags & SIESTA_F_ONLY_APPEARANCES) != 0;
        nfa.compile();
        std::vector<OEvent> stream(n);
        for (int i = 0; i < n; ++i) stream[i] = OEvent{type[i], id[i], ts[i], i, i};
        Engine eng(nfa, (flags & SIESTA_F_MODE_HEAD) != 0);
        eng.run(stream);
        for (const Match& m : eng.matches) {
            for (const OEvent* e : m.events) ids.push_back(e ? e->position : -1);
            off.push_back((int64_t)ids.size());
        }
    } catch (const RefThrow&) {
        *status = 2;
        off.assign(1, 0);
        ids.clear();
    }
    *n_matches = (int64_t)off.size() - 1;
    *match_off = dup(off);
    *match_ids = dup(ids);
    return 0;
}

void oracle_free(void* p) { std::free(p); }

/* message of the last modelled Java exception on this thread (debugging aid) */
const char* oracle_last_throw(void) { return g_last_throw; }

/* SaseConnector.evaluate + clearOccurrences over a CSR log; fills the same
 * siesta_matches layout the product returns.  n_threads > 1 splits the trace
 * loop over std::threads (the reference's loop is single-threaded:
 * SaseConnector.java:51-74). */
int oracle_detect(const int64_t* trace_off, const int32_t* act, const int64_t* ts_ms, int64_t n_traces,
                  const siesta_nfa* nfa_desc, const int64_t* cand, int64_t n_cand, uint32_t flags, int32_t n_threads,
                  siesta_matches** out) {
    Nfa nfa;
    nfa.d = *nfa_desc;
    nfa.ignore_preds = (flags & SIESTA_F_ONLY_APPEARANCES) != 0;
    try {
        nfa.compile();
    } catch (const RefThrow&) {
        return SIESTA_E_INVALID;
    }
    const bool evt_pos = (flags & SIESTA_F_EVT_POS) != 0;
    const int64_t n = cand ? n_cand : n_traces;
    std::vector<TraceResult> results(n);
    std::vector<std::vector<OEvent>> streams(n); /* selected pointers refer into these */
    auto work = [&](int64_t lo, int64_t hi) {
        for (int64_t i = lo; i < hi; ++i) {
            int64_t t = cand ? cand[i] : i;
            int64_t t0 = 0;
            build_stream(nfa, act, ts_ms, trace_off[t], trace_off[t + 1], evt_pos, streams[i], t0);
            run_trace(nfa, streams[i], flags, results[i]);
            results[i].t0_ms = t0;
        }
    };
    auto t_begin = std::chrono::steady_clock::now();
    if (n_threads <= 1 || n < 2) {
        work(0, n);
    } else {
        std::vector<std::thread> th;
        int64_t chunk = (n + n_threads - 1) / n_threads;
        for (int k = 0; k < n_threads; ++k) {
            int64_t lo = k * chunk, hi = std::min(n, lo + chunk);
            if (lo < hi) th.emplace_back(work, lo, hi);
        }
        for (auto& t : th) t.join();
    }
    double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_begin).count();

    std::vector<int64_t> trace_idx, occ_off{0}, ev_off{0}, ev_ts, err;
    std::vector<int32_t> ev_pos, ev_rank, ev_act;
    int64_t emitted = 0;
    for (int64_t i = 0; i < n; ++i) {
        int64_t t = cand ? cand[i] : i;
        const TraceResult& r = results[i];
        if (r.status == 2) {
            err.push_back(t);
            continue;
        }
        emitted += r.n_emitted;
        if (r.status != 1) continue;
        trace_idx.push_back(t);
        for (const auto& occ : r.selected) {
            for (const OEvent* e : occ) {
                ev_pos.push_back(e->src);
                ev_rank.push_back(e->rank);
                ev_act.push_back(e->type);
                /* SaseEvent.getEventBoth: timestamp*1000 + minTs when isTimestampSet (SaseEvent.java:94-106) */
                ev_ts.push_back(evt_pos ? ts_ms[trace_off[t] + e->src] : (int64_t)e->timestamp * 1000 + r.t0_ms);
            }
            ev_off.push_back((int64_t)ev_pos.size());
        }
        occ_off.push_back((int64_t)ev_off.size() - 1);
    }
    siesta_matches* m = (siesta_matches*)std::calloc(1, sizeof(siesta_matches));
    m->n_traces = (int64_t)trace_idx.size();
    m->n_occurrences = (int64_t)ev_off.size() - 1;
    m->n_events = (int64_t)ev_pos.size();
    m->n_matches_emitted = emitted;
    m->n_ref_errors = (int64_t)err.size();
    m->trace_idx = dup(trace_idx);
    m->occ_off = dup(occ_off);
    m->ev_off = dup(ev_off);
    m->ev_pos = dup(ev_pos);
    m->ev_rank = dup(ev_rank);
    m->ev_act = dup(ev_act);
    m->ev_ts_ms = dup(ev_ts);
    m->err_trace_idx = dup(err);
    m->kernel_ms = ms;
    m->detect_ms = ms;
    *out = m;
    return 0;
}

void oracle_matches_free(siesta_matches* m) {
    if (!m) return;
    std::free(m->trace_idx);
    std::free(m->occ_off);
    std::free(m->ev_off);
    std::free(m->ev_pos);
    std::free(m->ev_rank);
    std::free(m->ev_act);
    std::free(m->ev_ts_ms);
    std::free(m->err_trace_idx);
    std::free(m);
}

}  // extern "C"
