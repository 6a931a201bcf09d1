// detect_fast.cuh — closed-form evaluators of kernel K1 for two NFA classes (device + host harness).
//
// The general RunEngine (detect_engine.cuh) replays the reference's run list event by event.  For two classes of
// NFAs the selected occurrences have a closed form that needs no run list at all; both forms are derived from the
// same reference code the engine restates (S/engine/Engine.java:654-725, 933-982, 1102-1224; S/engine/Run.java:
// 196-327; J/model/Occurrences.java:58-89) and are checked bit for bit against the oracle by
// tests/test_fast_paths_host_vs_oracle.py and tests/soak_fast.py.  validate_nfa (nfa.cpp) decides the class; every
// NFA outside the two classes runs on the general engine.
//
// Class NK  — only normal / or / negative states (no Kleene state), first and last state positive, no two
//             negative states in a row, every predicate references an earlier positive state.
//   Without Kleene states a run never forks: each event of state 0's types starts one run, and a run takes, for each
//   state in turn, the FIRST later event of that state's types that passes the state's predicates (events that fail
//   are skipped, Engine.java:1102-1163).  At a negative state the first later event that is either of the negative
//   type (and passes its predicates: the run is deleted, Engine.java:679-682) or of the next state's types (and
//   passes the next state's predicates, Engine.java:1165-1180: the run moves on) decides.  Runs never change their
//   list position, so matches are emitted in ascending (completion event, start event) order; they all have the same
//   number of events, hence "first largest" (Occurrences.java:60-69) = the smallest (completion, start).
//   Monotone walks: if every attribute a predicate reads never decreases along the events (list index and in-trace
//   position always; relative seconds only when the caller's timestamps are sorted, which is not assumed), a later
//   start takes, state by state, events that are not earlier than those of an earlier start (a lower bound `>=` only
//   grows with the referenced event; an upper bound `<=` that an earlier start fails on an event, it fails on every
//   later event as well, so that start never completes).  Completed runs therefore complete in start order, and the
//   first-largest occurrence is the FIRST start whose walk completes: `first_only` stops there.
//   The evaluator is generic in the event accessor: TraceEvents (filtered list in shared memory, masks over list
//   indices) or PosEvents (masks over the trace's raw 64 position slots, list index = popcount of the relevant-event
//   mask below the slot) for kernel K1-P, which needs no shared memory at all.
//
// Class FK2 — two states `a+ b*` (kleeneClosure, kleeneClosure*), one type each, a != b, predicates only on state 1
//             and only referencing state 0, first-largest occurrence only (returnAll = false).
//   Let A[0..M] be the a-events.  The run list only ever holds (Engine.createNewRun :933-982 creates two runs per
//   a-event, each with its own value vector):
//     P(i)   = [a_i] proceeded at once to state 1; its value vector keeps a_i for ever, so it takes every later b with
//              pass(b, a_i);
//     Z(i)   = [a_i ..] staying in state 0; it takes EVERY later a (state 0 has no predicates) and forks each time:
//              the proceeding branch Q(i,m) = [a_i..a_m] keeps the list slot, the staying clone goes to the tail
//              (Engine.java:704-708);
//     Q(i,m) shares Z(i)'s value vector (Run.clone is shallow, Run.java:319-327), whose state-0 slot is the LATEST a
//              of the stream, so every Q takes exactly the b's with pass(b, last a before b) ("good" b's).
//   A state-1 run that takes b emits [run + b] and is replaced by a clone at the TAIL of the list (Engine.java:
//   704-713).  Q(i,m), i > 0, is always smaller than Q(0,m), so the first-largest occurrence is among P(i) and Q(0,m),
//   each reaching its final size at the last b it takes.  Ties (same size at the same event) go to the run that sits
//   earlier in the list.  List order = order of last placement at the tail; two runs placed at the same b keep their
//   previous order.  Hence: with pl(X) = the set of events at which X was placed (its creation slot + every b it
//   took), X precedes Y at event k iff the highest event below k in pl(X) xor pl(Y) belongs to pl(Y).  Q(0,m) is
//   created in Z(0)'s slot, which was placed at a_{m-1}; at a_q, q >= 1, Z(0) is re-appended before P(q) is created,
//   at a_0 P(0) is appended first.
//
// Class NP1 — normal / or states and exactly ONE kleeneClosure+ state (index k; k = 0 needs a second state; kleeneClosure*:
//             see the end of this block), no predicate on any state (SIESTA's pattern without constraints, or any pattern under onlyAppearances:
//             ComplexPattern.getNfaWithoutConstraints :253-283), e.g. `a b+ c`.  returnAll only on the EventPos route.
//   Without predicates there is no value vector, so runs do not interact; only their list order matters, for ties.
//   k >= 1: the run started at an event of state 0's types walks the prefix greedily (first later event of each state's
//   types) and reaches the Kleene state after its prefix end p.  There it takes EVERY later event b_1 < b_2 < ... of the
//   Kleene type; each time the run that took b_m moves on (Run.proceed) and a clone stays behind at the tail of the
//   list and takes b_{m+1} (Engine.java:702-713).  The run that moved on after b_m walks the suffix greedily from b_m.
//   The engine's matches are therefore (start i, m) = prefix_i + {b_1..b_m} + suffix(b_m) for every m whose suffix
//   completes, with n_states - 1 + m events.  Whether the suffix completes from a position is monotone: it does iff the
//   position lies below l = the state-(k+1) event of the LATEST embedding of the suffix (backward greedy); F = the
//   positions below l (everything when the Kleene state is the last one: the run is full when it moves on).
//   Largest: prefix ends never decrease with the start, so the first start has the most Kleene events after its prefix,
//   m* = |T[k] & above(p_0) & F| (no start completes a prefix if the first one does not).  Another start ties only if no
//   Kleene event lies between the two prefix ends; both then take the same b's and complete at the same event, and the
//   first start's run sits earlier in the list: (i, 1) is the start run itself, placed at its start event; (i, m) is
//   the clone placed while b_{m-1} was processed, in the list order of the runs it was cloned from.  By induction the
//   first start precedes.  Engine matches: the sum over the starts of |T[k] & above(p_i) & F|.
//   k = 0: Engine.createNewRun :933-982 creates two runs per event a_j of the Kleene type, [a_j] moved on to state 1 and
//   [a_j] staying, which takes every later a_l and leaves [a_j..a_l] at state 1 each time: matches (j, l >= j) =
//   {a_j..a_l} + suffix(a_l); the largest, {all a up to the last one in F} + suffix, is unique.
//   Constraints that reference PREFIX states (k >= 1, every predicate's referenced state < k, none on state 0; first-largest
//   only): the value vector is shared by a start run and all its clones (Run.clone is shallow), but the prefix slots
//   never change once the prefix is complete and nobody writes the others (a slot is written only for referenced
//   states), so the family of start i behaves as above on its own filtered masks: prefix walk with predicates as in
//   class NK, B_i = the later Kleene events that pass state k's predicates against prefix_i, suffix states take the first
//   later event that passes theirs.  The B_i are no longer nested, so all starts are evaluated: the largest m wins, then
//   the earlier completion event, then the earlier place in the run list.  (i, m) sits where its lineage was last placed:
//   pl(i, m) = {a_i, b_1 .. b_{m-1}}; a later placement is later in the list, clones placed at the same event keep their
//   parents' order, and a start run is appended after the clones of its event (Engine.java:207-224: createNewRun follows
//   the loop): with h = the highest event in pl(X) xor pl(Y), say in pl(X): Y precedes unless pl(Y) has nothing below h.
//   (A predicate that references the Kleene state or a suffix state reads a slot that siblings overwrite: not taken.)
//   returnAll (Occurrences.java:74-87), without constraints: every match starts at or after the largest one's first event and completes at
//   or before its last (suffix completion is monotone in the position it starts from), so on the EventPos route, where
//   Occurrence.overlaps compares positions, every other match overlaps it: the selection is the largest alone.  On the
//   EventTs route overlaps compares timestamps, which the caller may hand over unsorted: not taken.
//   kleeneClosure* instead of kleeneClosure+ (one type, NO predicate anywhere; the shapes of the reference's own tests: `a b* e`,
//   `a* b`, `a* b e`, `a b*`, `a b a*`, `(a|b) b* e`, EvaluateComplexQueries.java:101-103, 126-127, 150-152, 175-176, 199-201,
//   298-301).  A run at the state takes events exactly as above (Engine.checkPredicate :1102-1163 rejects every event of another
//   type for both kinds); what differs is how a run gets PAST the state without a Kleene event:
//   k = 1: the start run is created at state 1 by Engine.createNewRun (no proceed block there).  When it is offered its first
//     event b_1 of the Kleene type, the block at Engine.java:658-670 (kleeneClosure*, not initialised) appends a clone that
//     skipped the state; the clone walks the suffix from behind b_1: one more match per start, (i, 0) = prefix_i +
//     suffix(b_1), iff b_1 lies in F, i.e. iff (i, 1) exists.  It is smaller than (i, 1): the selection is the `+` one.  A trace
//     without any event of the Kleene type after the prefix has NO match (`a e` does not match `a b* e`).
//   k >= 2: the run enters the state inside evaluateEventForSkipTillNext, and the proceed block behind addEvent (:691-713) fires
//     at once (Engine.checkProceed :1205-1224 only holds back an untouched kleeneClosure+ state): the run itself moves on from
//     its prefix end p_i (k last: it is a match at once), the clone stays, initialised, and takes the b's as above.  One more
//     match per start, (i, 0) = prefix_i + suffix(p_i), iff p_i lies in F - with or without Kleene events.  If the first
//     start has none the suffix allows, every match has S - 1 events and the first emitted wins: completion is monotone in the
//     start and the first start's run sits first in the list: (0, 0).
//   k = 0: createNewRun :960-982 starts a run at state 1 for an event that is of state 1's types and not of the Kleene type (state 0
//     counts as passed); those runs never fork and emit S - 1 events, fewer than any match that holds an a.  They are
//     selected only when no a lies in F; the first of them (the earliest embedding of states 1 .. S-1) is then the first emitted.
//   returnAll: (i, 0) of k >= 2 and the runs started at state 1 may lie outside the largest match: only k = 1 is taken.
#pragma once
#include "detect_engine.cuh"

namespace siesta {

enum { FAST_NONE = 0, FAST_NK = 1, FAST_FK2 = 2, FAST_NP1 = 3 };

template <int W>
struct MaskX : MaskOps<W> {
    typedef typename MaskOps<W>::T T;
    static SIESTA_HD __forceinline__ T below(int j) { return ~(~(T)0 << j); }   // bits < j   (0 <= j < bits of T)
    static SIESTA_HD __forceinline__ T above(int j) { return ~(T)1 << j; }      // bits > j
};

// "attr(e) op attr(ref) + c" for every predicate of state s; vv(ref) gives the event stored for state `ref`.
// A predicate that references state `cur` itself is true (PredicateOptimized.java:331-340).
// Events of a trace addressed by raw slot: slot j = in-trace position j - lead (lead = events of the previous trace
// that share the first 32-byte sector); R = slots that hold an event of the pattern's types.  The SaseEvent attributes
// after Utils.transformToSaseEvents (J/model/Utils/Utils.java:48-65): EventTs route position = index in the filtered
// list (= rank); EventPos route position = in-trace position, timestamp = index in the filtered list.  Relative
// seconds are not available here: NFAs with a time predicate on the EventTs route never take this path.
struct PosEvents {
    unsigned long long R;
    int lead;
    bool evt_pos;
    SIESTA_HD __forceinline__ int rank(int j) const { return popc64(R & ((1ull << j) - 1ull)); }
    SIESTA_HD __forceinline__ int position(int j) const { return evt_pos ? j - lead : rank(j); }
    SIESTA_HD __forceinline__ int timestamp(int j) const { return evt_pos ? rank(j) : 0; }
    SIESTA_HD __forceinline__ int attr(int j, int a) const { return a == SIESTA_ATTR_POSITION ? position(j) : timestamp(j); }
};

template <class EV, class VV>
SIESTA_HD __forceinline__ bool fast_preds(const DevNfa& nfa, const EV& ev, int s, int e, int cur, const VV& vv) {
    const int np = nfa.n_preds[s];
    for (int k = 0; k < np; ++k) {
        const int ref = nfa.p_ref[s][k];
        if (ref == cur) continue;
        const int rj = vv(ref);
        const int a = nfa.p_attr[s][k];
        const long long lhs = ev.attr(e, a);
        const long long rhs = (long long)ev.attr(rj, a) + nfa.p_c[s][k];
        if (nfa.p_op[s][k] == SIESTA_OP_LE ? !(lhs <= rhs) : !(lhs >= rhs)) return false;
    }
    return true;
}

// ------------------------------------------------------------------------------------------------ class FK2
// Streaming form: kernel K1 feeds on_event() from the pass that builds the trace's shared-memory columns, so the
// events are not read back; finish() then needs the columns only for P(i).
// NPMAX = compile-time bound of state 1's predicate count (1 covers the usual single constraint; 4 = any)
template <int W, int NPMAX = SIESTA_MAX_PREDS>
struct Fk2Eval {
    typedef MaskX<W> MO;
    typedef typename MO::T mask_t;
    // state 1's predicates, decoded once: where the attribute of event j lives (0 = j itself, 1 = in-trace position,
    // 2 = relative seconds, 3 = constant 0; SaseEvent attributes after Utils.transformToSaseEvents, see TraceEvents)
    int np;
    int p_mode[NPMAX];
    bool p_le[NPMAX];
    long long p_c[NPMAX];
    // a's, b's that follow an a, and the good b's (pass against the latest a before them).
    // mono_le: every predicate is a `within` (<=) and its attribute never decreases along the events; then
    // pass(b, a_i) implies good(b), so P(i), i >= 1, is always smaller than Q(0,i) and only P(0) can be selected.
    mask_t am, bm, good;
    bool seen_a, mono_le;
    int prev_v[NPMAX];
    long long la_v[NPMAX];

    SIESTA_HD __forceinline__ void init(const DevNfa& nfa, bool evt_pos, bool has_ts) {
        np = nfa.n_preds[1];
        am = bm = good = 0;
        seen_a = false;
        mono_le = true;
#pragma unroll
        for (int k = 0; k < NPMAX; ++k) {
            const bool is_pos = nfa.p_attr[1][k] == SIESTA_ATTR_POSITION;
            p_mode[k] = is_pos ? (evt_pos ? 1 : 0) : (evt_pos ? 0 : (has_ts ? 2 : 3));
            p_le[k] = nfa.p_op[1][k] == SIESTA_OP_LE;
            p_c[k] = nfa.p_c[1][k];
            prev_v[k] = -2147483647 - 1;
            la_v[k] = 0;
            if (k < np) mono_le = mono_le && p_le[k];
        }
    }
    // event j of the filtered list: lut word, in-trace position, relative seconds
    SIESTA_HD __forceinline__ void on_event(int j, uint32_t w, int src, int rel) {
        const bool is_a = w & 1u;
        const bool is_b = !is_a && (w & 2u) && seen_a;  // b's before the first a meet no run
        bool ok = true;
#pragma unroll
        for (int k = 0; k < NPMAX; ++k) {
            if (k < np) {
                const int m = p_mode[k];
                const int vj = m == 2 ? rel : (m == 1 ? src : (m == 0 ? j : 0));
                mono_le = mono_le && vj >= prev_v[k];
                prev_v[k] = vj;
                ok = ok && (p_le[k] ? (long long)vj <= la_v[k] : (long long)vj >= la_v[k]);
                if (is_a) la_v[k] = (long long)vj + p_c[k];
            }
        }
        if (is_a) {
            am |= MO::bit(j);
            seen_a = true;
        } else if (is_b) {
            bm |= MO::bit(j);
            if (ok) good |= MO::bit(j);
        }
    }
    SIESTA_HD bool finish(const TraceEvents& ev, mask_t& out) const {
        if (!bm) return false;
        const int a_first = MO::lo(am);
        auto val = [&](int mode, int j) -> int {
            if (mode == 2) return ev.ts[j * ev.ts_stride];
            if (mode == 1) return ev.src(j);
            return mode == 0 ? j : 0;
        };
        auto pass = [&](int b, int a) -> bool {
            bool ok = true;
#pragma unroll
            for (int k = 0; k < NPMAX; ++k) {
                if (k < np) {
                    const long long lhs = val(p_mode[k], b);
                    const long long rhs = (long long)val(p_mode[k], a) + p_c[k];
                    ok = ok && (p_le[k] ? lhs <= rhs : lhs >= rhs);
                }
            }
            return ok;
        };

        int best_size = 0, best_k = 0;
        mask_t best_pl = 0, best_out = 0;
        bool best_is_p = false;
        auto consider = [&](int size, int k, mask_t pl, mask_t o, bool is_p) {
            bool better = size > best_size || (size == best_size && k < best_k);
            if (!better && size == best_size && k == best_k) {
                const mask_t x = (pl ^ best_pl) & MO::below(k);
                if (x) better = (best_pl >> MO::hi(x)) & 1;
                else if (is_p != best_is_p) {
                    // P(q) and Q(0,q+1) were both placed at a_q and moved together ever since
                    const bool q0 = (pl >> a_first) & 1;
                    better = is_p == q0;
                }
            }
            if (better) { best_size = size; best_k = k; best_pl = pl; best_out = o; best_is_p = is_p; }
        };

        // Q(0,m), m >= 1
        if (good) {
            const int kq = MO::hi(good);
            int m = 0, prev = a_first;
            for (mask_t r = am & MO::above(a_first); r; r &= r - 1) {
                const int j = MO::lo(r);
                ++m;
                const mask_t g = good & MO::above(j);
                if (!g) break;  // later a's have no good b either
                consider(m + 1 + MO::popc(g), kq, MO::bit(prev) | g, (am & MO::below(j)) | MO::bit(j) | g, false);
                prev = j;
            }
        }
        // P(i): only if it can still reach the best size
        for (mask_t r = am; r; r &= r - 1) {
            const int j = MO::lo(r);
            if (mono_le && j != a_first) break;
            const mask_t cand = bm & MO::above(j);
            if (1 + MO::popc(cand) < best_size) continue;
            mask_t pm = 0;
            for (mask_t c = cand; c; c &= c - 1) {
                const int b = MO::lo(c);
                if (pass(b, j)) pm |= MO::bit(b);
                else if (mono_le) break;  // the passing b's are a prefix
            }
            if (!pm) continue;
            consider(1 + MO::popc(pm), MO::hi(pm), MO::bit(j) | pm, MO::bit(j) | pm, true);
        }
        if (best_size == 0) return false;
        out = best_out;
        return true;
    }
};

template <int W, int NPMAX>
SIESTA_HD bool fk2_eval_n(const DevNfa& nfa, const TraceEvents& ev, typename MaskOps<W>::T& out) {
    Fk2Eval<W, NPMAX> f;
    f.init(nfa, ev.evt_pos, ev.ts != nullptr);
    for (int j = 0; j < ev.n; ++j) f.on_event(j, ev.word(j), ev.src(j), ev.ts ? ev.ts[j * ev.ts_stride] : 0);
    return f.finish(ev, out);
}
template <int W>
SIESTA_HD bool fk2_eval(const DevNfa& nfa, const TraceEvents& ev, typename MaskOps<W>::T& out) {
    if (nfa.n_preds[1] <= 1) return fk2_eval_n<W, 1>(nfa, ev, out);
    return fk2_eval_n<W, SIESTA_MAX_PREDS>(nfa, ev, out);
}

// ------------------------------------------------------------------------------------------------- class NK
// Greedy walk of the run started at event s.  Returns the run's events as a mask (0 = the run never completes or
// is deleted at a negative state).
template <int W, class EV>
SIESTA_HD __forceinline__ typename MaskOps<W>::T nk_walk(const DevNfa& nfa, const EV& ev, const typename MaskOps<W>::T* T, int s) {
    typedef MaskX<W> MO;
    typedef typename MO::T mask_t;
    const int S = nfa.n_states;
    unsigned long long vvw = (unsigned long long)s;  // byte k = event taken for state k
    auto vv = [&vvw](int ref) { return (int)((vvw >> (8 * ref)) & 0xFF); };
    mask_t taken = MO::bit(s);
    int p = s;
    int k_next = 1;
    // fully unrolled over the state index so that T[k] stays in registers
#pragma unroll
    for (int k = 1; k < SIESTA_MAX_STATES; ++k) {
        if (k >= S) break;
        if (k != k_next) continue;
        const bool neg = nfa.kind[k] == SIESTA_STATE_NEGATIVE;
        mask_t c = (neg ? (T[k] | T[k + 1]) : T[k]) & MO::above(p);
        int got = -1;
        while (c) {
            const int e = MO::lo(c);
            c &= c - 1;
            if (neg) {
                if ((T[k] >> e) & 1) {
                    if (fast_preds(nfa, ev, k, e, k, vv)) return 0;   // containsNegative: deleted (Engine.java:679-682)
                } else if (fast_preds(nfa, ev, k + 1, e, k, vv)) {    // Engine.checkPredicatesForNextState :1165-1180
                    got = e;
                    break;
                }
            } else if (fast_preds(nfa, ev, k, e, k, vv)) {
                got = e;
                break;
            }
        }
        if (got < 0) return 0;
        const int ks = neg ? k + 1 : k;  // the state the event was taken for
        taken |= MO::bit(got);
        vvw |= (unsigned long long)got << (8 * ks);
        p = got;
        k_next = ks + 1;
    }
    return taken;
}

// ------------------------------------------------------------------------------------------------ class NP1
// T[s] = events whose type belongs to state s.  out = the events of the first-largest occurrence (state order = index
// order); n_emitted = the engine's match count when `count`.
template <int W>
SIESTA_HD __forceinline__ bool np1_eval(const DevNfa& nfa, const typename MaskOps<W>::T* T, bool count, typename MaskOps<W>::T& out,
                                        unsigned& n_emitted) {
    typedef MaskX<W> MO;
    typedef typename MO::T mask_t;
    const int S = nfa.n_states;
    n_emitted = 0;
    int k = 0;
    bool star = false;
#pragma unroll
    for (int s = 0; s < SIESTA_MAX_STATES; ++s) {
        if (s >= S) break;
        if (nfa.kind[s] == SIESTA_STATE_KLEENE_PLUS || nfa.kind[s] == SIESTA_STATE_KLEENE_STAR) {
            k = s;
            star = nfa.kind[s] == SIESTA_STATE_KLEENE_STAR;
        }
    }
    const bool star0 = star && k == 0;   // `a* b ..`: an event of state 1's types starts a run without any a
    const bool star2 = star && k >= 2;   // the run is cloned past the state the moment it enters it: matches without any b
#pragma unroll
    for (int s = 0; s < SIESTA_MAX_STATES; ++s) {
        if (s >= S) break;
        if (T[s] == 0 && !(star0 && s == 0) && !(star2 && s == k)) return false;
    }
    // F: positions from which the suffix completes (backward greedy: the latest embedding of states S-1 .. k+1)
    mask_t F = ~(mask_t)0, F2 = ~(mask_t)0;   // F2: the same for states S-1 .. 2
#pragma unroll
    for (int s = SIESTA_MAX_STATES - 1; s >= 1; --s) {
        if (s >= S || s <= k) continue;
        const mask_t c = T[s] & F;
        if (!c) return false;
        F = MO::below(MO::hi(c));
        if (s == 2) F2 = F;
    }
    // prefix walk of the run started at event p; returns the mask of its events (0: does not complete) and its end in p
    auto prefix = [&](int& p) -> mask_t {
        mask_t m = MO::bit(p);
#pragma unroll
        for (int s = 1; s < SIESTA_MAX_STATES; ++s) {
            if (s >= k) break;
            const mask_t c = T[s] & MO::above(p);
            if (!c) return 0;
            p = MO::lo(c);
            m |= MO::bit(p);
        }
        return m;
    };
    mask_t B, taken;
    if (k == 0) {
        B = T[0] & F;
        taken = 0;
        if (count)
            for (mask_t r = B; r; r &= r - 1) n_emitted += (unsigned)MO::popc(T[0] & MO::below(MO::lo(r))) + 1u;   // starts at or before a_l
        if (star0) {
            // Engine.createNewRun :960-982: an event that cannot start state 0 but is of state 1's types starts a run at
            // state 1 (state 0 counts as passed).  Those runs never fork; each emits S - 1 events, fewer than any match
            // that holds an a, so they matter for the selection only when no a lies in F: the first of them then completes
            // first (walks are monotone) and is the first emitted.
            const mask_t bs = T[1] & ~T[0] & F2;
            if (count) n_emitted += (unsigned)MO::popc(bs);
            if (!B) {
                if (!bs) return false;
                int q = MO::lo(bs);
                taken = MO::bit(q);
#pragma unroll
                for (int s = 2; s < SIESTA_MAX_STATES; ++s) {
                    if (s >= S) break;
                    q = MO::lo(T[s] & MO::above(q));   // exists: q lies in F2
                    taken |= MO::bit(q);
                }
                out = taken;
                return true;
            }
        }
    } else {
        int p = MO::lo(T[0]);
        taken = prefix(p);
        if (!taken) return false;
        B = T[k] & MO::above(p) & F;
        // kleeneClosure*, one more match per start, the prefix and the suffix alone.  k = 1: the start run (created by
        // Engine.createNewRun at state 1) is cloned past the state when it is offered its first event of the Kleene type
        // (Engine.java:658-670): the suffix starts behind that event.  k >= 2: the run enters the state inside
        // evaluateEventForSkipTillNext, whose proceed block (:691-713) fires at once (a kleeneClosure* state counts as
        // initialised): the run itself moves on from its prefix end, the clone stays and takes the b's.
        const bool skip0 = star2 && ((F >> p) & 1);
        if (count) {
            n_emitted = (unsigned)MO::popc(B) + ((star && k == 1 && B) || skip0 ? 1u : 0u);
            for (mask_t r = T[0] & (T[0] - 1); r; r &= r - 1) {
                int q = MO::lo(r);
                if (!prefix(q)) break;                       // later starts do not complete their prefix either
                const unsigned n = (unsigned)MO::popc(T[k] & MO::above(q) & F);
                const unsigned extra = (star && k == 1 && n) || (star2 && ((F >> q) & 1)) ? 1u : 0u;
                if (!n && !extra) break;
                n_emitted += n + extra;
            }
        }
        if (!B && skip0) {   // no b the suffix allows: the first start's own run is the first match of the least size
            int q = p;
#pragma unroll
            for (int s = 1; s < SIESTA_MAX_STATES; ++s) {
                if (s >= S) break;
                if (s <= k) continue;
                q = MO::lo(T[s] & MO::above(q));   // exists: p lies in F
                taken |= MO::bit(q);
            }
            out = taken;
            return true;
        }
    }
    if (!B) return false;
    taken |= B;
    int q = MO::hi(B);
#pragma unroll
    for (int s = 1; s < SIESTA_MAX_STATES; ++s) {
        if (s >= S) break;
        if (s <= k) continue;
        q = MO::lo(T[s] & MO::above(q));   // exists: q lies in F
        taken |= MO::bit(q);
    }
    out = taken;
    return true;
}

// With constraints on prefix states (see the header): every start on its own filtered masks.
template <int W, class EV>
SIESTA_HD __forceinline__ bool np1p_eval(const DevNfa& nfa, const EV& ev, const typename MaskOps<W>::T* T, typename MaskOps<W>::T& out,
                                         unsigned& n_emitted) {
    typedef MaskX<W> MO;
    typedef typename MO::T mask_t;
    const int S = nfa.n_states;
    n_emitted = 0;
    int k = 0;
    for (int s = 0; s < S; ++s) {
        if (T[s] == 0) return false;
        if (nfa.kind[s] == SIESTA_STATE_KLEENE_PLUS) k = s;
    }
    mask_t best = 0, best_pl = 0;
    int best_m = 0, best_c = 0;
    for (mask_t r = T[0]; r; r &= r - 1) {
        const int s0 = MO::lo(r);
        unsigned long long vvw = (unsigned long long)s0;  // byte j = event taken for prefix state j
        auto vv = [&vvw](int ref) { return (int)((vvw >> (8 * ref)) & 0xFF); };
        mask_t taken = MO::bit(s0);
        int p = s0;
        bool ok = true;
        for (int j = 1; j < k && ok; ++j) {
            mask_t c = T[j] & MO::above(p);
            int got = -1;
            while (c) {
                const int e = MO::lo(c);
                c &= c - 1;
                if (fast_preds(nfa, ev, j, e, j, vv)) { got = e; break; }
            }
            if (got < 0) { ok = false; break; }
            taken |= MO::bit(got);
            vvw |= (unsigned long long)got << (8 * j);
            p = got;
        }
        if (!ok) continue;
        mask_t B = 0;
        for (mask_t c = T[k] & MO::above(p); c; c &= c - 1) {
            const int e = MO::lo(c);
            if (fast_preds(nfa, ev, k, e, k, vv)) B |= MO::bit(e);
        }
        if (!B) continue;
        // latest embedding of the suffix: the Kleene events below its first event can complete
        mask_t F = ~(mask_t)0;
        for (int j = S - 1; j > k && ok; --j) {
            mask_t c = T[j] & F;
            int got = -1;
            while (c) {
                const int e = MO::hi(c);
                c &= ~MO::bit(e);
                if (fast_preds(nfa, ev, j, e, j, vv)) { got = e; break; }
            }
            if (got < 0) { ok = false; break; }
            F = MO::below(got);
        }
        if (!ok) continue;
        B &= F;
        const int m = MO::popc(B);
        if (!m) continue;
        n_emitted += (unsigned)m;
        if (m < best_m) continue;
        int q = MO::hi(B);
        const mask_t pl = MO::bit(s0) | (B & ~MO::bit(q));
        taken |= B;
        for (int j = k + 1; j < S; ++j) {
            mask_t c = T[j] & MO::above(q);
            while (c) {   // exists: q lies in F
                const int e = MO::lo(c);
                c &= c - 1;
                if (fast_preds(nfa, ev, j, e, j, vv)) { q = e; break; }
            }
            taken |= MO::bit(q);
        }
        bool better = m > best_m || q < best_c;
        if (!better && q == best_c) {   // same size, same completion event: the run list's order
            const mask_t d = pl ^ best_pl;
            const int h = MO::hi(d);
            if ((pl >> h) & 1) better = (best_pl & MO::below(h)) == 0;
            else better = (pl & MO::below(h)) != 0;
        }
        if (better) { best = taken; best_pl = pl; best_m = m; best_c = q; }
    }
    if (!best) return false;
    out = best;
    return true;
}

// aux: scratch of NE masks (one per possible start), element i at aux[i * aux_stride]; only used when return_all.
// sel[0..nsel) = the selected occurrences (Occurrences.clearOccurrences); n_emitted = engine matches.
// T[k] = events whose type belongs to state k (T[SIESTA_MAX_STATES] stays 0)
template <int W>
struct NkMasks {
    typename MaskOps<W>::T T[SIESTA_MAX_STATES + 1];
    SIESTA_HD __forceinline__ void init() {
#pragma unroll
        for (int k = 0; k <= SIESTA_MAX_STATES; ++k) T[k] = 0;
    }
    SIESTA_HD __forceinline__ void on_event(int j, uint32_t w) {
        const typename MaskOps<W>::T b = MaskOps<W>::bit(j);
#pragma unroll
        for (int k = 0; k < SIESTA_MAX_STATES; ++k)
            if (w & (1u << k)) T[k] |= b;  // one bits-to-predicates move + predicated ORs
    }
};

// first_only: stop at the first start whose walk completes (valid for monotone walks, see the header; n_emitted is
// then not the engine's match count and must not be reported).
template <int W, class EV>
SIESTA_HD __forceinline__ bool nk_eval(const DevNfa& nfa, const EV& ev, const typename MaskOps<W>::T* T, bool return_all, bool by_pos,
                       typename MaskOps<W>::T* aux, int aux_stride, typename MaskOps<W>::T* sel, int& nsel, unsigned& n_emitted,
                       bool first_only = false) {
    typedef MaskX<W> MO;
    typedef typename MO::T mask_t;
    const int S = nfa.n_states;
    nsel = 0;
    n_emitted = 0;
    // every state's types must occur at all (cheap reject of most traces); unrolled: T[] stays in registers
#pragma unroll
    for (int k = 0; k < SIESTA_MAX_STATES; ++k)
        if (k < S && nfa.kind[k] != SIESTA_STATE_NEGATIVE && T[k] == 0) return false;

    mask_t best = 0, done = 0;  // done: starts whose run completed
    int best_c = 0;
    for (mask_t r = T[0]; r; r &= r - 1) {
        const int s = MO::lo(r);
        const mask_t m = nk_walk<W>(nfa, ev, T, s);
        if (!m) continue;
        ++n_emitted;
        const int c = MO::hi(m);
        if (!best || c < best_c) { best = m; best_c = c; }  // starts ascend: an equal completion keeps the earlier start
        if (first_only) break;
        if (return_all) {
            aux[s * aux_stride] = m;
            done |= MO::bit(s);
        }
    }
    if (!best) return false;
    sel[0] = best;
    nsel = 1;
    if (return_all && n_emitted > 1) {
        // Occurrences.java:74-87: matches 1..n-1 in emission order (completion, start), kept if they overlap nothing chosen.
        // M[0] (= best here) is skipped.
        done &= ~MO::bit(MO::lo(best));
        auto overlaps = [&](mask_t a, mask_t b) {
            const int af = MO::lo(a), al = MO::hi(a), bf = MO::lo(b), bl = MO::hi(b);
            bool not_ov;
            if (by_pos) not_ov = ev.position(al) < ev.position(bf) || ev.position(af) > ev.position(bl);
            else not_ov = ev.timestamp(al) < ev.timestamp(bf) || ev.timestamp(af) > ev.timestamp(bl);
            return !not_ov;
        };
        while (done) {
            int ps = -1, pc = 0;
            mask_t pm = 0;
            for (mask_t r = done; r; r &= r - 1) {
                const int s = MO::lo(r);
                const mask_t m = aux[s * aux_stride];
                const int c = MO::hi(m);
                if (ps < 0 || c < pc) { ps = s; pc = c; pm = m; }
            }
            done &= ~MO::bit(ps);
            bool ov = false;
            for (int o = 0; o < nsel && !ov; ++o) ov = overlaps(pm, sel[o]);
            if (!ov) sel[nsel++] = pm;
        }
    }
    return true;
}

// --------------------------------------------------------------------------------- class NK on traces of any length
// The mask evaluators above hold a trace's pattern-relevant events in 32 / 64 / 64-bit masks.  For class NK the greedy walk
// needs no mask at all: it moves forward through the filtered event list once per start.  LongEvents is that list in
// plain arrays (global memory on the device); nk_long_walk is nk_walk over it.  Kernel K1-L (detect.cu) runs one warp per
// trace beyond the mask kernels' limits, lanes = starts, so NO trace length is out of reach for a pattern without a
// Kleene state (the reference has no limit either: S/engine/Engine.java:207-224).
struct LongEvents {
    int n;                  // pattern-relevant events of the trace
    const uint16_t* word;   // [n] bit k <=> the event's type belongs to state k
    const int32_t* pos;     // [n] in-trace position
    const int32_t* sec;     // [n] relative seconds (EventTs route with a time predicate / returnAll), or nullptr
    bool evt_pos;
    // SaseEvent attributes after Utils.transformToSaseEvents (J/model/Utils/Utils.java:48-65)
    SIESTA_HD __forceinline__ int position(int j) const { return evt_pos ? pos[j] : j; }
    SIESTA_HD __forceinline__ int timestamp(int j) const { return evt_pos ? j : (sec ? sec[j] : 0); }
    SIESTA_HD __forceinline__ int attr(int j, int a) const { return a == SIESTA_ATTR_POSITION ? position(j) : timestamp(j); }
};

// Greedy walk of the run started at filtered event s: out[0..n_positive) = the events it takes (indices into the filtered
// list); returns the number of events taken, 0 = the run never completes or is deleted at a negative state.
SIESTA_HD inline int nk_long_walk(const DevNfa& nfa, const LongEvents& ev, int s, int* out) {
    const int S = nfa.n_states;
    int vvi[SIESTA_MAX_STATES];
    vvi[0] = s;
    auto vv = [&vvi](int ref) { return vvi[ref]; };
    int n_out = 0;
    out[n_out++] = s;
    int p = s, k = 1;
    while (k < S) {
        const bool neg = nfa.kind[k] == SIESTA_STATE_NEGATIVE;
        int got = -1;
        for (int e = p + 1; e < ev.n; ++e) {
            const uint32_t w = ev.word[e];
            if (neg) {
                if ((w >> k) & 1u) {
                    if (fast_preds(nfa, ev, k, e, k, vv)) return 0;   // containsNegative: deleted (Engine.java:679-682)
                } else if (((w >> (k + 1)) & 1u) && fast_preds(nfa, ev, k + 1, e, k, vv)) {   // Engine.checkPredicatesForNextState :1165-1180
                    got = e;
                    break;
                }
            } else if (((w >> k) & 1u) && fast_preds(nfa, ev, k, e, k, vv)) {
                got = e;
                break;
            }
        }
        if (got < 0) return 0;
        const int ks = neg ? k + 1 : k;
        vvi[ks] = got;
        out[n_out++] = got;
        p = got;
        k = ks + 1;
    }
    return n_out;
}

// Occurrence.overlaps (J/model/Occurrence.java:36-49) between two runs given by their first / last filtered events
SIESTA_HD __forceinline__ bool nk_long_overlaps(const LongEvents& ev, bool by_pos, int af, int al, int bf, int bl) {
    bool not_ov;
    if (by_pos) not_ov = ev.position(al) < ev.position(bf) || ev.position(af) > ev.position(bl);
    else not_ov = ev.timestamp(al) < ev.timestamp(bf) || ev.timestamp(af) > ev.timestamp(bl);
    return !not_ov;
}

// ------------------------------------------------------------------------- class NK as window walks (kernel K1-P)
// Kernel K1-P addresses a trace's events by ONE index space in which every attribute the NFA's predicates read IS the
// bit index of the event:
//   rank space (32-bit masks): bit r = the r-th event of the list filtered to the pattern's types.  After
//       Utils.transformToSaseEvents (J/model/Utils/Utils.java:48-65) that index is `position` on the EventTs route
//       (Utils.java:54-57) and `timestamp` on the EventPos route (Utils.java:59-62).
//   raw space (64-bit masks): bit j = position slot j of the trace (slot = in-trace position + lead, lead = events of
//       the previous trace in the first 32-byte sector); `position` on the EventPos route is slot - lead, and the lead
//       cancels on both sides of a predicate.
// A predicate `attr(e) <= attr($ref) + c` (SIESTAPattern.java:136-145) then says "e lies below bit idx(ref) + c + 1",
// `>=` says "e lies at or above bit idx(ref) + c": each predicate is a WINDOW mask built with one shift, and the first
// later event of a state's types that passes all predicates (the greedy step of class NK, header of this file) is the
// lowest set bit of `T[k] & above(previous) & windows` - no loop over candidate events, no popcount.  At a negative
// state the run moves on with the first passing event of the next state's types unless a passing event of the negative
// type lies before it (Engine.java:679-682, 1165-1180); an event of the negative type is never taken for the next state
// (nk_walk tests the negative type first).  nkw_build (nfa.cpp) decides the space, or that the NFA needs relative
// seconds / mixes attributes and stays on the staged kernel.
template <class MT>
struct NkwBits;
template <>
struct NkwBits<uint32_t> {
    // 1 << n, and 0 once n reaches the width (PTX shl clamps the amount; C++ leaves it undefined)
    static SIESTA_HD __forceinline__ uint32_t one_shl(int n) {
#ifdef __CUDA_ARCH__
        uint32_t r;
        asm("shl.b32 %0, 1, %1;" : "=r"(r) : "r"(n));
        return r;
#else
        return n >= 32 ? 0u : 1u << n;
#endif
    }
    static SIESTA_HD __forceinline__ int idx(uint32_t b) { return 31 - clz32(b); }   // index of an isolated bit
    static SIESTA_HD __forceinline__ int popc(uint32_t m) { return popc32(m); }
};
template <>
struct NkwBits<unsigned long long> {
    static SIESTA_HD __forceinline__ unsigned long long one_shl(int n) {
#ifdef __CUDA_ARCH__
        unsigned long long r;
        asm("shl.b64 %0, 1, %1;" : "=l"(r) : "r"(n));
        return r;
#else
        return n >= 64 ? 0ull : 1ull << n;
#endif
    }
    static SIESTA_HD __forceinline__ int idx(unsigned long long b) { return 63 - clz64(b); }
    static SIESTA_HD __forceinline__ int popc(unsigned long long m) { return popc64(m); }
};

template <class MT>
struct NkwWalk {
    typedef NkwBits<MT> B;
    // events that pass every predicate of state s, given the indices taken so far (byte k of vv = state k's event)
    static SIESTA_HD __forceinline__ MT window(const NkwProgram& P, int s, unsigned long long vv) {
        MT w = ~(MT)0;
#pragma unroll
        for (int j = 0; j < SIESTA_MAX_PREDS; ++j) {
            if (j < P.n_preds[s]) {
                const int n = (int)((vv >> P.sh[s][j]) & 0xFFull) + P.cc[s][j];
                const MT below = B::one_shl(n) - 1;
                w &= P.le[s][j] ? below : ~below;
            }
        }
        return w;
    }
    // run started at the event with bit sb: its events as a mask, 0 = it never completes / is deleted at a negative
    // state.  exhausted: no later event of some state's types exists at all, whatever the predicates say; walks are
    // monotone (header of this file), so no later start completes either.
    static SIESTA_HD __forceinline__ MT walk(const NkwProgram& P, const MT* T, MT sb, bool& exhausted) {
        unsigned long long vv = (unsigned long long)B::idx(sb);
        MT taken = sb, pb = sb;
        int k_next = 1;
#pragma unroll
        for (int k = 1; k < SIESTA_MAX_STATES; ++k) {
            if (k >= P.n_states) break;
            if (k != k_next) continue;
            const MT after = ~(pb | (pb - 1));
            MT got;
            int ks = k;
            if (P.neg[k]) {
                ks = k + 1;
                const MT cy0 = T[k + 1] & ~T[k] & after;
                const MT cy = cy0 & window(P, k + 1, vv);
                if (!cy) {
                    exhausted = exhausted || !cy0;
                    return 0;
                }
                got = cy & ((MT)0 - cy);
                if (T[k] & after & (got - 1) & window(P, k, vv)) return 0;   // containsNegative: deleted (Engine.java:679-682)
            } else {
                const MT c0 = T[k] & after;
                const MT c = c0 & window(P, k, vv);
                if (!c) {
                    exhausted = exhausted || !c0;
                    return 0;
                }
                got = c & ((MT)0 - c);
            }
            vv |= (unsigned long long)B::idx(got) << (8 * ks);
            taken |= got;
            pb = got;
            k_next = ks + 1;
        }
        return taken;
    }
};

// Markov form (NkwProgram::markov): every predicate of a positive state k references prev[k], the positive state before
// it, so whether a run completes from an event taken for prev[k] depends on that event alone, and ALL starts are
// evaluated together, backwards, on whole masks:
//   good[S-1] = T[S-1];  good[prev[k]] = { p in T[prev[k]] : the first event of k's types at least ge[k] steps after p
//                                          lies at most le[k] steps after p and is in good[k] }
// "the first marker above q is in X" for all q at once is a downward fill from X through non-marker positions
// (five doubling steps on 32 bits), the upper bound a smear of the markers by le - ge + 1, the lower bound a shift by
// ge - 1.  With a (predicate-free) negative state before k the markers are the negative type's events as well: the
// first of {negative, k} after p must be an event of k that is not of the negative type (Engine.java:679-682,
// 1165-1180).  good[0] is the set of starts whose run completes: its popcount is the engine's match count, its lowest
// bit the start of the first-largest occurrence, whose events ONE forward walk then collects.  No per-start loop.
template <class MT>
struct NkwMarkov {
    typedef NkwBits<MT> B;
    // g(q) = "the first bit of `markers` strictly above q is in X" (X subset of markers)
    static SIESTA_HD __forceinline__ MT fill_down(MT X, MT markers) {
        MT g = X >> 1;
        MT P = ~markers >> 1;   // P(q): position q + 1 is not a marker; after the step with shift s: q+1 .. q+2s are none
#pragma unroll
        for (int s = 1; s < (int)(8 * sizeof(MT)); s <<= 1) {
            g |= (g >> s) & P;
            P &= P >> s;
        }
        return g;
    }
    // n(q) = "a bit of `markers` lies in (q, q + L]", 1 <= L
    static SIESTA_HD __forceinline__ MT near_above(MT markers, int L) {
        MT n = markers >> 1;
        int covered = 1;
        while (covered < L) {
            const int step = covered < L - covered ? covered : L - covered;
            n |= n >> step;
            covered += step;
        }
        return n;
    }
    static SIESTA_HD __forceinline__ MT shr(MT x, int n) { return n >= (int)(8 * sizeof(MT)) ? (MT)0 : (MT)(x >> n); }
    // the starts whose run completes
    static SIESTA_HD __forceinline__ MT good_starts(const NkwProgram& P, const MT* T) {
        const int S = P.n_states;
        MT G = 0;
        int k_next = S - 1;
#pragma unroll
        for (int k = SIESTA_MAX_STATES - 1; k >= 1; --k) {
            if (k >= S) continue;
            if (k == S - 1) G = T[k];
            if (k != k_next) continue;
            const int j = P.prev[k];
            MT markers = T[k], X = G;
            if (j != k - 1) {          // a negative state sits between j and k
                markers |= T[k - 1];
                X &= ~T[k - 1];
            }
            const int ge = P.m_ge[k], le = P.m_le[k];
            MT g = 0;
            if (le >= ge) {
                g = fill_down(X, markers);
                if (le != 255) g &= near_above(markers, le - ge + 1);
                g = shr(g, ge - 1);
            }
            G = T[j] & g;
            k_next = j;
        }
        return S == 1 ? T[0] : G;
    }
};

// first-largest occurrence = smallest (completion, start); first_only as in nk_eval
template <class MT>
SIESTA_HD __forceinline__ bool nkw_eval_markov(const NkwProgram& P, const MT* T, MT& best, unsigned& n_emitted) {
    const MT G = NkwMarkov<MT>::good_starts(P, T);
    n_emitted = (unsigned)NkwBits<MT>::popc(G);
    best = 0;
    if (!G) return false;
    bool exhausted = false;
    best = NkwWalk<MT>::walk(P, T, G & ((MT)0 - G), exhausted);
    return best != 0;
}

template <class MT>
SIESTA_HD __forceinline__ bool nkw_eval(const NkwProgram& P, const MT* T, MT& best, unsigned& n_emitted, bool first_only) {
    typedef NkwBits<MT> B;
    n_emitted = 0;
    best = 0;
#pragma unroll
    for (int k = 0; k < SIESTA_MAX_STATES; ++k)
        if (k < P.n_states && !P.neg[k] && T[k] == 0) return false;
    bool exhausted = false;
    int best_c = 0;
    for (MT r = T[0]; r && !exhausted;) {
        const MT sb = r & ((MT)0 - r);
        r ^= sb;
        const MT m = NkwWalk<MT>::walk(P, T, sb, exhausted);
        if (!m) continue;
        ++n_emitted;
        const int c = B::idx(m);   // clz of the whole mask: its highest event
        if (!best || c < best_c) { best = m; best_c = c; }
        if (first_only) break;
    }
    return best != 0;
}

}  // namespace siesta
