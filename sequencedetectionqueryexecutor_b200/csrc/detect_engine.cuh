// detect_engine.cuh — per-trace run-set engine (device), the heart of kernel K1.
//
// One thread owns one trace.  The thread walks the trace's events that belong to
// the pattern (already compacted into shared memory by the warp-cooperative filter
// phase) and keeps the engine's active-run list in a packed form:
//
//   run.mask  bitmask over the trace's filtered events = Run.eventIds (ids only ever
//             grow in stream order, so the list is the set of bits, low to high)
//   run.meta  cur:4 | state[]:2 bits x 8 | kleeneClosureInitialized | containsNegative |
//             isFull | delete-marks:2            (S/engine/Run.java:40-111)
//   run.fam   index of the value vector the run shares with every clone descending
//             from the same initializeRun (Run.clone is shallow: Run.java:319-327)
//
// The transition rules restate S/engine/Engine.java (evaluateEventForSkipTillNext
// :654-725, createNewRun :933-982, checkPredicate :1102-1180, checkProceed
// :1205-1224, cleanRuns :1404-1418) and S/engine/Run.java (addEvent :196-278,
// proceed :307-315).  Every place where the Java code would throw sets `err`.
// Engine.createNewRun's trailing block (:983-996) is not implemented: the host
// rejects SIESTA_F_MODE_HEAD for NFAs it would change.
#pragma once
#include "common.cuh"

// The engine compiles for the device (product) and, unchanged, for the host: tests/host_harness
// runs it on the CPU against the oracle before any GPU time is spent.  Nothing in the library's
// product path calls the host instantiation.
#define SIESTA_HD __host__ __device__

namespace siesta {

SIESTA_HD __forceinline__ int popc32(uint32_t x) {
#ifdef __CUDA_ARCH__
    return __popc(x);
#else
    return __builtin_popcount(x);
#endif
}
SIESTA_HD __forceinline__ int popc64(unsigned long long x) {
#ifdef __CUDA_ARCH__
    return __popcll(x);
#else
    return __builtin_popcountll(x);
#endif
}
SIESTA_HD __forceinline__ int clz32(uint32_t x) {
#ifdef __CUDA_ARCH__
    return __clz(x);
#else
    return x ? __builtin_clz(x) : 32;
#endif
}
SIESTA_HD __forceinline__ int clz64(unsigned long long x) {
#ifdef __CUDA_ARCH__
    return __clzll(x);
#else
    return x ? __builtin_clzll(x) : 64;
#endif
}
SIESTA_HD __forceinline__ int ffs32(uint32_t x) {
#ifdef __CUDA_ARCH__
    return __ffs(x);
#else
    return __builtin_ffs((int)x);
#endif
}
SIESTA_HD __forceinline__ int ffs64(unsigned long long x) {
#ifdef __CUDA_ARCH__
    return __ffsll(x);
#else
    return __builtin_ffsll((long long)x);
#endif
}

template <int W>
struct MaskOps;
template <>
struct MaskOps<1> {
    typedef uint32_t T;
    static SIESTA_HD __forceinline__ T bit(int j) { return 1u << j; }
    static SIESTA_HD __forceinline__ int popc(T m) { return popc32(m); }
    static SIESTA_HD __forceinline__ int hi(T m) { return 31 - clz32(m); }
    static SIESTA_HD __forceinline__ int lo(T m) { return ffs32(m) - 1; }
};
template <>
struct MaskOps<2> {
    typedef unsigned long long T;
    static SIESTA_HD __forceinline__ T bit(int j) { return 1ull << j; }
    static SIESTA_HD __forceinline__ int popc(T m) { return popc64(m); }
    static SIESTA_HD __forceinline__ int hi(T m) { return 63 - clz64(m); }
    static SIESTA_HD __forceinline__ int lo(T m) { return ffs64(m) - 1; }
};

// meta layout
#define M_CUR(m) ((m) & 15u)
#define M_ST_SHIFT 4
#define M_ST(m) (((m) >> M_ST_SHIFT) & 0xFFFFu)
#define M_KINIT (1u << 20)
#define M_CNEG (1u << 21)
#define M_FULL (1u << 22)
#define M_MARK_SHIFT 23
#define M_MARKS(m) (((m) >> M_MARK_SHIFT) & 3u)

// Events of one trace, compacted to the pattern's event types.  Shared memory,
// transposed [event][thread] so that a warp stepping in lockstep is conflict-free.
//   meta word: smask (bit k: type belongs to state k)  | fmask << 8 (bit k: type is
//   state k's first type) | index-in-trace << 16
struct TraceEvents {
    const uint32_t* meta;  // stride = NT words between consecutive events of this thread
    const int32_t* ts;     // relative seconds (EventTs route) or nullptr
    int stride;
    int n;
    bool evt_pos;
    int ts_stride;         // words between consecutive events in ts
    SIESTA_HD __forceinline__ uint32_t word(int j) const { return meta[j * stride]; }
    SIESTA_HD __forceinline__ int src(int j) const { return (int)(meta[j * stride] >> 16); }
    // SaseEvent attributes (J/SaseConnection/SaseEvent.java:80-89) after
    // Utils.transformToSaseEvents (J/model/Utils/Utils.java:48-65)
    SIESTA_HD __forceinline__ int position(int j) const { return evt_pos ? src(j) : j; }
    SIESTA_HD __forceinline__ int timestamp(int j) const { return evt_pos ? j : (ts ? ts[j * ts_stride] : 0); }
    SIESTA_HD __forceinline__ int attr(int j, int a) const { return a == SIESTA_ATTR_POSITION ? position(j) : timestamp(j); }
};

// Backing store of one trace's run list and value-vector families.  The narrow (common) configuration lives in
// shared memory, lane-transposed (element i of lane l at [i * 32 + l]) so a warp in lockstep is conflict-free;
// the wide fallback and the host test harness use plain per-thread arrays (STRIDE 1).
template <int W, int R_, int NF_, int STRIDE_>
struct RunStore {
    static constexpr int R = R_, NF = NF_, STRIDE = STRIDE_, WORDS = W;
    typename MaskOps<W>::T* rmask;   // [R]  Run.eventIds as a bitmask over the trace's filtered events
    uint32_t* rmeta;                 // [R]  packed run state
    uint8_t* rfam;                   // [R]  value-vector family of the run
    unsigned long long* famvv;       // [NF] byte k = 1 + rank of the event last stored for state k; 0 = null
    uint8_t* fmin;                   // [NF] smallest cursor among the family's live runs
    uint8_t* fmin2;                  // [NF] second smallest
    uint8_t* fcnt;                   // [NF] live runs of the family
};
// per-thread arrays for STRIDE 1 stores
template <int W, int R, int NF>
struct RunArrays {
    typename MaskOps<W>::T rmask[R];
    uint32_t rmeta[R];
    uint8_t rfam[R];
    unsigned long long famvv[NF];
    uint8_t fmin[NF], fmin2[NF], fcnt[NF];
    SIESTA_HD RunStore<W, R, NF, 1> store() { return RunStore<W, R, NF, 1>{rmask, rmeta, rfam, famvv, fmin, fmin2, fcnt}; }
};

template <class Store>
struct RunEngine {
    static constexpr int W = Store::WORDS, R = Store::R, NF = Store::NF, STRIDE = Store::STRIDE;
    static_assert(NF <= 255, "family ids are stored in a byte");
    typedef MaskOps<W> MO;
    typedef typename MO::T mask_t;

    const DevNfa& nfa;
    const TraceEvents& ev;
    Store st;
    uint32_t fam_used[(NF + 31) / 32];  // allocated family ids (recycled once no live run refers to them)
    int nruns;
    bool dirty;      // a run moved, forked, was marked or created since the last reduction pass
    unsigned moved;  // bit r (r < 32): run r changed its packed state or is new in this event; ~0u = unknown
    bool err, ovf;
    unsigned n_emitted;

    SIESTA_HD RunEngine(const DevNfa& n, const TraceEvents& e, const Store& s) : nfa(n), ev(e), st(s) {}

    // ---- small helpers over a packed run -------------------------------------------------
    SIESTA_HD __forceinline__ static uint32_t get_st(uint32_t m, int k) { return (m >> (M_ST_SHIFT + 2 * k)) & 3u; }
    SIESTA_HD __forceinline__ static uint32_t set_st(uint32_t m, int k, uint32_t v) {
        return (m & ~(3u << (M_ST_SHIFT + 2 * k))) | (v << (M_ST_SHIFT + 2 * k));
    }
    SIESTA_HD __forceinline__ static uint32_t set_cur(uint32_t m, int c) { return (m & ~15u) | (uint32_t)c; }
    SIESTA_HD __forceinline__ bool is_kleene(int kind) const { return kind == SIESTA_STATE_KLEENE_PLUS || kind == SIESTA_STATE_KLEENE_STAR; }
    // Run.checkMatch (Run.java:181-191)
    SIESTA_HD __forceinline__ bool check_match(uint32_t m) const { return (m & M_FULL) && M_ST(m) == nfa.all2; }
    // Run.proceed (Run.java:307-315)
    SIESTA_HD __forceinline__ uint32_t proceed(uint32_t m) const {
        int cur = M_CUR(m);
        m = set_st(m, cur, 2);
        if (cur == nfa.n_states - 1) m |= M_FULL;
        else m = set_cur(m, cur + 1);
        return m;
    }
    SIESTA_HD __forceinline__ int new_family() {
        if (!nfa.need_vv) return 0;
#pragma unroll
        for (int w = 0; w < (NF + 31) / 32; ++w) {
            const uint32_t free_bits = ~fam_used[w] & (NF - 32 * w >= 32 ? 0xffffffffu : ((1u << (NF - 32 * w)) - 1u));
            if (free_bits) {
                const int b = ffs32(free_bits) - 1;
                fam_used[w] |= 1u << b;
                st.famvv[(32 * w + b) * STRIDE] = 0ull;
                return 32 * w + b;
            }
        }
        ovf = true;
        return 0;
    }
    SIESTA_HD __forceinline__ void append(mask_t mask, uint32_t meta, int fam) {
        if (nruns >= R) { ovf = true; return; }
        dirty = true;
        moved |= nruns < 32 ? (1u << nruns) : ~0u;
        st.rmask[nruns * STRIDE] = mask;
        st.rmeta[nruns * STRIDE] = meta;
        st.rfam[nruns * STRIDE] = (uint8_t)fam;
        nruns++;
    }
    SIESTA_HD __forceinline__ int vv_get(int fam, int k) const { return (int)((st.famvv[fam * STRIDE] >> (8 * k)) & 0xFF) - 1; }
    SIESTA_HD __forceinline__ void vv_set(int fam, int k, int j) {
        st.famvv[fam * STRIDE] = (st.famvv[fam * STRIDE] & ~(0xFFull << (8 * k))) | ((unsigned long long)(j + 1) << (8 * k));
    }

    // Edge.evaluatePredicate(Event, Run, EventBuffer) over PredicateOptimized.evaluate (S/query/
    // PredicateOptimized.java:331-368): self reference -> true; null value vector -> false.
    SIESTA_HD bool eval_preds(int s, int j, int cur, int fam) const {
        const int np = nfa.n_preds[s];
        for (int k = 0; k < np; ++k) {
            const int ref = nfa.p_ref[s][k];
            if (ref == cur) continue;
            const int rj = vv_get(fam, ref);
            if (rj < 0) return false;
            const int a = nfa.p_attr[s][k];
            const long long lhs = ev.attr(j, a);
            const long long rhs = (long long)ev.attr(rj, a) + nfa.p_c[s][k];
            if (nfa.p_op[s][k] == SIESTA_OP_LE ? !(lhs <= rhs) : !(lhs >= rhs)) return false;
        }
        return true;
    }
    // State.canStartWithEvent (S/query/State.java:302-314): predicates see the event as its own reference
    // (PredicateOptimized.evaluate(Event, Event) :302-323).
    SIESTA_HD bool can_start(int s, int j, uint32_t w) const {
        if (!((w >> s) & 1u)) return false;
        const int np = nfa.n_preds[s];
        for (int k = 0; k < np; ++k) {
            const long long v = ev.attr(j, nfa.p_attr[s][k]);
            const long long rhs = v + nfa.p_c[s][k];
            if (nfa.p_op[s][k] == SIESTA_OP_LE ? !(v <= rhs) : !(v >= rhs)) return false;
        }
        return true;
    }
    // Engine.checkProceed (Engine.java:1205-1224); only ever reached on Kleene states here.
    SIESTA_HD __forceinline__ bool check_proceed(mask_t mask, uint32_t meta) {
        if (mask == 0) { err = true; return false; }  // eventIds.get(count-1) with count == 0
        const int cur = M_CUR(meta);
        if (get_st(meta, cur) == 0 && nfa.kind[cur] == SIESTA_STATE_KLEENE_PLUS) return false;
        return true;  // SIESTA never attaches proceed-edge predicates
    }
    // Engine.checkPredicate (:1102-1163) + checkPredicatesForNextState (:1165-1180)
    SIESTA_HD bool check_predicate(int j, uint32_t w, uint32_t meta, int fam) const {
        const int cur = M_CUR(meta);
        const int kind = nfa.kind[cur];
        const bool mine = (w >> cur) & 1u;
        if (kind == SIESTA_STATE_NEGATIVE) {
            if (mine) return eval_preds(cur, j, cur, fam);
            else if (nfa.n_states > cur + 1) {
                if ((w >> (cur + 1)) & 1u) return eval_preds(cur + 1, j, cur, fam);
                return false;
            }
        }
        if (!mine) return false;
        return eval_preds(cur, j, cur, fam);  // begin and take edges carry the same list
    }
    // Run.addEventToNormalorOr (Run.java:247-262)
    SIESTA_HD __forceinline__ void add_normal(int j, mask_t& mask, uint32_t& meta, int fam) {
        int cur = M_CUR(meta);
        mask |= MO::bit(j);
        meta = set_st(meta, cur, 2);
        const uint32_t st = M_ST(meta);
        const int c = popc32(st & 0xAAAAu & ~((st & 0x5555u) << 1));
        if (cur == nfa.n_states - 1 || c == nfa.n_states) {
            meta |= M_FULL;
        } else {
            if ((nfa.has_vv >> cur) & 1u) vv_set(fam, cur, j);
            meta = set_cur(meta, cur + 1);
        }
    }
    // Run.addEventToKleene (Run.java:264-278)
    SIESTA_HD __forceinline__ void add_kleene(int j, mask_t& mask, uint32_t& meta, int fam) {
        const int cur = M_CUR(meta);
        mask |= MO::bit(j);
        if ((nfa.has_vv >> cur) & 1u) {
            if ((meta & M_KINIT) && vv_get(fam, cur) < 0) { err = true; return; }  // updateValueVector NPE (Run.java:361)
            vv_set(fam, cur, j);
        }
        meta |= M_KINIT;
        meta = set_st(meta, cur, 3);
    }
    // Run.addEventToNextState (Run.java:228-245)
    SIESTA_HD __forceinline__ void add_next(int j, mask_t& mask, uint32_t& meta, int fam) {
        const int cur = M_CUR(meta);
        if (cur + 1 >= nfa.n_states) { err = true; return; }  // getStates(currentState+1)
        const int nk = nfa.kind[cur + 1];
        if (nk == SIESTA_STATE_NORMAL || nk == SIESTA_STATE_OR) {
            meta = set_st(meta, cur, 2);
            meta = set_cur(meta, cur + 1);
            add_normal(j, mask, meta, fam);
        } else if (is_kleene(nk)) {
            meta = set_cur(meta, cur + 1);
            add_kleene(j, mask, meta, fam);
        } else {  // negative
            meta |= M_CNEG;
            meta = set_cur(meta, cur + 1);
        }
    }
    // Run.addEvent (Run.java:196-225)
    SIESTA_HD void add_event(int j, uint32_t w, mask_t& mask, uint32_t& meta, int fam) {
        const int cur = M_CUR(meta);
        const int k = nfa.kind[cur];
        const bool mine = (w >> cur) & 1u;
        if (k == SIESTA_STATE_NORMAL || k == SIESTA_STATE_OR) add_normal(j, mask, meta, fam);
        else if (k == SIESTA_STATE_NEGATIVE) {
            if (mine) { meta |= M_CNEG; meta = set_cur(meta, cur + 1); }
            else add_next(j, mask, meta, fam);
        } else if (k == SIESTA_STATE_KLEENE_STAR) {
            if (mine) add_kleene(j, mask, meta, fam);
            else add_next(j, mask, meta, fam);
        } else add_kleene(j, mask, meta, fam);
    }

    // Engine.evaluateEventForSkipTillNext (Engine.java:654-725)
    template <class Emit>
    SIESTA_HD void evaluate(int j, uint32_t w, int r, Emit& emit) {
        mask_t mask = st.rmask[r * STRIDE];
        uint32_t meta = st.rmeta[r * STRIDE];
        const int fam = st.rfam[r * STRIDE];
        int cur = M_CUR(meta);
        if (cur >= nfa.n_states) { err = true; return; }  // getStates(currentState) past a trailing negative
        if (nfa.kind[cur] == SIESTA_STATE_KLEENE_STAR && !(meta & M_KINIT)) {
            if (check_proceed(mask, meta) && ((w >> (8 + cur)) & 1u)) {
                const uint32_t cm = proceed(meta);
                if (check_match(cm)) emit(mask);  // emits r's ids (:665)
                else append(mask, cm, fam);
            }
            if (err) return;
        }
        if (!check_predicate(j, w, meta, fam)) return;
        add_event(j, w, mask, meta, fam);
        if (err) return;
        if (meta & M_CNEG) meta += (1u << M_MARK_SHIFT);
        if ((meta & M_FULL) && check_match(meta)) {
            emit(mask);
            meta += (1u << M_MARK_SHIFT);
        }
        cur = M_CUR(meta);
        if (cur >= nfa.n_states) { err = true; return; }
        if (is_kleene(nfa.kind[cur])) {
            if (check_proceed(mask, meta)) {
                // the clone is a new object: it is not in toDeleteRuns even if its parent already is
                append(mask, (meta | M_KINIT) & ~(3u << M_MARK_SHIFT), fam);
                meta = proceed(meta);
                if (check_match(meta)) {
                    emit(mask);
                    meta += (1u << M_MARK_SHIFT);
                }
            }
            if (err) return;
        }
        if (M_MARKS(meta) >= 2) { err = true; return; }  // cleanRuns: second resetRun NPEs (Run.java:163)
        dirty = true;
        moved |= r < 32 ? (1u << r) : ~0u;
        st.rmask[r * STRIDE] = mask;
        st.rmeta[r * STRIDE] = meta;
    }

    // Engine.createNewRun (Engine.java:933-982)
    template <class Emit>
    SIESTA_HD void create_new_run(int j, uint32_t w, Emit& emit) {
        const int k0 = nfa.kind[0];
        if (can_start(0, j, w)) {
            if (is_kleene(k0)) {
                const int fam = new_family();
                mask_t mask = 0;
                uint32_t meta = nfa.init_st << M_ST_SHIFT;
                add_event(j, w, mask, meta, fam);
                if (err) return;
                if (check_proceed(mask, meta)) append(mask, proceed(meta), fam);
                if (err) return;
            }
            const int fam = new_family();
            mask_t mask = 0;
            uint32_t meta = nfa.init_st << M_ST_SHIFT;
            add_event(j, w, mask, meta, fam);
            if (err) return;
            if (check_match(meta)) emit(mask);
            else append(mask, meta, fam);
        } else if (k0 == SIESTA_STATE_KLEENE_STAR || k0 == SIESTA_STATE_NEGATIVE) {
            if (nfa.n_states > 1 && can_start(1, j, w)) {
                const int fam = new_family();
                mask_t mask = 0;
                uint32_t meta = nfa.init_st << M_ST_SHIFT;
                add_event(j, w, mask, meta, fam);
                if (err) return;
                const int cur = M_CUR(meta);
                if (cur >= nfa.n_states) { err = true; return; }
                if (nfa.kind[cur] == SIESTA_STATE_KLEENE_STAR && check_proceed(mask, meta)) meta = proceed(meta);
                if (err) return;
                if (check_match(meta)) emit(mask);
                else append(mask, meta, fam);
            }
        }
    }

    // ---- exact reductions of the run list (not in the reference; they never change its output) -------------
    //
    // (1) inert runs.  A run that is full but not complete is skipped for ever (Engine.java:364).  A run whose
    //     current state carries a "within" predicate (attr <= $ref.attr + c) that fails at the current event fails
    //     at every later event too, because position and timestamp never decrease along the stream, PROVIDED the
    //     referenced value can no longer change: no run of the same family (the runs sharing one value vector)
    //     sits at or before the referenced state.  Such a run never takes another event, never forks, never
    //     emits and never writes the value vector; the reference keeps it only because SIESTA disables the
    //     engine's time window (NFAWrapper.java:21-25).  Negative states and not-yet-initialised kleeneClosure*
    //     states are never pruned (they can move without passing a predicate: Engine.java:658-671, Run.java:200-203).
    // (2) dominated runs (only when just the first-largest occurrence is wanted: returnAll=false and no exact
    //     match count).  Two runs with the same packed state and the same family take the same events, fork and
    //     emit at the same events for ever; their matches differ only by the events selected so far.  Of the two,
    //     the one with fewer events (or, at equal size, the later one in list order) can never be the FIRST
    //     occurrence of maximal size (Occurrences.java:60-69), so it is dropped.  The survivor keeps its own list
    //     position, so emission order among the remaining runs is untouched.  Runs that selected nothing yet are
    //     never merged (Engine.checkProceed throws on them: Run.java:285).  Because runs are evaluated one after
    //     the other inside an event and siblings share one value vector, "same events for ever" needs the static
    //     condition DevNfa::merge_safe (nfa.cpp): no event type both writes a referenced slot and reads it.
    //     Across families the same holds when the slots the run can still read (DevNfa::relmask) hold the same
    //     events and nobody but the two runs themselves can rewrite them (same_future below).
    bool opt_prune, opt_dedup, ts_monotone;

    SIESTA_HD bool is_inert(int j, uint32_t meta, int fam) const {
        if (meta & M_FULL) return true;
        if (!opt_prune || !nfa.need_vv) return false;
        const int cur = M_CUR(meta);
        if (cur >= nfa.n_states) return false;
        const int kind = nfa.kind[cur];
        if (kind == SIESTA_STATE_NEGATIVE) return false;
        if (kind == SIESTA_STATE_KLEENE_STAR && !(meta & M_KINIT)) return false;
        const int np = nfa.n_preds[cur];
        for (int k = 0; k < np; ++k) {
            if (nfa.p_op[cur][k] != SIESTA_OP_LE) continue;
            const int ref = nfa.p_ref[cur][k];
            if (ref >= cur) continue;                       // self / forward reference
            if (st.fmin[fam * STRIDE] <= ref) continue;     // a sibling can still rewrite the slot
            const int rj = vv_get(fam, ref);
            if (rj < 0) return true;                        // null for ever -> predicate false for ever
            const int a = nfa.p_attr[cur][k];
            if (a == SIESTA_ATTR_TIMESTAMP && !ts_monotone) continue;
            if ((long long)ev.attr(j, a) > (long long)ev.attr(rj, a) + nfa.p_c[cur][k]) return true;
        }
        return false;
    }

    // Is run q interchangeable with run r for everything r can still do?  (same packed state; the value-vector
    // slots r can still read hold the same events and will be rewritten identically in both families)
    SIESTA_HD __forceinline__ bool same_future(int q, uint32_t meta, int fam, int cur) const {
        if (st.rmeta[q * STRIDE] != meta) return false;  // marked / already dropped runs differ in meta
        const int fq = st.rfam[q * STRIDE];
        if (fq == fam || !nfa.need_vv) return true;
        if (!nfa.cross_safe) return false;
        const unsigned long long rel = nfa.relmask[cur];
        if (rel == 0) return true;                                   // neither lineage reads a value vector again
        if ((st.famvv[fam * STRIDE] ^ st.famvv[fq * STRIDE]) & rel) return false;
        if (st.fcnt[fam * STRIDE] != 1) return false;                // dropping r must not starve its siblings
        const int nq = st.fcnt[fq * STRIDE];
        if (nq == 1) return true;
        const int mn = st.fmin[fq * STRIDE];
        const int others = mn == cur ? st.fmin2[fq * STRIDE] : mn;   // smallest cursor among q's siblings
        return others > nfa.kmax[cur];                               // nobody else in q's family rewrites those slots
    }

    SIESTA_HD void reduce_runs(int j) {
        if (nfa.need_vv) {
#pragma unroll
            for (int w = 0; w < (NF + 31) / 32; ++w)
                for (uint32_t m = fam_used[w]; m; m &= m - 1) {
                    const int f = 32 * w + ffs32(m) - 1;
                    st.fmin[f * STRIDE] = 15;
                    st.fmin2[f * STRIDE] = 15;
                    st.fcnt[f * STRIDE] = 0;
                }
            for (int r = 0; r < nruns; ++r) {
                const uint32_t m = st.rmeta[r * STRIDE];
                if (M_MARKS(m) || (m & M_FULL)) continue;
                const int f = st.rfam[r * STRIDE];
                const int c = M_CUR(m);
                const int mn = st.fmin[f * STRIDE];
                if (c < mn) { st.fmin2[f * STRIDE] = (uint8_t)mn; st.fmin[f * STRIDE] = (uint8_t)c; }
                else if (c < st.fmin2[f * STRIDE]) st.fmin2[f * STRIDE] = (uint8_t)c;
                if (st.fcnt[f * STRIDE] < 255) ++st.fcnt[f * STRIDE];
            }
        }
        const bool dedup = opt_dedup && nfa.merge_safe;
        const bool all_moved = moved == ~0u || nruns > 32;
        bool any_drop = false;
        for (int r = 0; r < nruns; ++r) {
            const uint32_t meta = st.rmeta[r * STRIDE];
            if (M_MARKS(meta)) { any_drop = true; continue; }   // cleanRuns (:1404-1418)
            const int fam = st.rfam[r * STRIDE];
            bool drop = is_inert(j, meta, fam);
            if (!drop && dedup) {
                const mask_t mask = st.rmask[r * STRIDE];
                if (mask != 0) {
                    const int cnt = MO::popc(mask);
                    const int cur = M_CUR(meta);
                    // two runs that both sat still were already compared after an earlier event
                    const bool r_moved = all_moved || ((moved >> r) & 1u);
                    for (int q = 0; q < nruns && !drop; ++q) {
                        if (q == r || !(r_moved || ((moved >> q) & 1u))) continue;
                        if (!same_future(q, meta, fam, cur)) continue;
                        const mask_t mq = st.rmask[q * STRIDE];
                        if (mq == 0) continue;
                        const int cq = MO::popc(mq);
                        drop = cq > cnt || (cq == cnt && q < r);
                    }
                }
            }
            if (drop) { st.rmeta[r * STRIDE] = meta | (1u << M_MARK_SHIFT) | (1u << 31); any_drop = true; }
        }
        if (any_drop) {
            int k = 0;
            for (int r = 0; r < nruns; ++r) {
                const uint32_t m = st.rmeta[r * STRIDE];
                if (M_MARKS(m)) continue;
                if (k != r) {
                    st.rmask[k * STRIDE] = st.rmask[r * STRIDE];
                    st.rmeta[k * STRIDE] = m;
                    st.rfam[k * STRIDE] = st.rfam[r * STRIDE];
                }
                ++k;
            }
            nruns = k;
        }
        dirty = false;
        moved = 0;
        if (nfa.need_vv) {
            // recycle the ids of families no surviving run refers to
#pragma unroll
            for (int w = 0; w < (NF + 31) / 32; ++w) fam_used[w] = 0;
            for (int r = 0; r < nruns; ++r) {
                const int f = st.rfam[r * STRIDE];
                fam_used[f >> 5] |= 1u << (f & 31);
            }
        }
    }

    // Engine.runSkipTillNextEngine (Engine.java:207-224) over one trace.
    template <class Emit>
    SIESTA_HD void run(Emit& emit, bool prune, bool dedup) {
        nruns = 0;
#pragma unroll
        for (int w = 0; w < (NF + 31) / 32; ++w) fam_used[w] = 0;
        err = false;
        ovf = false;
        n_emitted = 0;
        opt_prune = prune;
        opt_dedup = dedup;
        ts_monotone = true;
        if (prune && ev.ts && !ev.evt_pos)
            for (int j = 1; j < ev.n; ++j) ts_monotone = ts_monotone && ev.ts[j * ev.ts_stride] >= ev.ts[(j - 1) * ev.ts_stride];
        dirty = false;
        moved = 0;
        for (int j = 0; j < ev.n; ++j) {
            const uint32_t w = ev.word(j);
            const int n0 = nruns;  // runs appended while evaluating this event are not visited (:361)
            for (int r = 0; r < n0; ++r) {
                if (st.rmeta[r * STRIDE] & M_FULL) continue;
                evaluate(j, w, r, emit);
                if (err || ovf) return;
            }
            create_new_run(j, w, emit);
            if (err || ovf) return;
            // cleanRuns (:1404-1418) happens before createNewRun in the reference; the order is immaterial because
            // marked runs are never looked at again, and folding it into the reduction pass saves one sweep.
            if (nruns && dirty) reduce_runs(j);
            else if (nruns == 0) {
#pragma unroll
                for (int w2 = 0; w2 < (NF + 31) / 32; ++w2) fam_used[w2] = 0;
            }
        }
    }
};

// Occurrences.clearOccurrences (J/model/Occurrences.java:58-89), streamed over the emission order.
template <int W>
struct BestEmit {
    typedef MaskOps<W> MO;
    typename MO::T best;
    int best_cnt;
    unsigned n;
    SIESTA_HD BestEmit() : best(0), best_cnt(-1), n(0) {}
    // the first occurrence of maximal size (:60-69)
    SIESTA_HD __forceinline__ void operator()(typename MO::T m) {
        const int c = MO::popc(m);
        if (c > best_cnt) { best = m; best_cnt = c; }
        ++n;
    }
};

template <int W, int NSEL>
struct GreedyEmit {
    typedef MaskOps<W> MO;
    typedef typename MO::T mask_t;
    const TraceEvents& ev;
    mask_t sel[NSEL];
    int nsel;
    unsigned idx;
    bool by_pos;
    bool ovf;
    SIESTA_HD GreedyEmit(const TraceEvents& e, mask_t best, bool bp) : ev(e), nsel(1), idx(0), by_pos(bp), ovf(false) { sel[0] = best; }
    // Occurrence.overlaps (J/model/Occurrence.java:36-49); a = this, b = argument
    SIESTA_HD __forceinline__ bool overlaps(mask_t a, mask_t b) const {
        const int af = MO::lo(a), al = MO::hi(a), bf = MO::lo(b), bl = MO::hi(b);
        bool not_ov;
        if (by_pos) not_ov = ev.position(al) < ev.position(bf) || ev.position(af) > ev.position(bl);
        else not_ov = ev.timestamp(al) < ev.timestamp(bf) || ev.timestamp(af) > ev.timestamp(bl);
        return !not_ov;
    }
    // returnAll branch (:74-87): matches 1..n-1 in emission order, kept if they overlap nothing chosen so far
    SIESTA_HD void operator()(mask_t m) {
        const unsigned i = idx++;
        if (i == 0) return;
        for (int o = 0; o < nsel; ++o)
            if (overlaps(m, sel[o])) return;
        if (nsel >= NSEL) { ovf = true; return; }
        sel[nsel++] = m;
    }
};

}  // namespace siesta
