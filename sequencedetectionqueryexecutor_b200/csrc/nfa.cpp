// nfa.cpp — host-side NFA validation and the per-activity lookup table of kernel K1.
#include <algorithm>
#include <cstring>
#include <vector>

#include "common.cuh"

namespace siesta {

// Which closed-form evaluator (detect_fast.cuh) may replace the run-list engine for this NFA and these flags?
// 0 = none, 1 = NK (no Kleene state), 2 = FK2 (`a+ b*`), 3 = NP1 (one `+` or `*` state).  The conditions are exactly the ones the derivations in
// detect_fast.cuh rely on; anything else keeps the general engine.
static int classify_fast(const siesta_nfa* nfa, const DevNfa& d, uint32_t flags) {
    if (flags & SIESTA_F_LITERAL_RUNS) return 0;
    const int S = nfa->n_states;
    auto positive = [&](int s) { return nfa->states[s].kind == SIESTA_STATE_NORMAL || nfa->states[s].kind == SIESTA_STATE_OR; };
    if (!d.any_kleene) {
        if (!positive(0) || !positive(S - 1) || d.n_preds[0] != 0) return 0;
        for (int s = 0; s < S; ++s) {
            if (s > 0 && !positive(s) && !positive(s - 1)) return 0;  // two negative states in a row
            for (int k = 0; k < d.n_preds[s]; ++k) {
                const int ref = d.p_ref[s][k];
                if (ref >= s || !positive(ref)) return 0;
            }
        }
        return 1;
    }
    if (S == 2 && nfa->states[0].kind == SIESTA_STATE_KLEENE_PLUS && nfa->states[1].kind == SIESTA_STATE_KLEENE_STAR &&
        nfa->states[0].n_types == 1 && nfa->states[1].n_types == 1 && nfa->states[0].types[0] != nfa->states[1].types[0] &&
        d.n_preds[0] == 0 && !(flags & (SIESTA_F_RETURN_ALL | SIESTA_F_COUNT_MATCHES | SIESTA_F_MODE_HEAD))) {
        for (int k = 0; k < d.n_preds[1]; ++k)
            if (d.p_ref[1][k] != 0) return 0;
        return 2;
    }
    {   // NP1: positive states and exactly one Kleene state.  kleeneClosure+: predicates (after onlyAppearances) only
        // reference states before the Kleene state.  kleeneClosure* (one type): no predicate at all; `a* b ..` has
        // matches that start at a b, `a b c* d` matches without any c: they may lie outside the largest one, so returnAll
        // keeps the engine unless the state is the second one.
        int n_kleene = 0, k = 0;
        for (int s = 0; s < S; ++s) {
            if (nfa->states[s].kind == SIESTA_STATE_KLEENE_PLUS || nfa->states[s].kind == SIESTA_STATE_KLEENE_STAR) { ++n_kleene; k = s; }
            else if (!positive(s)) return 0;
        }
        if (n_kleene != 1 || S < 2 || d.n_preds[0] != 0) return 0;
        if (nfa->states[k].kind == SIESTA_STATE_KLEENE_STAR) {
            if (d.need_vv || nfa->states[k].n_types != 1 || (flags & SIESTA_F_MODE_HEAD)) return 0;
            if (k != 1 && (flags & SIESTA_F_RETURN_ALL)) return 0;
        }
        for (int s = 0; s < S; ++s)
            for (int q = 0; q < d.n_preds[s]; ++q)
                if (d.p_ref[s][q] >= k) return 0;
        if (d.need_vv) return (flags & SIESTA_F_RETURN_ALL) ? 0 : 3;
        if ((flags & SIESTA_F_RETURN_ALL) && !(flags & SIESTA_F_EVT_POS)) return 0;
        return 3;
    }
}

int validate_nfa(const siesta_nfa* nfa, uint32_t flags, DevNfa* out) {
    if (!nfa || nfa->n_states < 1 || nfa->n_states > SIESTA_MAX_STATES) {
        set_error("NFA must have 1.." + std::to_string(SIESTA_MAX_STATES) + " states");
        return SIESTA_E_INVALID;
    }
    DevNfa d;
    std::memset(&d, 0, sizeof(d));
    d.n_states = nfa->n_states;
    const bool ignore_preds = (flags & SIESTA_F_ONLY_APPEARANCES) != 0;
    for (int s = 0; s < nfa->n_states; ++s) {
        const siesta_state& st = nfa->states[s];
        if (st.kind < SIESTA_STATE_NORMAL || st.kind > SIESTA_STATE_OR) {
            set_error("unknown state kind");
            return SIESTA_E_INVALID;
        }
        if (st.n_types < 1 || st.n_types > SIESTA_MAX_OR_TYPES) {
            set_error("state needs 1.." + std::to_string(SIESTA_MAX_OR_TYPES) + " event types");
            return SIESTA_E_INVALID;
        }
        if ((st.kind == SIESTA_STATE_NORMAL || st.kind == SIESTA_STATE_KLEENE_PLUS) && st.n_types != 1) {
            set_error("normal and kleeneClosure states carry exactly one event type (State.java:108-133)");
            return SIESTA_E_INVALID;
        }
        if (st.n_preds < 0 || st.n_preds > SIESTA_MAX_PREDS) {
            set_error("too many predicates on one state");
            return SIESTA_E_INVALID;
        }
        d.kind[s] = (uint8_t)st.kind;
        if (st.kind == SIESTA_STATE_NEGATIVE) d.init_st |= 2u << (2 * s);
        if (st.kind == SIESTA_STATE_KLEENE_STAR) d.init_st |= 3u << (2 * s);
        if (st.kind == SIESTA_STATE_KLEENE_STAR || st.kind == SIESTA_STATE_KLEENE_PLUS) d.any_kleene = 1;
        d.all2 |= 2u << (2 * s);
        const int np = ignore_preds ? 0 : st.n_preds;
        d.n_preds[s] = (uint8_t)np;
        for (int k = 0; k < np; ++k) {
            const siesta_pred& p = st.preds[k];
            if ((p.attr != SIESTA_ATTR_POSITION && p.attr != SIESTA_ATTR_TIMESTAMP) ||
                (p.op != SIESTA_OP_LE && p.op != SIESTA_OP_GE) || p.ref_state < 0 || p.ref_state >= nfa->n_states ||
                p.constant < 0) {
                set_error("malformed predicate (attr/op/ref_state/constant)");
                return SIESTA_E_INVALID;
            }
            d.p_attr[s][k] = (uint8_t)p.attr;
            d.p_op[s][k] = (uint8_t)p.op;
            d.p_ref[s][k] = (uint8_t)p.ref_state;
            d.p_c[s][k] = p.constant;
            d.has_vv |= (uint8_t)(1u << p.ref_state);
            d.need_vv = 1;
        }
    }
    // Dominated-run merging (detect_engine.cuh) treats two runs with the same packed state and the same
    // value-vector family as interchangeable.  Inside ONE event the engine evaluates runs sequentially and
    // a sibling may rewrite the shared value vector between them (Run.clone is shallow, Run.java:319-327),
    // so that only holds if no event type can both write a referenced slot (an event of state `ref`'s
    // types) and trigger a predicate that reads it (an event of the predicate's own state's types).
    d.merge_safe = 1;
    for (int s = 0; s < nfa->n_states && d.merge_safe; ++s)
        for (int k = 0; k < d.n_preds[s] && d.merge_safe; ++k) {
            const siesta_state& a = nfa->states[s];
            const siesta_state& b = nfa->states[nfa->states[s].preds[k].ref_state];
            for (int i = 0; i < a.n_types; ++i)
                for (int j = 0; j < b.n_types; ++j)
                    if (a.types[i] == b.types[j]) d.merge_safe = 0;
        }
    // kleeneClosureInitialized is sticky across states (Run.proceed never resets it), so a run entering a SECOND
    // Kleene state that owns a value vector calls updateValueVector on a slot only a sibling may have initialised
    // (Run.java:266-274, 361): whether it throws depends on its list position relative to that sibling.
    {
        bool seen_kleene = false;
        for (int s = 0; s < nfa->n_states; ++s) {
            const bool kl = nfa->states[s].kind == SIESTA_STATE_KLEENE_PLUS || nfa->states[s].kind == SIESTA_STATE_KLEENE_STAR;
            if (kl && seen_kleene && ((d.has_vv >> s) & 1u)) d.merge_safe = 0;
            seen_kleene = seen_kleene || kl;
        }
    }
    // Which value-vector slots can a run sitting at state c still read?  Those referenced by predicates of states
    // >= c (a negative state also evaluates the next state's predicates: Engine.java:1165-1180).
    d.cross_safe = d.merge_safe;
    for (int c = 0; c <= SIESTA_MAX_STATES; ++c) {
        d.relmask[c] = 0;
        d.kmax[c] = -1;
        for (int s = c; s < nfa->n_states; ++s)
            for (int k = 0; k < d.n_preds[s]; ++k) {
                const int ref = nfa->states[s].preds[k].ref_state;
                d.relmask[c] |= 0xFFull << (8 * ref);
                if (ref > d.kmax[c]) d.kmax[c] = (int8_t)ref;
                if (ref >= s) d.cross_safe = 0;
            }
    }
    if (flags & SIESTA_F_MODE_HEAD) {
        // Engine.createNewRun's trailing block (Engine.java:983-996) only acts when states[1] is
        // kleeneClosure*, and throws for single-state NFAs; both have no defined reference output.
        if (nfa->n_states < 2 || nfa->states[1].kind == SIESTA_STATE_KLEENE_STAR) {
            set_error("SIESTA_F_MODE_HEAD: the reference's HEAD engine has no defined output for this NFA "
                      "(state 1 is kleeneClosure* or the NFA has one state); see DESIGN.md");
            return SIESTA_E_UNSUPPORTED;
        }
    }
    d.fast_class = (uint8_t)classify_fast(nfa, d, flags);
    *out = d;
    return SIESTA_OK;
}


// Kernel K1-P's window program (detect_fast.cuh).  Eligible: class NK, only the first-largest occurrence, and every
// predicate reads the attribute that is the bit index of ONE index space:
//   EventTs route  (Utils.java:51-58): position = index in the filtered list -> rank space; a timestamp predicate
//                  needs relative seconds (staged kernel);
//   EventPos route (Utils.java:59-62): timestamp = index in the filtered list -> rank space, position = in-trace
//                  position -> raw slots; an NFA that mixes the two stays on the staged kernel.
int nkw_build(const DevNfa& dn, uint32_t flags, NkwProgram* out) {
    std::memset(out, 0, sizeof(*out));
    // (returnAll: kernel K1-P answers the traces with a single engine match - its occurrence IS clearOccurrences(true)'s
    //  selection - and hands the others to the staged kernel, which knows the overlap test)
    if (dn.fast_class != 1 /* FAST_NK */) return NKW_NONE;
    const bool evt_pos = (flags & SIESTA_F_EVT_POS) != 0;
    bool any_pos = false, any_ts = false;
    out->n_states = dn.n_states;
    for (int s = 0; s < dn.n_states; ++s) {
        out->neg[s] = dn.kind[s] == SIESTA_STATE_NEGATIVE;
        out->n_preds[s] = dn.n_preds[s];
        for (int k = 0; k < dn.n_preds[s]; ++k) {
            (dn.p_attr[s][k] == SIESTA_ATTR_POSITION ? any_pos : any_ts) = true;
            const bool le = dn.p_op[s][k] == SIESTA_OP_LE;
            const int64_t c = std::min<int64_t>(dn.p_c[s][k], 64);
            out->le[s][k] = le;
            out->sh[s][k] = (uint8_t)(8 * dn.p_ref[s][k]);
            out->cc[s][k] = (uint8_t)(c + (le ? 1 : 0));
        }
    }
    // Markov form?
    out->markov = 1;
    for (int k = 1; k < dn.n_states; ++k) {
        if (out->neg[k]) {
            if (dn.n_preds[k] || dn.n_preds[k + 1]) out->markov = 0;   // (class NK: a negative state is never the last one)
            continue;
        }
        const int j = out->neg[k - 1] ? k - 2 : k - 1;
        out->prev[k] = (uint8_t)j;
        int64_t ge = 1, le = 255;
        for (int q = 0; q < dn.n_preds[k]; ++q) {
            if (dn.p_ref[k][q] != j) out->markov = 0;
            const int64_t c = std::min<int64_t>(dn.p_c[k][q], 200);
            if (dn.p_op[k][q] == SIESTA_OP_LE) le = std::min(le, c);
            else ge = std::max(ge, c);
        }
        out->m_ge[k] = (uint8_t)std::min<int64_t>(ge, 200);
        out->m_le[k] = (uint8_t)le;
    }
    if (!evt_pos) return any_ts ? NKW_NONE : NKW_RANK;
    if (any_pos && any_ts) return NKW_NONE;
    return any_pos ? NKW_RAW : NKW_RANK;
}

// Per-activity lookup: bit k = the type belongs to state k (State.checkEventType / AdditionalState.checkEventType);
// bit 8+k = the type is state k's first type (State.getEventType, used by Engine.java:661).
void build_lut(const siesta_nfa* nfa, const DevNfa& dn, int32_t n_activities, uint32_t flags, std::vector<uint16_t>& lut,
               int* needs_ts, int* n_positive) {
    lut.assign((size_t)std::max(1, n_activities), 0);
    bool time_pred = false;
    int np = 0;
    for (int s = 0; s < nfa->n_states; ++s) {
        const siesta_state& st = nfa->states[s];
        if (st.kind != SIESTA_STATE_NEGATIVE) ++np;
        for (int k = 0; k < st.n_types; ++k) {
            const int ty = st.types[k];
            if (ty < 0 || ty >= n_activities) continue;  // activity absent from this log: never matches
            lut[ty] |= (uint16_t)(1u << s);
            if (k == 0) lut[ty] |= (uint16_t)(1u << (8 + s));
        }
        for (int k = 0; k < dn.n_preds[s]; ++k) time_pred |= st.preds[k].attr == SIESTA_ATTR_TIMESTAMP;
    }
    const bool evt_pos = (flags & SIESTA_F_EVT_POS) != 0;
    const bool return_all = (flags & SIESTA_F_RETURN_ALL) != 0;
    // relative seconds are needed by time predicates and by Occurrence.overlaps on the EventTs route
    *needs_ts = (!evt_pos && (time_pred || return_all)) ? 1 : 0;
    *n_positive = np;
}

}  // namespace siesta
