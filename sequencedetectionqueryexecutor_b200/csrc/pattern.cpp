// pattern.cpp — host-side pattern compiler: SIESTA pattern -> NFA states + predicates.
//
// Restates ComplexPattern.getNfa / getStatesWithoutConstraints / getItSimpler
// (J/model/Patterns/ComplexPattern.java:181-283), SimplePattern.getNfa (SimplePattern.java:96-104) and
// SIESTAPattern.generatePredicates (SIESTAPattern.java:131-149).  J/ = com/datalab/siesta/queryprocessor/.
#include <cstring>
#include <string>

#include "../../include/siesta_gpu.h"

namespace siesta {
void set_error(const std::string& msg);
}

namespace {

// SIESTAPattern.generatePredicates: every constraint whose posB is this state contributes one predicate.
int add_predicates(siesta_state* st, int m, const siesta_constraint* cs, int nc) {
    for (int i = 0; i < nc; ++i) {
        const siesta_constraint& c = cs[i];
        if (c.pos_b != m) continue;
        if (st->n_preds >= SIESTA_MAX_PREDS) {
            siesta::set_error("more than " + std::to_string(SIESTA_MAX_PREDS) + " constraints end on one state");
            return SIESTA_E_UNSUPPORTED;
        }
        siesta_pred& p = st->preds[st->n_preds++];
        p.ref_state = c.pos_a;
        p.reserved = 0;
        p.op = c.method == SIESTA_METHOD_WITHIN ? SIESTA_OP_LE : SIESTA_OP_GE;
        if (c.kind == SIESTA_CONSTRAINT_GAP) {
            p.attr = SIESTA_ATTR_POSITION;
            p.constant = c.value;
        } else {
            p.attr = SIESTA_ATTR_TIMESTAMP;
            // TimeConstraint.getConstraintInSeconds (J/model/Constraints/TimeConstraint.java:45-49)
            p.constant = c.granularity == SIESTA_GRAN_MINUTES ? c.value * 60 : c.granularity == SIESTA_GRAN_HOURS ? c.value * 3600 : c.value;
        }
    }
    return SIESTA_OK;
}

}  // namespace

extern "C" int siesta_pattern_compile(const siesta_event_symbol* symbols, int32_t n_symbols, const siesta_constraint* constraints,
                                      int32_t n_constraints, int32_t only_appearances, siesta_nfa* out) {
    if (!symbols || n_symbols < 1 || !out || n_constraints < 0 || (n_constraints && !constraints)) {
        siesta::set_error("siesta_pattern_compile: bad argument");
        return SIESTA_E_INVALID;
    }
    std::memset(out, 0, sizeof(*out));
    for (int i = 0; i < n_constraints; ++i) {
        const siesta_constraint& c = constraints[i];
        if (c.pos_a >= c.pos_b || c.pos_a < 0) {  // Constraint.hasError (Constraint.java:98-100) -> HTTP 400 upstream
            siesta::set_error("constraint needs 0 <= posA < posB");
            return SIESTA_E_INVALID;
        }
        if (c.value < 0 || (c.kind != SIESTA_CONSTRAINT_GAP && c.kind != SIESTA_CONSTRAINT_TIME)) {
            siesta::set_error("constraint value must be >= 0 and kind gap|time");
            return SIESTA_E_INVALID;
        }
    }
    int n_states = 0;
    int i = 0;
    // getItSimpler (:181-190) routes all-"_" patterns to SimplePattern; its states are the same "normal" states
    // getStatesWithoutConstraints (:216-272) produces for "_", so one loop covers both.
    while (i < n_symbols) {
        if (n_states >= SIESTA_MAX_STATES) {
            siesta::set_error("pattern has more than " + std::to_string(SIESTA_MAX_STATES) + " states");
            return SIESTA_E_UNSUPPORTED;
        }
        const siesta_event_symbol& es = symbols[i];
        siesta_state& st = out->states[n_states];
        st.n_types = 1;
        st.types[0] = es.activity;
        switch (es.symbol) {
            case SIESTA_SYM_NORMAL: st.kind = SIESTA_STATE_NORMAL; ++i; break;
            case SIESTA_SYM_PLUS: st.kind = SIESTA_STATE_KLEENE_PLUS; ++i; break;
            case SIESTA_SYM_STAR: st.kind = SIESTA_STATE_KLEENE_STAR; ++i; break;
            case SIESTA_SYM_NOT: st.kind = SIESTA_STATE_NEGATIVE; ++i; break;
            case SIESTA_SYM_OR: {
                // every following entry that shares this entry's position joins the "or" state (:242-255)
                st.kind = SIESTA_STATE_OR;
                st.n_types = 0;
                int k = i;
                while (k < n_symbols && symbols[k].position == es.position) {
                    if (st.n_types >= SIESTA_MAX_OR_TYPES) {
                        siesta::set_error("more than " + std::to_string(SIESTA_MAX_OR_TYPES) + " alternatives in one || group");
                        return SIESTA_E_UNSUPPORTED;
                    }
                    st.types[st.n_types++] = symbols[k].activity;
                    ++k;
                }
                i = k;
                break;
            }
            default:
                siesta::set_error("unknown symbol");
                return SIESTA_E_INVALID;
        }
        ++n_states;
    }
    out->n_states = n_states;
    if (!only_appearances) {
        for (int c = 0; c < n_constraints; ++c)
            if (constraints[c].pos_b < n_states && constraints[c].pos_a >= n_states) {
                siesta::set_error("constraint references a state beyond the pattern");
                return SIESTA_E_INVALID;
            }
        for (int m = 0; m < n_states; ++m) {
            int rc = add_predicates(&out->states[m], m, constraints, n_constraints);
            if (rc) return rc;
        }
    }
    return SIESTA_OK;
}

// ---------------------------------------------------------------------------------------------- pair extraction (P3)
// Restates ComplexPattern.extractPairsForPatternDetection / splitWithOr / generateCombinations
// (J/model/Patterns/ComplexPattern.java:75-171) over SIESTAPattern.extractPairsForPatternDetection
// (J/model/Patterns/SIESTAPattern.java:34-69).  EventPair equality is by the two names only
// (J/model/Events/EventPair.java:65-81), so a pair set is a set of (a, b) activity ids.
#include <algorithm>
#include <set>
#include <utility>
#include <vector>

namespace {
typedef std::pair<int32_t, int32_t> Pair;
struct Sym {
    int32_t activity, symbol;
    bool operator<(const Sym& o) const { return activity != o.activity ? activity < o.activity : symbol < o.symbol; }
    bool operator==(const Sym& o) const { return activity == o.activity && symbol == o.symbol; }
};
inline bool non_empty(int32_t symbol) { return symbol == SIESTA_SYM_NORMAL || symbol == SIESTA_SYM_PLUS; }
}  // namespace

extern "C" int siesta_pattern_extract_pairs(const siesta_event_symbol* symbols, int32_t n_symbols,
                                            const siesta_constraint* constraints, int32_t n_constraints, int32_t from_or_till_set,
                                            int32_t cap_expansions, int32_t cap_pairs, int32_t* n_expansions, int32_t* true_off,
                                            int32_t* true_a, int32_t* true_b, int32_t* all_off, int32_t* all_a, int32_t* all_b) {
    if (!symbols || n_symbols < 1 || n_constraints < 0 || (n_constraints && !constraints) || !n_expansions || !true_off || !all_off ||
        cap_expansions < 1 || cap_pairs < 0 || (cap_pairs && (!true_a || !true_b || !all_a || !all_b))) {
        siesta::set_error("siesta_pattern_extract_pairs: bad argument");
        return SIESTA_E_INVALID;
    }
    // splitWithOr (:130-152): one set of symbols per pattern position; "||" entries become "_"
    std::vector<std::vector<Sym>> same_pos;
    for (int i = 0; i < n_symbols; ++i) {
        Sym s{symbols[i].activity, symbols[i].symbol == SIESTA_SYM_OR ? SIESTA_SYM_NORMAL : symbols[i].symbol};
        const int pos = symbols[i].position;
        if ((int)same_pos.size() == pos) same_pos.push_back({s});
        else if (pos >= 0 && pos < (int)same_pos.size()) {
            std::vector<Sym>& v = same_pos[pos];
            if (std::find(v.begin(), v.end(), s) == v.end()) v.push_back(s);  // HashSet<EventSymbol>
        } else {
            siesta::set_error("symbol positions must be 0, 1, 2 ... without gaps (ComplexPattern.splitWithOr indexes by position)");
            return SIESTA_E_REFERENCE_THROWS;
        }
    }
    // generateCombinations (:161-171), duplicates removed (the reference returns a HashSet of the lists)
    std::set<std::vector<Sym>> seen;
    std::vector<std::vector<Sym>> expansions;
    std::vector<size_t> choice(same_pos.size(), 0);
    while (true) {
        std::vector<Sym> cur;
        for (size_t p = 0; p < same_pos.size(); ++p) cur.push_back(same_pos[p][choice[p]]);
        if (seen.insert(cur).second) expansions.push_back(cur);
        int p = (int)same_pos.size() - 1;
        while (p >= 0 && ++choice[p] == same_pos[p].size()) choice[p--] = 0;
        if (p < 0) break;
        if ((int)expansions.size() > cap_expansions) break;
    }
    if ((int)expansions.size() > cap_expansions) {
        siesta::set_error("more OR-expansions than cap_expansions");
        return SIESTA_E_INVALID;
    }
    std::set<int32_t> constraint_pos;
    for (int c = 0; c < n_constraints; ++c) {
        constraint_pos.insert(constraints[c].pos_a);
        constraint_pos.insert(constraints[c].pos_b);
    }
    int n_true = 0, n_all = 0;
    true_off[0] = all_off[0] = 0;
    for (size_t x = 0; x < expansions.size(); ++x) {
        const std::vector<Sym>& ev = expansions[x];
        std::set<Pair> tp, ap;
        std::vector<std::pair<int32_t, int>> l;  // (activity, position in the expansion) of the "_" and "+" events
        bool have_first = false;
        int32_t first_non_empty = 0;
        for (int i = 0; i < (int)ev.size(); ++i) {
            const Sym& es = ev[i];
            if (!have_first && non_empty(es.symbol)) {
                have_first = true;
                first_non_empty = es.activity;
            }
            switch (es.symbol) {
                case SIESTA_SYM_NORMAL: l.push_back({es.activity, i}); break;
                case SIESTA_SYM_PLUS:
                    l.push_back({es.activity, i});
                    ap.insert({es.activity, es.activity});
                    break;
                case SIESTA_SYM_NOT:
                case SIESTA_SYM_STAR:
                    if (have_first) ap.insert({first_non_empty, es.activity});
                    else if (i + 1 < (int)ev.size()) {
                        // the reference's search loop (:108-113) never advances k: it ends only if the very next symbol is "_" / "+"
                        if (!non_empty(ev[i + 1].symbol)) {
                            siesta::set_error("the reference loops forever on a leading '*' / '!' that is not followed by a '_' / '+' symbol "
                                              "(ComplexPattern.java:108-113)");
                            return SIESTA_E_REFERENCE_THROWS;
                        }
                        ap.insert({es.activity, ev[i + 1].activity});
                    }
                    ap.insert({es.activity, es.activity});
                    break;
                default:
                    siesta::set_error("unknown symbol");
                    return SIESTA_E_INVALID;
            }
        }
        // SIESTAPattern.extractPairsForPatternDetection (:34-69)
        for (int i = 0; i + 1 < (int)l.size(); ++i) {
            if (constraint_pos.count(l[i].second)) ap.insert({l[i].first, l[i].first});
            for (int j = i + 1; j < (int)l.size(); ++j) tp.insert({l[i].first, l[j].first});
        }
        if (from_or_till_set)
            for (const auto& e : l) ap.insert({e.first, e.first});
        ap.insert(tp.begin(), tp.end());
        if (n_true + (int)tp.size() > cap_pairs || n_all + (int)ap.size() > cap_pairs) {
            siesta::set_error("more pairs than cap_pairs");
            return SIESTA_E_INVALID;
        }
        for (const Pair& p : tp) { true_a[n_true] = p.first; true_b[n_true] = p.second; ++n_true; }
        for (const Pair& p : ap) { all_a[n_all] = p.first; all_b[n_all] = p.second; ++n_all; }
        true_off[x + 1] = n_true;
        all_off[x + 1] = n_all;
    }
    *n_expansions = (int32_t)expansions.size();
    return SIESTA_OK;
}
