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
