"""Pins the pattern compiler (siesta_pattern_compile, csrc/pattern.cpp) to the reference's OWN tests of
ComplexPattern.getNfa: the 14 scenarios of src/test/java/com/datalab/siesta/queryprocessor/SaseConnection/
EvaluateComplexQueries.java:45-404.  Each case below is the EventSymbol / GapConstraint list the Java test builds
(same names, positions and symbols, including the repeated positions the Java tests use) and the match count the Java
test asserts.  The pattern goes through siesta_pattern_compile; the compiled NFA then runs
  * on the oracle (CPU) - asserted count, and the compiled states must equal the hand-built ones of tests/kat.py;
  * through siesta_detect on the GPU (-m gpu) - the engine's match count (SIESTA_F_COUNT_MATCHES) must be the asserted
    one and the selected occurrences must equal the oracle's.
Nothing here feeds one compiled NFA to both sides of a comparison without the Java-asserted number in between."""
import numpy as np
import pytest

import oracle
from sequencedetectionqueryexecutor_b200 import _abi as abi
from sequencedetectionqueryexecutor_b200.sase import ActivityDictionary, ComplexPattern, EventSymbol, GapConstraint
from tests import kat

ES = EventSymbol
# (java test, line of the assertion, eventsWithSymbols, constraints, asserted matches, name of the same scenario in tests/kat.py)
JAVA_CASES = [
    ("testOrState", 66, [ES("A", 0, "_"), ES("C", 1, "||"), ES("D", 1, "_"), ES("B", 2, "_")], [], 2, "A,(C|D),B"),
    ("testNegativeState", 91, [ES("A", 0, "_"), ES("C", 1, "!"), ES("B", 2, "_")], [], 2, "A,!C,B"),
    ("testKleeneStarState", 116, [ES("A", 0, "_"), ES("B", 1, "*"), ES("E", 1, "_")], [], 7, "A,B*,E"),
    ("testKleeneStarBeginState", 140, [ES("A", 0, "*"), ES("B", 1, "_")], [], 8, "A*,B"),
    ("testKleeneStarBegin3State", 165, [ES("A", 0, "*"), ES("B", 1, "_"), ES("E", 1, "_")], [], 8, "A*,B,E"),
    ("testKleeneStarEndState", 189, [ES("A", 0, "_"), ES("B", 1, "*")], [], 7, "A,B*"),
    ("testKleeneStarEnd3State", 214, [ES("A", 0, "_"), ES("B", 1, "_"), ES("A", 1, "*")], [], 5, "A,B,A*"),
    ("testNegativeBeginState", 239, [ES("A", 0, "!"), ES("B", 1, "_"), ES("C", 1, "_")], [], 1, "!A,B,C"),
    ("testNegativeEndState", 264, [ES("B", 0, "_"), ES("C", 1, "_"), ES("A", 1, "!")], [], 1, "B,C,!A"),
    ("testNormalStates", 291, [ES("A", 0, "_"), ES("B", 1, "_"), ES("C", 2, "_")], [], 1, "A,B,C"),
    ("testOrBeginKleene", 317, [ES("A", 0, "||"), ES("B", 0, "_"), ES("B", 1, "*"), ES("E", 2, "_")], [], 9, "(A|B),B*,E"),
    ("testWithGapConstraints", 346, [ES("A", 0, "_"), ES("C", 1, "||"), ES("D", 1, "_"), ES("B", 2, "_")],
     [GapConstraint(0, 1, 2)], 1, "A,(C|D),B gap within 2 (0,1)"),
    ("testNegativeGapConstraint", 374, [ES("A", 0, "_"), ES("D", 1, "!"), ES("E", 2, "_")], [GapConstraint(0, 1, 2)], 2,
     "A,!D,E gap within 2 (0,1)"),
    ("testKleeneStarWithGapConstraint", 403, [ES("A", 0, "_"), ES("B", 1, "*"), ES("E", 1, "_")],
     [GapConstraint(0, 1, 1), GapConstraint(1, 2, 1)], 2, "A,B*,E gap within 1 (0,1),(1,2)"),
]
ACTS = ActivityDictionary(["A", "B", "C", "D", "E"])   # ids as in tests/kat.py
TYPES = np.array(kat.STREAM_TYPES, dtype=np.int32)      # A B A C D A B E (EvaluateComplexQueries.java:29-36)


def _states_of(nfa):
    out = []
    for s in range(nfa.n_states):
        st = nfa.states[s]
        out.append({"kind": st.kind, "types": [st.types[k] for k in range(st.n_types)],
                    "preds": [(st.preds[k].attr, st.preds[k].op, st.preds[k].ref_state, st.preds[k].constant) for k in range(st.n_preds)]})
    return out


@pytest.mark.parametrize("case", JAVA_CASES, ids=[c[0] for c in JAVA_CASES])
def test_compiled_nfa_reproduces_the_java_assertion_on_the_oracle(case):
    name, line, symbols, constraints, expected, kat_name = case
    nfa = ComplexPattern(symbols, constraints)._compile(ACTS, only_appearances=False)
    # EventPos stream: id = position, timestamp = list index (Utils.transformToSaseEvents, Utils.java:59-62)
    _, matches = oracle.run_stream(nfa, TYPES, np.arange(8), np.arange(8))
    assert len(matches) == expected, f"EvaluateComplexQueries.java:{line} asserts {expected}"
    hand = next(k for k in kat.KATS if k["name"] == kat_name)
    assert matches == hand["matches"]
    want_states = [{"kind": s["kind"], "types": list(s["types"]), "preds": [tuple(p) for p in s.get("preds", ())]} for s in hand["states"]]
    assert _states_of(nfa) == want_states, "compiled states differ from the NFA the Java compiler builds (tests/kat.py)"


@pytest.mark.gpu
@pytest.mark.parametrize("case", JAVA_CASES, ids=[c[0] for c in JAVA_CASES])
def test_compiled_nfa_reproduces_the_java_assertion_on_the_gpu(case):
    from sequencedetectionqueryexecutor_b200 import api
    name, line, symbols, constraints, expected, _ = case
    nfa = ComplexPattern(symbols, constraints)._compile(ACTS, only_appearances=False)
    off = np.array([0, 8], dtype=np.int64)
    ts = np.arange(8, dtype=np.int64) * 1000
    flags = abi.F_EVT_POS | abi.F_COUNT_MATCHES
    with api.Context(0) as ctx:
        log = ctx.load_log(off, TYPES, ts, 5)
        got = log.detect(nfa, flags=flags)
        got_all = log.detect(nfa, flags=flags | abi.F_RETURN_ALL)
        log.close()
    assert got.n_matches_emitted == expected, f"EvaluateComplexQueries.java:{line} asserts {expected}"
    ok, why = got.same_as(oracle.detect(off, TYPES, ts, nfa, flags=flags))
    assert ok, why
    ok, why = got_all.same_as(oracle.detect(off, TYPES, ts, nfa, flags=flags | abi.F_RETURN_ALL))
    assert ok, why
