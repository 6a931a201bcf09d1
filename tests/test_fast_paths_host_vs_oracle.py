"""The closed-form evaluators of kernel K1 (detect_fast.cuh: class NK = no Kleene state, class FK2 = `a+ b*`),
compiled for the host, must agree with the oracle bit for bit.  tests/soak_fast.py is the long-running version."""
import numpy as np
import pytest

import oracle
from sequencedetectionqueryexecutor_b200 import _abi as abi
from tests import gen, host_engine, soak_fast

N_, P_, S_, X_, O_ = abi.STATE_NORMAL, abi.STATE_KLEENE_PLUS, abi.STATE_KLEENE_STAR, abi.STATE_NEGATIVE, abi.STATE_OR


@pytest.mark.parametrize("seed0", [0, 500, 1000, 1500])
def test_random_nfas_of_the_fast_classes(seed0):
    assert soak_fast.main(seed0, 120) == 0


def _first_largest(matches):
    return max(matches, key=len) if matches else None  # max() returns the first maximal element


def test_fk2_ties_follow_the_run_list_order():
    """Streams built to tie: Q(0,m) vs Q(0,m'), P(i) vs Q(0,m).  Checked against the oracle's full emission list."""
    nfa = abi.make_nfa([dict(kind=P_, types=[0]), dict(kind=S_, types=[1], preds=[(abi.ATTR_TIMESTAMP, abi.OP_LE, 0, 3)])])
    rng = np.random.default_rng(7)
    for _ in range(400):
        n = int(rng.integers(2, 12))
        types = rng.integers(0, 2, size=n).astype(np.int32)
        ts_s = np.cumsum(rng.integers(0, 3, size=n))
        _, matches = oracle.run_stream(nfa, types, np.arange(n), ts_s)
        want = _first_largest(matches)
        off = np.array([0, n], dtype=np.int64)
        rc, got, _ = host_engine.detect(off, types, ts_s.astype(np.int64) * 1000, 2, nfa)
        assert rc == 0
        if want is None:
            assert got.n_traces == 0
        else:
            assert got.as_dict() == {0: [want]}


def test_class_boundaries_fall_back_to_the_engine():
    """NFAs just outside the two classes (same type twice, predicate on state 0, trailing negative ...) still agree
    with the oracle: they must be routed to the general engine."""
    off, act, ts = gen.make_log(200, 5, 40, 4, seed=11, max_gap_s=200)
    cases = [
        [dict(kind=P_, types=[0]), dict(kind=S_, types=[0])],                                      # a+ a*
        [dict(kind=P_, types=[0]), dict(kind=S_, types=[1]), dict(kind=N_, types=[2])],            # three states
        [dict(kind=X_, types=[0]), dict(kind=N_, types=[1])],                                      # leading negative
        [dict(kind=N_, types=[0]), dict(kind=X_, types=[1]), dict(kind=X_, types=[2]), dict(kind=N_, types=[3])],
        [dict(kind=N_, types=[0]), dict(kind=N_, types=[1], preds=[(abi.ATTR_POSITION, abi.OP_LE, 1, 2)])],  # self reference
    ]
    for states in cases:
        for flags in (0, abi.F_RETURN_ALL, abi.F_EVT_POS):
            nfa = abi.make_nfa(states)
            rc, got, _ = host_engine.detect(off, act, ts, 4, nfa, flags=flags)
            assert rc == 0
            want = oracle.detect(off, act, ts, nfa, flags=flags)
            ok, why = got.same_as(want)
            assert ok, (why, states, flags)


def test_np1_is_the_first_start_with_every_kleene_event_the_suffix_allows():
    """Class NP1 (one `+` or `*` state, no predicates): the first-largest occurrence and the engine's match count against the
    oracle's FULL emission list on dense two- and three-letter streams (ties between starts, Kleene type repeated in
    the prefix / suffix, Kleene state first / in the middle / last)."""
    shapes = [
        [dict(kind=N_, types=[0]), dict(kind=P_, types=[1]), dict(kind=N_, types=[2])],            # a b+ c
        [dict(kind=N_, types=[0]), dict(kind=P_, types=[1])],                                      # a b+
        [dict(kind=P_, types=[0]), dict(kind=N_, types=[1])],                                      # a+ b
        [dict(kind=P_, types=[0]), dict(kind=N_, types=[0]), dict(kind=N_, types=[1])],            # a+ a b
        [dict(kind=N_, types=[1]), dict(kind=P_, types=[1]), dict(kind=N_, types=[1])],            # b b+ b
        [dict(kind=N_, types=[0]), dict(kind=N_, types=[1]), dict(kind=P_, types=[0]), dict(kind=O_, types=[1, 2]), dict(kind=N_, types=[0])],
        # kleeneClosure*: the extra matches (a run cloned past the state, a run started at state 1) count and may be the result
        [dict(kind=N_, types=[0]), dict(kind=S_, types=[1]), dict(kind=N_, types=[2])],            # a b* c
        [dict(kind=N_, types=[0]), dict(kind=S_, types=[1])],                                      # a b*
        [dict(kind=S_, types=[0]), dict(kind=N_, types=[1])],                                      # a* b
        [dict(kind=S_, types=[0]), dict(kind=N_, types=[1]), dict(kind=N_, types=[2])],            # a* b c
        [dict(kind=N_, types=[0]), dict(kind=N_, types=[1]), dict(kind=S_, types=[0])],            # a b a*
        [dict(kind=N_, types=[0]), dict(kind=N_, types=[1]), dict(kind=S_, types=[2]), dict(kind=N_, types=[0])],   # a b c* a
        [dict(kind=O_, types=[0, 1]), dict(kind=S_, types=[1]), dict(kind=N_, types=[2])],         # (a|b) b* c
        [dict(kind=S_, types=[0]), dict(kind=N_, types=[0]), dict(kind=N_, types=[1])],            # a* a b
        [dict(kind=N_, types=[1]), dict(kind=S_, types=[1]), dict(kind=N_, types=[1])],            # b b* b
    ]
    rng = np.random.default_rng(23)
    for states in shapes:
        nfa = abi.make_nfa(states)
        for _ in range(150):
            n = int(rng.integers(2, 14))
            types = rng.integers(0, 3, size=n).astype(np.int32)
            ts_s = np.cumsum(rng.integers(0, 3, size=n))
            _, matches = oracle.run_stream(nfa, types, np.arange(n), ts_s)
            want = _first_largest(matches)
            off = np.array([0, n], dtype=np.int64)
            rc, got, _ = host_engine.detect(off, types, ts_s.astype(np.int64) * 1000, 3, nfa, flags=abi.F_COUNT_MATCHES)
            assert rc == 0
            if want is None:
                assert got.n_traces == 0 and got.n_matches_emitted == 0
            else:
                assert got.as_dict() == {0: [want]}, (states, types.tolist())
                assert got.n_matches_emitted == len(matches), (states, types.tolist())


@pytest.mark.parametrize("seed0", [900, 1420])
def test_np1_constraints_on_prefix_states_and_the_run_list_order(seed0):
    """Class NP1 with time constraints that reference prefix states on UNSORTED timestamps: the starts' Kleene-event sets
    are not nested, equal sizes completing at the same event are decided by the run list's order (seeds 919 and 1431 are
    decided by it: a later start's run sits earlier in the list)."""
    for seed in range(seed0, seed0 + 40):
        rng = np.random.default_rng(seed)
        n_act = int(rng.integers(2, 4))
        off, act, ts = gen.make_log(60, 3, int(rng.integers(6, 30)), n_act, seed=int(rng.integers(1 << 30)), max_gap_s=20)
        ts = ts.copy()
        rng.shuffle(ts)
        k = int(rng.integers(1, 3))
        n = k + 1 + int(rng.integers(0, 2))
        states = [dict(kind=P_ if s == k else N_, types=[int(rng.integers(0, n_act))], preds=[]) for s in range(n)]
        for _ in range(int(rng.integers(1, 4))):
            b = int(rng.integers(k, n))
            if len(states[b]["preds"]) < 4:
                states[b]["preds"].append((abi.ATTR_TIMESTAMP, abi.OP_LE if rng.random() < 0.5 else abi.OP_GE,
                                           int(rng.integers(0, k)), int(rng.integers(0, 400))))
        flags = abi.F_COUNT_MATCHES if rng.random() < 0.5 else 0
        nfa = abi.make_nfa(states)
        rc, got, _ = host_engine.detect(off, act, ts, n_act, nfa, flags=flags)
        assert rc == 0
        ok, why = got.same_as(oracle.detect(off, act, ts, nfa, flags=flags))
        assert ok, (seed, why, states)


def test_which_shapes_take_a_closed_form():
    """validate_nfa's class decision (csrc/nfa.cpp, the product's host code, through the harness): the reference's own test
    shapes (tests/kat.py: EvaluateNewQueries.java / EvaluateComplexQueries.java) and BASELINE.json's two /detection queries
    all take a closed form; shapes the derivations do not cover keep the run-list engine."""
    from tests.kat import KATS
    NONE, NK, FK2, NP1 = host_engine.FAST_NONE, host_engine.FAST_NK, host_engine.FAST_FK2, host_engine.FAST_NP1
    want = {"A,(C|D),B": NK, "A,!C,B": NK, "A,B*,E": NP1, "A*,B": NP1, "A*,B,E": NP1, "A,B*": NP1, "A,B,A*": NP1,
            "!A,B,C": NONE, "B,C,!A": NONE,   # leading / trailing negative state
            "A,B,C": NK, "A,E,C": NK, "A,B": NK, "(A|B),C": NK, "D,(A|B),E": NK, "A+,B,E": NP1, "A,B+,E": NP1, "A,B,A+": NP1,
            "(A|B),B*,E": NP1, "A,(C|D),B gap within 2 (0,1)": NK, "A,!D,E gap within 2 (0,1)": NK,
            "A,B*,E gap within 1 (0,1),(1,2)": NONE}   # constraints on and behind a `*` state
    assert set(want) == {k["name"] for k in KATS}
    for k in KATS:
        assert host_engine.fast_class(abi.make_nfa(k["states"]), abi.F_EVT_POS) == want[k["name"]], k["name"]
    gap6 = [dict(kind=N_, types=[0]), dict(kind=O_, types=[1, 2], preds=[(abi.ATTR_POSITION, abi.OP_LE, 0, 10)]),
            dict(kind=X_, types=[3]), dict(kind=N_, types=[4]), dict(kind=N_, types=[5], preds=[(abi.ATTR_POSITION, abi.OP_GE, 3, 2)])]
    kleene = [dict(kind=P_, types=[0]), dict(kind=S_, types=[1], preds=[(abi.ATTR_TIMESTAMP, abi.OP_LE, 0, 600)])]
    assert host_engine.fast_class(abi.make_nfa(gap6), 0) == NK                       # BASELINE configs[4]
    assert host_engine.fast_class(abi.make_nfa(gap6), abi.F_RETURN_ALL) == NK
    assert host_engine.fast_class(abi.make_nfa(kleene), 0) == FK2                    # BASELINE configs[1]
    assert host_engine.fast_class(abi.make_nfa(kleene), abi.F_RETURN_ALL) == NONE
    star = lambda k, n: [dict(kind=S_ if s == k else N_, types=[s]) for s in range(n)]   # noqa: E731
    for k, n in ((0, 2), (0, 3), (1, 2), (1, 3), (2, 3), (2, 4)):
        assert host_engine.fast_class(abi.make_nfa(star(k, n)), 0) == NP1
        assert host_engine.fast_class(abi.make_nfa(star(k, n)), abi.F_COUNT_MATCHES) == NP1
        # returnAll: only on the EventPos route, and for `*` only as the second state (other extra matches may lie outside the largest)
        assert host_engine.fast_class(abi.make_nfa(star(k, n)), abi.F_RETURN_ALL) == NONE
        assert host_engine.fast_class(abi.make_nfa(star(k, n)), abi.F_RETURN_ALL | abi.F_EVT_POS) == (NP1 if k == 1 else NONE)
    two = [dict(kind=N_, types=[0]), dict(kind=S_, types=[1]), dict(kind=P_, types=[2])]
    assert host_engine.fast_class(abi.make_nfa(two), 0) == NONE                      # two Kleene states
    c = [dict(kind=N_, types=[0]), dict(kind=S_, types=[1]), dict(kind=N_, types=[2], preds=[(abi.ATTR_POSITION, abi.OP_LE, 0, 5)])]
    assert host_engine.fast_class(abi.make_nfa(c), 0) == NONE                        # a constraint next to a `*` state
    assert host_engine.fast_class(abi.make_nfa(c), abi.F_ONLY_APPEARANCES) == NP1    # ... dropped by onlyAppearances
    assert host_engine.fast_class(abi.make_nfa(star(1, 3)), abi.F_LITERAL_RUNS) == NONE


def test_bench_workloads_run_on_the_kernels_their_labels_name():
    """bench.py labels every workload with the kernel it measures (`roofline.kernel`); the label must be the evaluator the
    library's own class decision picks for that NFA and those flags."""
    import bench
    label = {host_engine.FAST_NK: ("detect_nkp_kernel",), host_engine.FAST_FK2: ("FAST_FK2",), host_engine.FAST_NP1: ("FAST_NP1",),
             host_engine.FAST_NONE: ("FAST_NONE",)}
    assert bench.DEFAULT_WORKLOAD == "detection_gap6_100Mx50"   # BASELINE.json configs[4], the metric's configuration
    for name, wl in bench.WORKLOADS.items():
        c = host_engine.fast_class(abi.make_nfa(wl["states"]), wl.get("flags", 0))
        assert c >= 0, name
        assert any(t in wl["kernel"] for t in label[c]), (name, c, wl["kernel"])
