"""Why-not-match: the literal oracle against the reference's own tests, kernel W's evaluation
(csrc/wnm.cuh compiled for the host) against the oracle on the CPU, and the kernel through the C-ABI on the GPU."""
import numpy as np
import pytest

import oracle
from sequencedetectionqueryexecutor_b200 import _abi as abi
from tests import host_engine

GAP, TIME, WITHIN, ATLEAST = abi.WNM_GAP, abi.WNM_TIME, abi.WNM_WITHIN, abi.WNM_ATLEAST


def reference_trace():
    """WhyNotMatchSASETest.getEvents (:40-47): A at 101 s, B at 104 s, C at 109 s"""
    return np.array([0, 3], dtype=np.int64), np.array([0, 1, 2], dtype=np.int32), np.array([101000, 104000, 109000], dtype=np.int64)


def test_reference_uncertain_stream_sizes():
    """WhyNotMatchSASETest.testGetUncertainStream (:31-38): 7 x 3 events for u = 3, step = 1; 6 x 3 for u = 5, step = 2"""
    assert oracle.wnm_stream_size([101, 104, 109], 3, 1) == 7 * 3
    assert oracle.wnm_stream_size([101, 104, 109], 5, 2) == 6 * 3
    assert oracle.wnm_stream_size([1], 3, 1) == 5            # Math.max(original - u, 0): 0 .. 4


def test_reference_evaluation_known_answer():
    """WhyNotMatchSASETest.testEvaluation (:70-81): pattern A B C, (0,1) within 2 s, (1,2) at least 7 s, u = 3, step = 1, k = 3
    -> exactly one almost-match, not empty.  Its content, worked by hand from the engine's rules: start A@101 (no change);
    its runs take B@102 and B@103, so the value vector of state 2 holds B@103 when the C's arrive: C must be >= 110.
    A@101, B@103 (change 1), C@110 (change 1): total 2; every other start costs 3 or more."""
    off, act, ts = reference_trace()
    cons = [(0, 1, TIME, WITHIN, 2), (1, 2, TIME, ATLEAST, 7)]
    got = oracle.why_not_match(off, act, ts, [0, 1, 2], cons, 3, 1, 3)
    assert got.n_traces == 1 and got.trace_idx.tolist() == [0]
    assert got.total_change.tolist() == [2]
    assert got.ev_value.tolist() == [[101, 103, 110]] and got.ev_change.tolist() == [[0, 1, 1]] and got.ev_pos.tolist() == [[0, 1, 2]]
    mine = host_engine.wnm_eval(off, act, ts, [0, 1, 2], cons, 3, 1, 3)
    assert abi.AlmostMatchResult.same_as(mine, got)[0]
    # the true pattern holds nowhere near: without the constraints the trace matches as it is (total change 0, the LAST such match)
    free = oracle.why_not_match(off, act, ts, [0, 1, 2], [], 3, 1, 3)
    assert free.total_change.tolist() == [0] and free.ev_value.tolist() == [[101, 104, 109]]


def random_case(rng, n_traces=12):
    """Small traces over 4 activities with close, sometimes equal, sometimes unsorted timestamps; a pattern of 1 - 6 events
    (activities may repeat); constraints between any earlier and later event, also between two events of the same activity."""
    m = int(rng.integers(1, 7))
    pattern = rng.integers(0, 3, size=m).astype(np.int32)
    cons = []
    for _ in range(int(rng.integers(0, 4))):
        if m < 2:
            break
        b = int(rng.integers(1, m))
        a = int(rng.integers(0, b))
        kind = int(rng.integers(0, 2))
        cons.append((a, b, kind, int(rng.integers(0, 2)), int(rng.integers(0, 6 if kind == TIME else 9))))
    lens = rng.integers(0, 7, size=n_traces)
    off = np.zeros(n_traces + 1, dtype=np.int64)
    np.cumsum(lens, out=off[1:])
    act = rng.integers(0, 4, size=int(off[-1])).astype(np.int32)
    ts = np.zeros(int(off[-1]), dtype=np.int64)
    for t in range(n_traces):
        base = int(rng.integers(0, 4)) * 1000           # near 0: Math.max(original - u, 0) clamps
        steps = rng.integers(0, 4, size=int(lens[t])) * 1000 + rng.integers(0, 1000, size=int(lens[t]))
        col = base + np.cumsum(steps)
        if rng.random() < 0.15:
            rng.shuffle(col)                             # the reference sorts the uncertain stream, not the trace
        ts[off[t]:off[t + 1]] = col
    u, step, k = int(rng.integers(0, 4)), int(rng.integers(1, 4)), int(rng.integers(0, 4))
    flags = abi.F_EVT_POS if rng.random() < 0.3 else 0
    return off, act, ts, pattern, cons, u, step, k, flags


@pytest.mark.parametrize("seed", range(6))
def test_closed_form_equals_the_engine_on_random_cases(seed):
    """csrc/wnm.cuh (one sweep per start event) against the literal skip-till-any-match engine with shared value vectors."""
    rng = np.random.default_rng(1000 + seed)
    done = 0
    for _ in range(250):
        off, act, ts, pattern, cons, u, step, k, flags = random_case(rng)
        want = oracle.why_not_match(off, act, ts, pattern, cons, u, step, k, flags=flags, run_limit=300_000)
        if want is None:
            continue
        got = host_engine.wnm_eval(off, act, ts, pattern, cons, u, step, k, flags=flags)
        ok, why = abi.AlmostMatchResult.same_as(got, want)
        assert ok, (why, pattern.tolist(), cons, u, step, k, flags, off.tolist(), act.tolist(), ts.tolist(),
                    getattr(got, why).tolist(), getattr(want, why).tolist())
        done += 1
    assert done > 200


# ------------------------------------------------------------------------------------------------------- GPU
@pytest.fixture(scope="module")
def ctx():
    from sequencedetectionqueryexecutor_b200 import api
    with api.Context(0) as c:
        yield c


@pytest.mark.gpu
def test_kernel_equals_the_engine_on_random_cases(ctx):
    rng = np.random.default_rng(77)
    done = 0
    for _ in range(120):
        off, act, ts, pattern, cons, u, step, k, flags = random_case(rng, n_traces=40)
        want = oracle.why_not_match(off, act, ts, pattern, cons, u, step, k, flags=flags, run_limit=300_000)
        if want is None:
            continue
        log = ctx.load_log(off, act, ts, 4)
        got = log.why_not_match(pattern, cons, u, step, k, flags=flags)
        ok, why = got.same_as(want)
        assert ok, (why, pattern.tolist(), cons, u, step, k, flags)
        cand = np.sort(rng.choice(40, size=13, replace=False)).astype(np.int64)
        sub = log.why_not_match(pattern, cons, u, step, k, cand=cand, flags=flags)
        ok, why = sub.same_as(oracle.why_not_match(off, act, ts, pattern, cons, u, step, k, cand=cand, flags=flags, run_limit=300_000))
        assert ok, (why, "candidates")
        assert log.why_not_match(pattern, cons, u, step, k, cand=np.zeros(0, dtype=np.int64), flags=flags).n_traces == 0
        log.close()
        done += 1
    assert done > 90


@pytest.mark.gpu
def test_reference_known_answer_and_plan_on_the_gpu(ctx):
    """The reference's test trace through the C-ABI, then the plan of QueryPlanWhyNotMatch.execute (:54-100): true
    occurrences first, the why-not-match search over the other traces."""
    from sequencedetectionqueryexecutor_b200 import sase
    off, act, ts = reference_trace()
    # trace 1 holds the pattern as it is; trace 2 cannot be repaired within u = 3
    off = np.array([0, 3, 6, 9], dtype=np.int64)
    act = np.array([0, 1, 2] * 3, dtype=np.int32)
    ts = np.array([101000, 104000, 109000, 200000, 201000, 209000, 300000, 320000, 321000], dtype=np.int64)
    log = ctx.load_log(off, act, ts, 3)
    acts = sase.ActivityDictionary(["A", "B", "C"])
    pattern = sase.ComplexPattern([sase.EventSymbol("A", 0), sase.EventSymbol("B", 1), sase.EventSymbol("C", 2)],
                                  [sase.TimeConstraint(0, 1, 2), sase.TimeConstraint(1, 2, 7, method="atleast")])
    occ, almost = sase.why_not_match_plan(pattern, log, acts, 3, 1, 3)
    assert [o.traceID for o in occ] == [1]
    assert len(almost) == 1 and almost[0].trace_id == 0 and almost[0].totalChange == 2
    assert [(e.event_type, e.timestamp, e.change) for e in almost[0].match] == [("A", 101, 0), ("B", 103, 1), ("C", 110, 1)]
    log.close()


@pytest.mark.gpu
def test_limits_and_rejected_shapes(ctx):
    """A stream beyond SIESTA_WNM_MAX_STREAM is listed, the other traces are answered; malformed requests are refused (the
    reference would loop forever on step = 0)."""
    from sequencedetectionqueryexecutor_b200._lib import SiestaError
    off = np.array([0, 3, 3 + 200], dtype=np.int64)
    act = np.concatenate([[0, 1, 2], np.tile([0, 1], 100)]).astype(np.int32)
    ts = np.concatenate([[101000, 104000, 109000], 1000000 + 1000 * np.arange(200)]).astype(np.int64)
    log = ctx.load_log(off, act, ts, 3)
    cons = [(0, 1, TIME, WITHIN, 2)]
    got = log.why_not_match([0, 1], cons, 3, 1, 3)          # trace 1: 200 events x 7 variants = 1400 > 1024
    assert got.unsupported_trace_idx.tolist() == [1] and got.trace_idx.tolist() == [0]
    want = oracle.why_not_match(off, act, ts, [0, 1], cons, 3, 1, 3, cand=[0])
    assert got.same_as(want)[0]
    for bad in (dict(p=[0, 1], c=[(1, 1, GAP, WITHIN, 3)], step=1, code=abi.E_INVALID),
                dict(p=[0, 1, 1], c=[(2, 1, GAP, WITHIN, 3)], step=1, code=abi.E_INVALID),
                dict(p=[0, 1], c=[], step=0, code=abi.E_INVALID),
                dict(p=[], c=[], step=1, code=abi.E_INVALID)):
        with pytest.raises(SiestaError) as e:
            log.why_not_match(bad["p"], bad["c"], 3, bad["step"], 3)
        assert e.value.code == bad["code"]
    log.close()


@pytest.mark.gpu
def test_sharded_log_one_process_all_gpus():
    """siesta_multi_why_not_match: candidates are cut at the shard borders, every shard answers its own, trace indices stay global."""
    import torch

    from sequencedetectionqueryexecutor_b200 import api
    rng = np.random.default_rng(5)
    n_gpu = torch.cuda.device_count()
    for ids in ([0, 0, 0], list(range(n_gpu)) if n_gpu > 1 else [0]):
        with api.Multi(ids) as m:
            for _ in range(6):
                off, act, ts, pattern, cons, u, step, k, flags = random_case(rng, n_traces=60)
                want = oracle.why_not_match(off, act, ts, pattern, cons, u, step, k, flags=flags, run_limit=300_000)
                if want is None or off[-1] == 0:
                    continue
                log = m.load_log(off, act, ts, 4)
                assert log.why_not_match(pattern, cons, u, step, k, flags=flags).same_as(want)[0]
                cand = np.sort(rng.choice(60, size=25, replace=False)).astype(np.int64)
                sub = oracle.why_not_match(off, act, ts, pattern, cons, u, step, k, cand=cand, flags=flags, run_limit=300_000)
                assert log.why_not_match(pattern, cons, u, step, k, cand=cand, flags=flags).same_as(sub)[0]
                log.close()
