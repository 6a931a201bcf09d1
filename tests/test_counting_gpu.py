"""Parity of the CUDA counting kernels (declare counts K3, pair index + intersection K2) against the oracle."""
import numpy as np
import pytest

import oracle
from tests import gen

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    from sequencedetectionqueryexecutor_b200 import api
    c = api.Context(0)
    yield c
    c.close()


@pytest.mark.parametrize("n_act,n_traces,lens,zipf", [(3, 50, (0, 12), None), (20, 3000, (30, 70), None),
                                                      (20, 3000, (30, 70), 1.1), (100, 1500, (50, 50), None),
                                                      (33, 700, (0, 90), 1.1), (64, 300, (1, 300), None),
                                                      # K3 v2 (position masks, pair-owned counters): <= 32 activities and
                                                      # <= 64 / <= 128 events per trace; one event more falls back
                                                      (32, 2500, (0, 64), None), (24, 2000, (60, 64), 1.1), (23, 1500, (0, 128), None),
                                                      (32, 900, (100, 128), 1.1), (2, 3000, (0, 40), None), (1, 500, (0, 9), None),
                                                      (20, 800, (0, 129), None), (7, 4099, (33, 33), None),
                                                      # more than 104 activities: declare_any_kernel (lanes = the distinct
                                                      # activities of a trace, counts by global atomics), any trace length
                                                      (105, 1200, (0, 80), None), (600, 800, (20, 400), 1.1), (1500, 300, (0, 2500), 1.05)])
def test_declare_counts_match_oracle(ctx, n_act, n_traces, lens, zipf):
    off, act, ts = gen.make_log(n_traces, lens[0], lens[1], n_act, seed=100 + n_act, zipf=zipf)
    log = ctx.load_log(off, act, ts, n_act)
    got = log.declare_counts(k_cap=40)
    log.close()
    want = oracle.declare_counts(off, act, n_act, 40)
    for name in ("tot", "uniq", "first", "last", "hist", "co", "ordered", "response", "precedence", "alt_response",
                 "alt_precedence", "chain_response", "chain_precedence"):
        assert np.array_equal(getattr(got, name), getattr(want, name)), name
    assert got.hist_overflow == want.hist_overflow and got.n_nonempty == want.n_nonempty


def test_declare_counts_histogram_cap_and_limits(ctx):
    from sequencedetectionqueryexecutor_b200._lib import SiestaError
    off, act, ts = gen.make_log(500, 20, 60, 4, seed=8)
    log = ctx.load_log(off, act, ts, 4)
    got = log.declare_counts(k_cap=5)
    want = oracle.declare_counts(off, act, 4, 5)
    log.close()
    assert np.array_equal(got.packed, want.packed) and got.hist_overflow > 0
    log = ctx.load_log(*gen.make_log(10, 5, 5, 5000, seed=8), 5000)   # the eight A x A matrices of the result bound the alphabet (4096)
    with pytest.raises(SiestaError):
        log.declare_counts()
    log.close()


@pytest.mark.parametrize("n_act,lens", [(5, (0, 30)), (40, (0, 150)), (100, (10, 60))])
def test_declare_counts_any_alphabet_kernel_equals_the_others(ctx, n_act, lens, monkeypatch):
    """SIESTA_K3_ANY forces declare_any_kernel on alphabets the shared-memory kernels also take: all three agree with the oracle."""
    off, act, ts = gen.make_log(1500, lens[0], lens[1], n_act, seed=200 + n_act, zipf=1.1)
    want = oracle.declare_counts(off, act, n_act, 30)
    log = ctx.load_log(off, act, ts, n_act)
    a = log.declare_counts(k_cap=30)
    monkeypatch.setenv("SIESTA_K3_ANY", "1")
    b = log.declare_counts(k_cap=30)
    log.close()
    assert np.array_equal(a.packed, want.packed)
    assert np.array_equal(b.packed, want.packed)


@pytest.mark.parametrize("n_act,n_traces,lens", [(5, 2000, (0, 12)), (20, 20000, (30, 50)), (3, 777, (0, 4))])
def test_pair_index_and_intersection_match_oracle(ctx, n_act, n_traces, lens):
    off, act, ts = gen.make_log(n_traces, lens[0], lens[1], n_act, seed=300 + n_act)
    log = ctx.load_log(off, act, ts, n_act)
    pairs = [(0, 1), (1, 0), (0, 0), (1, 2), (0, 2), (2, 2), (0, n_act + 5)]  # the last names an unknown activity
    idx = log.build_index(pairs)
    lists = []
    for i, (a, b) in enumerate(pairs):
        got = idx.posting_list(i)
        want = oracle.posting_list(off, act, a, b)
        assert np.array_equal(got, want), (a, b)
        lists.append(want)
    # getCommonIds for A B C: all ordered pairs i<j must be present
    for sel in ([0, 3, 4], [0], [0, 1], [2, 5], [0, 1, 2, 3, 4, 5], [0, 6]):
        got = idx.intersect(sel)
        want = oracle.intersect([lists[i] for i in sel])
        assert np.array_equal(got, want), sel
    idx.close()
    # posting lists supplied by the caller (index.parquet route)
    idx2 = log.load_index(pairs[:3], lists[:3])
    assert np.array_equal(idx2.intersect(), oracle.intersect(lists[:3]))
    idx2.close()
    log.close()


def test_pruning_then_verification_pipeline(ctx):
    """P1 -> V: the intersection feeds siesta_detect's candidate list; result equals verifying every trace,
    because traces without all true pairs cannot match (QueryPlanPatternDetection.getMiddleResults :146-164)."""
    from sequencedetectionqueryexecutor_b200 import _abi as abi
    off, act, ts = gen.make_log(8000, 10, 30, 12, seed=41)
    log = ctx.load_log(off, act, ts, 12)
    nfa = abi.make_nfa([dict(kind=abi.STATE_NORMAL, types=[0]), dict(kind=abi.STATE_NORMAL, types=[1]),
                        dict(kind=abi.STATE_NORMAL, types=[2])])
    idx = log.build_index([(0, 1), (0, 2), (1, 2)])
    cand = idx.intersect()
    idx.close()
    pruned = log.detect(nfa, cand=cand)
    full = log.detect(nfa)
    log.close()
    assert len(cand) < 8000 and pruned.same_as(full)[0]
    want = oracle.detect(off, act, ts, nfa, cand=cand)
    assert pruned.same_as(want)[0]


def test_pair_stats_match_oracle(ctx):
    """Kernel K4 against the oracle's restatement of the stated pairing policy (both unpinned by the reference)."""
    off, act, ts = gen.make_log(4000, 0, 80, 9, seed=71, jitter_ms=True)
    pairs = [(a, b) for a in range(9) for b in range(9)][:32]
    log = ctx.load_log(off, act, ts, 9)
    got, ms = log.pair_stats(pairs)
    assert got == oracle.pair_stats(off, act, ts, pairs)
    every = [(a, b) for a in range(9) for b in range(9)]     # 81 pairs: served in passes of 32
    got_all, _ = log.pair_stats(every)
    assert got_all == oracle.pair_stats(off, act, ts, every)
    # a pair nobody holds, an activity outside the alphabet, durations of years (128-bit sum of squares)
    off2 = np.array([0, 4], dtype=np.int64)
    act2 = np.array([0, 1, 0, 1], dtype=np.int32)
    ts2 = np.array([0, 3_000_000_000_000, 3_000_000_000_001, 9_000_000_000_000], dtype=np.int64)
    log2 = ctx.load_log(off2, act2, ts2, 3)
    got2, _ = log2.pair_stats([(0, 1), (2, 0), (7, 1)])
    assert got2 == oracle.pair_stats(off2, act2, ts2, [(0, 1), (2, 0), (7, 1)])
    assert got2[0]["sum_squares"] > 2 ** 64 and got2[1]["count"] == 0
    log2.close()
    log.close()
