"""The counting oracle (declare / pair index / intersection) against hand-computed examples and identities.
The reference has no tests for its declare package: these pins are by code reading."""
import numpy as np

import oracle
from tests import gen


def test_hand_example():
    # t0 = A B A C B ; t1 = (empty) ; t2 = C C A          (A=0, B=1, C=2)
    off = np.array([0, 5, 5, 8])
    act = np.array([0, 1, 0, 2, 1, 2, 2, 0], dtype=np.int32)
    d = oracle.declare_counts(off, act, 3, 8)
    assert d.tot.tolist() == [3, 2, 3] and d.uniq.tolist() == [2, 1, 2]
    assert d.first.tolist() == [1, 0, 1] and d.last.tolist() == [1, 1, 0] and d.n_nonempty == 2
    assert d.hist[0, 1] == 1 and d.hist[0, 2] == 1 and d.hist[1, 2] == 1 and d.hist[2, 1] == 1 and d.hist[2, 2] == 1
    assert d.ordered.tolist() == [[1, 1, 1], [1, 1, 1], [1, 1, 1]]
    assert d.co.tolist() == [[1, 1, 2], [1, 1, 1], [2, 1, 1]]
    # countResponse(A,B) on t0: both A's have a later B -> 2 ; countPrecedence(A,C) on t0: the C has an earlier A -> 1
    assert d.response.tolist() == [[0, 2, 2], [1, 0, 1], [2, 1, 0]]
    assert d.precedence.tolist() == [[0, 2, 1], [1, 0, 1], [1, 1, 0]]
    assert oracle.posting_list(off, act, 0, 1).tolist() == [0]
    assert oracle.posting_list(off, act, 2, 0).tolist() == [2]
    assert oracle.posting_list(off, act, 2, 2).tolist() == [2]
    assert oracle.intersect([[0, 2, 5, 9], [2, 3, 9], [1, 2, 9, 11]]).tolist() == [2, 9]


def test_identities_on_random_log():
    off, act, _ = gen.make_log(400, 0, 30, 7, seed=3)
    A = 7
    d = oracle.declare_counts(off, act, A, 40)
    lens = np.diff(off)
    assert d.tot.sum() == lens.sum() and d.n_nonempty == (lens > 0).sum()
    assert d.first.sum() == d.n_nonempty and d.last.sum() == d.n_nonempty
    assert np.array_equal(d.hist.sum(axis=1), d.uniq) and d.hist_overflow == 0
    assert np.array_equal((d.hist * np.arange(41)).sum(axis=1), d.tot)
    assert np.array_equal(d.co, d.co.T)                                   # union of both orders is symmetric
    for a in range(A):
        for b in range(A):
            pl = oracle.posting_list(off, act, a, b)
            assert len(pl) == d.ordered[a, b]                              # ordered = posting-list length
            if a != b:
                both = len(np.union1d(pl, oracle.posting_list(off, act, b, a)))
                assert both == d.co[a, b]                                  # joinUnionTraces
                assert (d.response[a, b] > 0) == (d.ordered[a, b] > 0) == (d.precedence[a, b] > 0)
                assert d.response[a, b] <= d.tot[a] and d.precedence[a, b] <= d.tot[b]
    # histogram cap: occurrences beyond k_cap are reported, not dropped silently
    d2 = oracle.declare_counts(off, act, A, 2)
    assert d2.hist_overflow == d.hist[:, 3:].sum()
