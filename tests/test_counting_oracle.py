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
    # alternate: A=[0,2] B=[1,4] on t0 -> a B inside (0,2) and one after 2 -> 2 ; chain: only "A B" at 0,1 -> 1
    assert d.alt_response.tolist() == [[0, 2, 1], [1, 0, 1], [1, 1, 0]]
    assert d.alt_precedence.tolist() == d.alt_response.tolist()
    assert d.chain_response.tolist() == [[0, 1, 1], [1, 0, 0], [1, 1, 0]]
    assert d.chain_precedence.tolist() == d.chain_response.tolist()
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
    # the four literal restatements of the alternate / chain counters: response and precedence forms count the same
    # events, chain <= alternate <= simple
    assert np.array_equal(d.alt_response, d.alt_precedence) and np.array_equal(d.chain_response, d.chain_precedence)
    assert np.all(d.chain_response <= d.alt_response) and np.all(d.alt_response <= d.response)
    assert np.all(d.alt_precedence <= d.precedence)
    # chain = adjacent pairs of different activities
    adj = np.zeros((A, A), dtype=np.int64)
    for t in range(len(off) - 1):
        tr = act[off[t]:off[t + 1]]
        for x, y in zip(tr[:-1], tr[1:]):
            if x != y:
                adj[x, y] += 1
    assert np.array_equal(adj, d.chain_response)
    # histogram cap: occurrences beyond k_cap are reported, not dropped silently
    d2 = oracle.declare_counts(off, act, A, 2)
    assert d2.hist_overflow == d.hist[:, 3:].sum()


def test_pair_stats_hand_example():
    """Stated pairing policy (parity unpinned: the producer of count.parquet is not in the reference repository):
    non-overlapping skip-till-next-match pairs."""
    # t0 = A A B A B B (ts 0,1,3,6,10,15 s) ; t1 = B A ; t2 = A A A A   (A=0, B=1)
    off = np.array([0, 6, 8, 12])
    act = np.array([0, 0, 1, 0, 1, 1, 1, 0, 0, 0, 0, 0], dtype=np.int32)
    ts = np.array([0, 1, 3, 6, 10, 15, 0, 5, 0, 2, 5, 9], dtype=np.int64) * 1000
    ab, aa, ba = oracle.pair_stats(off, act, ts, [(0, 1), (0, 0), (1, 0)])
    # (A,B) on t0: (0 -> 3) and (6 -> 10): durations 3 s, 4 s
    assert ab == {"count": 2, "sum": 7000, "min": 3000, "max": 4000, "sum_squares": 3000 ** 2 + 4000 ** 2}
    # (A,A): t0 pairs (0,1) then the third A waits; t2 pairs (0,2) and (5,9)
    assert aa == {"count": 3, "sum": 1000 + 2000 + 4000, "min": 1000, "max": 4000, "sum_squares": 1000 ** 2 + 2000 ** 2 + 4000 ** 2}
    # (B,A): t0 (3 -> 6), then B at 10 waits for an A that never comes; t1 (0 -> 5)
    assert ba == {"count": 2, "sum": 8000, "min": 3000, "max": 5000, "sum_squares": 3000 ** 2 + 5000 ** 2}
