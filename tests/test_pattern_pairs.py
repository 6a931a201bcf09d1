"""Pair extraction (row P3): siesta_pattern_extract_pairs against the reference's own PatternTest
(src/test/java/com/datalab/siesta/queryprocessor/model/Patterns/PatternTest.java) and the reading of
ComplexPattern.extractPairsForPatternDetection (:75-128) / SIESTAPattern.extractPairsForPatternDetection (:34-69).
Host only: no GPU needed."""
import pytest

from sequencedetectionqueryexecutor_b200 import _abi as abi
from sequencedetectionqueryexecutor_b200 import sase
from sequencedetectionqueryexecutor_b200._lib import SiestaError

ACTS = sase.ActivityDictionary(["A", "B", "C", "D", "E"])
A, B, Cc, D, E = range(5)


def pat(*syms, constraints=()):
    return sase.ComplexPattern([sase.EventSymbol(n, p, s) for n, p, s in syms], list(constraints))


def test_extract_pairs_simple():          # PatternTest.extractPairsSimple :17-31
    x = pat(("A", 0, ""), ("B", 1, ""), ("C", 2, "")).extractPairsForPatternDetection(ACTS, False)
    assert len(x) == 1
    assert len(x[0].allPairs) == 3
    assert x[0].truePairs == [(A, B), (A, Cc), (B, Cc)]


def test_extract_pairs_or():              # PatternTest.extractPairsOr :32-50
    x = pat(("A", 0, ""), ("B", 1, ""), ("C", 1, ""), ("D", 2, "")).extractPairsForPatternDetection(ACTS, False)
    assert len(x) == 2
    assert [len(e.truePairs) for e in x] == [3, 3]
    assert {tuple(e.truePairs) for e in x} == {((A, B), (A, D), (B, D)), ((A, Cc), (A, D), (Cc, D))}


def test_extract_pairs_kleene():          # PatternTest.extractPairsKleene :52-69
    x = pat(("A", 0, ""), ("B", 1, "+"), ("C", 2, ""), ("D", 3, "*")).extractPairsForPatternDetection(ACTS, False)
    assert len(x) == 1
    assert x[0].truePairs == [(A, B), (A, Cc), (B, Cc)]
    # all pairs: true pairs + (B,B) for "+", (A,D) (firstNonEmpty, D) and (D,D) for "*"
    assert set(x[0].allPairs) == {(A, B), (A, Cc), (B, Cc), (B, B), (A, D), (D, D)}


def test_constraints_and_from_till_add_self_pairs():
    p = pat(("A", 0, "_"), ("B", 1, "_"), ("C", 2, "_"), constraints=[sase.GapConstraint(0, 1, 3)])
    x = p.extractPairsForPatternDetection(ACTS, False)[0]
    # positions 0 and 1 appear in a constraint; the self pair is added for i < size - 1 only (SIESTAPattern.java:42-44)
    assert set(x.allPairs) == {(A, B), (A, Cc), (B, Cc), (A, A), (B, B)}
    x = p.extractPairsForPatternDetection(ACTS, True)[0]
    assert set(x.allPairs) == {(A, B), (A, Cc), (B, Cc), (A, A), (B, B), (Cc, Cc)}


def test_or_symbol_and_negation():
    # BASELINE configs[4] shape: a, (b|c), !d, e  -> 2 expansions x C(3,2) true pairs; "!" fetches (a,d) and (d,d)
    p = pat(("A", 0, "_"), ("B", 1, "||"), ("C", 1, "_"), ("D", 2, "!"), ("E", 3, "_"))
    x = p.extractPairsForPatternDetection(ACTS, False)
    assert len(x) == 2 and all(len(e.truePairs) == 3 for e in x)
    assert all({(A, D), (D, D)} <= set(e.allPairs) for e in x)


def test_leading_star_quirks():
    # leading "*" followed by "_": pair (star, next) (ComplexPattern.java:104-118)
    x = pat(("A", 0, "*"), ("B", 1, "_")).extractPairsForPatternDetection(ACTS, False)[0]
    assert x.truePairs == [] and set(x.allPairs) == {(A, B), (A, A)}
    # leading "*" followed by another "*": the reference's while loop never advances -> reported, not imitated
    with pytest.raises(SiestaError) as e:
        pat(("A", 0, "*"), ("B", 1, "*"), ("C", 2, "_")).extractPairsForPatternDetection(ACTS, False)
    assert e.value.code == abi.E_REFERENCE_THROWS
