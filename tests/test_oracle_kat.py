"""Pins the CPU oracle to the reference's own engine known-answer tests."""
import pytest

import oracle
from sequencedetectionqueryexecutor_b200 import _abi as abi
from tests.kat import KATS, STREAM_TYPES

IDS = list(range(len(STREAM_TYPES)))


@pytest.mark.parametrize("kat", KATS, ids=[k["name"] for k in KATS])
def test_kat_counts_and_lists(kat):
    nfa = abi.make_nfa(kat["states"])
    status, matches = oracle.run_stream(nfa, STREAM_TYPES, IDS, IDS, flags=0)
    assert status == 0
    assert len(matches) == kat["expected"], f"reference asserts {kat['expected']} at {kat['where']}"
    assert matches == kat["matches"]


@pytest.mark.parametrize("kat", [k for k in KATS if "head" in k], ids=lambda k: k["name"])
def test_head_mode_disagrees_with_reference_tests(kat):
    """Engine.createNewRun's trailing block (Engine.java:983-996) doubles these counts at HEAD."""
    nfa = abi.make_nfa(kat["states"])
    status, matches = oracle.run_stream(nfa, STREAM_TYPES, IDS, IDS, flags=abi.F_MODE_HEAD)
    assert status == 0
    assert len(matches) == kat["head"]


@pytest.mark.parametrize("kat", [k for k in KATS if "head" not in k and "B*" not in k["name"]],
                         ids=lambda k: k["name"])
def test_head_mode_identical_when_state1_is_not_kleene_star(kat):
    nfa = abi.make_nfa(kat["states"])
    s0, m0 = oracle.run_stream(nfa, STREAM_TYPES, IDS, IDS, flags=0)
    s1, m1 = oracle.run_stream(nfa, STREAM_TYPES, IDS, IDS, flags=abi.F_MODE_HEAD)
    assert (s0, m0) == (s1, m1)
