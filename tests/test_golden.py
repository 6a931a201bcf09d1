"""Committed golden vectors (tests/golden/): the reference's own engine known-answer tests and outputs of the pinned
oracle on small seeded logs.  CPU: the oracle still reproduces them (guards the checker against drift).  GPU: the CUDA
path reproduces them through the C-ABI without consulting the live oracle."""
import json
import os

import numpy as np
import pytest

from sequencedetectionqueryexecutor_b200 import _abi as abi
from tests import gen, kat
from tests.golden import make_fixtures as mf

HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def fx():
    return np.load(os.path.join(HERE, "fixtures.npz"))


def test_reference_kats_file_matches_the_table_used_by_the_tests():
    doc = json.load(open(os.path.join(HERE, "reference_kats.json")))
    assert doc["stream_types"] == kat.STREAM_TYPES and len(doc["kats"]) == len(kat.KATS) == 21
    for a, b in zip(doc["kats"], kat.KATS):
        assert (a["name"], a["expected"], a["where"], a["matches"]) == (b["name"], b["expected"], b["where"], b["matches"])
        assert json.loads(json.dumps(b["states"])) == a["states"]


def test_reference_kats_through_the_oracle():
    """every golden vector of the reference's engine tests, replayed from the committed file"""
    import oracle
    doc = json.load(open(os.path.join(HERE, "reference_kats.json")))
    types = np.array(doc["stream_types"], dtype=np.int32)
    for k in doc["kats"]:
        states = [dict(kind=s["kind"], types=s["types"], preds=[tuple(p) for p in s["preds"]]) for s in k["states"]]
        status, matches = oracle.run_stream(abi.make_nfa(states), types, np.arange(len(types)), np.arange(len(types)))
        assert status == 0 and len(matches) == k["expected"], (k["name"], k["where"])
        assert [list(m) for m in matches] == k["matches"], k["name"]


def test_oracle_reproduces_the_fixtures(fx):
    want = mf.compute()
    assert sorted(want) == sorted(fx.files)
    for k in fx.files:
        assert np.array_equal(np.asarray(want[k]), fx[k]), k


@pytest.fixture(scope="module")
def ctx():
    from sequencedetectionqueryexecutor_b200 import api
    c = api.Context(0)
    yield c
    c.close()


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(mf.DETECT_CASES))
def test_gpu_detection_reproduces_the_fixtures(ctx, fx, name):
    lg, n_act, states, flags = mf.DETECT_CASES[name]
    off, act, ts = gen.make_log(**lg)
    log = ctx.load_log(off, act, ts, n_act)
    try:
        got = log.detect(abi.make_nfa(states), flags=flags)
        via_events = ctx.evaluate_events(off, act, ts, n_act, abi.make_nfa(states), flags=flags)
    finally:
        log.close()
    for r in (got, via_events):
        for k in mf.MATCH_KEYS:
            assert np.array_equal(np.asarray(getattr(r, k)), fx[f"detect/{name}/{k}"]), (name, k)
        if r.n_matches_emitted >= 0:
            assert r.n_matches_emitted == int(fx[f"detect/{name}/n_matches_emitted"][0])


@pytest.mark.gpu
def test_gpu_counting_paths_reproduce_the_fixtures(ctx, fx):
    off, act, ts = gen.make_log(**mf.COUNT_LOG)
    log = ctx.load_log(off, act, ts, mf.COUNT_LOG["n_act"])
    try:
        assert np.array_equal(log.declare_counts(k_cap=40).packed, fx["declare/packed"])
        st = log.pair_stats(mf.PAIRS)[0]
        assert np.array_equal(np.array([[s["count"], s["sum"], s["min"], s["max"]] for s in st], dtype=np.int64), fx["stats/count_sum_min_max"])
        assert [str(s["sum_squares"]) for s in st] == list(fx["stats/sum_squares_str"])
        idx = log.build_index(mf.PAIRS[:2] + [(0, 2)])
        for i in range(3):
            assert np.array_equal(idx.posting_list(i), fx[f"index/list{i}"])
        assert np.array_equal(idx.intersect(), fx["index/intersection"])
        idx.close()
    finally:
        log.close()
    off, act, ts = gen.make_log(**mf.EXPLORE["log"])
    log = ctx.load_log(off, act, ts, mf.EXPLORE["log"]["n_act"])
    try:
        comp, dur, _ = log.explore_accurate(mf.EXPLORE["pattern"], list(range(mf.EXPLORE["log"]["n_act"])))
    finally:
        log.close()
    assert np.array_equal(comp, fx["explore/completions"]) and np.array_equal(dur, fx["explore/sum_duration_ms"])


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(mf.WNM["cases"]))
def test_gpu_why_not_match_reproduces_the_fixtures(ctx, fx, name):
    off, act, ts = gen.make_log(**mf.WNM["log"])
    pattern, cons, u, step, k, flags = mf.WNM["cases"][name]
    log = ctx.load_log(off, act, ts, mf.WNM["log"]["n_act"])
    got = log.why_not_match(pattern, cons, u, step, k, flags=flags)
    log.close()
    for key in mf.WNM_KEYS:
        assert np.array_equal(np.asarray(getattr(got, key)), fx[f"wnm/{name}/{key}"]), (name, key)
