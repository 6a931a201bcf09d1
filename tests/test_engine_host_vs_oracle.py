"""The device engine (detect_engine.cuh) compiled for the host must agree with the oracle bit for bit:
matching traces, reference-throw traces, selected occurrences, every output column."""
import numpy as np
import pytest

import oracle
from sequencedetectionqueryexecutor_b200 import _abi as abi
from tests import gen, host_engine
from tests.kat import KATS, STREAM_TYPES


@pytest.mark.parametrize("kat", KATS, ids=[k["name"] for k in KATS])
def test_kat_through_device_engine(kat):
    nfa = abi.make_nfa(kat["states"])
    off = np.array([0, len(STREAM_TYPES)], dtype=np.int64)
    act = np.array(STREAM_TYPES, dtype=np.int32)
    ts = np.arange(len(STREAM_TYPES), dtype=np.int64) * 1000
    rc, res, _ = host_engine.detect(off, act, ts, 5, nfa, flags=abi.F_EVT_POS | abi.F_COUNT_MATCHES)
    assert rc == 0
    assert res.n_matches_emitted == kat["expected"]
    if kat["matches"]:
        best = max(kat["matches"], key=len)  # max() returns the first maximal element
        assert res.as_dict() == {0: [best]}
    else:
        assert res.n_traces == 0


def _compare(off, act, ts, n_act, states, flags):
    nfa = abi.make_nfa(states)
    rc, got, n_wide = host_engine.detect(off, act, ts, n_act, nfa, flags=flags)
    if rc == abi.E_UNSUPPORTED:
        return "unsupported", 0
    assert rc == 0, rc
    want = oracle.detect(off, act, ts, nfa, flags=flags)
    ok, why = got.same_as(want)
    assert ok, f"mismatch in {why}: states={states} flags={flags}"
    return ("err" if want.n_ref_errors else "ok"), n_wide


@pytest.mark.parametrize("seed", range(6))
def test_random_nfas_all_kinds(seed):
    rng = np.random.default_rng(1000 + seed)
    stats = {"ok": 0, "err": 0, "unsupported": 0}
    for it in range(150):
        n_act = int(rng.integers(3, 7))
        off, act, ts = gen.make_log(40, 0, 24, n_act, seed=int(rng.integers(1 << 30)), max_gap_s=300,
                                    jitter_ms=bool(rng.integers(0, 2)))
        states = gen.random_nfa(rng, n_act)
        flags = 0
        if rng.random() < 0.5:
            flags |= abi.F_EVT_POS
        if rng.random() < 0.5:
            flags |= abi.F_RETURN_ALL
        if rng.random() < 0.15:
            flags |= abi.F_ONLY_APPEARANCES
        if rng.random() < 0.25:
            flags |= abi.F_COUNT_MATCHES
        if rng.random() < 0.15:
            flags |= abi.F_LITERAL_RUNS
        r, _ = _compare(off, act, ts, n_act, states, flags)
        stats[r] += 1
    assert stats["ok"] > 50


@pytest.mark.parametrize("flags", [0, abi.F_RETURN_ALL, abi.F_EVT_POS, abi.F_EVT_POS | abi.F_RETURN_ALL])
def test_baseline_config_shapes(flags):
    N_, P_, S_, X_, O_ = abi.STATE_NORMAL, abi.STATE_KLEENE_PLUS, abi.STATE_KLEENE_STAR, abi.STATE_NEGATIVE, abi.STATE_OR
    # config 1: A_ B_
    off, act, ts = gen.make_log(300, 30, 50, 20, seed=0x51E57A01)
    assert _compare(off, act, ts, 20, [dict(kind=N_, types=[0]), dict(kind=N_, types=[1])], flags)[0] == "ok"
    # config 2: a+ b* within 10 minutes (0,1)
    off, act, ts = gen.make_log(300, 100, 100, 20, seed=0x51E57A02, max_gap_s=120)
    st = [dict(kind=P_, types=[0]), dict(kind=S_, types=[1], preds=[(abi.ATTR_TIMESTAMP, abi.OP_LE, 0, 600)])]
    assert _compare(off, act, ts, 20, st, flags)[0] == "ok"
    # config 5: a, (b|c), !d, e, f ; gap within 10 (0,1), gap atleast 2 (3,4)
    off, act, ts = gen.make_log(300, 50, 50, 20, seed=0x51E57A05)
    st = [dict(kind=N_, types=[0]), dict(kind=O_, types=[1, 2], preds=[(abi.ATTR_POSITION, abi.OP_LE, 0, 10)]),
          dict(kind=X_, types=[3]), dict(kind=N_, types=[4]),
          dict(kind=N_, types=[5], preds=[(abi.ATTR_POSITION, abi.OP_GE, 3, 2)])]
    assert _compare(off, act, ts, 20, st, flags)[0] == "ok"


def test_wide_engine_path_is_exercised(monkeypatch):
    """Few activity types -> many relevant events and O(n^2) Kleene runs: must fall over to the wide configuration."""
    monkeypatch.setenv("SIESTA_HARNESS_NO_LONG", "1")   # (the harness otherwise takes class NK to the long-trace evaluator from 25 events on)
    N_, P_, S_ = abi.STATE_NORMAL, abi.STATE_KLEENE_PLUS, abi.STATE_KLEENE_STAR
    off, act, ts = gen.make_log(60, 20, 45, 3, seed=77)
    r, n_wide = _compare(off, act, ts, 3, [dict(kind=P_, types=[0]), dict(kind=S_, types=[1])], 0)
    assert r == "ok" and n_wide > 0
    r, n_wide = _compare(off, act, ts, 3, [dict(kind=N_, types=[0]), dict(kind=N_, types=[1]), dict(kind=N_, types=[2])],
                         abi.F_RETURN_ALL)
    assert r == "ok" and n_wide > 0
