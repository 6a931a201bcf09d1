"""Derived logs built on the device (csrc/logview.cu): the from / till window (Trace.filter, Trace.java:25-29) and the
group streams (SparkDatabaseRepository.querySingleTableGroups :307-336), each compared with a plain numpy restatement
of the Java code, and /detection on them compared with the oracle run on the restated log."""
import numpy as np
import pytest

import oracle
from sequencedetectionqueryexecutor_b200 import _abi as abi
from sequencedetectionqueryexecutor_b200 import ingest
from tests import gen

pytestmark = pytest.mark.gpu
N_, P_, S_, X_, O_ = abi.STATE_NORMAL, abi.STATE_KLEENE_PLUS, abi.STATE_KLEENE_STAR, abi.STATE_NEGATIVE, abi.STATE_OR


@pytest.mark.parametrize("window", [(None, None), (0.2, None), (None, 0.7), (0.3, 0.6), (0.9, 0.1)])
def test_from_till_on_the_device_equals_trace_filter(window):
    from sequencedetectionqueryexecutor_b200 import api
    off, act, ts = gen.make_log(3000, 0, 70, 8, seed=21, max_gap_s=400, jitter_ms=True)
    lo, hi = int(ts.min()), int(ts.max())
    frm = None if window[0] is None else lo + int((hi - lo) * window[0])
    til = None if window[1] is None else lo + int((hi - lo) * window[1])
    w_off, w_act, w_ts, kept = ingest.filter_time_range(off, act, ts, frm, til)   # the host restatement of Trace.filter
    nfa = abi.make_nfa([dict(kind=N_, types=[0]), dict(kind=O_, types=[1, 2], preds=[(abi.ATTR_TIMESTAMP, abi.OP_LE, 0, 900)]),
                        dict(kind=X_, types=[3]), dict(kind=N_, types=[4])])
    with api.Context(0) as ctx:
        log = ctx.load_log(off, act, ts, 8)
        view = log.filter_time(frm, til)
        assert (view.n_traces, view.n_events) == (len(off) - 1, len(w_act))
        assert np.array_equal(view.source_events(), kept)
        for flags in (0, abi.F_RETURN_ALL):
            got = view.detect(nfa, flags=flags)
            ok, why = got.same_as(oracle.detect(w_off, w_act, w_ts, nfa, flags=flags))
            assert ok, (window, flags, why)
        assert np.array_equal(view.declare_counts(20).packed, oracle.declare_counts(w_off, w_act, 8, 20).packed)
        view.close()
        log.close()


def _java_groups(off, act, ts, groups, types):
    """querySingleTableGroups restated: first group wins, groups must cover all types, events sorted by timestamp
    (ties: order of the trace in the group's list, then position)."""
    owner = {}
    for g, members in enumerate(groups):
        for t in members:
            owner.setdefault(t, g)
    g_off, g_act, g_ts, g_src, ids = [0], [], [], [], []
    for g, members in enumerate(groups):
        rows = []
        seen = []
        for order, t in enumerate(members):
            if owner[t] != g or t in seen:
                continue
            seen.append(t)
            for e in range(off[t], off[t + 1]):
                rows.append((int(ts[e]), order, e - off[t], e))
        if not rows or not set(types) <= {int(act[r[3]]) for r in rows}:
            continue
        rows.sort()
        ids.append(g + 1)
        g_act += [int(act[r[3]]) for r in rows]
        g_ts += [r[0] for r in rows]
        g_src += [r[3] for r in rows]
        g_off.append(len(g_act))
    return (np.array(g_off, dtype=np.int64), np.array(g_act, dtype=np.int32), np.array(g_ts, dtype=np.int64),
            np.array(g_src, dtype=np.int64), ids)


def test_group_streams_equal_the_java_restatement_and_detect_on_them():
    from sequencedetectionqueryexecutor_b200 import api
    off, act, ts = gen.make_log(400, 0, 30, 6, seed=8, max_gap_s=50)      # coarse gaps: plenty of equal timestamps across traces
    ts = (ts // 20000) * 20000
    ts = np.concatenate([np.sort(ts[off[t]:off[t + 1]]) for t in range(len(off) - 1)]) if len(ts) else ts
    rng = np.random.default_rng(5)
    groups = [list(rng.choice(400, size=int(rng.integers(0, 7)), replace=True)) for _ in range(120)]   # overlaps, repeats, empty groups
    groups = [[int(x) for x in g] for g in groups]
    nfa = abi.make_nfa([dict(kind=N_, types=[0]), dict(kind=P_, types=[1]), dict(kind=N_, types=[2])])
    types = [0, 1, 2]
    w_off, w_act, w_ts, w_src, ids = _java_groups(off, act, ts, groups, types)
    with api.Context(0) as ctx:
        log = ctx.load_log(off, act, ts, 6)
        glog, gids = log.group(groups, types)
        assert gids.tolist() == ids
        assert np.array_equal(glog.source_events(), w_src)
        for flags in (0, abi.F_RETURN_ALL, abi.F_COUNT_MATCHES):
            got = glog.detect(nfa, flags=flags)          # SaseConnector.evaluateGroups + clearOccurrences
            ok, why = got.same_as(oracle.detect(w_off, w_act, w_ts, nfa, flags=flags))
            assert ok, (flags, why)
        # the window and the groups compose: groups of the filtered log
        lo, hi = int(ts.min()), int(ts.max())
        frm, til = lo + (hi - lo) // 4, hi - (hi - lo) // 4
        f_off, f_act, f_ts, _ = ingest.filter_time_range(off, act, ts, frm, til)
        view = log.filter_time(frm, til)
        glog2, gids2 = view.group(groups, types)
        w2 = _java_groups(f_off, f_act, f_ts, groups, types)
        assert gids2.tolist() == w2[4]
        ok, why = glog2.detect(nfa, flags=0).same_as(oracle.detect(w2[0], w2[1], w2[2], nfa, flags=0))
        assert ok, why
        for x in (glog2, view, glog, log):
            x.close()


def test_group_and_window_edge_cases():
    from sequencedetectionqueryexecutor_b200 import api
    from sequencedetectionqueryexecutor_b200._lib import SiestaError
    off, act, ts = gen.make_log(50, 1, 10, 3, seed=2)
    with api.Context(0) as ctx:
        log = ctx.load_log(off, act, ts, 3)
        glog, gids = log.group([], [0])                       # no groups at all
        assert glog.n_traces == 0 and len(gids) == 0
        glog.close()
        glog, gids = log.group([[1, 2], [3]], [0, 1, 2, 7])   # a type the log never holds: every group is dropped
        assert glog.n_traces == 0 and len(gids) == 0
        glog.close()
        with pytest.raises(SiestaError):
            log.group([[0, 50]], [0])                         # trace index out of range
        empty = log.filter_time(int(ts.max()) + 1, None)      # nothing survives: all traces empty
        assert empty.n_traces == 50 and empty.n_events == 0
        assert empty.detect(abi.make_nfa([dict(kind=N_, types=[0])]), flags=0).n_traces == 0
        empty.close()
        log.close()
