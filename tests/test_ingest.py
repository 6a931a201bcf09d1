"""SIESTA's bucket tables (seq.parquet, index.parquet; schema of S3Connector.java:226-314) ->
the CSR log and the posting lists of the GPU path.  Round trips through real parquet files; the posting lists read from
index.parquet must equal the oracle's SeqTable-view lists of the same log."""
import numpy as np
import pyarrow as pa
import pyarrow.parquet as pq
import pytest

import oracle
from sequencedetectionqueryexecutor_b200 import ingest
from tests import gen


def _bucket(tmp_path, n_traces=300, n_act=7, seed=5, shuffle=True):
    off, act, ts = gen.make_log(n_traces, 0, 25, n_act, seed=seed, jitter_ms=True)
    names = [f"Act{i:02d}" for i in range(n_act)]
    tids = [f"case-{i * 7919 % 100003}" for i in range(n_traces)]
    tb = ingest.seq_table_from_csr(off, act, ts, tids, names)
    if shuffle:   # Spark gives no row order: the reader must restore it from `position`
        perm = np.random.default_rng(seed).permutation(tb.num_rows)
        tb = tb.take(pa.array(perm))
    pq.write_table(tb, tmp_path / "seq.parquet")
    return off, act, ts, names, tids


def test_seq_table_round_trip(tmp_path):
    off, act, ts, names, tids = _bucket(tmp_path)
    log = ingest.read_seq_table(str(tmp_path / "seq.parquet"))
    # traces are numbered by first appearance in the (shuffled) table: compare through the trace ids
    assert sorted(log.trace_ids) == sorted(t for t, n in zip(tids, np.diff(off)) if n > 0)
    by_id = {t: i for i, t in enumerate(tids)}
    for k, t in enumerate(log.trace_ids):
        i = by_id[t]
        a = log.act[log.trace_off[k]:log.trace_off[k + 1]]
        assert [log.activities.names[x] for x in a] == [names[x] for x in act[off[i]:off[i + 1]]]
        assert np.array_equal(log.ts_ms[log.trace_off[k]:log.trace_off[k + 1]], ts[off[i]:off[i + 1]])


def test_activity_names_fold_case_and_timestamps_parse(tmp_path):
    tb = pa.table({"trace_id": ["t1", "t1", "t2", "t1"], "event_type": ["Pay", "PAY", "ship", "Ship"],
                   "timestamp": ["2020-01-01 00:00:00", "2020-01-01 00:00:01.5", "2020-03-01 12:30:45.123456789", "1970-01-01 00:00:00.001"],
                   "position": pa.array([0, 1, 0, 2], type=pa.int32())})
    log = ingest.read_seq_table(tb)
    assert log.trace_ids == ["t1", "t2"] and len(log.activities) == 2   # Pay == PAY, ship == Ship (equalsIgnoreCase)
    assert log.act.tolist() == [0, 0, 1, 1] and log.trace_off.tolist() == [0, 3, 4]
    assert log.ts_ms.tolist() == [1577836800000, 1577836801500, 1, 1583065845123]
    assert ingest.read_seq_table(tb, tz_offset_ms=3_600_000).ts_ms[0] == 1577836800000 - 3_600_000
    # arrow timestamps (the timestamps mode of index.parquet) go through the same function
    col = pa.array([1577836800000, 1], type=pa.timestamp("ms"))
    assert ingest.timestamps_to_ms(col).tolist() == [1577836800000, 1]
    empty = ingest.read_seq_table(tb.slice(0, 0))
    assert empty.n_traces == 0 and len(empty.act) == 0


def test_index_table_gives_the_oracles_posting_lists(tmp_path):
    off, act, ts, names, tids = _bucket(tmp_path, shuffle=False)
    log = ingest.read_seq_table(str(tmp_path / "seq.parquet"))
    # without a shuffle and with every trace non-empty or skipped, dense indices follow the order of `tids`
    nonempty = [i for i in range(len(tids)) if off[i + 1] > off[i]]
    assert log.trace_ids == [tids[i] for i in nonempty]
    pairs = [(0, 1), (1, 0), (2, 2), (3, 6), (6, 6)]
    tb = ingest.index_table_from_csr(off, act, tids, names, pairs)
    tb = pa.concat_tables([tb, tb.slice(0, 50)])            # duplicate rows (several occurrence pairs per trace)
    tb = tb.take(pa.array(np.random.default_rng(1).permutation(tb.num_rows)))
    pq.write_table(tb, tmp_path / "index.parquet")
    got_pairs, lists = ingest.read_index_table(str(tmp_path / "index.parquet"), log)
    # the ingested log numbers the activities by first appearance: compare through the names
    orig = {n: i for i, n in enumerate(names)}
    got_pairs = [(orig[log.activities.names[a]], orig[log.activities.names[b]]) for a, b in got_pairs]
    assert sorted(got_pairs) == sorted(pairs)
    remap = {o: k for k, o in enumerate(nonempty)}
    for (a, b), l in zip(got_pairs, lists):
        want = np.array([remap[int(t)] for t in oracle.posting_list(off, act, a, b)], dtype=np.int64)
        assert np.array_equal(l, want), (a, b)
    # restricted read (the where-clause of getAllEventPairs) and names the log does not know
    only, l2 = ingest.read_index_table(str(tmp_path / "index.parquet"), log, pairs=[("act00", "ACT01"), ("nope", "Act01")])
    assert [(orig[log.activities.names[a]], orig[log.activities.names[b]]) for a, b in only] == [(0, 1)]
    assert np.array_equal(l2[0], lists[got_pairs.index((0, 1))])
    # the intersection of the lists is what getCommonIds returns
    common = oracle.intersect([lists[got_pairs.index((0, 1))], lists[got_pairs.index((1, 0))]])
    both = [k for k, i in enumerate(nonempty)
            if any(x == 0 for x in act[off[i]:off[i + 1]]) and any(x == 1 for x in act[off[i]:off[i + 1]])]
    assert set(common.tolist()) <= set(both)


def test_stats_is_a_lookup_in_the_count_table(tmp_path):
    """Row S: QueryPlanStats answers from count.parquet.  The records written from the oracle's pair statistics come back
    exactly, in the order of the pattern's consecutive pairs, missing pairs skipped, first record per pair."""
    off, act, ts = gen.make_log(500, 5, 40, 6, seed=21, jitter_ms=True)
    names = [f"act{i:02d}" for i in range(6)]
    pairs_id = [(0, 1), (1, 2), (2, 2), (4, 0)]
    st = oracle.pair_stats(off, act, ts, pairs_id)
    recs = [ingest.Count(names[a], names[b], s["sum"], s["count"], s["min"], s["max"], float(s["sum_squares"]))
            for (a, b), s in zip(pairs_id, st)]
    dup = ingest.Count("act00", "act01", 1, 1, 1, 1, 1.0)            # a second record of a pair is never returned
    pq.write_table(ingest.count_table_from_stats(recs + [dup]), tmp_path / "count.parquet")
    assert ingest.consecutive_pairs(["act00", "act01", "act02", "act02"]) == [("act00", "act01"), ("act01", "act02"), ("act02", "act02")]
    got = ingest.read_count_table(str(tmp_path / "count.parquet"), ingest.consecutive_pairs(["act00", "act01", "act02", "act02"]))
    assert got == recs[:3]
    got = ingest.read_count_table(str(tmp_path / "count.parquet"), [("act04", "act00"), ("act05", "act00"), ("ACT00", "act01"), ("act00", "act01")])
    assert got == [recs[3], recs[0]]                                   # unknown pair and wrong case skipped (String.equals)
    assert isinstance(got[0].count, int) and isinstance(got[0].sum_squares, float)


def test_time_range_filter_equals_per_trace_filtering():
    """Trace.filter(from, till): the vectorised CSR filter equals filtering every trace's list, and the oracle run on the
    filtered log equals the oracle run on per-trace filtered lists."""
    from sequencedetectionqueryexecutor_b200 import _abi as abi
    off, act, ts = gen.make_log(400, 0, 30, 5, seed=12, max_gap_s=3600)
    lo, hi = int(np.percentile(ts, 30)), int(np.percentile(ts, 80))
    for f, t in [(lo, hi), (None, hi), (lo, None), (None, None), (hi, lo)]:
        n_off, n_act, n_ts, kept = ingest.filter_time_range(off, act, ts, f, t)
        assert len(n_off) == len(off) and n_off[-1] == len(n_act) == len(kept)
        for i in range(len(off) - 1):
            seg = slice(off[i], off[i + 1])
            m = np.ones(off[i + 1] - off[i], dtype=bool)
            if f is not None:
                m &= ts[seg] >= f
            if t is not None:
                m &= ts[seg] <= t
            assert np.array_equal(n_act[n_off[i]:n_off[i + 1]], act[seg][m])
            assert np.array_equal(n_ts[n_off[i]:n_off[i + 1]], ts[seg][m])
    n_off, n_act, n_ts, _ = ingest.filter_time_range(off, act, ts, lo, hi)
    nfa = abi.make_nfa([dict(kind=abi.STATE_NORMAL, types=[0]), dict(kind=abi.STATE_NORMAL, types=[1])])
    got = oracle.detect(n_off, n_act, n_ts, nfa)
    assert got.n_traces > 0 and np.all(got.ev_ts_ms >= lo - 999) and np.all(got.ev_ts_ms <= hi)


@pytest.mark.gpu
def test_ingested_bucket_through_the_gpu_path(tmp_path):
    """seq.parquet + index.parquet -> load_log / load_index -> intersection -> detection on the candidates = the oracle."""
    from sequencedetectionqueryexecutor_b200 import _abi as abi, api
    off, act, ts, names, tids = _bucket(tmp_path, n_traces=2000, shuffle=True)
    log = ingest.read_seq_table(str(tmp_path / "seq.parquet"))
    pairs = [(0, 1), (0, 2), (1, 2)]
    pq.write_table(ingest.index_table_from_csr(off, act, tids, names, pairs), tmp_path / "index.parquet")
    got_pairs, lists = ingest.read_index_table(str(tmp_path / "index.parquet"), log)
    nfa = abi.make_nfa([dict(kind=abi.STATE_NORMAL, types=[x]) for x in (0, 1, 2)])                       # original ids (oracle)
    nfa_ing = abi.make_nfa([dict(kind=abi.STATE_NORMAL, types=[log.activities.id(names[x])]) for x in (0, 1, 2)])   # ingested ids
    with api.Context(0) as ctx:
        glog = ctx.load_log(log.trace_off, log.act, log.ts_ms, len(log.activities))
        idx = glog.load_index(got_pairs, lists)
        cand = idx.intersect()
        got = glog.detect(nfa_ing, cand=cand)
        idx.close()
        glog.close()
    want = oracle.detect(off, act, ts, nfa)
    assert sorted(log.trace_ids[int(t)] for t in got.trace_idx) == sorted(tids[int(t)] for t in want.trace_idx)
    assert got.n_events == want.n_events
