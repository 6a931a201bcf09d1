"""Writes the committed golden fixtures of this directory.  Run from the repo root: `python tests/golden/make_fixtures.py`.

* reference_kats.json - the reference's own engine known-answer tests (stream A B A C D A B E; match counts asserted by
  src/test/java/.../SaseConnection/EvaluateNewQueries.java and EvaluateComplexQueries.java, file:line in `where`),
  serialised from tests/kat.py.  These are the vectors that pin the oracle.
* fixtures.npz - outputs of the pinned CPU oracle on small seeded logs for every entry point of the hot path
  (detection in the shapes of BASELINE configs[0], [1], [4] and a run-list-engine pattern, declare counts, pair
  statistics, posting lists + intersection, /explore, why-not-match).  tests/test_golden.py checks that the oracle still reproduces
  them (CPU) and that the CUDA path reproduces them through the C-ABI (-m gpu).  The reference itself cannot run here
  (no JVM), so these are oracle outputs, not outputs of the Java code.
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402
from sequencedetectionqueryexecutor_b200 import _abi as abi  # noqa: E402
from tests import gen, kat  # noqa: E402

N_, P_, S_, X_, O_ = abi.STATE_NORMAL, abi.STATE_KLEENE_PLUS, abi.STATE_KLEENE_STAR, abi.STATE_NEGATIVE, abi.STATE_OR

# name -> (log arguments of tests/gen.make_log, n_activities, states, flags)
DETECT_CASES = {
    "cfg0_ab": (dict(n_traces=400, min_len=30, max_len=50, n_act=20, seed=0x51E57A01), 20,
                [dict(kind=N_, types=[0]), dict(kind=N_, types=[1])], 0),
    "cfg1_kleene_within": (dict(n_traces=300, min_len=100, max_len=100, n_act=20, seed=0x51E57A02, max_gap_s=120), 20,
                           [dict(kind=P_, types=[0]), dict(kind=S_, types=[1], preds=[(abi.ATTR_TIMESTAMP, abi.OP_LE, 0, 600)])], 0),
    "cfg4_gap6": (dict(n_traces=400, min_len=50, max_len=50, n_act=20, seed=0x51E57A05), 20,
                  [dict(kind=N_, types=[0]), dict(kind=O_, types=[1, 2], preds=[(abi.ATTR_POSITION, abi.OP_LE, 0, 10)]),
                   dict(kind=X_, types=[3]), dict(kind=N_, types=[4]),
                   dict(kind=N_, types=[5], preds=[(abi.ATTR_POSITION, abi.OP_GE, 3, 2)])], 0),
    "cfg4_gap6_return_all_pos": (dict(n_traces=300, min_len=40, max_len=70, n_act=12, seed=77), 12,
                                 [dict(kind=N_, types=[0]), dict(kind=O_, types=[1, 2], preds=[(abi.ATTR_POSITION, abi.OP_LE, 0, 10)]),
                                  dict(kind=X_, types=[3]), dict(kind=N_, types=[4])], abi.F_RETURN_ALL | abi.F_EVT_POS),
    "engine_a_bplus_c": (dict(n_traces=300, min_len=5, max_len=30, n_act=6, seed=78, jitter_ms=True), 6,
                         [dict(kind=N_, types=[0]), dict(kind=P_, types=[1]), dict(kind=N_, types=[2])], abi.F_RETURN_ALL),
}
COUNT_LOG = dict(n_traces=500, min_len=30, max_len=70, n_act=20, seed=0x51E57A03)
PAIRS = [(0, 1), (1, 2), (2, 2), (5, 0)]
EXPLORE = dict(log=dict(n_traces=600, min_len=10, max_len=50, n_act=12, seed=81, jitter_ms=True), pattern=[0, 1])
# why-not-match: (pattern, constraints (pos_a, pos_b, kind, method, value), uncertainty, step, k, flags)
WNM = dict(log=dict(n_traces=400, min_len=5, max_len=25, n_act=8, seed=83, max_gap_s=6, jitter_ms=True),
           cases={"time3": ([0, 1, 2], [(0, 1, abi.WNM_TIME, abi.WNM_WITHIN, 4), (1, 2, abi.WNM_TIME, abi.WNM_ATLEAST, 6)], 3, 1, 3, 0),
                  "gap_far": ([0, 1, 0, 3], [(0, 2, abi.WNM_GAP, abi.WNM_WITHIN, 30), (1, 3, abi.WNM_TIME, abi.WNM_ATLEAST, 9)], 4, 2, 2, 0),
                  "positions": ([2, 1], [(0, 1, abi.WNM_TIME, abi.WNM_ATLEAST, 4)], 2, 1, 1, abi.F_EVT_POS)})
WNM_KEYS = ("trace_idx", "total_change", "ev_pos", "ev_value", "ev_change", "ev_stream_pos")

MATCH_KEYS = ("trace_idx", "occ_off", "ev_off", "ev_pos", "ev_rank", "ev_act", "ev_ts_ms", "err_trace_idx")


def explore_by_detection(off, act, ts, pattern, n_act):
    comp, dur = [], []
    for c in range(n_act):
        nfa = abi.make_nfa([dict(kind=N_, types=[x]) for x in pattern + [c]])
        w = oracle.detect(off, act, ts, nfa, flags=abi.F_RETURN_ALL)
        comp.append(w.n_occurrences)
        dur.append(sum(int(w.ev_ts_ms[w.ev_off[o + 1] - 1] - w.ev_ts_ms[w.ev_off[o]]) for o in range(w.n_occurrences)))
    return np.array(comp, dtype=np.int64), np.array(dur, dtype=np.int64)


def compute():
    out = {}
    for name, (lg, n_act, states, flags) in DETECT_CASES.items():
        off, act, ts = gen.make_log(**lg)
        r = oracle.detect(off, act, ts, abi.make_nfa(states), flags=flags)
        for k in MATCH_KEYS:
            out[f"detect/{name}/{k}"] = np.asarray(getattr(r, k))
        out[f"detect/{name}/n_matches_emitted"] = np.array([r.n_matches_emitted], dtype=np.int64)
    off, act, ts = gen.make_log(**COUNT_LOG)
    out["declare/packed"] = np.asarray(oracle.declare_counts(off, act, COUNT_LOG["n_act"], 40).packed)
    st = oracle.pair_stats(off, act, ts, PAIRS)
    out["stats/count_sum_min_max"] = np.array([[s["count"], s["sum"], s["min"], s["max"]] for s in st], dtype=np.int64)
    out["stats/sum_squares_str"] = np.array([str(s["sum_squares"]) for s in st])
    lists = [oracle.posting_list(off, act, a, b) for a, b in PAIRS[:2] + [(0, 2)]]
    for i, l in enumerate(lists):
        out[f"index/list{i}"] = np.asarray(l, dtype=np.int64)
    out["index/intersection"] = np.asarray(oracle.intersect(lists), dtype=np.int64)
    off, act, ts = gen.make_log(**EXPLORE["log"])
    comp, dur = explore_by_detection(off, act, ts, EXPLORE["pattern"], EXPLORE["log"]["n_act"])
    out["explore/completions"], out["explore/sum_duration_ms"] = comp, dur
    off, act, ts = gen.make_log(**WNM["log"])
    for name, (pattern, cons, u, step, k, flags) in WNM["cases"].items():
        r = oracle.why_not_match(off, act, ts, pattern, cons, u, step, k, flags=flags, run_limit=3_000_000)
        assert r is not None and r.n_traces > 0, name
        for key in WNM_KEYS:
            out[f"wnm/{name}/{key}"] = np.asarray(getattr(r, key))
    return out


def main():
    kats = [dict(name=k["name"], states=k["states"], expected=k["expected"], where=k["where"], matches=k["matches"],
                 head=k.get("head")) for k in kat.KATS]
    with open(os.path.join(HERE, "reference_kats.json"), "w") as f:
        json.dump({"stream_types": kat.STREAM_TYPES, "source": "src/test/java/com/datalab/siesta/queryprocessor/SaseConnection/"
                   "EvaluateNewQueries.java (N), EvaluateComplexQueries.java (C)", "kats": kats}, f, indent=1)
    np.savez_compressed(os.path.join(HERE, "fixtures.npz"), **compute())


if __name__ == "__main__":
    main()
