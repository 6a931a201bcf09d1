"""Secondary measurement: why-not-match (row f-4) on the configs[4] log shape: 20 activities, 50 events per trace, a
4-event simple pattern with two time constraints; the traces WITHOUT a true occurrence go to siesta_why_not_match
(u = 3 s, step = 1 s, k = 3: 7 uncertain events per event).  `python tools/bench_wnm.py [--traces N]` (one GPU, one JSON line).
Parity against the literal oracle (exponential: the reference's engine materialises every combination of uncertain events as
a run) on a small prefix whose traces the oracle finishes; the oracle's time on that prefix is the CPU figure."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import oracle  # noqa: E402
from sequencedetectionqueryexecutor_b200 import _abi as abi, api  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--traces", type=int, default=1_000_000)
    ap.add_argument("--sample", type=int, default=300)
    args = ap.parse_args()
    import torch
    dev = torch.device("cuda", 0)
    n_act = 20
    off, act, ts = bench.make_log_fast(args.traces, 50, 50, n_act, 0x51E57A07, 5)   # gaps of a few seconds: shifts of +-3 s matter
    T, E = len(off) - 1, len(act)
    d = [torch.from_numpy(x).to(dev) for x in (off, act, ts)]
    ctx = api.Context(0)
    log = ctx.wrap_log(*d, n_act, max_trace_len=50)
    pattern = [0, 1, 2, 3]
    cons = [(0, 1, abi.WNM_TIME, abi.WNM_WITHIN, 2), (1, 2, abi.WNM_TIME, abi.WNM_ATLEAST, 4), (2, 3, abi.WNM_TIME, abi.WNM_WITHIN, 3)]
    nfa = abi.make_nfa([dict(kind=abi.STATE_NORMAL, types=[0]),
                        dict(kind=abi.STATE_NORMAL, types=[1], preds=[(abi.ATTR_TIMESTAMP, abi.OP_LE, 0, 2)]),
                        dict(kind=abi.STATE_NORMAL, types=[2], preds=[(abi.ATTR_TIMESTAMP, abi.OP_GE, 1, 4)]),
                        dict(kind=abi.STATE_NORMAL, types=[3], preds=[(abi.ATTR_TIMESTAMP, abi.OP_LE, 2, 3)])])
    true = log.detect(nfa, flags=abi.F_NO_EVENT_COLUMNS)
    rest = np.setdiff1d(np.arange(T, dtype=np.int64), true.trace_idx)
    u, step, k = 3, 1, 3
    log.why_not_match(pattern, cons, u, step, k, cand=rest)   # warm
    best, wall = 1e9, 1e9
    for _ in range(3):
        t0 = time.perf_counter()
        got = log.why_not_match(pattern, cons, u, step, k, cand=rest)
        wall = min(wall, (time.perf_counter() - t0) * 1e3)
        best = min(best, got.kernel_ms)
    rel = int(np.isin(act, pattern).sum() * len(rest) / T)      # relevant events of the rest traces (estimate from the whole log)
    S = min(args.sample, len(rest))
    t0 = time.perf_counter()
    want = oracle.why_not_match(off, act, ts, pattern, cons, u, step, k, cand=rest[:S], run_limit=3_000_000)
    cpu = time.perf_counter() - t0
    sub = log.why_not_match(pattern, cons, u, step, k, cand=rest[:S])
    print(json.dumps({"kernel": "wnm_kernel (warp per trace, lanes = start events, one sweep per start)",
                      "workload": f"why-not-match, 20 activities x 50 events, pattern of 4 with 3 time constraints, u={u} step={step} k={k}",
                      "traces": T, "traces_without_occurrence": int(len(rest)), "almost_matches": int(got.n_traces),
                      "unsupported": int(len(got.unsupported_trace_idx)), "kernel_ms": best, "request_ms_wall": wall,
                      "traces_per_s": len(rest) / (best * 1e-3), "uncertain_events_per_s": rel * (2 * u // step + 1) / (best * 1e-3),
                      "cpu_oracle_traces_per_s": (S / cpu) if want is not None else None, "cpu_oracle_sample": S,
                      "parity_on_sample": bool(want is not None and sub.same_as(want)[0])}))


if __name__ == "__main__":
    main()
