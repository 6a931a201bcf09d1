"""Secondary measurements (not the headline bench): kernels K2 (pair index + intersection), K3 (declare counts) and
K4 (pair statistics) at the sizes of BASELINE configs[2] / configs[3], with the CPU oracle on a sample beside them.
    python tools/bench_counting.py [--traces N]        # one GPU; prints one JSON line per kernel
Byte counts are the algorithmic bytes of DESIGN.md / SURVEY.md §8(d)."""
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
from sequencedetectionqueryexecutor_b200 import api  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--traces", type=int, default=2_000_000)
    ap.add_argument("--sample", type=int, default=20_000)
    args = ap.parse_args()
    import torch
    dev = torch.device("cuda", 0)
    peak, _ = bench.measured_peak_gbs()
    out = []
    for name, n_act, lo, hi in (("cfg3_declare_20act", 20, 30, 70), ("declare_20act_len50", 20, 50, 50), ("cfg4_stats_100act", 100, 50, 50), ("declare_400act_len50", 400, 50, 50)):
        off, act, ts = bench.make_log_fast(args.traces, lo, hi, n_act, 0x51E57A03, 600)
        T, E = len(off) - 1, len(act)
        d = [torch.from_numpy(x).to(dev) for x in (off, act, ts)]
        ctx = api.Context(0)
        log = ctx.wrap_log(*d, n_act, max_trace_len=hi)
        S = min(args.sample, T)
        s_off, s_act, s_ts = off[:S + 1], act[:int(off[S])], ts[:int(off[S])]
        # K3
        ms = min(log.declare_counts(k_cap=64).kernel_ms for _ in range(4))
        t0 = time.perf_counter(); want = oracle.declare_counts(s_off, s_act, n_act, 64); cpu = time.perf_counter() - t0
        slog = ctx.load_log(s_off, s_act, s_ts, n_act)
        ok = bool(np.array_equal(slog.declare_counts(k_cap=64).packed, want.packed))
        out.append({"kernel": "K3 declare_pairs_kernel (<= 32 activities, <= 128 events/trace)" if n_act <= 32 and hi <= 128 else
                    ("K3 declare_any_kernel (lanes = distinct activities of the trace, global atomics)" if n_act > 104 else "K3 declare_kernel + declare_alt_chain_kernel (serial fallback)"),
                    "workload": name, "traces": T, "events": E, "kernel_ms": ms,
                    "events_per_s": E / (ms * 1e-3), "algorithmic_GBps": (4 * E + 8 * T) / (ms * 1e-3) / 1e9, "frac_of_hbm_peak": (4 * E + 8 * T) / (ms * 1e-3) / 1e9 / peak,
                    "cpu_oracle_events_per_s": int(off[S]) / cpu, "parity_on_sample": ok})
        # K4: the consecutive pairs of a 3-event pattern + 30 more pairs
        pairs = [(0, 1), (1, 2)] + [(a, b) for a in range(3, 9) for b in range(5)]
        pairs = pairs[:32]
        ms = min(log.pair_stats(pairs)[1] for _ in range(4))
        t0 = time.perf_counter(); want = oracle.pair_stats(s_off, s_act, s_ts, pairs); cpu = time.perf_counter() - t0
        ok = slog.pair_stats(pairs)[0] == want
        out.append({"kernel": "K4 pair_stats_kernel (32 pairs)", "workload": name, "traces": T, "events": E, "kernel_ms": ms, "events_per_s": E / (ms * 1e-3),
                    "algorithmic_GBps": (12 * E + 8 * T) / (ms * 1e-3) / 1e9, "frac_of_hbm_peak": (12 * E + 8 * T) / (ms * 1e-3) / 1e9 / peak,
                    "cpu_oracle_events_per_s": int(off[S]) * len(pairs) / cpu / len(pairs), "parity_on_sample": bool(ok)})
        # K2: index of the 6 true pairs of a 4-event pattern, then their intersection
        tp = [(0, 1), (0, 2), (0, 3), (1, 2), (1, 3), (2, 3)]
        t0 = time.perf_counter(); idx = log.build_index(tp); torch.cuda.synchronize(); build_s = time.perf_counter() - t0
        t0 = time.perf_counter(); cand = idx.intersect(); inter_s = time.perf_counter() - t0
        lens = [len(idx.posting_list(i)) for i in range(len(tp))]
        sidx = slog.build_index(tp)
        ok = bool(np.array_equal(sidx.intersect(), oracle.intersect([oracle.posting_list(s_off, s_act, a, b) for a, b in tp])))
        out.append({"kernel": "K2 index build (6 pairs) + intersection", "workload": name, "traces": T, "events": E, "build_ms_wall": build_s * 1e3,
                    "intersect_ms_wall": inter_s * 1e3, "intersect_kernel_ms": min(idx.intersect_device_ms()[1] for _ in range(4)),
                    "intersect_algorithmic_GBps": 8 * (sum(lens) + len(cand)) / (min(idx.intersect_device_ms()[1] for _ in range(2)) * 1e-3) / 1e9, "list_lengths": lens, "result": int(len(cand)),
                    "build_algorithmic_GBps": (4 * E + 8 * T) / build_s / 1e9, "parity_on_sample": ok})
        sidx.close(); idx.close(); slog.close(); log.close(); ctx.close()
    for o in out:
        print(json.dumps(o))


if __name__ == "__main__":
    main()
