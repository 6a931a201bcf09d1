"""Secondary measurement: BASELINE configs[2] - /declare counting (existences + ordered relations) on 10 M traces x 50 events,
20 activities, sharded over N GPUs; every rank counts its shard with kernel K3 and the packed count arrays are all-reduced over
the peer regions (siesta_exchange_allreduce_i64).  One process per GPU:
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29531 tools/bench_declare_multi.py
A step = counting + all-reduce; time = wall clock around a barrier + device synchronise, max over ranks.  Parity: the reduced
array restricted to what a prefix contributes is not separable, so rank 0 also counts the first `--sample` traces alone and
compares THAT with the oracle; the all-reduce itself is checked against the sum of the ranks' arrays gathered through NCCL."""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import oracle  # noqa: E402
from sequencedetectionqueryexecutor_b200 import api  # noqa: E402
from sequencedetectionqueryexecutor_b200 import distributed as D  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--traces", type=int, default=10_000_000)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--sample", type=int, default=20_000)
    args = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    n_act, k_cap = 20, 40
    T = args.traces // world
    off, act, ts = bench.make_log_fast(T, 50, 50, n_act, 0x51E57A03, 600, rank=rank)
    d = [torch.from_numpy(x).to(dev) for x in (off, act, ts)]
    ctx = api.Context(local)
    log = ctx.wrap_log(*d, n_act, max_trace_len=50)
    join = D.MatchExchange(ctx, dev) if world > 1 else None
    n = len(oracle.declare_counts(off[:2], act[:off[1]], n_act, k_cap).packed)
    buf = torch.zeros(n, dtype=torch.int64, device=dev)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step():
        ms = log.declare_counts_device(buf, k_cap)
        if join is not None:
            torch.cuda.synchronize()
            join.allreduce_counts(buf)
        return ms

    for _ in range(3):
        step()
    barrier()
    t0 = time.perf_counter()
    k_ms = 0.0
    for _ in range(args.steps):
        k_ms += step()
    barrier()
    sec = (time.perf_counter() - t0) / args.steps
    t = torch.tensor([sec], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    sec = float(t.item())
    # the all-reduce against NCCL's sum of the ranks' own arrays
    own = torch.zeros(n, dtype=torch.int64, device=dev)
    log.declare_counts_device(own, k_cap)
    ok_sum = True
    if world > 1:
        dist.all_reduce(own)
        ok_sum = bool(torch.equal(own, buf))
    S = min(args.sample, T)
    ok_oracle = True
    if rank == 0:
        slog = ctx.load_log(off[:S + 1], act[:int(off[S])], ts[:int(off[S])], n_act)
        ok_oracle = bool(np.array_equal(slog.declare_counts(k_cap).packed, oracle.declare_counts(off[:S + 1], act[:int(off[S])], n_act, k_cap).packed))
        slog.close()
        E = int(off[-1]) * world
        print(json.dumps({"workload": "BASELINE configs[2]: /declare counting, 20 activities x 50 events, traces sharded over the GPUs, counts all-reduced over peer regions",
                          "n_gpus": world, "traces": T * world, "events": E, "ms_per_step": sec * 1e3, "k3_kernel_ms_rank0": k_ms / args.steps,
                          "events_per_s": E / sec, "count_array_int64": n, "allreduce_equals_nccl_sum": ok_sum, "parity_on_sample": ok_oracle}))
    if join is not None:
        join.close()
    log.close()
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
