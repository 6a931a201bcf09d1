"""Multi-process check of the library's exchange (one process per GPU, CUDA IPC handles carried by torch.distributed):
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29517 tests/exchange_worker.py
Every rank verifies its shard, joins the match lists with siesta_detect_allgather and compares the JOINED list with the
oracle's result on the unsharded log; then the declare counts are all-reduced over the peer regions and compared too.
Needs N GPUs with peer access (not collected by pytest; tests/test_exchange_gpu.py covers the same code on one device)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle  # noqa: E402
from sequencedetectionqueryexecutor_b200 import _abi as abi  # noqa: E402
from sequencedetectionqueryexecutor_b200 import api  # noqa: E402
from sequencedetectionqueryexecutor_b200 import distributed as D  # noqa: E402
from tests import gen  # noqa: E402
from tests.test_exchange_gpu import ABC, AB, GAP6, KLEENE  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    ctx = api.Context(local)
    join = D.MatchExchange(ctx, dev)
    bad = 0
    for name, states, flags, shape in [("gap6", GAP6, 0, dict(lo=50, hi=50, n_act=20, n=40000)),
                                       ("kleene", KLEENE, 0, dict(lo=100, hi=100, n_act=20, n=20000, gap=120)),
                                       ("abc", ABC, 0, dict(lo=5, hi=40, n_act=6, n=20000)),
                                       ("ab returnAll", AB, abi.F_RETURN_ALL, dict(lo=5, hi=60, n_act=5, n=20000))]:
        off, act, ts = gen.make_log(shape["n"], shape["lo"], shape["hi"], shape["n_act"], seed=99, max_gap_s=shape.get("gap", 300))
        nfa = abi.make_nfa(states)
        s_off, s_act, s_ts, lo = D.local_shard(off, act, ts, rank, world)
        log = ctx.load_log(s_off, s_act, s_ts, shape["n_act"])
        log.set_first_trace(lo)
        for it in range(3):
            dm, st = join.detect_allgather(log, nfa, flags)
            got = D.to_match_result(dm.tensors(local), dm.n_matches_emitted)
            dm.close()
        want = oracle.detect(off, act, ts, nfa, flags=flags)
        ok, why = got.same_as(want)
        print(f"[rank {rank}] {name}: joined {got.n_traces} traces / {got.n_events} events, local {st.local_traces}, "
              f"pulled {st.pulled_bytes} B, scan {st.scan_ms:.3f} wait {st.wait_ms:.3f} pull {st.pull_ms:.3f} ms -> "
              f"{'OK' if ok else 'MISMATCH in ' + why}", flush=True)
        bad += not ok
        # counts: all-reduce over the peer regions
        n = len(oracle.declare_counts(off[:2], act[:off[1]], shape["n_act"], 20).packed)
        buf = torch.zeros(n, dtype=torch.int64, device=dev)
        log.declare_counts_device(buf, 20)
        torch.cuda.synchronize()
        join.allreduce_counts(buf)
        okc = np.array_equal(buf.cpu().numpy(), oracle.declare_counts(off, act, shape["n_act"], 20).packed)
        print(f"[rank {rank}] {name}: all-reduced declare counts -> {'OK' if okc else 'MISMATCH'}", flush=True)
        bad += not okc
        log.close()
    join.close()
    ctx.close()
    t = torch.tensor([bad], device=dev)
    dist.all_reduce(t)
    dist.destroy_process_group()
    if int(t.item()):
        raise SystemExit(1)
    if rank == 0:
        print("exchange_worker: all ranks OK", flush=True)


if __name__ == "__main__":
    main()
