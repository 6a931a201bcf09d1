"""Correctness + timing of PeerExchanger against exchange_blocks (NCCL).  torchrun --nproc-per-node N tools/peer_exchange_check.py"""
import os, sys, time
import torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from sequencedetectionqueryexecutor_b200 import _abi as abi, api, distributed as D

rank = int(os.environ.get("RANK", 0)); lr = int(os.environ.get("LOCAL_RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(lr); dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
wl = bench.WORKLOADS[bench.DEFAULT_WORKLOAD]
n = 200_000 + 1000 * rank   # different sizes per rank
off, act, ts = bench.make_log_fast(n, wl["min_len"], wl["max_len"], wl["n_act"], wl["seed"], wl["max_gap_s"], rank)
d = [torch.from_numpy(x).to(dev) for x in (off, act, ts)]
ctx = api.Context(lr); log = ctx.wrap_log(*d, wl["n_act"], max_trace_len=100); log.set_first_trace(rank * 10_000_000)
nfa = abi.make_nfa(wl["states"])
peer = D.PeerExchanger(dev, capacity=1 << 20)   # small: forces one collective re-allocation
ok = True
for it in range(5):
    dm = log.detect_device(nfa)
    blk, hdr = dm.packed_block(log, 0, rank * 10_000_000, lr) if it % 2 == 0 else dm.block(lr)
    a = D.exchange_blocks(blk, hdr, async_op=False).concatenated()
    b = peer.exchange(blk, hdr).concatenated()
    torch.cuda.synchronize()
    for k in a:
        ok = ok and torch.equal(a[k], b[k])
    dm.close()
flag = torch.tensor([1 if ok else 0], device=dev); dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0: print("peer exchange equals NCCL exchange on all ranks:", bool(flag.item()), "traces joined:", int(a["trace_idx"].numel()), flush=True)
dist.destroy_process_group()
