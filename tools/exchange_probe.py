"""Exploration: where does the multi-GPU step spend its time?  torchrun --nproc-per-node N tools/exchange_probe.py"""
import os, sys, time
import torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from sequencedetectionqueryexecutor_b200 import _abi as abi, api, distributed as D

rank = int(os.environ.get("RANK", 0)); lr = int(os.environ.get("LOCAL_RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(lr); dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
wl = bench.WORKLOADS[bench.DEFAULT_WORKLOAD]
off, act, ts = bench.make_log_fast(wl["n_traces"], wl["min_len"], wl["max_len"], wl["n_act"], wl["seed"], wl["max_gap_s"], rank)
d = [torch.from_numpy(x).to(dev) for x in (off, act, ts)]
ctx = api.Context(lr); log = ctx.wrap_log(*d, wl["n_act"], max_trace_len=100); log.set_first_trace(rank * wl["n_traces"])
nfa = abi.make_nfa(wl["states"])

def timed(name, fn, n=20):
    for _ in range(3): fn()
    dist.barrier(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): fn()
    dist.barrier(); torch.cuda.synchronize()
    if rank == 0: print(f"{name:40s} {(time.perf_counter() - t0) / n * 1e3:8.3f} ms", flush=True)

def detect_only():
    dm = log.detect_device(nfa); dm.close()
timed("detect only", detect_only)
dm = log.detect_device(nfa); block, header = dm.block(lr)
def hdr():
    h = torch.tensor(list(header) + [block.numel()], dtype=torch.int64, device=dev)
    hs = torch.empty(world * 6, dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(hs, h); return hs.cpu()
timed("header all-gather + .cpu()", hdr)
recv = torch.empty((world, block.numel() + 4096), dtype=torch.uint8, device=dev)
def cp():
    recv[rank, :block.numel()].copy_(block); torch.cuda.current_stream().synchronize()
timed("copy own block + stream sync", cp)
def ag():
    dist.all_gather_into_tensor(recv.view(-1), recv[rank]); torch.cuda.synchronize()
timed(f"block all-gather ({block.numel() / 1e6:.0f} MB/rank), sync", ag)
def ag_then_detect():
    w = dist.all_gather_into_tensor(recv.view(-1), recv[rank], async_op=True)
    dm2 = log.detect_device(nfa); dm2.close()
    w.wait(); torch.cuda.synchronize()
timed("AG async || detect, then wait", ag_then_detect)
side = torch.cuda.Stream()
def copy_then_detect():
    with torch.cuda.stream(side):
        recv[rank, :block.numel()].copy_(block)
        for _ in range(8): recv[rank, :block.numel()].copy_(block)
    dm2 = log.detect_device(nfa); dm2.close()
    torch.cuda.synchronize()
timed("9 device copies on side stream || detect", copy_then_detect)
def copies_only():
    for _ in range(9): recv[rank, :block.numel()].copy_(block)
    torch.cuda.synchronize()
timed("9 device copies alone", copies_only)
def alloc():
    x = torch.empty((world, block.numel() + 4096), dtype=torch.uint8, device=dev); return x
timed("recv alloc", alloc)
def full_sync():
    dm = log.detect_device(nfa); b, h = dm.block(lr); j = D.exchange_blocks(b, h, async_op=False); dm.close()
timed("detect + exchange (sync)", full_sync)
fl = []
def full_async():
    dm = log.detect_device(nfa); b, h = dm.block(lr); j = D.exchange_blocks(b, h); dm.close(); fl.append(j)
    if len(fl) > 2: fl.pop(0).wait()
timed("detect + exchange (async, 2 in flight)", full_async)
dist.destroy_process_group()
