"""Multi-GPU plumbing: one process per GPU, traces sharded by contiguous trace range.

The scan itself has no exchange step: every rank verifies / counts its own shard.  Two collectives join the
results over NCCL (NVLink 5 / NVSwitch): an all-gather of the per-rank match lists (/detection) and a sum
all-reduce of the packed int64 count array (/declare, /stats).  Both are latency-bound (outputs are a few
percent of the scan), so they are issued once per request on padded buffers.

Works on CPU tensors with the gloo backend too (tests/test_distributed_cpu.py).
"""
import numpy as np
import torch
import torch.distributed as dist

from . import _abi


def shard_bounds(trace_off, world):
    """Contiguous trace ranges balanced by event count: rank r owns traces [b[r], b[r+1])."""
    trace_off = np.asarray(trace_off, dtype=np.int64)
    T = len(trace_off) - 1
    E = int(trace_off[-1])
    targets = (np.arange(1, world, dtype=np.int64) * E) // world
    cuts = np.searchsorted(trace_off, targets, side="left").astype(np.int64)
    b = np.concatenate(([0], np.clip(cuts, 0, T), [T])).astype(np.int64)
    return np.maximum.accumulate(b)


def local_shard(trace_off, act, ts_ms, rank, world):
    """The CSR slice of rank `rank` (offsets rebased to 0) and the global index of its first trace."""
    b = shard_bounds(trace_off, world)
    lo, hi = int(b[rank]), int(b[rank + 1])
    e0, e1 = int(trace_off[lo]), int(trace_off[hi])
    return (np.asarray(trace_off[lo:hi + 1], dtype=np.int64) - e0, np.asarray(act[e0:e1]), np.asarray(ts_ms[e0:e1]), lo)


def _gather_var(t, group=None):
    """all-gather of 1-D tensors of different lengths: sizes first, then one padded all-gather."""
    world = dist.get_world_size(group)
    n = torch.tensor([t.numel()], dtype=torch.int64, device=t.device)
    sizes = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(sizes, n, group=group)
    sizes = [int(s.item()) for s in sizes]
    m = max(max(sizes), 1)
    pad = torch.zeros(m, dtype=t.dtype, device=t.device)
    pad[:t.numel()] = t
    outs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(outs, pad, group=group)
    return [o[:s] for o, s in zip(outs, sizes)]


def allgather_matches(local, first_trace, group=None):
    """Join per-rank match lists (dict of 1-D tensors: trace_idx, occ_off, ev_off, ev_pos[, ev_rank, ev_act, ev_ts_ms],
    err_trace_idx; indices local to the shard) into the global CSR, identical on every rank.  Ranks own ascending,
    disjoint trace ranges, so concatenating in rank order keeps trace_idx ascending."""
    cols = [k for k in ("ev_pos", "ev_rank", "ev_act", "ev_ts_ms") if local.get(k) is not None]
    tr = _gather_var(local["trace_idx"] + first_trace, group)
    err = _gather_var(local["err_trace_idx"] + first_trace, group)
    occ_cnt = _gather_var(local["occ_off"][1:] - local["occ_off"][:-1], group)
    ev_cnt = _gather_var(local["ev_off"][1:] - local["ev_off"][:-1], group)
    out = {"trace_idx": torch.cat(tr), "err_trace_idx": torch.cat(err)}
    for name, parts in (("occ_off", occ_cnt), ("ev_off", ev_cnt)):
        c = torch.cat(parts)
        off = torch.zeros(c.numel() + 1, dtype=torch.int64, device=c.device)
        torch.cumsum(c, 0, out=off[1:])
        out[name] = off
    for k in cols:
        out[k] = torch.cat(_gather_var(local[k], group))
    return out


def allreduce_counts(packed, group=None):
    """Sum all-reduce of the packed int64 count array of siesta_declare_counts (in place)."""
    dist.all_reduce(packed, op=dist.ReduceOp.SUM, group=group)
    return packed


def to_match_result(g, n_matches_emitted=-1):
    """dict of tensors -> host MatchResult (for comparison with the oracle)."""
    r = _abi.MatchResult()
    h = {k: v.cpu().numpy() for k, v in g.items()}
    r.trace_idx, r.occ_off, r.ev_off, r.err_trace_idx = h["trace_idx"], h["occ_off"], h["ev_off"], h["err_trace_idx"]
    r.ev_pos = h.get("ev_pos")
    r.ev_rank, r.ev_act, r.ev_ts_ms = h.get("ev_rank"), h.get("ev_act"), h.get("ev_ts_ms")
    r.n_traces, r.n_occurrences, r.n_events = len(r.trace_idx), len(r.ev_off) - 1, int(r.ev_off[-1])
    r.n_ref_errors, r.n_matches_emitted, r.kernel_ms, r.detect_ms = len(r.err_trace_idx), n_matches_emitted, 0.0, 0.0
    return r


def match_result_to_tensors(res, device="cpu"):
    d = {k: torch.from_numpy(np.ascontiguousarray(getattr(res, k))).to(device)
         for k in ("trace_idx", "occ_off", "ev_off", "ev_pos", "ev_rank", "ev_act", "ev_ts_ms", "err_trace_idx")
         if getattr(res, k) is not None}
    return d
