"""Multi-GPU plumbing: one process per GPU, traces sharded by contiguous trace range.

The scan itself has no exchange step: every rank verifies / counts its own shard.  Two collectives join the
results over NCCL (NVLink 5 / NVSwitch): an all-gather of the per-rank match lists (/detection) and a sum
all-reduce of the packed int64 count array (/declare, /stats).  Both are latency-bound (outputs are a few
percent of the scan), so they are issued once per request on padded buffers.

Works on CPU tensors with the gloo backend too (tests/test_distributed_cpu.py).
"""
import numpy as np
import os

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


# ---------------------------------------------------------------------------------------------- block exchange
# siesta_dev_matches keeps a result in ONE allocation (include/siesta_gpu.h): eight arrays at 256-byte aligned
# offsets.  The /detection exchange ships that block as it is: a header all-gather (5 int64 per rank), one copy of
# the own block into the receive buffer, one padded all-gather of bytes.  The joined match list is the sequence of
# per-rank parts (each a self-contained CSR with GLOBAL trace indices, see siesta_log_set_first_trace); ranks own
# ascending disjoint trace ranges, so the parts in rank order are the trace-ordered result.  Nothing is re-packed,
# and the all-gather may still be in flight when the next request's kernels start (JoinedMatches.wait()).

def _align256(n):
    return (n + 255) & ~255


def block_layout(n_tr, n_occ, n_ev, n_err, all_cols):
    """Byte offsets of the arrays inside a result block (mirror of detect_device_impl's carve order)."""
    off, o = {}, 0
    for name, nbytes in (("trace_idx", n_tr * 8), ("occ_off", (n_tr + 1) * 8), ("ev_off", (n_occ + 1) * 8), ("ev_pos", n_ev * 4),
                         ("err_trace_idx", n_err * 8)):
        off[name] = (o, nbytes)
        o += _align256(nbytes)
    if all_cols:
        for name, nbytes in (("ev_rank", n_ev * 4), ("ev_act", n_ev * 4), ("ev_ts_ms", n_ev * 8)):
            off[name] = (o, nbytes)
            o += _align256(nbytes)
    return off, o


_DT = {"trace_idx": torch.int64, "occ_off": torch.int64, "ev_off": torch.int64, "ev_pos": torch.int32,
       "err_trace_idx": torch.int64, "ev_rank": torch.int32, "ev_act": torch.int32, "ev_ts_ms": torch.int64}


def packed_layout(n_tr, n_occ, n_ev, n_err, all_cols):
    """Byte offsets inside a compact block (mirror of siesta_dev_matches_pack, csrc/explore.cu)."""
    off, o = {}, 0
    secs = [("trace_local", n_tr * 4), ("occ_cnt", n_tr), ("ev_cnt", n_occ), ("ev_pos", n_ev * 2), ("err_trace_idx", n_err * 8)]
    if all_cols:
        secs += [("ev_rank", n_ev), ("ev_act", n_ev * 2), ("ts_base", n_tr * 8), ("ts_delta", n_ev * 4)]
    for name, nbytes in secs:
        off[name] = (o, nbytes)
        o += _align256(nbytes)
    return off, o


_PDT = {"trace_local": torch.int32, "occ_cnt": torch.uint8, "ev_cnt": torch.uint8, "ev_pos": torch.int16,
        "err_trace_idx": torch.int64, "ev_rank": torch.uint8, "ev_act": torch.int16, "ts_base": torch.int64,
        "ts_delta": torch.int32}


def unpack_block(row, header):
    """Compact block -> the standard columns (dict of tensors, global trace indices)."""
    n_tr, n_occ, n_ev, n_err, all_cols, fmt, trace_base, seconds = header[:8]
    lay, _ = packed_layout(n_tr, n_occ, n_ev, n_err, all_cols)
    v = {k: row[o:o + nb].view(_PDT[k]) for k, (o, nb) in lay.items()}
    dev = row.device

    def offsets(cnt, n):
        off = torch.zeros(n + 1, dtype=torch.int64, device=dev)
        torch.cumsum(cnt.to(torch.int64), 0, out=off[1:])
        return off

    out = {"trace_idx": v["trace_local"].to(torch.int64) + trace_base, "occ_off": offsets(v["occ_cnt"], n_tr),
           "ev_off": offsets(v["ev_cnt"], n_occ), "ev_pos": v["ev_pos"].to(torch.int32) & 0xFFFF,
           "err_trace_idx": v["err_trace_idx"]}
    if all_cols:
        out["ev_rank"] = v["ev_rank"].to(torch.int32)
        out["ev_act"] = v["ev_act"].to(torch.int32) & 0xFFFF
        # event -> its trace: events per trace = sum of ev_cnt over the trace's occurrences
        ev_per_occ = v["ev_cnt"].to(torch.int64)
        occ_trace = torch.repeat_interleave(torch.arange(n_tr, device=dev), v["occ_cnt"].to(torch.int64))
        ev_trace = torch.repeat_interleave(occ_trace, ev_per_occ)
        out["ev_ts_ms"] = v["ts_base"][ev_trace] + v["ts_delta"].to(torch.int64) * (1000 if seconds else 1)
    return out


class JoinedMatches:
    """The match lists of all ranks after exchange_blocks: parts[r] = dict of tensors (views into the receive buffer
    for plain blocks, decoded columns for compact ones)."""

    def __init__(self, recv, headers, work):
        self.recv, self.headers, self._work = recv, headers, work
        self.n_traces = int(sum(h[0] for h in headers))
        self.n_occurrences = int(sum(h[1] for h in headers))
        self.n_events = int(sum(h[2] for h in headers))

    def wait(self):
        """Block the current stream (not the host) until the all-gather has delivered every part."""
        if self._work is not None:
            self._work.wait()
            self._work = None
        return self

    @property
    def parts(self):
        self.wait()
        out = []
        for r, h in enumerate(self.headers):
            n_tr, n_occ, n_ev, n_err, all_cols, fmt = h[:6]
            row = self.recv[r]
            if fmt == 1:
                out.append(unpack_block(row, h))
            else:
                lay, _ = block_layout(n_tr, n_occ, n_ev, n_err, all_cols)
                out.append({k: row[o:o + nb].view(_DT[k]) for k, (o, nb) in lay.items()})
        return out

    def concatenated(self):
        """One CSR over all ranks (copies; for consumers that want a single array per column)."""
        parts = self.parts
        out = {"trace_idx": torch.cat([p["trace_idx"] for p in parts]),
               "err_trace_idx": torch.cat([p["err_trace_idx"] for p in parts])}
        for name in ("occ_off", "ev_off"):
            c = torch.cat([p[name][1:] - p[name][:-1] for p in parts])
            off = torch.zeros(c.numel() + 1, dtype=torch.int64, device=c.device)
            torch.cumsum(c, 0, out=off[1:])
            out[name] = off
        for k in ("ev_pos", "ev_rank", "ev_act", "ev_ts_ms"):
            if k in parts[0]:
                out[k] = torch.cat([p[k] for p in parts])
        return out


def exchange_blocks(block, header, group=None, async_op=True):
    """All-gather of result blocks.  block: uint8 tensor; header: (n_traces, n_occurrences, n_events, n_ref_errors,
    has_event_columns[, format, trace_base, seconds]) with format 0 = the library's plain block (DeviceMatches.block(),
    pack_block()) and 1 = the compact wire format (DeviceMatches.packed_block()).  After this call returns the caller
    may free `block` (it has been copied into the receive buffer); the all-gather itself may still be running."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    dev = block.device
    header = (list(header) + [0, 0, 0])[:8]
    h = torch.tensor(header + [block.numel()], dtype=torch.int64, device=dev)
    hs = torch.empty(world * 9, dtype=torch.int64, device=dev)
    if dev.type == "cuda":
        dist.all_gather_into_tensor(hs, h, group=group)
    else:
        parts = [torch.empty_like(h) for _ in range(world)]
        dist.all_gather(parts, h, group=group)
        hs = torch.cat(parts)
    hs = hs.view(world, 9).cpu().tolist()          # the only host synchronisation of the exchange
    maxb = max(max(x[8] for x in hs), 256)
    recv = torch.empty((world, maxb), dtype=torch.uint8, device=dev)
    recv[rank, :block.numel()].copy_(block)
    work = None
    if dev.type == "cuda":
        torch.cuda.current_stream(dev).synchronize()  # the copy is done: `block` may be freed by the caller
        work = dist.all_gather_into_tensor(recv.view(-1), recv[rank], group=group, async_op=async_op)
        if not async_op:
            work = None
    else:
        rows = [torch.empty(maxb, dtype=torch.uint8) for _ in range(world)]
        dist.all_gather(rows, recv[rank].clone(), group=group)
        for r in range(world):
            recv[r].copy_(rows[r])
    return JoinedMatches(recv, [tuple(x[:8]) for x in hs], work)


class PeerExchanger:
    """exchange_blocks with the payload moved by the copy engines over NVLink peer memory instead of an SM-based NCCL
    kernel: every rank stages its block in a symmetric-memory buffer and PULLS the other ranks' blocks with plain
    device-to-device copies on a side stream (the pattern of torch's low-contention all-gather).  The copies use no SM,
    so the next request's verification kernel runs beside them at full speed; only the 9-value header still goes
    through NCCL.  One node only (symmetric memory needs peer access); `depth` requests may be in flight."""

    def __init__(self, device, group=None, capacity=64 << 20, depth=3, n_streams=4):
        import torch.distributed._symmetric_memory as symm
        self._symm = symm
        self.group = group if group is not None else dist.group.WORLD
        self.device, self.depth = device, depth
        self.rank, self.world = dist.get_rank(self.group), dist.get_world_size(self.group)
        self.stream = torch.cuda.Stream(device)
        # one copy engine does not fill an NVLink port: the pulls of one request are spread over several streams
        self.pull_streams = [torch.cuda.Stream(device) for _ in range(max(1, int(os.environ.get("SIESTA_PEER_STREAMS", n_streams))))]
        self.k = 0
        self._alloc(capacity)

    def _alloc(self, capacity):
        self.capacity = int(capacity)
        self.bufs = [self._symm.empty(self.capacity, dtype=torch.uint8, device=self.device) for _ in range(self.depth)]
        self.hdls = [self._symm.rendezvous(b, self.group) for b in self.bufs]
        self.free = [None] * self.depth   # event: every rank has finished pulling from this slot

    def exchange(self, block, header):
        """Same contract as exchange_blocks: on return `block` may be freed; the pulls may still be running."""
        world, rank, dev = self.world, self.rank, self.device
        header = (list(header) + [0, 0, 0])[:8]
        h = torch.tensor(header + [block.numel()], dtype=torch.int64, device=dev)
        hs = torch.empty(world * 9, dtype=torch.int64, device=dev)
        dist.all_gather_into_tensor(hs, h, group=self.group)
        hs = hs.view(world, 9).cpu().tolist()
        maxb = max(max(x[8] for x in hs), 256)
        if maxb > self.capacity:            # every rank sees the same sizes: a collective, deterministic decision
            torch.cuda.synchronize(dev)
            self._alloc(max(2 * maxb, 2 * self.capacity))
        slot = self.k % self.depth
        self.k += 1
        cur = torch.cuda.current_stream(dev)
        if self.free[slot] is not None:
            cur.wait_event(self.free[slot])
        self.bufs[slot][:block.numel()].copy_(block)
        cur.synchronize()                    # the own block is staged: the caller may free it
        recv = torch.empty((world, maxb), dtype=torch.uint8, device=dev)
        hdl = self.hdls[slot]
        self.stream.wait_stream(cur)
        with torch.cuda.stream(self.stream):
            hdl.barrier()                    # every rank has staged its block
            staged = torch.cuda.Event()
            staged.record(self.stream)
        for step in range(world):
            r = (rank - step) % world
            n = hs[r][8]
            if n:
                st = self.pull_streams[step % len(self.pull_streams)]
                st.wait_event(staged)
                with torch.cuda.stream(st):
                    src = hdl.get_buffer(r, (self.capacity,), torch.uint8)
                    recv[r, :n].copy_(src[:n], non_blocking=True)
        for st in self.pull_streams:
            self.stream.wait_stream(st)
            recv.record_stream(st)
        with torch.cuda.stream(self.stream):
            hdl.barrier()                    # every rank has pulled: the slot may be overwritten
            done = torch.cuda.Event()
            done.record(self.stream)
        self.free[slot] = done
        recv.record_stream(self.stream)
        return JoinedMatches(recv, [tuple(x[:8]) for x in hs], _EventWork(done))


class _EventWork:
    def __init__(self, event):
        self.event = event

    def wait(self):
        torch.cuda.current_stream().wait_event(self.event)


def pack_block(tensors):
    """dict of 1-D tensors (as DeviceMatches.tensors() / match_result_to_tensors) -> (uint8 block, header): the layout
    the library produces natively; used where the result did not come from the library (CPU tests)."""
    n_tr, n_occ = tensors["trace_idx"].numel(), tensors["ev_off"].numel() - 1
    n_ev, n_err = tensors["ev_pos"].numel(), tensors["err_trace_idx"].numel()
    all_cols = 1 if tensors.get("ev_rank") is not None else 0
    lay, total = block_layout(n_tr, n_occ, n_ev, n_err, all_cols)
    block = torch.zeros(max(total, 1), dtype=torch.uint8, device=tensors["trace_idx"].device)
    for k, (o, nb) in lay.items():
        if nb:
            block[o:o + nb].copy_(tensors[k].contiguous().view(torch.uint8))
    return block, (n_tr, n_occ, n_ev, n_err, all_cols)


def allreduce_counts(packed, group=None):
    """Sum all-reduce of the packed int64 count array of siesta_declare_counts (in place)."""
    dist.all_reduce(packed, op=dist.ReduceOp.SUM, group=group)
    return packed


def allreduce_pair_stats(packed, group=None):
    """Combine siesta_pair_stats_device results (int64[8 * n_pairs]: count, sum, min, max, four 32-bit limbs of the
    sum of squares) across ranks: SUM for count / sum / limbs, MIN / MAX for the extremes, then carry-normalise the
    limbs.  In place; returns the tensor."""
    v = packed.view(-1, 8)
    sums = v[:, [0, 1, 4, 5, 6, 7]].contiguous()
    mn = v[:, 2].contiguous()
    mx = v[:, 3].contiguous()
    dist.all_reduce(sums, op=dist.ReduceOp.SUM, group=group)
    dist.all_reduce(mn, op=dist.ReduceOp.MIN, group=group)
    dist.all_reduce(mx, op=dist.ReduceOp.MAX, group=group)
    limbs = sums[:, 2:6].clone()
    for k in range(3):
        carry = limbs[:, k] >> 32
        limbs[:, k] &= 0xFFFFFFFF
        limbs[:, k + 1] += carry
    v[:, 0], v[:, 1], v[:, 2], v[:, 3] = sums[:, 0], sums[:, 1], mn, mx
    v[:, 4:8] = limbs
    return packed


def pair_stats_records(packed):
    """packed int64[8 * n_pairs] (after allreduce_pair_stats) -> list of dicts like EventLog.pair_stats."""
    out = []
    for r in packed.view(-1, 8).cpu().tolist():
        cnt = r[0]
        sq = r[4] | (r[5] << 32) | (r[6] << 64) | (r[7] << 96)
        out.append({"count": cnt, "sum": r[1], "min": r[2] if cnt else 0, "max": r[3] if cnt else 0, "sum_squares": sq})
    return out


def allreduce_explore(completions, sum_duration_ms, group=None):
    """Combine siesta_explore_accurate results across ranks (both are plain sums over traces)."""
    dist.all_reduce(completions, op=dist.ReduceOp.SUM, group=group)
    dist.all_reduce(sum_duration_ms, op=dist.ReduceOp.SUM, group=group)
    return completions, sum_duration_ms


# ------------------------------------------------------------------------------------- exchange inside the library
class MatchExchange:
    """One process per GPU (torchrun): the library's own exchange (siesta_exchange_*, csrc/multi.cu) wired up over
    torch.distributed.  torch only carries the 64-byte IPC handles and the capacity agreement at set-up; every request
    afterwards runs inside libsiesta_gpu: the scan places its compact block in the rank's region, sizes travel in the
    block's header, one kernel pulls all peers' blocks over NVLink and decodes them to the joined columns."""

    def __init__(self, ctx, device, group=None):
        self.ctx, self.device, self.group = ctx, device, group
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.x = None
        self.capacity = 0

    def _ensure(self, need):
        """Collective: (re)create the exchanges when any rank needs a larger region."""
        t = torch.tensor([need], dtype=torch.int64, device=self.device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=self.group)
        need = int(t.item())
        if self.x is not None and need <= self.capacity:
            return
        from . import api
        if self.x is not None:
            torch.cuda.synchronize(self.device)
            dist.barrier(group=self.group)
            self.x.close()
        self.capacity = int(need * 1.25) + (1 << 20)
        self.x = api.Exchange(self.ctx, self.world, self.rank, self.capacity)
        mine = torch.frombuffer(bytearray(self.x.export()), dtype=torch.uint8).to(self.device)
        handles = torch.empty(self.world * mine.numel(), dtype=torch.uint8, device=self.device)
        dist.all_gather_into_tensor(handles, mine, group=self.group)
        handles = handles.cpu().view(self.world, -1)
        for p in range(self.world):
            if p != self.rank:
                self.x.import_peer(p, bytes(handles[p].numpy().tobytes()))
        dist.barrier(group=self.group)

    def detect_allgather(self, log, nfa, flags=0):
        """-> (api.DeviceMatches holding the match list of ALL ranks, ExchangeStats).  Collective."""
        from . import api
        key = (log.n_traces, log.n_events, bytes(nfa), flags)
        if getattr(self, "_sized_for", None) != key:
            self._ensure(api.exchange_required_bytes(log, nfa, flags))
            self._sized_for = key
        return self.x.detect_allgather(log, nfa, flags)

    def allreduce_counts(self, packed, op=_abi.REDUCE_SUM):
        if self.x is None or packed.numel() * 8 > self.capacity:
            self._ensure(packed.numel() * 8)
            self._sized_for = None
        return self.x.allreduce_i64(packed, op, torch.cuda.current_stream(self.device).cuda_stream)

    def close(self):
        if self.x is not None:
            torch.cuda.synchronize(self.device)
            dist.barrier(group=self.group)
            self.x.close()
            self.x = None


def joined_prefix(tensors, n_first_traces):
    """The part of a joined match list (dict of device tensors, DeviceMatches.tensors()) that belongs to global traces
    [0, n_first_traces): prefix slices, because the list is in trace order."""
    tr = tensors["trace_idx"]
    n = int(torch.searchsorted(tr, torch.tensor([n_first_traces], dtype=tr.dtype, device=tr.device)).item())
    n_occ = int(tensors["occ_off"][n].item()) if tensors["occ_off"].numel() else 0
    n_ev = int(tensors["ev_off"][n_occ].item()) if tensors["ev_off"].numel() else 0
    out = {"trace_idx": tr[:n], "occ_off": tensors["occ_off"][:n + 1], "ev_off": tensors["ev_off"][:n_occ + 1],
           "err_trace_idx": tensors["err_trace_idx"][tensors["err_trace_idx"] < n_first_traces]}
    for k in ("ev_pos", "ev_rank", "ev_act", "ev_ts_ms"):
        if k in tensors:
            out[k] = tensors[k][:n_ev]
    return out


_COLS = ("trace_idx", "occ_off", "ev_off", "err_trace_idx", "ev_pos", "ev_rank", "ev_act", "ev_ts_ms")


def send_columns(cols, dst, group=None):
    """Point-to-point transfer of a dict of 1-D tensors (sizes first); the receiver calls recv_columns."""
    dev = cols["trace_idx"].device
    sizes = torch.tensor([cols[k].numel() if k in cols else -1 for k in _COLS], dtype=torch.int64, device=dev)
    dist.send(sizes, dst, group=group)
    for k in _COLS:
        if k in cols and cols[k].numel():
            dist.send(cols[k].contiguous(), dst, group=group)


def recv_columns(src, device, group=None):
    sizes = torch.zeros(len(_COLS), dtype=torch.int64, device=device)
    dist.recv(sizes, src, group=group)
    out = {}
    for k, n in zip(_COLS, sizes.tolist()):
        if n < 0:
            continue
        t = torch.zeros(n, dtype=_DT[k], device=device)
        if n:
            dist.recv(t, src, group=group)
        out[k] = t
    return out


def to_match_result(g, n_matches_emitted=-1):
    """dict of tensors -> host MatchResult (for comparison with the oracle)."""
    r = _abi.MatchResult()
    h = {k: v.cpu().numpy() for k, v in g.items()}
    r.trace_idx, r.occ_off, r.ev_off, r.err_trace_idx = h["trace_idx"], h["occ_off"], h["ev_off"], h["err_trace_idx"]
    r.ev_pos = h.get("ev_pos")
    r.ev_rank, r.ev_act, r.ev_ts_ms = h.get("ev_rank"), h.get("ev_act"), h.get("ev_ts_ms")
    r.n_traces, r.n_occurrences, r.n_events = len(r.trace_idx), len(r.ev_off) - 1, int(r.ev_off[-1])
    r.n_ref_errors, r.n_matches_emitted, r.kernel_ms, r.detect_ms = len(r.err_trace_idx), n_matches_emitted, 0.0, 0.0
    r.unsupported_trace_idx = h.get("unsupported_trace_idx", np.zeros(0, dtype=np.int64))
    r.n_unsupported = len(r.unsupported_trace_idx)
    return r


def match_result_to_tensors(res, device="cpu"):
    d = {k: torch.from_numpy(np.ascontiguousarray(getattr(res, k))).to(device)
         for k in ("trace_idx", "occ_off", "ev_off", "ev_pos", "ev_rank", "ev_act", "ev_ts_ms", "err_trace_idx")
         if getattr(res, k) is not None}
    return d
