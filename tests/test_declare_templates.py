"""Row D4: the /declare templates (sequencedetectionqueryexecutor_b200/declare.py, fed by the count matrices of kernel K3)
against an independent restatement of the reference's Spark jobs that works on SETS OF TRACE IDS and per-trace position
lists, the way QueryPlanExistences / QueryPlanOrderedRelations / QueryPlanPositions do (tables under the SeqTable view).
The reference has no tests for declare/: parity is pinned by code reading; these tests pin the count-matrix formulation
to the set formulation, on the oracle's counts (CPU) and on the GPU's (-m gpu)."""
import itertools
from collections import Counter, defaultdict

import numpy as np
import pytest

import oracle
from sequencedetectionqueryexecutor_b200.declare import DeclareMiner
from tests import gen


class SetRestatement:
    """the Spark jobs, literally, on python sets"""

    def __init__(self, off, act, names):
        self.names = names
        self.N = len(off) - 1
        self.single = defaultdict(dict)          # type -> {trace: [positions]}       (single.parquet)
        self.index = defaultdict(set)            # (a, b) -> traces with an a before a b (index.parquet, Seq view)
        self.first, self.last = Counter(), Counter()
        for t in range(self.N):
            seg = act[off[t]:off[t + 1]].tolist()
            if not seg:
                continue
            self.first[names[seg[0]]] += 1
            self.last[names[seg[-1]]] += 1
            for i, x in enumerate(seg):
                self.single[names[x]].setdefault(t, []).append(i)
            present = {names[x] for x in seg}
            for a in present:
                for b in present:
                    pa, pb = self.single[a][t], self.single[b][t]
                    if (a == b and len(pa) >= 2) or (a != b and pa[0] < pb[-1]):
                        self.index[(a, b)].add(t)
        self.group_times = {a: Counter(len(p) for p in tr.values()) for a, tr in self.single.items()}
        self.unique = {a: sum(g.values()) for a, g in self.group_times.items()}
        self.joined = {k: len(v | self.index.get((k[1], k[0]), set())) for k, v in self.index.items()}
        types = set(self.group_times)
        self.not_found = {(a, b) for a in types for b in types if a != b} - set(self.joined)

    def never_together(self):
        return {(a, b) if a > b else (b, a) for a, b in self.not_found if (b, a) in self.not_found}

    def existences(self, support):
        N, U, out = self.N, self.unique, {}
        out["existence"] = {(a, time, sum(v for k, v in g.items() if k >= time) / N) for a, g in self.group_times.items()
                            for time in (3, 2, 1) if sum(v for k, v in g.items() if k >= time) / N >= support}
        ab = set()
        for a, g in self.group_times.items():
            g0 = dict(g)
            g0[0] = N - sum(g.values())
            for time in (3, 2):
                s = sum(v for k, v in g0.items() if k < time) / N
                if s >= support:
                    ab.add((a, time, s))
        out["absence"] = ab
        out["exactly"] = {(a, k, v / N) for a, g in self.group_times.items() for k, v in g.items() if v >= support * N}
        out["co-existence"] = {(a, b, float(N - U[a] - U[b] + 2 * n) / N) for (a, b), n in self.joined.items()
                               if a != b and a <= b and n >= support * N and float(N - U[a] - U[b] + 2 * n) >= support * N}
        nce = {(a, b, 1 - n / N) for (a, b), n in self.joined.items() if a != b and a <= b and n <= (1 - support) * N}
        out["not-co-existence"] = nce | {(a, b, 1.0) for a, b in self.never_together()}
        ch = set()
        for a, b in itertools.product(self.single, self.single):
            if a < b and len(self.single[a]) + len(self.single[b]) >= support * N:
                s = len(set(self.single[a]) | set(self.single[b])) / N
                if s >= support:
                    ch.add((a, b, s))
        out["choice"] = ch
        ex = {(a, b, (U[a] + U[b] - 2 * n) / N) for (a, b), n in self.joined.items() if a < b and (U[a] + U[b] - 2 * n) / N >= support}
        out["exclusive-choice"] = ex | {(a, b, (U[a] + U[b]) / N) for a, b in self.never_together() if (U[a] + U[b]) / N >= support}
        re_ = set()
        for (a, b), n in self.joined.items():
            if a != b:
                re_.add((a, b, (float(n) + N - U[a]) / N))
                re_.add((b, a, (float(n) + N - U[b]) / N))
        out["responded-existence"] = {x for x in re_ if x[2] >= support}
        return out

    def count(self, kind, mode, la, lb):
        """OrderedRelationsUtilityFunctions :25-102 on the two position lists of a listed trace"""
        if mode == "simple":
            return sum(any(y > x for y in lb) for x in la) if kind == "r" else sum(any(x < y for x in la) for y in lb)
        if mode == "chain":
            return sum((x + 1) in lb for x in la) if kind == "r" else sum((y - 1) in la for y in lb)
        if kind == "r":
            return sum(any(la[i] < y < la[i + 1] for y in lb) for i in range(len(la) - 1)) + any(y > la[-1] for y in lb)
        return sum(any(lb[i - 1] < x < lb[i] for x in la) for i in range(1, len(lb))) + any(x < lb[0] for x in la)

    def ordered_relations(self, mode, constraint, support):
        tot = {a: sum(len(p) for p in tr.values()) for a, tr in self.single.items()}
        recs = {}
        for (a, b), traces in self.index.items():
            if a == b:
                continue
            for kind in ("p", "r"):
                if (kind == "p" and constraint == "response") or (kind == "r" and constraint == "precedence"):
                    continue
                recs[(kind, a, b)] = sum(self.count(kind, mode, self.single[a][t], self.single[b][t]) for t in traces)
        out = {"response": set(), "precedence": set(), "succession": set(), "not-succession": set()}
        found = {(a, b) for _, a, b in recs}
        out["not-succession"] |= {(a, b, 1.0) for a in tot for b in tot if a != b and (a, b) not in found}
        inter = {(k, a, b): n / (tot[a] if k == "r" else tot[b]) for (k, a, b), n in recs.items()}
        resp = {(a, b): s for (k, a, b), s in inter.items() if k == "r" and s >= support}
        prec = {(a, b): s for (k, a, b), s in inter.items() if k == "p" and s >= support}
        if resp and prec:
            out["response"] = {(a, b, s) for (a, b), s in resp.items()}
            out["precedence"] = {(a, b, s) for (a, b), s in prec.items()}
            out["succession"] = {(a, b, s * prec[(a, b)]) for (a, b), s in resp.items() if (a, b) in prec}
            if mode != "alternate":
                lr = {(a, b): s for (k, a, b), s in inter.items() if k == "r" and s <= 1 - support}
                lp = {(a, b): s for (k, a, b), s in inter.items() if k == "p" and s <= 1 - support}
                out["not-succession"] |= {(a, b, (1 - s) * (1 - lp[(a, b)])) for (a, b), s in lr.items() if (a, b) in lp}
        elif not prec:
            out["response"] = {(a, b, s) for (a, b), s in resp.items()}
        else:
            out["precedence"] = {(a, b, s) for (a, b), s in prec.items()}
        return out


def as_sets(d):
    return {k: {tuple(x.__dict__.values()) for x in v} for k, v in d.items()}


MODES = ["existence", "absence", "exactly", "co-existence", "not-co-existence", "choice", "exclusive-choice", "responded-existence"]
LOGS = [dict(n_traces=300, min_len=0, max_len=14, n_act=6, seed=31),             # sparse: pairs that never co-occur
        dict(n_traces=400, min_len=20, max_len=50, n_act=8, seed=32, zipf=1.3),  # skewed: high and low supports
        dict(n_traces=60, min_len=1, max_len=4, n_act=9, seed=33)]


def check(counts, off, act, names):
    ref = SetRestatement(off, act, names)
    miner = DeclareMiner(counts, names, len(off) - 1)
    for support in (0.0, 0.1, 0.5, 0.9, 1.0):
        got = as_sets(miner.existences(MODES, support))
        want = ref.existences(support)
        for m in MODES:
            assert got[m] == want[m], (m, support, sorted(got[m] ^ want[m])[:4])
        for mode in ("simple", "alternate", "chain"):
            for constraint in ("response", "precedence", "succession"):
                got = as_sets(miner.ordered_relations(mode, constraint, support))
                want = ref.ordered_relations(mode, constraint, support)
                for k in want:
                    assert got[k] == want[k], (mode, constraint, k, support, sorted(got[k] ^ want[k])[:4])
        pos = miner.positions("both", support)
        assert {(x.ev, x.support) for x in pos["first"]} == {(a, n / ref.N) for a, n in ref.first.items() if n / ref.N >= support}
        assert {(x.ev, x.support) for x in pos["last"]} == {(a, n / ref.N) for a, n in ref.last.items() if n / ref.N >= support}
        assert set(miner.positions("first", support)) == {"first"} and set(miner.positions("last", support)) == {"last"}


@pytest.mark.parametrize("lg", LOGS, ids=["sparse", "zipf", "tiny"])
def test_templates_on_the_oracles_counts(lg):
    off, act, ts = gen.make_log(**lg)
    names = [f"ev{chr(ord('H') - i)}{i}" for i in range(lg["n_act"])]   # name order differs from id order
    check(oracle.declare_counts(off, act, lg["n_act"], 64), off, act, names)


def test_truncated_histogram_is_refused():
    off, act, ts = gen.make_log(50, 20, 30, 2, seed=3)
    with pytest.raises(ValueError):
        DeclareMiner(oracle.declare_counts(off, act, 2, 3), ["a", "b"], 50)


@pytest.mark.gpu
def test_templates_on_the_gpu_counts():
    from sequencedetectionqueryexecutor_b200 import api
    lg = LOGS[1]
    off, act, ts = gen.make_log(**lg)
    names = [f"ev{chr(ord('H') - i)}{i}" for i in range(lg["n_act"])]
    with api.Context(0) as ctx:
        log = ctx.load_log(off, act, ts, lg["n_act"])
        counts = log.declare_counts(k_cap=64)
        log.close()
    check(counts, off, act, names)
