"""Row D4: the /declare templates on the host, fed by the integer count matrices of kernel K3 (siesta_declare_counts).

The reference computes the counts with Spark jobs and then applies, on the driver, one double division and a threshold
per constraint.  The counting is kernel K3; this module is the driver part, restated from
  J/declare/queryPlans/existence/QueryPlanExistences.java        existence :188-205, absence :213-233, exactly :241-256,
                                                                 coExistence :265-291, notCoExistence :300-330, choice :338-362,
                                                                 exclusiveChoice :372-410, respondedExistence :418-438
  J/declare/queryPlans/orderedRelations/QueryPlanOrderedRelations.java   execute :54-86, filterBasedOnSupport :161-233,
                                                                 extendNotSuccession :266-280 (+ ...Alternate, ...Chain)
  J/declare/queryPlans/position/QueryPlanPositions.java          execute :51-87
  J/declare/DeclareUtilities.java                                extractNotFoundPairs :25-42
with the same operand order of every double expression (IEEE doubles: a Python float is a Java double).  The reference
collects from HashMaps / RDDs, so the ORDER of a result list is unspecified there; here lists follow activity-id order.
Quirks kept on purpose (they decide which pairs are reported): co-existence / not-co-existence / exclusive-choice only look at
index keys (A, B) with A <= B by name, so a pair that only ever occurs as "B before A" is skipped; the never-co-occurring
pairs are reported as (larger name, smaller name).

Parity: the reference has no tests for `declare/` - pinned by code reading; tests/test_declare_templates.py
checks this module against an independent set-based restatement of the same Spark jobs on small logs.
"""
from dataclasses import dataclass
from typing import Dict, List, Sequence


@dataclass(frozen=True)
class EventN:            # J/declare/model/EventN.java
    ev: str
    n: int
    support: float


@dataclass(frozen=True)
class EventSupport:      # J/declare/model/EventSupport.java
    ev: str
    support: float


@dataclass(frozen=True)
class EventPairSupport:  # J/declare/model/EventPairSupport.java
    evA: str
    evB: str
    support: float


class DeclareMiner:
    """counts: _abi.DeclareCounts (GPU or oracle); names: activity id -> name; total_traces: Metadata.getTraces()."""

    def __init__(self, counts, names: Sequence[str], total_traces: int):
        if counts.hist_overflow:
            raise ValueError("the occurrence histogram was truncated (k_cap too small): 'exactly' would miss records")
        self.c, self.names, self.N = counts, list(names), int(total_traces)
        A = counts.n_activities
        # groupTimes.keySet(): the event types of the single table; U = extractUniqueTracesSingle
        self.types = [a for a in range(A) if counts.uniq[a] > 0]
        self.U = {a: int(counts.uniq[a]) for a in self.types}
        # joined = joinUnionTraces: one record per index key (A, B) with |traces(A,B) U traces(B,A)|
        self.joined = [(a, b, int(counts.co[a, b])) for a in self.types for b in self.types if counts.ordered[a, b] > 0]
        found = {(a, b) for a, b, _ in self.joined}
        # DeclareUtilities.extractNotFoundPairs over the single table's event types
        self.not_found = {(a, b) for a in self.types for b in self.types if a != b and (a, b) not in found}

    def _group_times(self, a) -> Dict[int, int]:
        h = self.c.hist[a]
        return {k: int(h[k]) for k in range(1, len(h)) if h[k] > 0}

    # ------------------------------------------------------------------------------------------- existence templates
    def existence(self, support) -> List[EventN]:
        out = []
        for a in self.types:
            t = self._group_times(a)
            for time in (3, 2, 1):
                s = float(sum(v for k, v in t.items() if k >= time)) / self.N
                if s >= support:
                    out.append(EventN(self.names[a], time, s))
        return out

    def absence(self, support) -> List[EventN]:
        out = []
        for a in self.types:
            t = self._group_times(a)
            t[0] = self.N - sum(t.values())
            for time in (3, 2):
                s = float(sum(v for k, v in t.items() if k < time)) / self.N
                if s >= support:
                    out.append(EventN(self.names[a], time, s))
        return out

    def exactly(self, support) -> List[EventN]:
        out = []
        for a in self.types:
            for k, v in self._group_times(a).items():          # keys > 0 only (the reference skips key 0)
                if v >= support * self.N:
                    out.append(EventN(self.names[a], k, float(v) / self.N))
        return out

    def _ordered_by_name(self, a, b):
        return self.names[a] <= self.names[b]

    def _never_together(self):
        """pairs neither order of which occurs, each once, as (larger name, smaller name)"""
        out = set()
        for a, b in self.not_found:
            if (b, a) in self.not_found:
                out.add((a, b) if self.names[a] > self.names[b] else (b, a))
        return sorted(out)

    def co_existence(self, support) -> List[EventPairSupport]:
        out = []
        for a, b, n in self.joined:
            if a == b or not self._ordered_by_name(a, b) or not n >= support * self.N:
                continue
            sup = float(self.N - self.U[a] - self.U[b] + 2 * n)
            if sup >= support * self.N:
                out.append(EventPairSupport(self.names[a], self.names[b], sup / self.N))
        return out

    def not_co_existence(self, support) -> List[EventPairSupport]:
        out = [EventPairSupport(self.names[a], self.names[b], 1 - float(n) / self.N)
               for a, b, n in self.joined
               if a != b and self._ordered_by_name(a, b) and n <= (1 - support) * self.N]
        out += [EventPairSupport(self.names[a], self.names[b], 1.0) for a, b in self._never_together()]
        return out

    def choice(self, support) -> List[EventPairSupport]:
        out = []
        for a in self.types:
            for b in self.types:
                if not self.names[a] < self.names[b]:
                    continue
                if not self.U[a] + self.U[b] >= support * self.N:        # early pruning on the two list sizes
                    continue
                both = int(self.c.co[a, b])                               # traces that hold both
                s = float(self.U[a] + self.U[b] - both) / self.N
                if s >= support:
                    out.append(EventPairSupport(self.names[a], self.names[b], s))
        return out

    def exclusive_choice(self, support) -> List[EventPairSupport]:
        out = []
        for a, b in self._never_together():
            s = float(self.U[a] + self.U[b]) / self.N
            if s >= support:
                out.append(EventPairSupport(self.names[a], self.names[b], s))
        for a, b, n in self.joined:
            if a != b and self.names[a] < self.names[b]:
                s = float(self.U[a] + self.U[b] - 2 * n) / self.N
                if s >= support:
                    out.append(EventPairSupport(self.names[a], self.names[b], s))
        return out

    def responded_existence(self, support) -> List[EventPairSupport]:
        seen, out = set(), []
        for a, b, n in self.joined:
            if a == b:
                continue
            for x, y in ((a, b), (b, a)):
                s = (float(n) + self.N - self.U[x]) / self.N
                e = EventPairSupport(self.names[x], self.names[y], s)
                if e not in seen:                                           # .distinct()
                    seen.add(e)
                    if s >= support:
                        out.append(e)
        return out

    def existences(self, modes, support) -> Dict[str, list]:
        """QueryPlanExistences.execute: the requested modes -> response lists."""
        fn = {"existence": self.existence, "absence": self.absence, "exactly": self.exactly,
              "co-existence": self.co_existence, "not-co-existence": self.not_co_existence, "choice": self.choice,
              "exclusive-choice": self.exclusive_choice, "responded-existence": self.responded_existence}
        return {m: fn[m](support) for m in modes if m in fn}

    # ------------------------------------------------------------------------------------------- positions
    def positions(self, mode, support) -> Dict[str, List[EventSupport]]:
        def side(v):
            return [EventSupport(self.names[a], float(int(v[a])) / self.N) for a in range(self.c.n_activities)
                    if v[a] > 0 and float(int(v[a])) / self.N >= support]
        out = {}
        if mode != "last":
            out["first"] = side(self.c.first)
        if mode != "first":
            out["last"] = side(self.c.last)
        return out

    # ------------------------------------------------------------------------------------------- ordered relations
    def ordered_relations(self, mode, constraint, support) -> Dict[str, List[EventPairSupport]]:
        """mode: simple | alternate | chain; constraint: response | precedence | anything else (= succession: both).
        Returns response / precedence / succession / not-succession lists."""
        c = self.c
        R, P = {"simple": (c.response, c.precedence), "alternate": (c.alt_response, c.alt_precedence),
                "chain": (c.chain_response, c.chain_precedence)}[mode]
        tot = {a: int(c.tot[a]) for a in range(c.n_activities) if c.tot[a] > 0}   # occurrences per event type
        keys = [(a, b) for a in tot for b in tot if a != b and c.ordered[a, b] > 0]  # index rows, eventA != eventB
        recs = []                                                                     # (mode, a, b, occurrences)
        if constraint != "response":
            recs += [("p", a, b, int(P[a, b])) for a, b in keys]
        if constraint != "precedence":
            recs += [("r", a, b, int(R[a, b])) for a, b in keys]
        out = {"response": [], "precedence": [], "succession": [], "not-succession": []}
        # extendNotSuccession: the pairs that never occur in this order hold not-succession with support 1
        found = {(a, b) for _, a, b, _ in recs}
        out["not-succession"] += [EventPairSupport(self.names[a], self.names[b], 1.0)
                                  for a in tot for b in tot if a != b and (a, b) not in found]
        inter = [(m, a, b, float(n) / (tot[a] if m == "r" else tot[b])) for m, a, b, n in recs]
        resp = [(a, b, s) for m, a, b, s in inter if m == "r" and s >= support]
        prec = [(a, b, s) for m, a, b, s in inter if m == "p" and s >= support]
        eps = lambda t: EventPairSupport(self.names[t[0]], self.names[t[1]], t[2])   # noqa: E731
        if prec and resp:
            out["response"] += [eps(t) for t in resp]
            out["precedence"] += [eps(t) for t in prec]
            pmap = {(a, b): s for a, b, s in prec}
            out["succession"] += [EventPairSupport(self.names[a], self.names[b], s * pmap[(a, b)])
                                  for a, b, s in resp if (a, b) in pmap]
            if mode != "alternate":   # QueryPlanOrderedRelationsAlternate.filterBasedOnSupport has no such branch
                low_r = {(a, b): s for m, a, b, s in inter if m == "r" and s <= (1 - support)}
                low_p = {(a, b): s for m, a, b, s in inter if m == "p" and s <= (1 - support)}
                out["not-succession"] += [EventPairSupport(self.names[a], self.names[b], (1 - s) * (1 - low_p[(a, b)]))
                                          for (a, b), s in low_r.items() if (a, b) in low_p]
        elif not prec:
            out["response"] += [eps(t) for t in resp]
        else:
            out["precedence"] += [eps(t) for t in prec]
        return out
