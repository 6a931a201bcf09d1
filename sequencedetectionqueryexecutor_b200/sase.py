"""Host-side mirror of the reference's interface for the verification path.

Names, argument meaning and error behaviour follow the Java classes so the parity tests read like the
reference's own (paths relative to src/main/java/com/datalab/siesta/queryprocessor/):
  EventSymbol            model/Events/EventSymbol.java
  GapConstraint / TimeConstraint   model/Constraints/*.java
  ComplexPattern         model/Patterns/ComplexPattern.java  (getNfa / getNfaWithoutConstraints)
  Occurrence(s)          model/Occurrence.java, model/Occurrences.java
  SaseConnector          SaseConnection/SaseConnector.java   (evaluate)
All compute happens in libsiesta_gpu (CUDA); nothing here evaluates patterns on the CPU.
"""
import ctypes as C
from dataclasses import dataclass, field
from typing import Dict, List, Optional

import numpy as np

from . import _abi
from ._lib import check, lib

_SYM = {"_": _abi.SYM_NORMAL, "": _abi.SYM_NORMAL, "+": _abi.SYM_PLUS, "*": _abi.SYM_STAR, "!": _abi.SYM_NOT,
        "||": _abi.SYM_OR}
_GRAN = {"seconds": _abi.GRAN_SECONDS, "minutes": _abi.GRAN_MINUTES, "hours": _abi.GRAN_HOURS}


class ActivityDictionary:
    """Activity name <-> dense id.  Names are folded case-insensitively because the engine compares event
    types with equalsIgnoreCase (edu/umass/cs/sase/query/State.java:135-137)."""

    def __init__(self, names=()):
        self._ids: Dict[str, int] = {}
        self.names: List[str] = []
        for n in names:
            self.add(n)

    def add(self, name):
        k = name.casefold()
        if k not in self._ids:
            self._ids[k] = len(self.names)
            self.names.append(name)
        return self._ids[k]

    def id(self, name):
        """-1 for a name the log has never seen (it can then never match an event)."""
        return self._ids.get(name.casefold(), -1)

    def __len__(self):
        return len(self.names)


@dataclass
class EventSymbol:
    name: str
    position: int
    symbol: str = "_"


@dataclass
class Constraint:
    posA: int
    posB: int
    constraint: int
    method: str = "within"  # "within" | "atleast"


@dataclass
class GapConstraint(Constraint):
    pass


@dataclass
class TimeConstraint(Constraint):
    granularity: str = "seconds"


@dataclass
class ComplexPattern:
    eventsWithSymbols: List[EventSymbol] = field(default_factory=list)
    constraints: List[Constraint] = field(default_factory=list)

    def getEventTypes(self):
        return {e.name for e in self.eventsWithSymbols}

    def _compile(self, activities: ActivityDictionary, only_appearances):
        n = len(self.eventsWithSymbols)
        syms = (_abi.EventSymbolC * max(n, 1))()
        for i, e in enumerate(self.eventsWithSymbols):
            if e.symbol not in _SYM:
                raise ValueError(f"unknown symbol {e.symbol!r}")
            syms[i].activity, syms[i].position, syms[i].symbol = activities.id(e.name), e.position, _SYM[e.symbol]
        nc = len(self.constraints)
        cs = (_abi.ConstraintC * max(nc, 1))()
        for i, c in enumerate(self.constraints):
            cs[i].pos_a, cs[i].pos_b, cs[i].value = c.posA, c.posB, c.constraint
            cs[i].kind = _abi.CONSTRAINT_TIME if isinstance(c, TimeConstraint) else _abi.CONSTRAINT_GAP
            # SIESTAPattern.generatePredicates: "within" -> <=, anything else -> >= (SIESTAPattern.java:136-145)
            cs[i].method = _abi.METHOD_WITHIN if c.method == "within" else _abi.METHOD_ATLEAST
            cs[i].granularity = _GRAN.get(getattr(c, "granularity", "seconds"), _abi.GRAN_SECONDS)
        nfa = _abi.Nfa()
        check(lib().siesta_pattern_compile(syms, n, cs, nc, 1 if only_appearances else 0, C.byref(nfa)))
        return nfa

    def _c_arrays(self, activities):
        n = len(self.eventsWithSymbols)
        syms = (_abi.EventSymbolC * max(n, 1))()
        for i, e in enumerate(self.eventsWithSymbols):
            if e.symbol not in _SYM:
                raise ValueError(f"unknown symbol {e.symbol!r}")
            syms[i].activity, syms[i].position, syms[i].symbol = activities.id(e.name), e.position, _SYM[e.symbol]
        nc = len(self.constraints)
        cs = (_abi.ConstraintC * max(nc, 1))()
        for i, c in enumerate(self.constraints):
            cs[i].pos_a, cs[i].pos_b, cs[i].value = c.posA, c.posB, c.constraint
            cs[i].kind = _abi.CONSTRAINT_TIME if isinstance(c, TimeConstraint) else _abi.CONSTRAINT_GAP
            cs[i].method = _abi.METHOD_WITHIN if c.method == "within" else _abi.METHOD_ATLEAST
            cs[i].granularity = _GRAN.get(getattr(c, "granularity", "seconds"), _abi.GRAN_SECONDS)
        return syms, n, cs, nc

    def extractPairsForPatternDetection(self, activities, fromOrTillSet=False):
        """List of ExtractedPairsForPatternDetection, one per OR-expansion (ComplexPattern.java:75-128)."""
        syms, n, cs, nc = self._c_arrays(activities)
        cap_x, cap_p = 256, 4096
        n_x = C.c_int32(0)
        t_off, a_off = (C.c_int32 * (cap_x + 1))(), (C.c_int32 * (cap_x + 1))()
        ta, tb, aa, ab = ((C.c_int32 * cap_p)() for _ in range(4))
        check(lib().siesta_pattern_extract_pairs(syms, n, cs, nc, 1 if fromOrTillSet else 0, cap_x, cap_p, C.byref(n_x),
                                                 t_off, ta, tb, a_off, aa, ab))
        out = []
        for x in range(n_x.value):
            out.append(ExtractedPairsForPatternDetection(
                truePairs=[(ta[k], tb[k]) for k in range(t_off[x], t_off[x + 1])],
                allPairs=[(aa[k], ab[k]) for k in range(a_off[x], a_off[x + 1])]))
        return out

    def getNfa(self, activities):
        return self._compile(activities, False)

    def getNfaWithoutConstraints(self, activities):
        return self._compile(activities, True)


@dataclass
class ExtractedPairsForPatternDetection:
    """model/ExtractedPairsForPatternDetection.java: pairs as (activity id, activity id), sorted."""
    truePairs: List[tuple]
    allPairs: List[tuple]


def pattern_candidates(pattern, log, activities, fromOrTillSet=False):
    """DatabaseRepository.patterDetectionTraceIds under the SeqTable view (SparkDatabaseRepository.java:243-253 ->
    getCommonIds :160-178, merged over the OR-expansions as QueryPlanPatternDetection.getMiddleResults :146-164 does):
    ascending indices of the traces that hold every true pair of at least one expansion.  None when the pattern has
    no true pair (the reference then takes the Single plan and verifies every trace holding the pattern's types)."""
    exps = pattern.extractPairsForPatternDetection(activities, fromOrTillSet)
    if any(len(x.truePairs) == 0 for x in exps):
        return None
    pairs = sorted({p for x in exps for p in x.truePairs})
    if any(a < 0 or b < 0 for a, b in pairs):
        return np.zeros(0, dtype=np.int64)  # a pattern activity the log has never seen: no trace can hold the pair
    idx = log.build_index(pairs)
    try:
        pid = {p: i for i, p in enumerate(pairs)}
        return idx.candidates([[pid[p] for p in x.truePairs] for x in exps])
    finally:
        idx.close()


@dataclass
class EventBoth:
    name: str
    position: int
    timestamp_ms: Optional[int]


@dataclass
class Occurrence:
    occurrence: List[EventBoth]


@dataclass
class Occurrences:
    traceID: object
    occurrences: List[Occurrence]


class SaseConnector:
    """evaluate(pattern, log, onlyAppearances) -> List[Occurrences], already passed through
    Occurrences.clearOccurrences(returnAll) like QueryPlanPatternDetection.execute does (:121-122)."""

    def __init__(self, activities: ActivityDictionary, trace_ids=None, positions_mode=False):
        self.activities = activities
        self.trace_ids = trace_ids
        self.positions_mode = positions_mode  # events are EventPos (metadata mode "positions") instead of EventTs

    def flags(self, only_appearances=False, return_all=False):
        f = 0
        if only_appearances:
            f |= _abi.F_ONLY_APPEARANCES
        if return_all:
            f |= _abi.F_RETURN_ALL
        if self.positions_mode:
            f |= _abi.F_EVT_POS
        return f

    def evaluate_raw(self, pattern, log, only_appearances=False, return_all=False, cand=None):
        nfa = pattern.getNfaWithoutConstraints(self.activities) if only_appearances else pattern.getNfa(self.activities)
        res = log.detect(nfa, cand=cand, flags=self.flags(only_appearances, return_all))
        if res.n_ref_errors:
            # SaseConnector.java:60-62 rethrows engine exceptions as RuntimeException: the whole request fails
            raise RuntimeError(f"the reference engine throws on {res.n_ref_errors} trace(s), first: "
                               f"{int(res.err_trace_idx[0])}")
        return res

    def evaluate(self, pattern, log, onlyAppearances=False, returnAll=False, cand=None):
        res = self.evaluate_raw(pattern, log, onlyAppearances, returnAll, cand)
        out = []
        names = self.activities.names
        for i, t in enumerate(res.trace_idx):
            occs = []
            for o in range(res.occ_off[i], res.occ_off[i + 1]):
                evs = []
                for e in range(res.ev_off[o], res.ev_off[o + 1]):
                    pos = int(res.ev_pos[e]) if self.positions_mode else int(res.ev_rank[e])
                    ts = None if self.positions_mode else int(res.ev_ts_ms[e])
                    evs.append(EventBoth(names[res.ev_act[e]], pos, ts))
                occs.append(Occurrence(evs))
            tid = self.trace_ids[t] if self.trace_ids is not None else int(t)
            out.append(Occurrences(tid, occs))
        return out


@dataclass
class Proposition:
    """model/Proposition.java: a possible continuation, its exact completions and the average duration (seconds)."""
    event: str
    completions: int
    averageDuration: float

    def score(self):
        # Proposition.compareTo (:51-60): completions / averageDuration (IEEE: x / 0.0 = inf, as in Java)
        return self.completions / self.averageDuration if self.averageDuration != 0 else float("inf")


def explore_accurate(pattern_names, log, activities: ActivityDictionary, candidates=None, positions_mode=False):
    """QueryPlanExplorationAccurate.execute (model/Queries/QueryPlans/Exploration/QueryPlanExplorationAccurate.java:
    54-72): for every possible next event after the pattern's last event, the exact number of occurrences of the
    extended pattern and their average duration, sorted best first (Collections.reverseOrder over compareTo).
    `candidates` defaults to the activities that follow the last event in some trace (the CountTable rows with
    eventA = last event)."""
    ids = [activities.id(n) for n in pattern_names]
    if candidates is None:
        d = log.declare_counts()
        candidates = [b for b in range(len(activities)) if ids[-1] >= 0 and d.ordered[ids[-1], b] > 0]
    else:
        candidates = [activities.id(c) if isinstance(c, str) else int(c) for c in candidates]
    comp, dur, _ = log.explore_accurate(ids, candidates, _abi.F_EVT_POS if positions_mode else 0)
    props = [Proposition(activities.names[c], int(n), (int(ms) / 1000.0) / int(n))
             for c, n, ms in zip(candidates, comp, dur) if n > 0]
    # reverse order of compareTo: higher score first; equal scores: compareTo falls back to t.event.compareTo(this.event),
    # reversed once more by reverseOrder -> ascending event name
    props.sort(key=lambda p: (-p.score(), p.event))
    return props


# ------------------------------------------------------------------------------------------------ why-not-match
@dataclass
class UncertainTimeEvent:
    """J/model/WhyNotMatch/UsingSase/UncertainTimeEvent.java: an event of the uncertain stream as the response carries it."""
    event_type: str
    timestamp: int      # shifted primary metric: epoch seconds (or the position in positions mode)
    change: int
    position: int       # index in the uncertain stream
    source_position: int = -1   # index of the original event inside its trace (not a field of the Java class)


@dataclass
class AlmostMatch:
    """J/model/WhyNotMatch/AlmostMatch.java"""
    trace_id: object
    match: List[UncertainTimeEvent]
    totalChange: int


def _wnm_constraints(constraints):
    """WhyNotMatchSASE.generatePredicatesFromConstraints (:136-158): gap constraints in positions, time constraints in seconds"""
    mult = {"seconds": 1, "minutes": 60, "hours": 3600}
    out = []
    for c in constraints:
        if isinstance(c, TimeConstraint):
            out.append((c.posA, c.posB, _abi.WNM_TIME, _abi.WNM_WITHIN if c.method == "within" else _abi.WNM_ATLEAST,
                        c.constraint * mult.get(c.granularity, 1)))   # TimeConstraint.getConstraintInSeconds :45-49
        else:
            out.append((c.posA, c.posB, _abi.WNM_GAP, _abi.WNM_WITHIN if c.method == "within" else _abi.WNM_ATLEAST, c.constraint))
    return out


class WhyNotMatchSASE:
    """evaluate(simple pattern, rest traces, uncertaintyPerEvent, step, k) -> List[AlmostMatch]
    (WhyNotMatchSASE.java:37-55) on the GPU (siesta_why_not_match)."""

    def __init__(self, activities: ActivityDictionary, trace_ids=None, positions_mode=False):
        self.activities = activities
        self.trace_ids = trace_ids
        self.positions_mode = positions_mode

    def evaluate_raw(self, event_names, constraints, log, rest, uncertaintyPerEvent, step, k):
        pattern = [self.activities.id(n) for n in event_names]
        return log.why_not_match(pattern, _wnm_constraints(constraints), uncertaintyPerEvent, step, k, cand=rest,
                                 flags=_abi.F_EVT_POS if self.positions_mode else 0)

    def evaluate(self, event_names, constraints, log, rest, uncertaintyPerEvent, step, k):
        res = self.evaluate_raw(event_names, constraints, log, rest, uncertaintyPerEvent, step, k)
        out = []
        for i, t in enumerate(res.trace_idx):
            evs = [UncertainTimeEvent(event_names[j], int(res.ev_value[i, j]), int(res.ev_change[i, j]), int(res.ev_stream_pos[i, j]),
                                      int(res.ev_pos[i, j])) for j in range(res.n_states)]
            out.append(AlmostMatch(self.trace_ids[t] if self.trace_ids is not None else int(t), evs, int(res.total_change[i])))
        return out


def why_not_match_plan(pattern: "ComplexPattern", log, activities: ActivityDictionary, uncertainty, step, k, returnAll=False,
                       cand=None, trace_ids=None, positions_mode=False):
    """QueryPlanWhyNotMatch.execute (:54-100) after the pruning: the true occurrences of the candidates, then the
    why-not-match search over the candidates WITHOUT an occurrence (:78-89).  Only simple patterns (every symbol "_", :66-70).
    -> (List[Occurrences], List[AlmostMatch])"""
    if any(e.symbol != "_" for e in pattern.eventsWithSymbols):
        raise ValueError("why-not-match takes a simple pattern (QueryResponseBadRequestWhyNotMatch.setSimple(false))")
    conn = SaseConnector(activities, trace_ids, positions_mode)
    raw = conn.evaluate_raw(pattern, log, False, returnAll, cand)
    occurrences = conn.evaluate(pattern, log, False, returnAll, cand)
    import numpy as np
    universe = np.arange(log.n_traces, dtype=np.int64) if cand is None else np.asarray(cand, dtype=np.int64)
    rest = np.setdiff1d(universe, raw.trace_idx, assume_unique=False)
    names = [e.name for e in sorted(pattern.eventsWithSymbols, key=lambda e: e.position)]
    almost = WhyNotMatchSASE(activities, trace_ids, positions_mode).evaluate(names, pattern.constraints, log, rest, uncertainty, step, k)
    return occurrences, almost
