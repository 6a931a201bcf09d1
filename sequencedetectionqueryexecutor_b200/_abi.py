"""ctypes mirror of include/siesta_gpu.h (struct layouts and constants only)."""
import ctypes as C

MAX_STATES = 8
MAX_OR_TYPES = 8
MAX_PREDS = 4

STATE_NORMAL, STATE_KLEENE_PLUS, STATE_KLEENE_STAR, STATE_NEGATIVE, STATE_OR = range(5)
ATTR_POSITION, ATTR_TIMESTAMP = 0, 1
OP_LE, OP_GE = 0, 1

SYM_NORMAL, SYM_PLUS, SYM_STAR, SYM_NOT, SYM_OR = range(5)
CONSTRAINT_GAP, CONSTRAINT_TIME = 0, 1
METHOD_WITHIN, METHOD_ATLEAST = 0, 1
GRAN_SECONDS, GRAN_MINUTES, GRAN_HOURS = 0, 1, 2

F_RETURN_ALL = 1
F_ONLY_APPEARANCES = 2
F_MODE_HEAD = 4
F_EVT_POS = 8
F_NO_EVENT_COLUMNS = 16
F_COUNT_MATCHES = 32
F_LITERAL_RUNS = 64

E_INVALID, E_CUDA, E_UNSUPPORTED, E_NOMEM, E_REFERENCE_THROWS = -1, -2, -3, -4, -5


class Pred(C.Structure):
    _fields_ = [("attr", C.c_int32), ("op", C.c_int32), ("ref_state", C.c_int32),
                ("reserved", C.c_int32), ("constant", C.c_int64)]


class State(C.Structure):
    _fields_ = [("kind", C.c_int32), ("n_types", C.c_int32), ("types", C.c_int32 * MAX_OR_TYPES),
                ("n_preds", C.c_int32), ("reserved", C.c_int32), ("preds", Pred * MAX_PREDS)]


class Nfa(C.Structure):
    _fields_ = [("n_states", C.c_int32), ("reserved", C.c_int32), ("states", State * MAX_STATES)]


class EventSymbolC(C.Structure):
    _fields_ = [("activity", C.c_int32), ("position", C.c_int32), ("symbol", C.c_int32)]


class ConstraintC(C.Structure):
    _fields_ = [("pos_a", C.c_int32), ("pos_b", C.c_int32), ("kind", C.c_int32), ("method", C.c_int32),
                ("value", C.c_int64), ("granularity", C.c_int32), ("reserved", C.c_int32)]


class Matches(C.Structure):
    _fields_ = [("n_traces", C.c_int64), ("n_occurrences", C.c_int64), ("n_events", C.c_int64),
                ("n_matches_emitted", C.c_int64), ("n_ref_errors", C.c_int64),
                ("trace_idx", C.POINTER(C.c_int64)), ("occ_off", C.POINTER(C.c_int64)),
                ("ev_off", C.POINTER(C.c_int64)), ("ev_pos", C.POINTER(C.c_int32)),
                ("ev_rank", C.POINTER(C.c_int32)), ("ev_act", C.POINTER(C.c_int32)),
                ("ev_ts_ms", C.POINTER(C.c_int64)), ("err_trace_idx", C.POINTER(C.c_int64)),
                ("kernel_ms", C.c_double), ("detect_ms", C.c_double),
                ("n_unsupported", C.c_int64), ("unsupported_trace_idx", C.POINTER(C.c_int64))]


class WnmConstraint(C.Structure):
    """siesta_wnm_constraint"""
    _fields_ = [("pos_a", C.c_int32), ("pos_b", C.c_int32), ("kind", C.c_int32), ("method", C.c_int32), ("value", C.c_int64)]


class AlmostMatches(C.Structure):
    """siesta_almost_matches"""
    _fields_ = [("n_traces", C.c_int64), ("n_states", C.c_int32), ("trace_idx", C.POINTER(C.c_int64)),
                ("total_change", C.POINTER(C.c_int32)), ("ev_pos", C.POINTER(C.c_int32)), ("ev_value", C.POINTER(C.c_int32)),
                ("ev_change", C.POINTER(C.c_int32)), ("ev_stream_pos", C.POINTER(C.c_int32)), ("n_unsupported", C.c_int64),
                ("unsupported_trace_idx", C.POINTER(C.c_int64)), ("kernel_ms", C.c_double)]


WNM_GAP, WNM_TIME, WNM_WITHIN, WNM_ATLEAST = 0, 1, 0, 1


def make_wnm_constraints(constraints):
    """[(pos_a, pos_b, kind, method, value)] -> (ctypes array, n)"""
    arr = (WnmConstraint * max(len(constraints), 1))()
    for i, (a, b, kind, method, value) in enumerate(constraints):
        arr[i] = WnmConstraint(int(a), int(b), int(kind), int(method), int(value))
    return arr, len(constraints)


class AlmostMatchResult:
    """Host copy of siesta_almost_matches: per trace with an almost-match its m uncertain events."""

    def __init__(self, st):
        import numpy as np
        n, m = int(st.n_traces), int(st.n_states)
        self.n_traces, self.n_states = n, m
        self.kernel_ms = float(st.kernel_ms)

        def col(ptr, k, dt):
            return np.ctypeslib.as_array(ptr, shape=(max(k, 1),))[:k].astype(dt, copy=True) if k else np.zeros(0, dtype=dt)
        self.trace_idx = col(st.trace_idx, n, np.int64)
        self.total_change = col(st.total_change, n, np.int32)
        self.ev_pos = col(st.ev_pos, n * m, np.int32).reshape(n, m)
        self.ev_value = col(st.ev_value, n * m, np.int32).reshape(n, m)
        self.ev_change = col(st.ev_change, n * m, np.int32).reshape(n, m)
        self.ev_stream_pos = col(st.ev_stream_pos, n * m, np.int32).reshape(n, m)
        self.unsupported_trace_idx = col(st.unsupported_trace_idx, int(st.n_unsupported), np.int64)

    def same_as(self, other):
        import numpy as np
        for k in ("trace_idx", "total_change", "ev_pos", "ev_value", "ev_change", "ev_stream_pos"):
            if not np.array_equal(getattr(self, k), getattr(other, k)):
                return False, k
        return True, ""


class PairCount(C.Structure):
    _fields_ = [("count", C.c_int64), ("sum_duration_ms", C.c_int64), ("min_duration_ms", C.c_int64),
                ("max_duration_ms", C.c_int64), ("sum_squares_lo", C.c_uint64), ("sum_squares_hi", C.c_uint64)]


class DevMatches(C.Structure):
    _fields_ = [("n_traces", C.c_int64), ("n_occurrences", C.c_int64), ("n_events", C.c_int64),
                ("n_matches_emitted", C.c_int64), ("n_ref_errors", C.c_int64),
                ("d_trace_idx", C.c_void_p), ("d_occ_off", C.c_void_p), ("d_ev_off", C.c_void_p),
                ("d_ev_pos", C.c_void_p), ("d_ev_rank", C.c_void_p), ("d_ev_act", C.c_void_p),
                ("d_ev_ts_ms", C.c_void_p), ("d_err_trace_idx", C.c_void_p),
                ("kernel_ms", C.c_double), ("detect_ms", C.c_double), ("d_block", C.c_void_p), ("block_bytes", C.c_int64),
                ("impl", C.c_void_p), ("n_unsupported", C.c_int64), ("d_unsupported_trace_idx", C.c_void_p)]


class ExchangeStats(C.Structure):
    _fields_ = [("local_traces", C.c_int64), ("local_occurrences", C.c_int64), ("local_events", C.c_int64),
                ("pulled_bytes", C.c_int64), ("k1_ms", C.c_double), ("scan_ms", C.c_double), ("wait_ms", C.c_double), ("pull_ms", C.c_double), ("host_gap_ms", C.c_double),
                ("join_ms", C.c_double), ("n_blocks", C.c_int32), ("eager", C.c_int32)]


EXCHANGE_HANDLE_BYTES = 64
REDUCE_SUM, REDUCE_MIN, REDUCE_MAX = 0, 1, 2


def make_nfa(states):
    """states: list of dicts {kind, types:[...], preds:[(attr, op, ref_state, constant), ...]}"""
    if not 1 <= len(states) <= MAX_STATES:
        raise ValueError(f"NFA must have 1..{MAX_STATES} states")
    nfa = Nfa()
    nfa.n_states = len(states)
    for i, s in enumerate(states):
        st = nfa.states[i]
        st.kind = s["kind"]
        types = list(s["types"])
        if not 1 <= len(types) <= MAX_OR_TYPES:
            raise ValueError("bad type list")
        st.n_types = len(types)
        for k, t in enumerate(types):
            st.types[k] = t
        preds = list(s.get("preds", ()))
        if len(preds) > MAX_PREDS:
            raise ValueError("too many predicates on one state")
        st.n_preds = len(preds)
        for k, (attr, op, ref, const) in enumerate(preds):
            st.preds[k].attr, st.preds[k].op, st.preds[k].ref_state, st.preds[k].constant = attr, op, ref, const
    return nfa


def _arr(ptr, n, dtype, copy=True):
    import numpy as np
    if n == 0 or not ptr:
        return np.zeros(0, dtype=dtype)
    a = np.ctypeslib.as_array(ptr, shape=(n,))
    return a.astype(dtype, copy=True) if copy else a


class MatchResult:
    """Host copy of a siesta_matches (CSR of selected occurrences per matching trace)."""

    __slots__ = ("n_traces", "n_occurrences", "n_events", "n_matches_emitted", "n_ref_errors", "trace_idx",
                 "occ_off", "ev_off", "ev_pos", "ev_rank", "ev_act", "ev_ts_ms", "err_trace_idx", "kernel_ms", "detect_ms",
                 "n_unsupported", "unsupported_trace_idx", "_free")

    def close(self):
        """Release the library-owned block behind zero-copy views (from_struct(copy=False)); the arrays die with it."""
        f = getattr(self, "_free", None)
        if f is not None:
            self._free = None
            for k in ("trace_idx", "occ_off", "ev_off", "ev_pos", "ev_rank", "ev_act", "ev_ts_ms", "err_trace_idx"):
                setattr(self, k, None)
            f()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @classmethod
    def from_struct(cls, m, copy=True, free=None):
        """copy=False: the arrays are views of the library's (pinned) result block, which `free` releases on close()."""
        import numpy as np
        r = cls()
        r._free = None if copy else free
        r.n_traces, r.n_occurrences, r.n_events = m.n_traces, m.n_occurrences, m.n_events
        r.n_matches_emitted, r.n_ref_errors, r.kernel_ms = m.n_matches_emitted, m.n_ref_errors, m.kernel_ms
        r.detect_ms = m.detect_ms
        r.trace_idx = _arr(m.trace_idx, m.n_traces, np.int64, copy)
        r.occ_off = _arr(m.occ_off, m.n_traces + 1, np.int64, copy)
        r.ev_off = _arr(m.ev_off, m.n_occurrences + 1, np.int64, copy)
        r.ev_pos = _arr(m.ev_pos, m.n_events, np.int32, copy) if m.ev_pos else None
        r.ev_rank = _arr(m.ev_rank, m.n_events, np.int32, copy) if m.ev_rank else None
        r.ev_act = _arr(m.ev_act, m.n_events, np.int32, copy) if m.ev_act else None
        r.ev_ts_ms = _arr(m.ev_ts_ms, m.n_events, np.int64, copy) if m.ev_ts_ms else None
        r.err_trace_idx = _arr(m.err_trace_idx, m.n_ref_errors, np.int64, copy)
        r.n_unsupported = getattr(m, "n_unsupported", 0)
        r.unsupported_trace_idx = _arr(m.unsupported_trace_idx, r.n_unsupported, np.int64, True) if r.n_unsupported else np.zeros(0, dtype=np.int64)
        return r

    def occurrences_of(self, i):
        """List of occurrences (each a list of in-trace positions) of the i-th matching trace."""
        out = []
        for o in range(self.occ_off[i], self.occ_off[i + 1]):
            out.append(self.ev_pos[self.ev_off[o]:self.ev_off[o + 1]].tolist())
        return out

    def as_dict(self):
        return {int(t): self.occurrences_of(i) for i, t in enumerate(self.trace_idx)}

    def same_as(self, other, columns=("ev_pos", "ev_rank", "ev_act", "ev_ts_ms")):
        import numpy as np
        keys = ["trace_idx", "occ_off", "ev_off", "err_trace_idx"] + list(columns)
        for k in ("n_traces", "n_occurrences", "n_events", "n_ref_errors"):
            if getattr(self, k) != getattr(other, k):
                return False, k
        if self.n_matches_emitted >= 0 and other.n_matches_emitted >= 0 and self.n_matches_emitted != other.n_matches_emitted:
            return False, "n_matches_emitted"
        for k in keys:
            a, b = getattr(self, k), getattr(other, k)
            if a is None or b is None:
                continue
            if a.shape != b.shape or not np.array_equal(a, b):
                return False, k
        return True, ""


class DeclareCounts:
    """Named views over the packed int64 result of siesta_declare_counts (layout: include/siesta_gpu.h)."""

    def __init__(self, packed, n_activities, k_cap, kernel_ms=0.0):
        import numpy as np
        A, K = n_activities, k_cap + 1
        self.packed = np.asarray(packed, dtype=np.int64)
        self.n_activities, self.k_cap, self.kernel_ms = A, k_cap, kernel_ms
        o = 0
        self.tot = self.packed[o:o + A]; o += A
        self.uniq = self.packed[o:o + A]; o += A
        self.first = self.packed[o:o + A]; o += A
        self.last = self.packed[o:o + A]; o += A
        self.hist = self.packed[o:o + A * K].reshape(A, K); o += A * K
        self.co = self.packed[o:o + A * A].reshape(A, A); o += A * A
        self.ordered = self.packed[o:o + A * A].reshape(A, A); o += A * A
        self.response = self.packed[o:o + A * A].reshape(A, A); o += A * A
        self.precedence = self.packed[o:o + A * A].reshape(A, A); o += A * A
        self.alt_response = self.packed[o:o + A * A].reshape(A, A); o += A * A
        self.alt_precedence = self.packed[o:o + A * A].reshape(A, A); o += A * A
        self.chain_response = self.packed[o:o + A * A].reshape(A, A); o += A * A
        self.chain_precedence = self.packed[o:o + A * A].reshape(A, A); o += A * A
        self.hist_overflow = int(self.packed[o])
        self.n_nonempty = int(self.packed[o + 1])
