"""CPU ORACLE — test infrastructure, not the product.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this package.  It wraps oracle/_build/liboracle.so, a literal C++ restatement of the
reference's verification engine (oracle/sase_oracle.cpp) and counting paths
(oracle/counting_oracle.cpp).  Pinned against the reference's own known-answer tests by
tests/test_oracle_kat.py.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from sequencedetectionqueryexecutor_b200 import _abi

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "liboracle.so")
_lib = None


def build(force=False):
    srcs = [os.path.join(_HERE, f) for f in ("sase_oracle.cpp", "counting_oracle.cpp", "wnm_oracle.cpp", "Makefile")]
    srcs.append(os.path.join(_HERE, "..", "include", "siesta_gpu.h"))
    stale = (not os.path.exists(_LIB_PATH)) or any(
        os.path.exists(s) and os.path.getmtime(s) > os.path.getmtime(_LIB_PATH) for s in srcs)
    if force or stale:
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        try:
            build()
        except Exception:
            if not os.path.exists(_LIB_PATH):
                raise
        _lib = C.CDLL(_LIB_PATH)
        _lib.oracle_run_stream.restype = C.c_int
        _lib.oracle_detect.restype = C.c_int
        _lib.oracle_free.argtypes = [C.c_void_p]
        _lib.oracle_matches_free.argtypes = [C.POINTER(_abi.Matches)]
    return _lib


def _p(a, ct):
    return a.ctypes.data_as(C.POINTER(ct))


def run_stream(nfa, types, ids, ts, flags=0):
    """Engine-level: one explicit stream -> (status, [match event-id lists] in emission order)."""
    L = lib()
    types = np.ascontiguousarray(types, dtype=np.int32)
    ids = np.ascontiguousarray(ids, dtype=np.int32)
    ts = np.ascontiguousarray(ts, dtype=np.int32)
    status, n = C.c_int32(0), C.c_int64(0)
    off, mid = C.POINTER(C.c_int64)(), C.POINTER(C.c_int32)()
    rc = L.oracle_run_stream(C.byref(nfa), _p(types, C.c_int32), _p(ids, C.c_int32), _p(ts, C.c_int32),
                             C.c_int32(len(types)), C.c_uint32(flags), C.byref(status), C.byref(n),
                             C.byref(off), C.byref(mid))
    assert rc == 0
    offs = [off[i] for i in range(n.value + 1)]
    matches = [[mid[j] for j in range(offs[i], offs[i + 1])] for i in range(n.value)]
    L.oracle_free(off)
    L.oracle_free(mid)
    return status.value, matches


def detect(trace_off, act, ts_ms, nfa, cand=None, flags=0, n_threads=1):
    """SaseConnector.evaluate + clearOccurrences over a CSR log -> _abi.MatchResult."""
    L = lib()
    trace_off = np.ascontiguousarray(trace_off, dtype=np.int64)
    act = np.ascontiguousarray(act, dtype=np.int32)
    ts_ms = np.ascontiguousarray(ts_ms, dtype=np.int64)
    T = len(trace_off) - 1
    if cand is not None:
        cand = np.ascontiguousarray(cand, dtype=np.int64)
        cp, nc = _p(cand, C.c_int64), len(cand)
    else:
        cp, nc = None, 0
    out = C.POINTER(_abi.Matches)()
    rc = L.oracle_detect(_p(trace_off, C.c_int64), _p(act, C.c_int32), _p(ts_ms, C.c_int64), C.c_int64(T),
                         C.byref(nfa), cp, C.c_int64(nc), C.c_uint32(flags), C.c_int32(n_threads), C.byref(out))
    if rc != 0:
        raise ValueError(f"oracle_detect failed: {rc}")
    res = _abi.MatchResult.from_struct(out.contents)
    L.oracle_matches_free(out)
    return res


def declare_counts(trace_off, act, n_activities, k_cap=64):
    """Literal restatement of the declare counting jobs -> _abi.DeclareCounts."""
    L = lib()
    trace_off = np.ascontiguousarray(trace_off, dtype=np.int64)
    act = np.ascontiguousarray(act, dtype=np.int32)
    L.oracle_declare_size.restype = C.c_int64
    n = L.oracle_declare_size(C.c_int32(n_activities), C.c_int32(k_cap))
    out = np.zeros(n, dtype=np.int64)
    L.oracle_declare_counts(_p(trace_off, C.c_int64), _p(act, C.c_int32), C.c_int64(len(trace_off) - 1),
                            C.c_int32(n_activities), C.c_int32(k_cap), _p(out, C.c_int64))
    return _abi.DeclareCounts(out, n_activities, k_cap)


def pair_stats(trace_off, act, ts_ms, pairs):
    """Pair statistics under the stated pairing policy (parity unpinned, see counting_oracle.cpp) ->
    list of dicts count / sum / min / max / sum_squares (python int, exact)."""
    L = lib()
    trace_off = np.ascontiguousarray(trace_off, dtype=np.int64)
    act = np.ascontiguousarray(act, dtype=np.int32)
    ts_ms = np.ascontiguousarray(ts_ms, dtype=np.int64)
    pa = np.array([p[0] for p in pairs], dtype=np.int32)
    pb = np.array([p[1] for p in pairs], dtype=np.int32)
    out = np.zeros(6 * len(pairs), dtype=np.int64)
    L.oracle_pair_stats(_p(trace_off, C.c_int64), _p(act, C.c_int32), _p(ts_ms, C.c_int64), C.c_int64(len(trace_off) - 1),
                        _p(pa, C.c_int32), _p(pb, C.c_int32), C.c_int32(len(pairs)), _p(out, C.c_int64))
    res = []
    for p in range(len(pairs)):
        o = out[6 * p:6 * p + 6]
        res.append({"count": int(o[0]), "sum": int(o[1]), "min": int(o[2]), "max": int(o[3]),
                    "sum_squares": (int(o[4]) & (2 ** 64 - 1)) | ((int(o[5]) & (2 ** 64 - 1)) << 64)})
    return res


def posting_list(trace_off, act, a, b):
    L = lib()
    trace_off = np.ascontiguousarray(trace_off, dtype=np.int64)
    act = np.ascontiguousarray(act, dtype=np.int32)
    out = np.zeros(len(trace_off) - 1, dtype=np.int64)
    L.oracle_posting_list.restype = C.c_int64
    n = L.oracle_posting_list(_p(trace_off, C.c_int64), _p(act, C.c_int32), C.c_int64(len(trace_off) - 1), C.c_int32(a),
                              C.c_int32(b), _p(out, C.c_int64))
    return out[:n].copy()


def intersect(lists):
    L = lib()
    lists = [np.ascontiguousarray(x, dtype=np.int64) for x in lists]
    if not lists:
        return np.zeros(0, dtype=np.int64)
    ptrs = (C.POINTER(C.c_int64) * len(lists))(*[_p(x, C.c_int64) for x in lists])
    lens = np.array([len(x) for x in lists], dtype=np.int64)
    out = np.zeros(min(len(x) for x in lists), dtype=np.int64)
    L.oracle_intersect.restype = C.c_int64
    n = L.oracle_intersect(ptrs, _p(lens, C.c_int64), C.c_int32(len(lists)), _p(out, C.c_int64))
    return out[:n].copy()


def wnm_stream_size(primary, uncertainty, step):
    """WhyNotMatchSASE.getUnCertainStream(...).getSize() for a list of primary metrics (oracle/wnm_oracle.cpp)."""
    L = lib()
    L.oracle_wnm_stream_size.restype = C.c_int64
    primary = np.ascontiguousarray(primary, dtype=np.int64)
    return int(L.oracle_wnm_stream_size(_p(primary, C.c_int64), C.c_int32(len(primary)), C.c_int32(uncertainty), C.c_int32(step)))


def why_not_match(trace_off, act, ts_ms, pattern, constraints, uncertainty, step, k, cand=None, flags=0, run_limit=2_000_000):
    """WhyNotMatchSASE.evaluate + createResponse, literally (exponential: small cases) -> _abi.AlmostMatchResult,
    or None when a trace needs more than run_limit runs."""
    L = lib()
    trace_off = np.ascontiguousarray(trace_off, dtype=np.int64)
    act = np.ascontiguousarray(act, dtype=np.int32)
    ts_ms = np.ascontiguousarray(ts_ms, dtype=np.int64)
    pattern = np.ascontiguousarray(pattern, dtype=np.int32)
    cons, n_cons = _abi.make_wnm_constraints(constraints)
    if cand is not None:
        cand = np.ascontiguousarray(cand, dtype=np.int64)
        cp, nc = _p(cand, C.c_int64), len(cand)
    else:
        cp, nc = None, 0
    out = C.POINTER(_abi.AlmostMatches)()
    L.oracle_why_not_match.restype = C.c_int
    L.oracle_almost_matches_free.argtypes = [C.POINTER(_abi.AlmostMatches)]
    rc = L.oracle_why_not_match(_p(trace_off, C.c_int64), _p(act, C.c_int32), _p(ts_ms, C.c_int64), C.c_int64(len(trace_off) - 1),
                                _p(pattern, C.c_int32), C.c_int32(len(pattern)), cons, C.c_int32(n_cons), C.c_int32(uncertainty),
                                C.c_int32(step), C.c_int32(k), cp, C.c_int64(nc), C.c_uint32(flags), C.c_int64(run_limit), C.byref(out))
    if rc == 1:
        return None
    assert rc == 0, rc
    res = _abi.AlmostMatchResult(out.contents)
    L.oracle_almost_matches_free(out)
    return res
