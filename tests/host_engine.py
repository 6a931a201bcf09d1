"""ctypes wrapper of tests/host_harness (the device engine compiled for the host; test-only)."""
import ctypes as C
import os
import subprocess

import numpy as np

from sequencedetectionqueryexecutor_b200 import _abi

_HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "host_harness")
_LIB = os.path.join(_HERE, "_build", "libengine_host.so")
_lib = None


def lib():
    global _lib
    if _lib is None:
        subprocess.check_call(["make", "-C", _HERE, "-s"])
        _lib = C.CDLL(_LIB)
        _lib.engine_host_detect.restype = C.c_int
        _lib.engine_host_last_error.restype = C.c_char_p
    return _lib


FAST_NONE, FAST_NK, FAST_FK2, FAST_NP1 = 0, 1, 2, 3   # detect_fast.cuh


def fast_class(nfa, flags=0):
    """The evaluator the product's validate_nfa picks for this NFA and these flags (FAST_*), or the negative error code."""
    L = lib()
    L.engine_host_fast_class.restype = C.c_int
    return int(L.engine_host_fast_class(C.byref(nfa), C.c_uint32(flags)))


def _p(a, ct):
    return a.ctypes.data_as(C.POINTER(ct))


def detect(trace_off, act, ts_ms, n_act, nfa, cand=None, flags=0):
    import oracle
    L = lib()
    trace_off = np.ascontiguousarray(trace_off, dtype=np.int64)
    act = np.ascontiguousarray(act, dtype=np.int32)
    ts_ms = np.ascontiguousarray(ts_ms, dtype=np.int64)
    if cand is not None:
        cand = np.ascontiguousarray(cand, dtype=np.int64)
        cp, nc = _p(cand, C.c_int64), len(cand)
    else:
        cp, nc = None, 0
    out = C.POINTER(_abi.Matches)()
    n_wide = C.c_int64(0)
    rc = L.engine_host_detect(_p(trace_off, C.c_int64), _p(act, C.c_int32), _p(ts_ms, C.c_int64),
                              C.c_int64(len(trace_off) - 1), C.c_int32(n_act), C.byref(nfa), cp, C.c_int64(nc),
                              C.c_uint32(flags), C.byref(n_wide), C.byref(out))
    if rc != 0:
        return rc, None, 0
    res = _abi.MatchResult.from_struct(out.contents)
    oracle.lib().oracle_matches_free(out)
    return 0, res, n_wide.value


def wnm_eval(trace_off, act, ts_ms, pattern, constraints, uncertainty, step, k, flags=0):
    """Kernel W's per-trace evaluation (csrc/wnm.cuh) on the host -> _abi.AlmostMatchResult-like object."""
    L = lib()
    trace_off = np.ascontiguousarray(trace_off, dtype=np.int64)
    act = np.ascontiguousarray(act, dtype=np.int32)
    ts_ms = np.ascontiguousarray(ts_ms, dtype=np.int64)
    pattern = np.ascontiguousarray(pattern, dtype=np.int32)
    cons, n_cons = _abi.make_wnm_constraints(constraints)
    T, m = len(trace_off) - 1, len(pattern)
    status = np.zeros(max(T, 1), dtype=np.int32)
    total = np.zeros(max(T, 1), dtype=np.int32)
    cols = [np.zeros(max(T * m, 1), dtype=np.int32) for _ in range(4)]
    L.wnm_host_eval.restype = C.c_int
    rc = L.wnm_host_eval(_p(trace_off, C.c_int64), _p(act, C.c_int32), _p(ts_ms, C.c_int64), C.c_int64(T), _p(pattern, C.c_int32),
                         C.c_int32(m), cons, C.c_int32(n_cons), C.c_int32(uncertainty), C.c_int32(step), C.c_int32(k), C.c_uint32(flags),
                         _p(status, C.c_int32), _p(total, C.c_int32), *[_p(c, C.c_int32) for c in cols])
    assert rc == 0, "wnm_rank did not produce a permutation"

    class R:
        pass
    r = R()
    hit = np.nonzero(status[:T] == 1)[0]
    r.trace_idx = hit.astype(np.int64)
    r.total_change = total[hit]
    r.ev_pos, r.ev_value, r.ev_change, r.ev_stream_pos = (c[:T * m].reshape(T, m)[hit] for c in cols)
    return r
