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
