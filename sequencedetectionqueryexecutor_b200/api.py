"""Thin object layer over the C-ABI: Context (one per GPU), EventLog (CSR log resident in HBM), detect()."""
import ctypes as C

import numpy as np

from . import _abi
from ._lib import check, lib


def _ptr(a):
    return C.c_void_p(a.ctypes.data) if a is not None else None


class Context:
    """siesta_ctx: one per (process, device)."""

    def __init__(self, device=0):
        self._h = C.c_void_p()
        check(lib().siesta_init(device, C.byref(self._h)))
        self.device = device

    def close(self):
        if self._h:
            lib().siesta_shutdown(self._h)
            self._h = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def load_log(self, trace_off, act, ts_ms, n_activities):
        return EventLog(self, trace_off, act, ts_ms, n_activities)

    def wrap_log(self, d_trace_off, d_act, d_ts_ms, n_activities, max_trace_len=0):
        """Borrow torch CUDA tensors (int64, int32, int64) as a resident log."""
        return EventLog(self, None, None, None, n_activities, device_tensors=(d_trace_off, d_act, d_ts_ms),
                        max_trace_len=max_trace_len)

    def evaluate_events(self, trace_off, act, ts_ms, n_activities, nfa, flags=0):
        """Literal SaseConnector.evaluate: events travel host -> device inside the call."""
        trace_off = np.ascontiguousarray(trace_off, dtype=np.int64)
        act = np.ascontiguousarray(act, dtype=np.int32)
        ts_ms = np.ascontiguousarray(ts_ms, dtype=np.int64)
        out = C.POINTER(_abi.Matches)()
        check(lib().siesta_evaluate_events(self._h, _ptr(trace_off), _ptr(act), _ptr(ts_ms), len(trace_off) - 1, len(act),
                                           n_activities, C.byref(nfa), flags, C.byref(out)))
        res = _abi.MatchResult.from_struct(out.contents)
        lib().siesta_matches_free(out)
        return res


class EventLog:
    """siesta_log: CSR event log (trace_off int64[T+1], act int32[E], ts_ms int64[E]) in HBM."""

    def __init__(self, ctx, trace_off, act, ts_ms, n_activities, device_tensors=None, max_trace_len=0):
        self.ctx = ctx
        self._h = C.c_void_p()
        self.n_activities = int(n_activities)
        if device_tensors is None:
            trace_off = np.ascontiguousarray(trace_off, dtype=np.int64)
            act = np.ascontiguousarray(act, dtype=np.int32)
            ts_ms = np.ascontiguousarray(ts_ms, dtype=np.int64)
            if len(act) != len(ts_ms):
                raise ValueError("act and ts_ms must have the same length")
            check(lib().siesta_log_load(ctx._h, _ptr(trace_off), _ptr(act), _ptr(ts_ms), len(trace_off) - 1, len(act),
                                        self.n_activities, C.byref(self._h)))
            self._keep = None
        else:
            d_off, d_act, d_ts = device_tensors
            self._keep = device_tensors  # the caller's tensors must outlive the log
            check(lib().siesta_log_wrap_device(ctx._h, C.c_void_p(d_off.data_ptr()), C.c_void_p(d_act.data_ptr()),
                                               C.c_void_p(d_ts.data_ptr()), d_off.numel() - 1, d_act.numel(),
                                               self.n_activities, int(max_trace_len), C.byref(self._h)))
        self.n_traces = lib().siesta_log_n_traces(self._h)
        self.n_events = lib().siesta_log_n_events(self._h)

    def close(self):
        if self._h:
            lib().siesta_log_free(self._h)
            self._h = C.c_void_p()
            self._keep = None

    def detect(self, nfa, cand=None, flags=0):
        """siesta_detect: host-buffer results (MatchResult)."""
        if cand is not None:
            cand = np.ascontiguousarray(cand, dtype=np.int64)
        out = C.POINTER(_abi.Matches)()
        check(lib().siesta_detect(self._h, C.byref(nfa), _ptr(cand), 0 if cand is None else len(cand), flags, C.byref(out)))
        res = _abi.MatchResult.from_struct(out.contents)
        lib().siesta_matches_free(out)
        return res

    def detect_device(self, nfa, d_cand=None, flags=0, stream=None):
        """siesta_detect_device: results stay in HBM; returns a DeviceMatches (free with .close())."""
        dm = _abi.DevMatches()
        cp = C.c_void_p(d_cand.data_ptr()) if d_cand is not None else None
        n = 0 if d_cand is None else d_cand.numel()
        check(lib().siesta_detect_device(self._h, C.byref(nfa), cp, n, flags, C.c_void_p(stream) if stream else None,
                                         C.byref(dm)))
        return DeviceMatches(dm)


class DeviceMatches:
    def __init__(self, dm):
        self.dm = dm
        for k in ("n_traces", "n_occurrences", "n_events", "n_matches_emitted", "n_ref_errors", "kernel_ms", "detect_ms"):
            setattr(self, k, getattr(dm, k))

    def close(self):
        if self.dm is not None:
            lib().siesta_dev_matches_free(C.byref(self.dm))
            self.dm = None


def kernel_launches():
    return int(lib().siesta_kernel_launches())
