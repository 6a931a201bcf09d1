"""B200-native SIESTA pattern-query hot path (verification, intersection, declare/stats counting).

The compute lives in libsiesta_gpu.so (hand-written sm_100a CUDA behind the C-ABI of include/siesta_gpu.h);
this package is the ctypes binding plus a host-side mirror of the reference's Java interface."""
from . import _abi  # noqa: F401
from ._abi import (F_COUNT_MATCHES, F_EVT_POS, F_LITERAL_RUNS, F_MODE_HEAD, F_NO_EVENT_COLUMNS, F_ONLY_APPEARANCES, F_RETURN_ALL, MatchResult,  # noqa: F401
                   make_nfa)
