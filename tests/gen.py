"""Seeded synthetic logs and random NFAs shared by the parity tests."""
import numpy as np

from sequencedetectionqueryexecutor_b200 import _abi as abi

T0_MS = 1577836800000  # 2020-01-01T00:00:00Z


def make_log(n_traces, min_len, max_len, n_act, seed, max_gap_s=600, jitter_ms=False, zipf=None):
    """CSR log: lengths ~ U{min_len..max_len}, uniform (or Zipf) activities, strictly increasing timestamps."""
    rng = np.random.default_rng(seed)
    lens = rng.integers(min_len, max_len + 1, size=n_traces, dtype=np.int64)
    off = np.zeros(n_traces + 1, dtype=np.int64)
    np.cumsum(lens, out=off[1:])
    E = int(off[-1])
    if zipf:
        p = 1.0 / np.arange(1, n_act + 1) ** zipf
        act = rng.choice(n_act, size=E, p=p / p.sum()).astype(np.int32)
    else:
        act = rng.integers(0, n_act, size=E, dtype=np.int32)
    gaps = rng.integers(1, max_gap_s + 1, size=E, dtype=np.int64) * 1000
    if jitter_ms:
        gaps += rng.integers(0, 1000, size=E, dtype=np.int64)
    start = T0_MS + rng.integers(0, 30 * 86400, size=n_traces, dtype=np.int64) * 1000
    ts = np.cumsum(gaps)
    # restart the running sum at every trace start
    first = off[:-1][lens > 0]
    base = np.zeros(E, dtype=np.int64)
    trace_of = np.repeat(np.arange(n_traces), lens)
    before = np.concatenate(([0], ts))[first] if len(first) else np.zeros(0, dtype=np.int64)
    sub = np.zeros(n_traces, dtype=np.int64)
    sub[lens > 0] = before
    base = ts - sub[trace_of] + start[trace_of]
    return off, act, base.astype(np.int64)


KINDS = [abi.STATE_NORMAL, abi.STATE_KLEENE_PLUS, abi.STATE_KLEENE_STAR, abi.STATE_NEGATIVE, abi.STATE_OR]


def random_nfa(rng, n_act, max_states=5, p_constraint=0.5, kinds=KINDS, max_const=6, time_const=(0, 900)):
    """Random NFA: 1..max_states states from `kinds`, 0-2 constraints (posA < posB)."""
    n = int(rng.integers(1, max_states + 1))
    states = []
    for _ in range(n):
        kind = int(rng.choice(kinds))
        if kind == abi.STATE_OR:
            k = int(rng.integers(2, 4))
            types = [int(x) for x in rng.choice(n_act, size=min(k, n_act), replace=False)]
        else:
            types = [int(rng.integers(0, n_act))]
        states.append({"kind": kind, "types": types, "preds": []})
    if n >= 2:
        for _ in range(2):
            if rng.random() < p_constraint:
                b = int(rng.integers(1, n))
                a = int(rng.integers(0, b))
                if len(states[b]["preds"]) >= abi.MAX_PREDS:
                    continue
                if rng.random() < 0.5:
                    states[b]["preds"].append((abi.ATTR_POSITION, int(rng.integers(0, 2)), a, int(rng.integers(0, max_const + 1))))
                else:
                    states[b]["preds"].append((abi.ATTR_TIMESTAMP, int(rng.integers(0, 2)), a,
                                               int(rng.integers(time_const[0], time_const[1] + 1))))
    return states
