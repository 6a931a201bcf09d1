"""Long-running randomized soak of the device engine (host build) against the oracle.  Not collected by pytest;
run `python tests/soak_engine.py <seed0> <n_seeds>`.  Exercises the run-list reductions (inert pruning, dominated-run
merging) on NFAs with distinct types per state (merge_safe) and `within` constraints."""
import sys

import numpy as np

sys.path.insert(0, __file__.rsplit("/", 2)[0])
import oracle  # noqa: E402
from sequencedetectionqueryexecutor_b200 import _abi as abi  # noqa: E402
from tests import gen, host_engine  # noqa: E402


def distinct_nfa(rng, n_act):
    n = int(rng.integers(2, 6))
    perm = rng.permutation(n_act)
    states, used = [], 0
    for _ in range(n):
        kind = int(rng.choice(gen.KINDS, p=[0.3, 0.25, 0.2, 0.1, 0.15]))
        k = int(rng.integers(2, 4)) if kind == abi.STATE_OR else 1
        if used + k > n_act:
            kind, k = abi.STATE_NORMAL, 1
        if used + k > n_act:
            used = 0  # out of distinct types: reuse (merge_safe may then be false; still a valid case)
        states.append({"kind": kind, "types": [int(x) for x in perm[used:used + k]], "preds": []})
        used += k
    for _ in range(3):
        if rng.random() < 0.7:
            b = int(rng.integers(1, n))
            a = int(rng.integers(0, b))
            if len(states[b]["preds"]) >= abi.MAX_PREDS:
                continue
            op = abi.OP_LE if rng.random() < 0.7 else abi.OP_GE
            if rng.random() < 0.5:
                states[b]["preds"].append((abi.ATTR_POSITION, op, a, int(rng.integers(0, 8))))
            else:
                states[b]["preds"].append((abi.ATTR_TIMESTAMP, op, a, int(rng.integers(0, 1500))))
    return states


def main(seed0, n_seeds):
    bad = 0
    stats = {"ok": 0, "err": 0, "unsupported": 0}
    for seed in range(seed0, seed0 + n_seeds):
        rng = np.random.default_rng(seed)
        n_act = int(rng.integers(6, 12))
        sorted_ts = rng.random() < 0.9
        off, act, ts = gen.make_log(60, 0, 40, n_act, seed=int(rng.integers(1 << 30)), max_gap_s=300,
                                    jitter_ms=bool(rng.integers(0, 2)))
        if not sorted_ts:  # unsorted timestamps must switch the timestamp pruning off, not break parity
            ts = ts.copy()
            rng.shuffle(ts)
        states = distinct_nfa(rng, n_act) if rng.random() < 0.7 else gen.random_nfa(rng, n_act)
        flags = 0
        if rng.random() < 0.4:
            flags |= abi.F_EVT_POS
        if rng.random() < 0.3:
            flags |= abi.F_RETURN_ALL
        if rng.random() < 0.1:
            flags |= abi.F_ONLY_APPEARANCES
        if rng.random() < 0.15:
            flags |= abi.F_COUNT_MATCHES
        nfa = abi.make_nfa(states)
        rc, got, _ = host_engine.detect(off, act, ts, n_act, nfa, flags=flags)
        if rc == abi.E_UNSUPPORTED:
            stats["unsupported"] += 1
            continue
        assert rc == 0
        want = oracle.detect(off, act, ts, nfa, flags=flags)
        ok, why = got.same_as(want)
        stats["err" if want.n_ref_errors else "ok"] += 1
        if not ok:
            bad += 1
            print("MISMATCH seed", seed, why, states, flags, "sorted" if sorted_ts else "unsorted", flush=True)
            if bad > 5:
                break
    print("done", seed0, n_seeds, stats, "mismatches", bad, flush=True)
    return bad


if __name__ == "__main__":
    sys.exit(1 if main(int(sys.argv[1]), int(sys.argv[2])) else 0)
