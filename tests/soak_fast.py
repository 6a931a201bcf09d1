"""Randomized soak of the closed-form evaluators (detect_fast.cuh, host build) against the oracle.  Not collected by
pytest; run `python tests/soak_fast.py <seed0> <n_seeds>`.  Draws only NFAs of the NK, FK2 and NP1 classes (few activity
types so that traces are dense in pattern events, ties and negatives are frequent)."""
import sys

import numpy as np

sys.path.insert(0, __file__.rsplit("/", 2)[0])
import oracle  # noqa: E402
from sequencedetectionqueryexecutor_b200 import _abi as abi  # noqa: E402
from tests import gen, host_engine  # noqa: E402

N_, P_, S_, X_, O_ = abi.STATE_NORMAL, abi.STATE_KLEENE_PLUS, abi.STATE_KLEENE_STAR, abi.STATE_NEGATIVE, abi.STATE_OR


def rand_pred(rng, ref):
    op = abi.OP_LE if rng.random() < 0.6 else abi.OP_GE
    if rng.random() < 0.5:
        return (abi.ATTR_POSITION, op, ref, int(rng.integers(0, 8)))
    return (abi.ATTR_TIMESTAMP, op, ref, int(rng.integers(0, 1500)))


def nk_nfa(rng, n_act):
    """normal / or / negative states; first and last positive; no two negatives in a row; types may repeat."""
    n = int(rng.integers(1, 7))
    states = []
    for s in range(n):
        neg_ok = 0 < s < n - 1 and states[-1]["kind"] != X_
        r = rng.random()
        if neg_ok and r < 0.25:
            kind = X_
        elif r < 0.5:
            kind = O_
        else:
            kind = N_
        k = int(rng.integers(2, 4)) if kind == O_ else (int(rng.integers(1, 3)) if kind == X_ else 1)
        types = [int(x) for x in rng.choice(n_act, size=min(k, n_act), replace=False)]
        states.append({"kind": kind, "types": types, "preds": []})
    positive = [s for s in range(n) if states[s]["kind"] != X_]
    if rng.random() < 0.4:   # Markov shape: position predicates on the positive state before (kernel K1-P's backward evaluation)
        for b in range(1, n):
            if states[b]["kind"] == X_ or states[b - 1]["kind"] == X_:
                continue
            for _ in range(int(rng.integers(0, 3))):
                op = abi.OP_LE if rng.random() < 0.5 else abi.OP_GE
                states[b]["preds"].append((abi.ATTR_POSITION, op, b - 1, int(rng.integers(0, 12))))
        return states
    for _ in range(3):
        if n >= 2 and rng.random() < 0.6:
            b = int(rng.integers(1, n))
            refs = [s for s in positive if s < b]
            if not refs or len(states[b]["preds"]) >= abi.MAX_PREDS:
                continue
            states[b]["preds"].append(rand_pred(rng, int(rng.choice(refs))))
    return states


def fk2_nfa(rng, n_act):
    a, b = [int(x) for x in rng.choice(n_act, size=2, replace=False)]
    preds = [rand_pred(rng, 0) for _ in range(int(rng.integers(0, 3)))]
    return [{"kind": P_, "types": [a], "preds": []}, {"kind": S_, "types": [b], "preds": preds}]


def np1_nfa(rng, n_act):
    """normal / or states and exactly one kleeneClosure+ (or kleeneClosure*) state anywhere; types may repeat."""
    n = int(rng.integers(2, 7))
    k = int(rng.integers(0, n))
    star = rng.random() < 0.4   # kleeneClosure* instead: in the class without any predicate
    states = []
    for s in range(n):
        if s == k:
            kind = S_ if star else P_
        else:
            kind = O_ if rng.random() < 0.3 else N_
        m = int(rng.integers(2, 4)) if kind == O_ else 1
        types = [int(x) for x in rng.choice(n_act, size=min(m, n_act), replace=False)]
        states.append({"kind": kind, "types": types, "preds": []})
    if k >= 1 and not star and rng.random() < 0.6:   # constraints that reference states before the Kleene state
        for _ in range(int(rng.integers(1, 4))):
            b = int(rng.integers(1, n))
            if len(states[b]["preds"]) < abi.MAX_PREDS:
                states[b]["preds"].append(rand_pred(rng, int(rng.integers(0, min(b, k)))))
    return states


def main(seed0, n_seeds):
    bad = 0
    stats = {"ok": 0, "err": 0, "unsupported": 0, "wide": 0, "matches": 0}
    for seed in range(seed0, seed0 + n_seeds):
        rng = np.random.default_rng(seed)
        n_act = int(rng.integers(2, 9))
        sorted_ts = rng.random() < 0.85
        off, act, ts = gen.make_log(80, 0, int(rng.integers(4, 60)), n_act, seed=int(rng.integers(1 << 30)),
                                    max_gap_s=300, jitter_ms=bool(rng.integers(0, 2)))
        if not sorted_ts:
            ts = ts.copy()
            rng.shuffle(ts)
        which = rng.random()
        fk2 = which < 0.25
        np1 = 0.25 <= which < 0.6
        states = fk2_nfa(rng, n_act) if fk2 else (np1_nfa(rng, n_act) if np1 else nk_nfa(rng, n_act))
        only = False
        if np1 and rng.random() < 0.1:    # any pattern under onlyAppearances is in the class as well
            b = int(rng.integers(1, len(states)))
            states[b]["preds"].append(rand_pred(rng, int(rng.integers(0, b + 1))))
            only = True
        flags = 0
        if rng.random() < 0.4:
            flags |= abi.F_EVT_POS
        if not fk2 and rng.random() < (0.15 if np1 else 0.4):
            flags |= abi.F_RETURN_ALL
        if rng.random() < 0.1 or only:
            flags |= abi.F_ONLY_APPEARANCES
        if not fk2 and rng.random() < 0.2:
            flags |= abi.F_COUNT_MATCHES
        nfa = abi.make_nfa(states)
        rc, got, n_wide = host_engine.detect(off, act, ts, n_act, nfa, flags=flags)
        if rc == abi.E_UNSUPPORTED:
            stats["unsupported"] += 1
            continue
        assert rc == 0
        want = oracle.detect(off, act, ts, nfa, flags=flags)
        ok, why = got.same_as(want)
        stats["err" if want.n_ref_errors else "ok"] += 1
        stats["wide"] += n_wide
        stats["matches"] += want.n_traces
        if not ok:
            bad += 1
            print("MISMATCH seed", seed, why, states, flags, "sorted" if sorted_ts else "unsorted", flush=True)
            if bad > 5:
                break
    print("done", seed0, n_seeds, stats, "mismatches", bad, flush=True)
    return bad


if __name__ == "__main__":
    sys.exit(1 if main(int(sys.argv[1]), int(sys.argv[2])) else 0)
