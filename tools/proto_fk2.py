"""Prototype (exploration only): closed form of the first-largest occurrence for NFA `a+ b*` with predicates on
state 1 referencing state 0, validated against the oracle's full emission list."""
import sys
import numpy as np
sys.path.insert(0, "/root/repo")
import oracle
from sequencedetectionqueryexecutor_b200 import _abi as abi

P_, S_ = abi.STATE_KLEENE_PLUS, abi.STATE_KLEENE_STAR


def closed_form(types, pos, ts, preds):
    n = len(types)
    def attr(j, a): return pos[j] if a == abi.ATTR_POSITION else ts[j]
    def pas(b, a):
        for (at, op, ref, c) in preds:
            l, r = attr(b, at), attr(a, at) + c
            if (op == abi.OP_LE and not l <= r) or (op == abi.OP_GE and not l >= r):
                return False
        return True
    A = [j for j in range(n) if types[j] == 0]
    if not A:
        return None
    M = len(A) - 1
    lastA = [-1] * n
    la = -1
    for j in range(n):
        if types[j] == 0: la = j
        lastA[j] = la
    good = [types[j] == 1 and lastA[j] >= 0 and pas(j, lastA[j]) for j in range(n)]
    # candidates: (size, k, kind, idx)
    cands = []
    goods = [j for j in range(n) if good[j]]
    for m in range(1, M + 1):
        g = [j for j in goods if j > A[m]]
        if g:
            cands.append((m + 1 + len(g), g[-1], 'Q', m))
    for i in range(M + 1):
        h = [j for j in range(A[i] + 1, n) if types[j] == 1 and pas(j, A[i])]
        if h:
            cands.append((1 + len(h), h[-1], 'P', i))
    if not cands:
        return None
    S = max(c[0] for c in cands)
    tied = [c for c in cands if c[0] == S]
    k = min(c[1] for c in tied)
    tied = [c for c in tied if c[1] == k]
    def placed(c, t):
        kind, x = c[2], c[3]
        if kind == 'P':
            return t == A[x] or (types[t] == 1 and t > A[x] and pas(t, A[x]))
        return t == A[x - 1] or (good[t] and t > A[x])
    def precedes(X, Y):  # X before Y in list at event k
        for t in range(k - 1, -1, -1):
            px, py = placed(X, t), placed(Y, t)
            if px and py:
                if types[t] == 0:  # creation tie at a_q: P(q) vs Q(q+1)
                    q = A.index(t)
                    if X[2] == 'P':
                        return q == 0
                    return q != 0
                continue
            if px: return False
            if py: return True
        raise AssertionError("no order")
    best = tied[0]
    for c in tied[1:]:
        if precedes(c, best): best = c
    kind, x = best[2], best[3]
    if kind == 'P':
        return [A[x]] + [j for j in range(A[x] + 1, n) if types[j] == 1 and pas(j, A[x])]
    return A[:x + 1] + [j for j in goods if j > A[x]]


def main(seed, iters):
    rng = np.random.default_rng(seed)
    bad = 0
    for it in range(iters):
        n = int(rng.integers(0, 14))
        types = rng.integers(0, 2, size=n)
        pos = np.cumsum(rng.integers(1, 4, size=n)) if rng.random() < 0.5 else np.arange(n)
        mono = rng.random() < 0.8
        ts = np.cumsum(rng.integers(0, 5, size=n)) if mono else rng.integers(0, 20, size=n)
        preds = []
        for _ in range(int(rng.integers(0, 3))):
            preds.append((int(rng.integers(0, 2)), int(rng.integers(0, 2)), 0, int(rng.integers(0, 8))))
        nfa = abi.make_nfa([dict(kind=P_, types=[0]), dict(kind=S_, types=[1], preds=preds)])
        status, matches = oracle.run_stream(nfa, types, pos, ts)
        assert status == 0
        want = None
        if matches:
            want = max(matches, key=len)
        got = closed_form(list(types), list(pos), list(ts), preds)
        gotp = None if got is None else [int(pos[j]) for j in got]
        if gotp != want:
            bad += 1
            if bad < 6:
                print("MISMATCH", list(types), list(pos), list(ts), preds, "want", want, "got", gotp)
    print("seed", seed, "iters", iters, "bad", bad)


if __name__ == "__main__":
    main(int(sys.argv[1]), int(sys.argv[2]))
