"""Soak: kernel W's closed form (host build) against the literal oracle on many random cases.  python -m tests.soak_wnm [seeds] [cases]"""
import sys

import numpy as np

import oracle
from sequencedetectionqueryexecutor_b200 import _abi as abi
from tests import host_engine
from tests.test_wnm import random_case

if __name__ == "__main__":
    seeds = int(sys.argv[1]) if len(sys.argv) > 1 else 10
    cases = int(sys.argv[2]) if len(sys.argv) > 2 else 500
    n_cases = n_hit = n_tr = n_cons = skipped = 0
    for seed in range(seeds):
        rng = np.random.default_rng(50_000 + seed)
        for _ in range(cases):
            off, act, ts, pattern, cons, u, step, k, flags = random_case(rng)
            want = oracle.why_not_match(off, act, ts, pattern, cons, u, step, k, flags=flags, run_limit=400_000)
            if want is None:
                skipped += 1
                continue
            got = host_engine.wnm_eval(off, act, ts, pattern, cons, u, step, k, flags=flags)
            ok, why = abi.AlmostMatchResult.same_as(got, want)
            if not ok:
                print("MISMATCH", why, pattern.tolist(), cons, u, step, k, flags, off.tolist(), act.tolist(), ts.tolist())
                sys.exit(1)
            n_cases += 1
            n_hit += want.n_traces
            n_tr += len(off) - 1
            n_cons += len(cons) > 0
    print(f"{n_cases} cases ({skipped} skipped as too large), {n_cons} with constraints, {n_hit} almost-matches over {n_tr} traces: all equal")
