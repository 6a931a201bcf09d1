import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle
from sequencedetectionqueryexecutor_b200 import _abi as abi, api
from tests import gen
N_, O_, X_, P_, S_ = abi.STATE_NORMAL, abi.STATE_OR, abi.STATE_NEGATIVE, abi.STATE_KLEENE_PLUS, abi.STATE_KLEENE_STAR
off, act, ts = gen.make_log(4000, 0, 70, 200, seed=77, max_gap_s=300, jitter_ms=True)
act = (act % 7 + (act % 3 == 0) * 190).astype(np.int32)
act8 = act.astype(np.uint8)
cases = [([dict(kind=N_, types=[0]), dict(kind=O_, types=[1, 2], preds=[(abi.ATTR_POSITION, abi.OP_LE, 0, 10)]),
           dict(kind=X_, types=[3]), dict(kind=N_, types=[190])], 0),
         ([dict(kind=P_, types=[0]), dict(kind=S_, types=[191], preds=[(abi.ATTR_TIMESTAMP, abi.OP_LE, 0, 600)])], 0),
         ([dict(kind=N_, types=[0]), dict(kind=P_, types=[1]), dict(kind=N_, types=[2])], abi.F_RETURN_ALL)]
with api.Context(0) as ctx:
    for chunk in ("1000", "4099", None):
        if chunk: os.environ["SIESTA_CHUNK_EVENTS"] = chunk
        else: os.environ.pop("SIESTA_CHUNK_EVENTS", None)
        for ci, (states, flags) in enumerate(cases):
            nfa = abi.make_nfa(states)
            want = oracle.detect(off, act, ts, nfa, flags=flags)
            import torch
            p_off, p_act, p_act8, p_ts = (torch.from_numpy(x).pin_memory() for x in (off, act, act8, ts))
            for name, a, col, c in (("int32", off, act, ts), ("act8", off, act8, ts), ("int32 pinned", p_off.numpy(), p_act.numpy(), p_ts.numpy()),
                                    ("act8 pinned", p_off.numpy(), p_act8.numpy(), p_ts.numpy())):
                print("chunk", chunk, "case", ci, name, flush=True)
                try:
                    got = ctx.evaluate_events(a, col, c, 200, nfa, flags=flags)
                    print("   ", got.same_as(want), flush=True)
                except Exception as ex:
                    print("   FAILED", ex, flush=True)
                    sys.exit(1)
