"""The reference's engine known-answer tests.

Stream "A B A C D A B E" as EventPos (positions 0-7; ids = positions, timestamps = indices):
src/test/java/com/datalab/siesta/queryprocessor/SaseConnection/EvaluateNewQueries.java:26-43.
`expected` is the match count the reference asserts (file:line in `where`; N = EvaluateNewQueries.java,
C = EvaluateComplexQueries.java).  `matches` are the match lists (event positions, emission order)
derived at survey time from a transliteration of the Java engine with Engine.createNewRun's trailing
block disabled; `head` is the count HEAD code gives where it differs from the asserted one.
"""
from sequencedetectionqueryexecutor_b200 import _abi as abi

A, B, Cc, D, E = range(5)
STREAM_TYPES = [A, B, A, Cc, D, A, B, E]

N_, P_, S_, X_, O_ = abi.STATE_NORMAL, abi.STATE_KLEENE_PLUS, abi.STATE_KLEENE_STAR, abi.STATE_NEGATIVE, abi.STATE_OR


def st(kind, *types, preds=()):
    return {"kind": kind, "types": list(types), "preds": list(preds)}


def gap_within(ref, c):
    return (abi.ATTR_POSITION, abi.OP_LE, ref, c)


KATS = [
    dict(name="A,(C|D),B", states=[st(N_, A), st(O_, Cc, D), st(N_, B)], expected=2, where="N:68 C:66",
         matches=[[0, 3, 6], [2, 3, 6]]),
    dict(name="A,!C,B", states=[st(N_, A), st(X_, Cc), st(N_, B)], expected=2, where="N:91 C:91",
         matches=[[0, 1], [5, 6]]),
    dict(name="A,B*,E", states=[st(N_, A), st(S_, B), st(N_, E)], expected=7, where="N:114 C:116", head=14,
         matches=[[0, 1, 7], [0, 7], [0, 1, 6, 7], [2, 6, 7], [5, 6, 7], [2, 7], [5, 7]]),
    dict(name="A*,B", states=[st(S_, A), st(N_, B)], expected=8, where="N:136 C:140",
         matches=[[0, 1], [1], [0, 2, 6], [0, 2, 5, 6], [2, 6], [2, 5, 6], [5, 6], [6]]),
    dict(name="A*,B,E", states=[st(S_, A), st(N_, B), st(N_, E)], expected=8, where="N:159 C:165",
         matches=[[0, 1, 7], [0, 2, 6, 7], [1, 7], [0, 2, 5, 6, 7], [2, 6, 7], [2, 5, 6, 7], [5, 6, 7], [6, 7]]),
    dict(name="A,B*", states=[st(N_, A), st(S_, B)], expected=7, where="N:181 C:189", head=15,
         matches=[[0], [0, 1], [0, 1, 6], [2], [2, 6], [5], [5, 6]]),
    dict(name="A,B,A*", states=[st(N_, A), st(N_, B), st(S_, A)], expected=5, where="N:205 C:214",
         matches=[[0, 1], [0, 1, 2], [0, 1, 2, 5], [2, 6], [5, 6]]),
    dict(name="!A,B,C", states=[st(X_, A), st(N_, B), st(N_, Cc)], expected=1, where="N:228 C:239",
         matches=[[1, 3]]),
    dict(name="B,C,!A", states=[st(N_, B), st(N_, Cc), st(X_, A)], expected=1, where="N:251 C:264",
         matches=[[1, 3]]),
    dict(name="A,B,C", states=[st(N_, A), st(N_, B), st(N_, Cc)], expected=1, where="N:274 C:291",
         matches=[[0, 1, 3]]),
    dict(name="A,E,C", states=[st(N_, A), st(N_, E), st(N_, Cc)], expected=0, where="N:297", matches=[]),
    dict(name="A,B", states=[st(N_, A), st(N_, B)], expected=3, where="N:319", matches=[[0, 1], [2, 6], [5, 6]]),
    dict(name="(A|B),C", states=[st(O_, A, B), st(N_, Cc)], expected=3, where="N:343",
         matches=[[0, 3], [1, 3], [2, 3]]),
    dict(name="D,(A|B),E", states=[st(N_, D), st(O_, A, B), st(N_, E)], expected=1, where="N:368",
         matches=[[4, 5, 7]]),
    dict(name="A+,B,E", states=[st(P_, A), st(N_, B), st(N_, E)], expected=6, where="N:390",
         matches=[[0, 1, 7], [0, 2, 6, 7], [0, 2, 5, 6, 7], [2, 6, 7], [2, 5, 6, 7], [5, 6, 7]]),
    dict(name="A,B+,E", states=[st(N_, A), st(P_, B), st(N_, E)], expected=4, where="N:412",
         matches=[[0, 1, 7], [0, 1, 6, 7], [2, 6, 7], [5, 6, 7]]),
    dict(name="A,B,A+", states=[st(N_, A), st(N_, B), st(P_, A)], expected=2, where="N:434",
         matches=[[0, 1, 2], [0, 1, 2, 5]]),
    dict(name="(A|B),B*,E", states=[st(O_, A, B), st(S_, B), st(N_, E)], expected=9, where="C:317", head=16,
         matches=[[0, 1, 7], [0, 7], [0, 1, 6, 7], [1, 6, 7], [2, 6, 7], [5, 6, 7], [1, 7], [2, 7], [5, 7]]),
    dict(name="A,(C|D),B gap within 2 (0,1)",
         states=[st(N_, A), st(O_, Cc, D, preds=[gap_within(0, 2)]), st(N_, B)], expected=1, where="C:346",
         matches=[[2, 3, 6]]),
    dict(name="A,!D,E gap within 2 (0,1)",
         states=[st(N_, A), st(X_, D, preds=[gap_within(0, 2)]), st(N_, E)], expected=2, where="C:374",
         matches=[[0, 7], [5, 7]]),
    dict(name="A,B*,E gap within 1 (0,1),(1,2)",
         states=[st(N_, A), st(S_, B, preds=[gap_within(0, 1)]), st(N_, E, preds=[gap_within(1, 1)])], expected=2,
         where="C:403", matches=[[5, 6, 7], [5, 7]]),
]
