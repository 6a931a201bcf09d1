"""Class NP1 with ONE kleeneClosure* state and no constraints (detect_fast.cuh): the Kleene shapes of the reference's own tests
(EvaluateComplexQueries.java:101-103, 126-127, 150-152, 175-176, 199-201, 298-301) on the GPU, against the oracle: occurrences,
the engine's match count (the `*` state adds matches without any Kleene event), EventPos route, returnAll (closed form only when
the `*` state is the second one, the run-list engine otherwise), traces beyond 64 relevant events listed."""
import pytest

from sequencedetectionqueryexecutor_b200 import _abi as abi
from tests import gen
from tests.test_detect_gpu import _check, ctx  # noqa: F401  (the module-scoped context fixture)

pytestmark = pytest.mark.gpu

N_, P_, S_, X_, O_ = abi.STATE_NORMAL, abi.STATE_KLEENE_PLUS, abi.STATE_KLEENE_STAR, abi.STATE_NEGATIVE, abi.STATE_OR

STAR_SHAPES = [
    # second state, first state, last state, behind a longer prefix
    ("a b* c", [dict(kind=N_, types=[0]), dict(kind=S_, types=[1]), dict(kind=N_, types=[2])]),
    ("a* b", [dict(kind=S_, types=[0]), dict(kind=N_, types=[1])]),
    ("a* b c", [dict(kind=S_, types=[0]), dict(kind=N_, types=[1]), dict(kind=N_, types=[2])]),
    ("a b*", [dict(kind=N_, types=[0]), dict(kind=S_, types=[1])]),
    ("a b a*", [dict(kind=N_, types=[0]), dict(kind=N_, types=[1]), dict(kind=S_, types=[0])]),
    ("(a|b) b* c", [dict(kind=O_, types=[0, 1]), dict(kind=S_, types=[1]), dict(kind=N_, types=[2])]),
    ("a (b|c) d* (a|e) b", [dict(kind=N_, types=[0]), dict(kind=O_, types=[1, 2]), dict(kind=S_, types=[3]), dict(kind=O_, types=[0, 4]), dict(kind=N_, types=[1])]),
]


@pytest.mark.parametrize("name,states", STAR_SHAPES, ids=[s[0] for s in STAR_SHAPES])
def test_one_kleene_star_state_without_constraints_closed_form(ctx, name, states):
    for n_act, max_len in ((5, 30), (7, 120), (3, 90)):
        off, act, ts = gen.make_log(3000, 0, max_len, n_act, seed=177 + n_act, max_gap_s=100, jitter_ms=True)
        for flags in (0, abi.F_COUNT_MATCHES, abi.F_EVT_POS | abi.F_COUNT_MATCHES, abi.F_EVT_POS | abi.F_RETURN_ALL, abi.F_NO_EVENT_COLUMNS):
            got = _check(ctx, off, act, ts, n_act, states, flags)
            if n_act == 5:
                assert got.n_unsupported == 0
