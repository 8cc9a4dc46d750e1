"""GPU: the forward-time classes (TimeIteration<true>, ForwardConstantStepper,
ForwardRannacher / CrankNicolson / BDFOne .. BDFSix; IterativeMethod.hpp:1494-1495, Stepper.hpp:41)
through the C ABI (batch.forward = 1).  The reference fixtures (forward_* cases of
tests/golden/reference_cases.npz, the classes themselves run by qpde_ref) go through
test_parity_gpu.py::test_golden_cases for both kernel families; here: varied batches against the
oracle and the relation to the reverse sweep."""
import numpy as np
import pytest

from helpers import REL_TOL, mask_mismatches, rel_err

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("kernel", ["resident", "interleaved"])
@pytest.mark.parametrize("refine,scheme,american,E", [
    (2, "rannacher", True, 0), (4, "bdf2", False, 4), (6, "rannacher", True, 3), (6, "cn", False, 0),
    (3, "bdf3", False, 2),
])
def test_forward_batch_vs_oracle(qpde, handle, oracle, kernel, refine, scheme, american, E):
    Cn = 5
    rng = np.random.default_rng(3)
    K = rng.uniform(80, 120, Cn); vol = rng.uniform(.15, .45, Cn); T = rng.uniform(.5, 1.5, Cn); r = rng.uniform(.01, .06, Cn)
    steps = 120
    axis = oracle.special_axis()[None, :] * K[:, None]
    spec = qpde.BatchSpec(axis=axis, refine=refine, r=r, sigma=vol, q=0., T=T, dt=T / steps, strike=K, scheme=scheme,
                          american=american, payoff="put", n_exercise=E, exercise="k_minus_s" if E else "none",
                          kernel=kernel, forward=True)
    try:
        plan = qpde.Plan(handle, spec)
    except qpde.QpdeError as e:
        if e.code == qpde.ERR_UNSUPPORTED:
            pytest.skip("kernel family does not take this scheme")
        raise
    with plan:
        assert plan.kernel == kernel
        plan.run(); res = plan.fetch()
    for c in range(Cn):
        o = oracle.forward_option(K[c], r[c], vol[c], T[c], steps, refine, american=american, scheme=scheme, E=E,
                                  exercise=E > 0)
        assert res.status[c] == 0 and res.n_steps[c] == o["n_steps"]
        assert np.array_equal(res.S[c], o["S"])
        err = rel_err(res.V[c], o["V"]).max()
        assert err <= REL_TOL, "contract %d: %.3e" % (c, err)
        if american:
            assert np.array_equal(res.inner[c, :o["n_steps"]], o["inner"])
            assert mask_mismatches(res.active[c], o["mask"], o["V"], np.maximum(K[c] - o["S"], 0.))[0] == 0


def test_forward_and_reverse_sweeps_agree_to_rounding(qpde, handle, oracle):
    """With time-independent coefficients the two directions take the same steps up to the rounding of
    the accumulated times: same step counts here, values equal to ~1e-12, not bit for bit."""
    K = np.array([100., 90.])
    common = dict(axis=oracle.special_axis()[None, :] * K[:, None], refine=4, r=.04, sigma=.25, q=0., T=1., dt=1. / 128,
                  strike=K, scheme="rannacher", american=True, payoff="put")
    out = []
    for fwd in (False, True):
        with qpde.Plan(handle, qpde.BatchSpec(forward=fwd, **common)) as plan:
            plan.run(); out.append(plan.fetch())
    assert np.array_equal(out[0].n_steps, out[1].n_steps)
    assert rel_err(out[0].V, out[1].V).max() <= 1e-10


def test_forward_with_variable_steps_is_rejected(qpde, handle, oracle):
    spec = qpde.BatchSpec(axis=oracle.special_axis()[None, :] * 100., refine=2, r=.04, sigma=.2, q=0., T=1., dt=.01,
                          strike=[100.], forward=True, variable_target=[.1], max_steps=1000)
    with pytest.raises(qpde.QpdeError) as e:
        qpde.Plan(handle, spec)
    assert e.value.code == qpde.ERR_UNSUPPORTED
