"""GPU: ReverseBDFThree .. ReverseBDFSix (Core/BDF.hpp:139-470,595-942) on the
INTERLEAVED family: start-up through the lower orders, restart after every event, variable-step
coefficients on a clipped last step, with and without the penalty iteration, policy iteration,
the variable stepper."""
import numpy as np
import pytest

from helpers import REL_TOL, rel_err

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("order", [3, 4, 5, 6])
def test_batch_vs_oracle(qpde, handle, oracle, order):
    scheme = "bdf%d" % order
    rng = np.random.default_rng(100 + order)
    Cn, refine, steps = 37, 3, 50
    K = rng.uniform(50, 150, Cn); vol = rng.uniform(.1, .6, Cn); T = rng.uniform(.25, 2., Cn)
    r = rng.uniform(.01, .08, Cn)
    sp = oracle.special_axis()
    axis = np.stack([oracle.axis_union(sp * k, sp * k) for k in K])
    for american in (False, True):
        spec = qpde.BatchSpec(axis=axis, refine=refine, r=r, sigma=vol, q=0., T=T, dt=T / steps, strike=K,
                              scheme=scheme, american=american, payoff="put")
        with qpde.Plan(handle, spec) as plan:
            assert plan.kernel == "interleaved"                 # AUTO: the RESIDENT family stops at BDF2
            plan.run()
            res = plan.fetch()
        assert np.all(res.status == 0)
        for c in (0, 17, Cn - 1):
            o = oracle.american_put(K[c], r[c], vol[c], T[c], steps, refine, american=american, scheme=scheme)
            assert res.n_steps[c] == o["n_steps"]
            err = rel_err(res.V[c], o["V"]).max()
            assert err <= REL_TOL, "%s contract %d: %.3e" % (scheme, c, err)
            assert np.array_equal(res.inner[c, :o["n_steps"]], o["inner"])
    with pytest.raises(qpde.QpdeError) as e:
        qpde.Plan(handle, qpde.BatchSpec(axis=axis, refine=refine, r=r, sigma=vol, q=0., T=T, dt=T / steps,
                                         strike=K, scheme=scheme, payoff="put", kernel="resident"))
    assert e.value.code == qpde.ERR_UNSUPPORTED


@pytest.mark.parametrize("order", [3, 6])
def test_bermudan_restart_after_events(qpde, handle, oracle, order):
    scheme = "bdf%d" % order
    Cn, refine, steps, E = 5, 2, 90, 6
    K = np.linspace(85., 115., Cn); alpha = np.linspace(12., 28., Cn); r = np.linspace(.02, .06, Cn)
    axis = np.stack([oracle.TUTORIAL_AXIS * (k / 100.) for k in K])
    spec = qpde.BatchSpec(axis=axis, refine=refine, r=r, sigma=alpha, q=0., T=1., dt=1. / steps, strike=K,
                          scheme=scheme, american=False, payoff="put", vol_model="inverse_s", n_exercise=E,
                          exercise="k_minus_s")
    with qpde.Plan(handle, spec) as plan:
        plan.run()
        res = plan.fetch()
    for c in range(Cn):
        o = oracle.bermudan_put(K[c], r[c], alpha[c], 1., steps, E, refine, scheme=scheme)
        assert res.n_steps[c] == o["n_steps"]
        assert rel_err(res.V[c], o["V"]).max() <= REL_TOL


def test_policy_iteration_and_variable_steps_under_bdf4(qpde, handle, oracle):
    K, vol, T, steps, refine = 100., .3, 1., 48, 2
    sp = oracle.special_axis()
    axis = oracle.axis_union(sp * K, sp * K)[None, :]
    o = oracle.hjb_straddle(K, .03, .05, vol, T, steps, refine, long_position=False, scheme="bdf4")
    spec = qpde.BatchSpec(axis=axis, refine=refine, r=0., sigma=vol, q=0., T=T, dt=T / steps, strike=[K],
                          scheme="bdf4", payoff="straddle", controls=np.array([[.03, .05]]), policy_max=False)
    with qpde.Plan(handle, spec) as plan:
        plan.run()
        res = plan.fetch()
    assert rel_err(res.V[0], o["V"]).max() <= REL_TOL
    assert np.array_equal(res.inner[0, :o["n_steps"]], o["inner"])
    # ReverseVariableStepper: every step has its own size, so every BDF4 coefficient is a variable-step one
    o = oracle.american_put(K, .04, .2, T, 40, refine, scheme="bdf4", variable_target=0.3)
    spec = qpde.BatchSpec(axis=axis, refine=refine, r=.04, sigma=.2, q=0., T=T, dt=T / 40, strike=[K],
                          scheme="bdf4", american=True, payoff="put", variable_target=0.3, max_steps=4000)
    with qpde.Plan(handle, spec) as plan:
        plan.run()
        res = plan.fetch()
    assert res.n_steps[0] == o["n_steps"] and res.status[0] == 0
    assert rel_err(res.V[0], o["V"]).max() <= REL_TOL
