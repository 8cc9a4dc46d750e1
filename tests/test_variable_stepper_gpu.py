"""GPU: ReverseVariableStepper (Core/Stepper.hpp:60-126, examples/vanilla_options.cpp:52-58;
SURVEY.md §8(f) item 2) through the C ABI.  The reference fixtures (variable_* cases of
tests/golden/reference_cases.npz) are covered by test_parity_gpu.py::test_golden_cases for
both kernel families; here: per-contract targets against the oracle, the step-capacity
status, validation."""
import numpy as np
import pytest

from helpers import REL_TOL, rel_err

pytestmark = pytest.mark.gpu


def _spec(qpde, po, K, vol, T, steps, target, refine, kernel, american=True, scheme="rannacher", max_steps=2000):
    axis = po.special_axis()[None, :] * np.asarray(K)[:, None]
    return qpde.BatchSpec(axis=axis, refine=refine, r=.04, sigma=vol, q=0., T=T, dt=np.asarray(T) / steps, strike=K,
                          scheme=scheme, american=american, payoff="put", kernel=kernel,
                          variable_target=target, max_steps=max_steps)


@pytest.mark.parametrize("kernel", ["interleaved", "resident"])
@pytest.mark.parametrize("scheme", ["rannacher", "bdf2"])
def test_per_contract_targets_vs_oracle(qpde, handle, oracle, kernel, scheme):
    rng = np.random.default_rng(5)
    Cn = 16
    K = rng.uniform(60, 140, Cn); vol = rng.uniform(.15, .5, Cn); T = rng.uniform(.3, 2., Cn)
    target = rng.uniform(.05, .5, Cn)
    spec = _spec(qpde, oracle, K, vol, T, 60, target, 2, kernel, scheme=scheme)
    with qpde.Plan(handle, spec) as plan:
        assert plan.kernel == kernel
        plan.run(); res = plan.fetch()
    assert len(set(res.n_steps.tolist())) > 4              # the contracts really take different numbers of steps
    for c in range(Cn):
        o = oracle.american_put(K[c], .04, vol[c], T[c], 60, 2, scheme=scheme, variable_target=target[c])
        assert res.status[c] == 0
        assert res.n_steps[c] == o["n_steps"], c
        err = rel_err(res.V[c], o["V"]).max()
        assert err <= REL_TOL, "contract %d: %.3e" % (c, err)
        assert np.array_equal(res.inner[c, :o["n_steps"]], o["inner"])


def test_step_capacity_is_reported(qpde, handle, oracle):
    K = np.array([100., 100.]); vol = np.array([.2, .2]); T = np.array([1., 1.])
    spec = _spec(qpde, oracle, K, vol, T, 192, np.array([1. / 16, 1. / 16]), 2, "auto", max_steps=10)   # needs 78 steps
    with qpde.Plan(handle, spec) as plan:
        plan.run(); res = plan.fetch()
    assert np.all(res.status == 3) and np.all(res.n_steps == 10)


def test_variable_stepper_validation(qpde, handle, oracle):
    K = np.array([100.]); vol = np.array([.2]); T = np.array([1.])
    with pytest.raises(qpde.QpdeError):       # max_steps is required
        qpde.Plan(handle, _spec(qpde, oracle, K, vol, T, 48, np.array([.25]), 1, "auto", max_steps=0))
    with pytest.raises(qpde.QpdeError):       # target > 0
        qpde.Plan(handle, _spec(qpde, oracle, K, vol, T, 48, np.array([0.]), 1, "auto"))
