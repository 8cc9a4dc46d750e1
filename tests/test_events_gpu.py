"""GPU: general Event<1> transforms as event programs (Core/Event.hpp:60-184,
examples/experimental/glwb.cpp:84-121) through the C ABI.  The reference fixtures (event_glwb_* cases of
tests/golden/reference_cases.npz: the reference's own Event<1> with the GLWB three-way lambda) go through
test_parity_gpu.py::test_golden_cases for both kernel families; here: per-contract parameter tables against
the oracle, the tagged transforms re-expressed as programs, and validation."""
import numpy as np
import pytest

from helpers import REL_TOL, rel_err

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("kernel", ["resident", "interleaved"])
@pytest.mark.parametrize("refine,scheme,american,forward", [(2, "bdf2", False, False), (4, "rannacher", True, False),
                                                            (6, "bdf1", False, False), (3, "cn", False, True)])
def test_glwb_event_with_per_contract_parameters(qpde, handle, oracle, kernel, refine, scheme, american, forward):
    Cn, E, steps = 5, 6, 90
    rng = np.random.default_rng(5)
    K = rng.uniform(80, 120, Cn); vol = rng.uniform(.15, .4, Cn)
    ev = oracle.glwb_event(E)
    params = np.stack([ev["params"] * np.array([1. - .02 * c, 1. + .1 * c]) for c in range(Cn)])      # [C][E][2]
    axis = oracle.special_axis()[None, :] * K[:, None]
    spec = qpde.BatchSpec(axis=axis, refine=refine, r=.04, sigma=vol, q=0., T=1., dt=1. / steps, strike=K, scheme=scheme,
                          american=american, payoff="put", n_exercise=E, exercise="none", kernel=kernel, forward=forward,
                          event=dict(program=ev["program"], consts=ev["consts"], params=params))
    with qpde.Plan(handle, spec) as plan:
        assert plan.kernel == kernel
        plan.run(); res = plan.fetch()
    for c in range(Cn):
        evc = dict(program=ev["program"], consts=ev["consts"], params=params[c])
        if forward:
            S = oracle.vanilla_grid(K[c], None, refine)
            lo, di, up = oracle.bs_coeffs(S, .04, vol[c], 0.)
            payoff = np.where(S < K[c], K[c] - S, 0.)
            sched, n = oracle.schedule(1., 1. / steps, [1. / E * e for e in range(1, E + 1)], forward=True)
            o = oracle.sweep(S, lo, di, up, payoff, 1., sched, n, scheme=scheme, american=american, obstacle=payoff,
                             forward=True, event=evc)
            o["n_steps"] = n
        else:
            o = oracle.american_put(K[c], .04, vol[c], 1., steps, refine, american=american, scheme=scheme, E=E, event=evc)
        assert res.status[c] == 0 and res.n_steps[c] == o["n_steps"]
        err = rel_err(res.V[c], o["V"]).max()
        assert err <= REL_TOL, "contract %d: %.3e" % (c, err)
        if american:
            assert np.array_equal(res.inner[c, :o["n_steps"]], o["inner"])


@pytest.mark.parametrize("kernel", ["resident", "interleaved"])
def test_tagged_transforms_as_programs_give_the_same_bits(qpde, handle, oracle, kernel):
    """max(V(S), K - S) and V(max(S - D, 0)) written as programs (K and D per contract through event_params)
    against the tagged exercise / cash-dividend transforms: same interpolation, same bits."""
    Cn, E = 4, 5
    K = np.array([95., 100., 105., 110.]); D = np.array([.5, 1., 1.5, 2.])
    common = dict(axis=oracle.TUTORIAL_AXIS[None, :] * (K[:, None] / 100.), refine=3, r=.04, sigma=20., vol_model="inverse_s",
                  q=0., T=1., dt=1. / 100, strike=K, scheme="bdf2", payoff="put", n_exercise=E, kernel=kernel)
    pe = np.repeat(K[:, None, None], E, axis=1)
    pd = np.repeat(D[:, None, None], E, axis=1)
    cases = [(dict(exercise="k_minus_s"),
              dict(program=qpde.event_program("S", "V", ("param", 0), "S", "sub", "max"), params=pe)),
             (dict(exercise="none", dividend=dict(kind="cash", amount=D)),
              dict(program=qpde.event_program("S", ("param", 0), "sub", ("const", 0), "max", "V"), consts=[0.], params=pd))]
    for tagged, program in cases:
        with qpde.Plan(handle, qpde.BatchSpec(**common, **tagged)) as plan:
            plan.run(); a = plan.fetch()
        with qpde.Plan(handle, qpde.BatchSpec(**common, exercise="none", event=program)) as plan:
            plan.run(); b = plan.fetch()
        assert np.array_equal(a.n_steps, b.n_steps)
        assert np.array_equal(a.V, b.V)


def test_event_program_validation(qpde, handle, oracle):
    base = dict(axis=oracle.special_axis()[None, :] * 100., refine=1, r=.04, sigma=.2, q=0., T=1., dt=.02, strike=[100.],
                n_exercise=2, exercise="none")
    bad = [qpde.event_program("add"),                                     # stack underflow
           qpde.event_program("S", "S"),                                  # two values left
           qpde.event_program(("const", 3), "V"),                         # constant index out of range
           np.array([1, 99, 0], dtype=np.int32),                          # unknown op
           qpde.event_program(*["S"] * 9 + ["add"] * 8)]                  # deeper than QPDE_EV_MAX_STACK
    for prog in bad:
        with pytest.raises(qpde.QpdeError):
            qpde.Plan(handle, qpde.BatchSpec(event=dict(program=prog, consts=[1.], params=np.zeros((2, 0))), **base))
    with pytest.raises(qpde.QpdeError):                                   # a program and a tagged transform
        qpde.Plan(handle, qpde.BatchSpec(**dict(base, exercise="k_minus_s"),
                                         event=dict(program=qpde.event_program("S", "V"), params=np.zeros((2, 0)))))
