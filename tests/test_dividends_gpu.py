"""GPU: discrete dividend events (general interpolating events, SURVEY.md §8(f) item 1;
README.md:357-375, Core/Event.hpp:100-112) through the C ABI.  The reference fixtures
(dividend_* cases of tests/golden/reference_cases.npz) are covered by
test_parity_gpu.py::test_golden_cases for both kernel families; here: per-contract
amounts against the oracle, the American + dividend combination, and properties."""
import numpy as np
import pytest

from helpers import REL_TOL, mask_mismatches, rel_err

pytestmark = pytest.mark.gpu


def _spec(qpde, po, K, r, alpha, T, steps, E, refine, kernel, dividend, scheme="bdf2", exercise="k_minus_s", **kw):
    Cn = len(K)
    axis = po.TUTORIAL_AXIS[None, :] * (np.asarray(K)[:, None] / 100.0)
    return qpde.BatchSpec(axis=axis, refine=refine, r=r, sigma=alpha, q=0.0, T=T, dt=np.asarray(T) / steps,
                          strike=K, scheme=scheme, american=False, payoff="put", vol_model="inverse_s",
                          n_exercise=E, exercise=exercise, kernel=kernel, dividend=dividend, **kw)


@pytest.mark.parametrize("kernel", ["interleaved", "resident"])
@pytest.mark.parametrize("kind,first", [("cash", False), ("cash", True), ("proportional", False)])
def test_per_contract_dividends_vs_oracle(qpde, handle, oracle, kernel, kind, first):
    rng = np.random.default_rng(7)
    Cn = 24
    K = rng.uniform(80, 120, Cn); r = rng.uniform(.01, .08, Cn); alpha = rng.uniform(10, 30, Cn)
    T = rng.uniform(.5, 1.5, Cn)
    D = rng.uniform(0., 3., Cn) if kind == "cash" else rng.uniform(0., .05, Cn)
    D[0] = 0.                                                    # a zero dividend is a valid event
    spec = _spec(qpde, oracle, K, r, alpha, T, 150, 6, 2, kernel, dict(kind=kind, amount=D, first=first))
    with qpde.Plan(handle, spec) as plan:
        plan.run(); res = plan.fetch()
    for c in range(Cn):
        o = oracle.bermudan_put(K[c], r[c], alpha[c], T[c], 150, 6, 2, scheme="bdf2", dividend=D[c],
                                dividend_kind=kind, dividend_first=first)
        assert res.status[c] == 0 and res.n_steps[c] == o["n_steps"]
        assert np.array_equal(res.S[c], o["S"])
        err = rel_err(res.V[c], o["V"]).max()
        assert err <= REL_TOL, "contract %d: %.3e" % (c, err)


@pytest.mark.parametrize("kernel", ["interleaved", "resident"])
def test_zero_dividend_is_the_identity_and_dividends_raise_put_values(qpde, handle, oracle, kernel):
    Cn = 8
    K = np.full(Cn, 100.); r = np.full(Cn, .04); alpha = np.full(Cn, 20.); T = np.full(Cn, 1.)
    D = np.linspace(0., 3.5, Cn)
    base = _spec(qpde, oracle, K, r, alpha, T, 100, 10, 1, kernel, None)
    div = _spec(qpde, oracle, K, r, alpha, T, 100, 10, 1, kernel, dict(kind="cash", amount=D))
    with qpde.Plan(handle, base) as p0, qpde.Plan(handle, div) as p1:
        p0.run(); p1.run()
        r0, r1 = p0.fetch(), p1.fetch()
    # V(max(S - 0, 0)) interpolates at the nodes themselves: weight 1 on the node, bit-identical
    assert np.array_equal(r1.V[0], r0.V[0])
    # a cash dividend lowers the stock, which can only help the put; more dividend, more value
    assert np.all(np.diff(r1.value) > 0)
    assert np.all(r1.V >= r0.V - 1e-12)


@pytest.mark.parametrize("kernel", ["resident", "interleaved"])
@pytest.mark.parametrize("refine,scheme,kind,exercise", [
    (2, "rannacher", "cash", "none"),          # 129 nodes: one warp per contract
    (4, "bdf2", "proportional", "none"),       # 513 nodes
    (6, "rannacher", "cash", "none"),          # 2049 nodes: the README tutorial's case at the BASELINE grid size
    (6, "bdf1", "cash", "k_minus_s"),          # an exercise transform on top of the penalty constraint
])
def test_american_put_with_dividends_vs_oracle(qpde, handle, oracle, kernel, refine, scheme, kind, exercise):
    """Penalty iteration + events (README.md:357-375 on an American put): both kernel families — the RESIDENT
    sweep rebuilds the active set from the transformed solution after every event."""
    Cn = 6
    rng = np.random.default_rng(11)
    K = rng.uniform(90, 110, Cn); vol = rng.uniform(.15, .4, Cn); T = np.full(Cn, 1.)
    D = rng.uniform(.5, 2., Cn) if kind == "cash" else rng.uniform(.005, .03, Cn)
    nsteps = 100 if refine < 6 else 200
    axis = oracle.special_axis()[None, :] * K[:, None]
    spec = qpde.BatchSpec(axis=axis, refine=refine, r=.04, sigma=vol, q=0., T=T, dt=T / nsteps, strike=K,
                          scheme=scheme, american=True, payoff="put", n_exercise=4, exercise=exercise,
                          dividend=dict(kind=kind, amount=D), kernel=kernel)
    with qpde.Plan(handle, spec) as plan:
        assert plan.kernel == kernel
        plan.run(); res = plan.fetch()
    for c in range(Cn):
        S = oracle.refine(oracle.axis_union(oracle.special_axis() * K[c], oracle.special_axis() * K[c]), refine)
        lo, di, up = oracle.bs_coeffs(S, .04, vol[c], 0.)
        payoff = np.where(S < K[c], K[c] - S, 0.)
        steps, n = oracle.schedule(1., 1. / nsteps, [1. / 4 * e for e in range(4)])
        o = oracle.sweep(S, lo, di, up, payoff, 1., steps, n, scheme=scheme, american=True, obstacle=payoff,
                         event_obstacle=(K[c] - S) if exercise != "none" else None, dividend=(kind, D[c], False))
        assert res.status[c] == 0 and res.n_steps[c] == n
        err = rel_err(res.V[c], o["V"]).max()
        assert err <= REL_TOL, "contract %d: %.3e" % (c, err)
        assert np.array_equal(res.inner[c, :n], o["inner"])
        assert mask_mismatches(res.active[c], o["mask"], o["V"], payoff)[0] == 0


def test_american_with_events_takes_the_resident_sweep_by_default(qpde, handle, oracle):
    K = np.array([100., 105.])
    spec = qpde.BatchSpec(axis=oracle.special_axis()[None, :] * K[:, None], refine=6, r=.04, sigma=.2, q=0., T=1.,
                          dt=1. / 50, strike=K, scheme="rannacher", american=True, payoff="put", n_exercise=2,
                          exercise="none", dividend=dict(kind="cash", amount=np.array([1., 1.])), kernel="auto")
    with qpde.Plan(handle, spec) as plan:
        assert plan.kernel == "resident"


def test_dividend_validation(qpde, handle, oracle):
    K = np.array([100.])
    with pytest.raises(qpde.QpdeError):       # proportional dividends must be < 1
        qpde.Plan(handle, _spec(qpde, oracle, K, .04, 20., np.array([1.]), 50, 4, 0, "auto",
                                dict(kind="proportional", amount=1.5)))
    with pytest.raises(qpde.QpdeError):       # dividends need event times
        qpde.Plan(handle, _spec(qpde, oracle, K, .04, 20., np.array([1.]), 50, 0, 0, "auto",
                                dict(kind="cash", amount=1.)))
