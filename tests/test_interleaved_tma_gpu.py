"""GPU: the TMA-staged sweep of the INTERLEAVED family (qpde_interleaved_tma.cu) against the
plain-load sweep (bit for bit: same arithmetic, different staging) and against the oracle,
on ragged batches (last warp of contracts partly empty, last tile of rows partly empty),
every scheme, with and without the penalty iteration, with exercise events."""
import os

import numpy as np
import pytest

from helpers import REL_TOL, rel_err

pytestmark = pytest.mark.gpu

FIELDS = ("V", "S", "value", "n_steps", "inner_total", "inner", "active", "status")


def _run(qpde, handle, spec, tma):
    old = os.environ.get("QPDE_INTERLEAVED_NO_TMA")
    os.environ["QPDE_INTERLEAVED_NO_TMA"] = "0" if tma else "1"      # read at plan creation
    try:
        with qpde.Plan(handle, spec) as plan:
            assert plan.kernel == "interleaved" and plan.tma_staged == tma
            plan.run()
            return plan.fetch()
    finally:
        if old is None:
            del os.environ["QPDE_INTERLEAVED_NO_TMA"]
        else:
            os.environ["QPDE_INTERLEAVED_NO_TMA"] = old


def _batch(oracle, Cn, seed):
    rng = np.random.default_rng(seed)
    K = rng.uniform(50, 150, Cn); vol = rng.uniform(.1, .6, Cn)
    T = rng.uniform(.25, 2., Cn); r = rng.uniform(.01, .08, Cn)
    sp = oracle.special_axis()
    axis = np.stack([oracle.axis_union(sp * k, sp * k) for k in K])
    return K, vol, T, r, axis


@pytest.mark.parametrize("scheme", ["rannacher", "cn", "bdf1", "bdf2"])
@pytest.mark.parametrize("american", [True, False])
def test_tma_staging_is_bit_identical_to_plain_loads(qpde, handle, oracle, scheme, american):
    Cn, refine, steps = 70, 3, 40                   # 257 nodes = 64 tiles + 1 row; 70 = 2 warps + 6 contracts
    K, vol, T, r, axis = _batch(oracle, Cn, 11)
    spec = qpde.BatchSpec(axis=axis, refine=refine, r=r, sigma=vol, q=0.01, T=T, dt=T / steps, strike=K,
                          scheme=scheme, american=american, payoff="put", kernel="interleaved")
    a = _run(qpde, handle, spec, True)
    p = _run(qpde, handle, spec, False)
    for name in FIELDS:
        assert np.array_equal(getattr(a, name), getattr(p, name)), name
    assert np.all(a.status == 0)
    c = Cn - 1                                      # a contract of the ragged last warp against the oracle
    o = oracle.american_put(K[c], r[c], vol[c], T[c], steps, refine, q=0.01, american=american, scheme=scheme)
    assert np.array_equal(a.S[c], o["S"]) and a.n_steps[c] == o["n_steps"]
    assert rel_err(a.V[c], o["V"]).max() <= REL_TOL
    assert np.array_equal(a.inner[c, :o["n_steps"]], o["inner"])


@pytest.mark.parametrize("scheme", ["bdf2", "rannacher"])
def test_tma_bermudan_events(qpde, handle, oracle, scheme):
    """exercise dates (events between steps, BDF restart) on the tutorial grid, local volatility:
    260 nodes = exactly 65 tiles; contracts of one warp hit their events at the same steps"""
    Cn, refine, steps, E = 33, 3, 60, 5
    rng = np.random.default_rng(5)
    K = rng.uniform(80., 120., Cn); alpha = rng.uniform(10., 30., Cn); r = rng.uniform(.01, .08, Cn)
    axis = np.stack([oracle.TUTORIAL_AXIS * (k / 100.) for k in K])
    spec = qpde.BatchSpec(axis=axis, refine=refine, r=r, sigma=alpha, q=0., T=1., dt=1. / steps, strike=K,
                          scheme=scheme, american=False, payoff="put", vol_model="inverse_s", n_exercise=E,
                          exercise="k_minus_s", kernel="interleaved")
    a = _run(qpde, handle, spec, True)
    p = _run(qpde, handle, spec, False)
    for name in FIELDS:
        assert np.array_equal(getattr(a, name), getattr(p, name)), name
    for c in (0, Cn - 1):
        o = oracle.bermudan_put(K[c], r[c], alpha[c], 1., steps, E, refine, scheme=scheme)
        assert a.n_steps[c] == o["n_steps"]
        assert rel_err(a.V[c], o["V"]).max() <= REL_TOL


def test_tma_contracts_with_different_step_counts_share_a_warp(qpde, handle, oracle):
    """T / dt differs per contract: lanes of one warp finish after different numbers of steps and
    penalty iterations; the finished lanes ride along without storing."""
    Cn, refine = 40, 4                              # 513 nodes
    K, vol, T, r, axis = _batch(oracle, Cn, 3)
    steps = np.random.default_rng(3).integers(10, 60, Cn)
    spec = qpde.BatchSpec(axis=axis, refine=refine, r=r, sigma=vol, q=0., T=T, dt=T / steps, strike=K,
                          american=True, payoff="put", kernel="interleaved")
    a = _run(qpde, handle, spec, True)
    p = _run(qpde, handle, spec, False)
    for name in FIELDS:
        assert np.array_equal(getattr(a, name), getattr(p, name)), name
    assert np.all(a.n_steps >= steps) and np.all(a.n_steps <= steps + 1)
    for c in (int(np.argmin(steps)), int(np.argmax(steps))):
        o = oracle.american_put(K[c], r[c], vol[c], T[c], int(steps[c]), refine)
        assert rel_err(a.V[c], o["V"]).max() <= REL_TOL
        assert np.array_equal(a.inner[c, :o["n_steps"]], o["inner"])


@pytest.mark.parametrize("kind,first", [("cash", False), ("cash", True), ("proportional", False)])
def test_tma_dividend_events(qpde, handle, oracle, kind, first):
    """discrete dividends at the exercise dates (interpolating events, ordinary loads inside the
    TMA-staged sweep), per-contract amounts, both orders relative to the exercise transform"""
    Cn, refine, steps, E = 35, 3, 60, 4
    rng = np.random.default_rng(21)
    K = rng.uniform(80., 120., Cn); alpha = rng.uniform(10., 30., Cn); r = rng.uniform(.01, .08, Cn)
    D = rng.uniform(.2, 2., Cn) if kind == "cash" else rng.uniform(.005, .03, Cn)
    axis = np.stack([oracle.TUTORIAL_AXIS * (k / 100.) for k in K])
    spec = qpde.BatchSpec(axis=axis, refine=refine, r=r, sigma=alpha, q=0., T=1., dt=1. / steps, strike=K,
                          scheme="bdf2", american=False, payoff="put", vol_model="inverse_s", n_exercise=E,
                          exercise="k_minus_s", kernel="interleaved",
                          dividend=dict(kind=kind, amount=D, first=first))
    a = _run(qpde, handle, spec, True)
    p = _run(qpde, handle, spec, False)
    for name in FIELDS:
        assert np.array_equal(getattr(a, name), getattr(p, name)), name
    c = Cn - 1
    o = oracle.bermudan_put(K[c], r[c], alpha[c], 1., steps, E, refine, scheme="bdf2", dividend=D[c],
                            dividend_kind=kind, dividend_first=first)
    assert a.n_steps[c] == o["n_steps"] and rel_err(a.V[c], o["V"]).max() <= REL_TOL


@pytest.mark.parametrize("scheme", ["rannacher", "bdf2"])
def test_tma_variable_stepper(qpde, handle, oracle, scheme):
    """ReverseVariableStepper: every contract of a warp has its own step sizes and step count"""
    Cn, refine = 36, 3
    K, vol, T, r, axis = _batch(oracle, Cn, 9)
    target = np.random.default_rng(9).uniform(.1, .4, Cn)
    spec = qpde.BatchSpec(axis=axis, refine=refine, r=r, sigma=vol, q=0., T=T, dt=T / 30, strike=K, scheme=scheme,
                          american=True, payoff="put", kernel="interleaved", variable_target=target, max_steps=600)
    a = _run(qpde, handle, spec, True)
    p = _run(qpde, handle, spec, False)
    for name in FIELDS:
        assert np.array_equal(getattr(a, name), getattr(p, name)), name
    assert np.all(a.status == 0) and len(set(a.n_steps.tolist())) > 3
    for c in (0, Cn - 1):
        o = oracle.american_put(K[c], r[c], vol[c], T[c], 30, refine, scheme=scheme, variable_target=target[c])
        assert a.n_steps[c] == o["n_steps"]
        assert rel_err(a.V[c], o["V"]).max() <= REL_TOL
        assert np.array_equal(a.inner[c, :o["n_steps"]], o["inner"])


def test_tma_full_occupancy_is_bit_identical(qpde, handle):
    """Many waves of warps (seven CTAs per SM, every SM busy, HBM saturated): the ordering between the
    ordinary stores of one pass and the TMA loads of the next must hold under load, not only in a
    quiet machine.  20,011 contracts x 513 nodes x 60 steps, every output against the plain-load sweep."""
    Cn, refine, steps = 20011, 4, 60
    rng = np.random.default_rng(77)
    K = rng.uniform(50, 150, Cn); vol = rng.uniform(.1, .6, Cn); T = rng.uniform(.25, 2., Cn); r = rng.uniform(.01, .08, Cn)
    from bench import SPECIAL
    spec = qpde.BatchSpec(axis=np.ascontiguousarray(K[:, None] * SPECIAL[None, :]), refine=refine, r=r, sigma=vol, q=0., T=T,
                          dt=T / steps, strike=K, american=True, payoff="put", kernel="interleaved",
                          outputs=qpde.OUT_V | qpde.OUT_VALUE | qpde.OUT_INNER_TOTAL | qpde.OUT_STATUS | qpde.OUT_STEPS)
    a = _run(qpde, handle, spec, True)
    p = _run(qpde, handle, spec, False)
    for name in ("V", "value", "n_steps", "inner_total", "status"):
        assert np.array_equal(getattr(a, name), getattr(p, name)), name
    assert np.all(a.status == 0) and a.inner_total.min() >= steps


@pytest.mark.parametrize("n_nodes", [32, 33, 34, 35, 37, 63])
@pytest.mark.parametrize("Cn", [1, 31, 33])
def test_tma_ragged_tiles_and_warps(qpde, handle, n_nodes, Cn):
    """every residue of N modulo the 4-row tile (incl. the smallest grid the ring takes, 32 nodes) and
    batches below / above one warp of contracts: equal to the plain-load sweep bit for bit"""
    rng = np.random.default_rng(1000 * n_nodes + Cn)
    K = rng.uniform(80., 120., Cn)
    base = np.concatenate([[0.], np.sort(rng.uniform(.3, 2.5, n_nodes - 2)), [6.]])
    axis = K[:, None] * base[None, :]
    spec = qpde.BatchSpec(axis=axis, refine=0, r=.04, sigma=rng.uniform(.15, .4, Cn), q=0., T=1., dt=1. / 25, strike=K,
                          scheme="rannacher", american=True, payoff="put", kernel="interleaved")
    a = _run(qpde, handle, spec, True)
    p = _run(qpde, handle, spec, False)
    for name in FIELDS:
        assert np.array_equal(getattr(a, name), getattr(p, name)), name
    assert np.all(a.status == 0) and np.all(np.isfinite(a.V))
