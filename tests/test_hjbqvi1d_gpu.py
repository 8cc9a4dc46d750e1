"""GPU: the 1-D impulse-control HJBQVI of examples/hjbqvi/exchange_rate.cpp through the C ABI
(qpde_hjbqvi1d_solve, one CTA per problem) against the reference's own outputs
(tests/golden/hjbqvi1d_reference.npz, made by `qpde_ref exchange` on the unmodified reference headers)
and against the oracle."""
import os

import numpy as np
import pytest

from helpers import REL_TOL
from test_oracle_golden import _er_kwargs

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "hjbqvi1d_reference.npz")
DEFAULTS = dict(rho=.02, sigma=.3, a=.25, b=3., m=0., lam=1., c=.1)


def _params(kw):
    p = dict(DEFAULTS); p.update({k: v for k, v in kw.items() if k in DEFAULTS})
    return [p[k] for k in ("rho", "sigma", "a", "b", "m", "lam", "c")]


def test_exchange_rate_against_reference_fixtures(qpde, handle):
    g = np.load(GOLDEN)
    for name in g["names"]:
        name = str(name)
        kw = _er_kwargs(eval(str(g[name + "/args"])))
        res = qpde.exchange_rate_solve(handle, g[name + "/X0"], g[name + "/Q0"], g[name + "/Z0"], [_params(kw)],
                                       kw["refine_times"], kw["timesteps"] * 2 ** kw["refine_times"], T=kw.get("T", 10.))
        assert res["status"][0] == 0 and res["n_steps"][0] == int(g[name + "/steps"]), name
        assert np.array_equal(res["X"][0], g[name + "/X"]), name
        ref = g[name + "/V"]
        err = np.abs(res["V"][0] - ref) / np.maximum(1., np.abs(ref))
        assert err.max() <= REL_TOL, "%s: %.3e" % (name, err.max())
        assert res["mean_inner"][0] == float(g[name + "/mean_inner"]), name     # "Mean Policy Iterations"
        # the impulse region and both control vectors, NaN where the other control acts
        assert np.array_equal(res["mask"][0] != 0, np.isnan(g[name + "/Q"])), name
        for key in ("Q", "Z"):
            assert np.array_equal(res[key][0], g[name + "/" + key], equal_nan=True), (name, key)


def test_exchange_rate_batch_of_problems_against_oracle(qpde, handle, oracle):
    """A batch with per-problem parameters and per-problem grids (the parity m moves the clustering) in one
    launch; refinement 5 (993 nodes, 512 steps) for one of them against the oracle."""
    g = np.load(GOLDEN)
    X0, Q0, Z0 = g["er_k0/X0"], g["er_k0/Q0"], g["er_k0/Z0"]
    rng = np.random.default_rng(7)
    P = 12
    par = np.array([[.02 + .03 * rng.random(), .2 + .2 * rng.random(), .1 + .4 * rng.random(), 1. + 3. * rng.random(),
                     0., .5 + rng.random(), .05 + .1 * rng.random()] for _ in range(P)])
    res = qpde.exchange_rate_solve(handle, X0, Q0, Z0, par, 2, 64)
    for p in range(P):
        o = oracle.exchange_rate(X0, Q0, Z0, refine_times=2, **dict(zip(("rho", "sigma", "a", "b", "m", "lam", "c"), par[p])))
        assert res["status"][p] == 0 and o["status"] == 0 and res["n_steps"][p] == o["n_steps"]
        err = np.abs(res["V"][p] - o["V"]) / np.maximum(1., np.abs(o["V"]))
        assert err.max() <= REL_TOL, "problem %d: %.3e" % (p, err.max())
        assert res["mean_inner"][p] == o["mean_inner"], p
        assert np.array_equal(res["mask"][p], o["mask"]), p
    # per-problem axes: the off-centre fixture next to the centred one
    Xs = np.stack([g["er_k1/X0"], g["er_parity_off_centre/X0"]]); Zs = np.stack([g["er_k1/Z0"], g["er_parity_off_centre/Z0"]])
    two = qpde.exchange_rate_solve(handle, Xs, g["er_k1/Q0"], Zs, [_params({}), _params(dict(m=.25))], 1, 32)
    for p, name in enumerate(("er_k1", "er_parity_off_centre")):
        err = np.abs(two["V"][p] - g[name + "/V"]) / np.maximum(1., np.abs(g[name + "/V"]))
        assert err.max() <= REL_TOL and two["mean_inner"][p] == float(g[name + "/mean_inner"]), name
    big = qpde.exchange_rate_solve(handle, X0, Q0, Z0, [_params({})], 5, 512)
    o = oracle.exchange_rate(X0, Q0, Z0, refine_times=5)
    err = np.abs(big["V"][0] - o["V"]) / np.maximum(1., np.abs(o["V"]))
    assert big["status"][0] == 0 and err.max() <= REL_TOL and big["mean_inner"][0] == o["mean_inner"]


def test_exchange_rate_errors(qpde, handle):
    g = np.load(GOLDEN)
    X0, Q0, Z0 = g["er_k0/X0"], g["er_k0/Q0"], g["er_k0/Z0"]
    with pytest.raises(qpde.QpdeError):      # 7 refinements: 3969 nodes
        qpde.exchange_rate_solve(handle, X0, Q0, Z0, [_params({})], 7, 2048)
    with pytest.raises(qpde.QpdeError):
        qpde.exchange_rate_solve(handle, X0[::-1], Q0, Z0, [_params({})], 0, 16)
    with pytest.raises(qpde.QpdeError):
        qpde.exchange_rate_solve(handle, X0, Q0, Z0, [[.02, .3, .25]], 0, 16)


def test_example_program_reproduces_the_reference_programs_stdout():
    """examples/bin/exchange_rate_b200 against the recorded stdout of the reference's own, unmodified
    examples/hjbqvi/exchange_rate.cpp (oracle/_ref/examples/exchange_rate: 7 minutes on one host core;
    tests/golden/tables_hjbqvi/exchange_rate.txt): the convergence table over the refinement levels 0 .. 6
    (2049 nodes, 513 + 1025 controls, 1024 steps at the last) and the solution with both controls at every node
    of the finest level.  Equal character for character except the two columns that cannot be ("Mean Solver
    Iterations": BiCGSTAB's count there, tridiagonal solves here; the execution time); "Change" / "Ratio",
    differences of values that agree to ~1e-12, to 2e-6 relative; and the 12th printed digit of a node value
    (the two linear solvers stop at different iterates of the same system: 1e-12 relative apart)."""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    subprocess.run(["make", "-s", "-C", os.path.join(root, "examples")], check=True)
    got = subprocess.run([os.path.join(root, "examples", "bin", "exchange_rate_b200")], check=True, capture_output=True,
                         text=True).stdout.splitlines()
    want = open(os.path.join(root, "tests", "golden", "tables_hjbqvi", "exchange_rate.txt")).read().splitlines()
    assert len(got) == len(want)
    assert got[0] == want[0]
    blank = want.index("")
    for g, w in zip(got[1:blank], want[1:blank]):
        assert len(g) == len(w) == 12 * 23
        gf, wf = [g[23 * k:23 * (k + 1)] for k in range(12)], [w[23 * k:23 * (k + 1)] for k in range(12)]
        assert gf[:7] == wf[:7], (g, w)                          # nodes, steps, scaling, tolerance, policy iterations
        assert gf[8] == wf[8] or abs(float(gf[8]) - float(wf[8])) <= 2e-11 * abs(float(wf[8])), (g, w)   # value: 12th digit
        for a, b in zip(gf[9:11], wf[9:11]):
            assert a == b or abs(float(a) - float(b)) <= 2e-6 * abs(float(b)), (g, w)
    assert got[blank:blank + 3] == want[blank:blank + 3]             # blank lines and the header of the node dump
    off = 0
    for g, w in zip(got[blank + 3:], want[blank + 3:]):              # 2049 rows: x, u, q*, z*
        if g == w:
            continue
        off += 1
        assert len(g) == len(w) and g[:23] == w[:23] and g[46:] == w[46:], (g, w)          # node and both controls
        assert abs(float(g[23:46]) - float(w[23:46])) <= 2e-11 * abs(float(w[23:46])), (g, w)
    assert off <= len(want) // 4


def test_exchange_rate_edge_parameters_against_oracle(qpde, handle, oracle):
    """Corners of the model in one batch: interventions that never pay (c = 5) or nearly always do (tiny costs),
    no diffusion at all (sigma = 0: the rows are pure upwinding), no discounting, a free stochastic control
    (b = 0).  The oracle agrees with the reference on each of them (checked when the fixtures were made:
    <= 8e-13, identical policy-iteration counts and impulse regions)."""
    g = np.load(GOLDEN)
    X0, Q0, Z0 = g["er_k0/X0"], g["er_k0/Q0"], g["er_k0/Z0"]
    cases = [dict(c=5.), dict(c=1e-3, lam=.05), dict(sigma=0.), dict(sigma=1e-3, a=2.), dict(lam=0., c=1e-6), dict(rho=0.), dict(b=0.)]
    res = qpde.exchange_rate_solve(handle, X0, Q0, Z0, [_params(kw) for kw in cases], 2, 64)
    for p, kw in enumerate(cases):
        o = oracle.exchange_rate(X0, Q0, Z0, refine_times=2, **kw)
        assert res["status"][p] == 0 and o["status"] == 0, kw
        err = np.abs(res["V"][p] - o["V"]) / np.maximum(1., np.abs(o["V"]))
        assert err.max() <= REL_TOL, "%s: %.3e" % (kw, err.max())
        assert res["mean_inner"][p] == o["mean_inner"], kw
        assert np.array_equal(res["mask"][p], o["mask"]), kw
    # one-point control grids: no stochastic control to choose, a single intervention target
    one = qpde.exchange_rate_solve(handle, X0, [0.], [0.], [_params({})], 1, 32)
    o = oracle.exchange_rate(X0, [0.], [0.], refine_times=1)
    err = np.abs(one["V"][0] - o["V"]) / np.maximum(1., np.abs(o["V"]))
    assert one["status"][0] == 0 and err.max() <= REL_TOL and one["mean_inner"][0] == o["mean_inner"]
    assert one["n_controls"] == 1 and one["n_impulses"] == 1


@pytest.mark.parametrize("nodes", [5, 33, 512, 513, 768, 769, 1280, 2048, 2049, 2304])
def test_exchange_rate_grid_sizes_at_the_geometry_boundaries(qpde, handle, oracle, nodes):
    """Every rows-per-thread geometry (2, 3, 5, 8, 9 rows x 256 threads) with and without padding rows, down to a
    5-node grid, against the oracle: uniform axes, 4 time steps."""
    X0 = np.linspace(-2., 2., nodes)
    Q0 = np.linspace(-.07, .07, 3)
    Z0 = np.linspace(-1.5, 1.5, 7)
    res = qpde.exchange_rate_solve(handle, X0, Q0, Z0, [_params({}), _params(dict(c=.02, lam=.3))], 0, 4)
    for p, kw in enumerate((dict(), dict(c=.02, lam=.3))):
        o = oracle.exchange_rate(X0, Q0, Z0, refine_times=0, timesteps=4, **kw)
        assert res["status"][p] == 0 and o["status"] == 0
        err = np.abs(res["V"][p] - o["V"]) / np.maximum(1., np.abs(o["V"]))
        assert err.max() <= REL_TOL, "%d nodes: %.3e" % (nodes, err.max())
        assert res["mean_inner"][p] == o["mean_inner"]
        assert np.array_equal(res["mask"][p], o["mask"])


def test_exchange_rate_rejects_grids_beyond_the_largest_geometry(qpde, handle):
    with pytest.raises(qpde.QpdeError):
        qpde.exchange_rate_solve(handle, np.linspace(-2., 2., 2305), [0.], [0.], [_params({})], 0, 4)
