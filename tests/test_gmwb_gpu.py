"""GPU: the 2-D GMWB impulse-control HJBQVI (BASELINE.json configs[4] path) through the
C ABI against the reference's own outputs (tests/golden/gmwb_reference.npz, made from
examples/hjbqvi/gmwb.cpp on the unmodified reference headers) and against the oracle."""
import os

import numpy as np
import pytest

from helpers import REL_TOL

pytestmark = pytest.mark.gpu


def _solve(qpde, handle, oracle, refine, a_points=51, timesteps=32, **kw):
    return qpde.gmwb_solve(handle, oracle.GMWB_W_AXIS, oracle.uniform_axis(0., 100., a_points), [0., 10.],
                           oracle.uniform_axis(0., 1., 3), refine, timesteps * 2 ** refine, **kw)


def test_gmwb_against_reference_fixtures(qpde, handle, oracle):
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "gmwb_reference.npz"))
    for name, kw in (("gmwb_small", dict(refine=0, a_points=11, timesteps=8)),
                     ("gmwb_k0", dict(refine=0)), ("gmwb_k1", dict(refine=1))):
        res = _solve(qpde, handle, oracle, **kw)
        ref = g[name + "/V"].reshape(res["V"].shape)
        assert res["status"] == 0 and res["n_steps"] == int(g[name + "/steps"])
        assert np.array_equal(res["W"], g[name + "/W"]) and np.array_equal(res["A"], g[name + "/A"])
        err = np.abs(res["V"] - ref) / np.maximum(1., np.abs(ref))
        assert err.max() <= REL_TOL, "%s: %.3e" % (name, err.max())
        assert res["mean_inner"] == float(g[name + "/mean_inner"]), name     # "Mean Policy Iterations"


def test_gmwb_refinement_two_against_oracle(qpde, handle, oracle):
    """257 x 201 nodes, 128 steps: the W-line solve uses the tail equation (257 = 32 * 8 + 1)."""
    res = _solve(qpde, handle, oracle, refine=2)
    o = oracle.gmwb(refine_times=2)
    assert res["status"] == 0 and o["status"] == 0 and res["n_steps"] == o["n_steps"]
    err = np.abs(res["V"] - o["V"]) / np.maximum(1., np.abs(o["V"]))
    assert err.max() <= REL_TOL, "%.3e" % err.max()
    assert res["mean_inner"] == o["mean_inner"]
    # convergence table value at (w, a) = (100, 100): the no-fee GMWB value of the literature
    iw, ia = np.searchsorted(res["W"], 100.), np.searchsorted(res["A"], 100.)
    assert res["W"][iw] == 100. and res["A"][ia] == 100. and 107.6 < res["V"][ia, iw] < 107.8


def test_gmwb_properties_and_errors(qpde, handle, oracle):
    res = _solve(qpde, handle, oracle, refine=1, sigma=0.3)
    payoff = np.maximum(res["W"][None, :], 0.9 * res["A"][:, None])
    assert np.all(res["V"] >= payoff - 1e-6 * np.maximum(1., payoff))      # no fee: V dominates the exit payoff
    assert np.all(np.diff(res["V"], axis=1) >= -1e-9)                      # increasing in the account value
    fee = _solve(qpde, handle, oracle, refine=1, sigma=0.3, eta=0.01)
    assert np.all(fee["V"] <= res["V"] + 1e-9)                             # a fee can only lower the value
    with pytest.raises(qpde.QpdeError):
        qpde.gmwb_solve(handle, oracle.GMWB_W_AXIS + 1.0, oracle.uniform_axis(0., 100., 11), [0., 10.],
                        oracle.uniform_axis(0., 1., 3), 0, 8)
