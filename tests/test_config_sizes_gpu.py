"""GPU: node-level parity AT the BASELINE.json config sizes ("pin parity at the
BASELINE config sizes").  The fixtures of the reference itself at these sizes (configs[0] k = 4, 5;
configs[2] 2113 x 6400 x 10 events; configs[3] 1057 x 385 and 2113 x 768) run through
test_parity_gpu.py::test_golden_cases; this file holds the live comparisons with the oracle, which
is bit-identical to the reference on every one of those fixtures (test_oracle_golden.py)."""
import os
import sys

import numpy as np
import pytest

from helpers import REL_TOL, mask_mismatches, rel_err

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def test_config0_table_k6_against_oracle(qpde, handle, oracle):
    """examples/vanilla_options.cpp with maximum_refinement 6: 2049 nodes x 49,152 steps
    (vanilla_options.cpp:36-42,139), every node, every per-step iteration count."""
    k = 6
    steps = 12 * 4 ** k
    o = oracle.american_put(100., .04, .2, 1., steps, k)
    axis = oracle.axis_union(oracle.special_axis() * 100., oracle.special_axis() * 100.)
    spec = qpde.BatchSpec(axis=np.tile(axis, (2, 1)), refine=k, r=.04, sigma=.2, q=0., T=1., dt=1. / steps,
                          strike=[100., 100.], american=True, payoff="put", outputs=qpde.OUT_ALL)
    with qpde.Plan(handle, spec) as plan:
        plan.run()
        res = plan.fetch()
    assert plan.N == 2049 and res.n_steps[0] == o["n_steps"] and res.status[0] == 0
    assert rel_err(res.V[0], o["V"]).max() <= REL_TOL
    assert np.array_equal(res.inner[0, :o["n_steps"]], o["inner"])
    assert mask_mismatches(res.active[0], o["mask"], o["V"], np.maximum(100. - o["S"], 0.))[0] == 0
    assert np.array_equal(res.V[0], res.V[1])
    # the convergence table's last column: ratio of successive changes -> 4 (second order)
    v = [float(oracle.interp(*[oracle.american_put(100., .04, .2, 1., 12 * 4 ** j, j)[q] for q in ("S", "V")], 100.))
         for j in (3, 4, 5)] + [float(res.value[0])]
    ratio = (v[2] - v[1]) / (v[3] - v[2])
    assert 3.0 < ratio < 5.0, ratio


def test_config1_stratified_sample_of_the_bench_batch(qpde, handle, oracle):
    """configs[1] as bench.py runs it: the full 65,536-contract batch in one launch, then >= 256 of its
    contracts (first / last wave, the extremes of strike, volatility, expiry and rate, evenly spaced
    rest) node for node against the oracle, with inner-iteration totals and step counts; active sets
    and per-step counts from a plan of the sampled contracts alone, whose V must equal the batch's bits."""
    sys.path.insert(0, ROOT)
    import bench
    Cn = 65536
    K, vol, T, r = bench.contract_parameters(Cn)
    axis = np.ascontiguousarray(K[:, None] * bench.SPECIAL[None, :])
    spec = qpde.BatchSpec(axis=axis, refine=6, r=r, sigma=vol, q=0., T=T, dt=T / 1000, strike=K, american=True,
                          payoff="put", outputs=qpde.OUT_V | qpde.OUT_INNER_TOTAL | qpde.OUT_STATUS | qpde.OUT_STEPS
                          | qpde.OUT_COMPUTED)
    with qpde.Plan(handle, spec) as plan:
        plan.run()
        res = plan.fetch()
    assert np.all(res.status == 0)
    assert np.all(res.computed <= res.inner_total) and np.all(res.computed >= res.n_steps)
    idx = bench.parity_sample_indices(Cn, K, vol, T, r, 256)
    assert len(idx) == 256 and idx[0] == 0 and idx[-1] == Cn - 1
    assert {int(np.argmin(vol)), int(np.argmax(vol)), int(np.argmin(T)), int(np.argmax(T))} <= set(idx.tolist())
    sub = qpde.BatchSpec(axis=axis[idx], refine=6, r=r[idx], sigma=vol[idx], q=0., T=T[idx], dt=T[idx] / 1000,
                         strike=K[idx], american=True, payoff="put")
    with qpde.Plan(handle, sub) as plan:
        plan.run()
        sres = plan.fetch()
    assert np.array_equal(sres.V, res.V[idx])            # a contract's bits do not depend on its batch
    p = bench.parity_sample(idx, K, vol, T, r, res.V[idx], res.inner_total[idx], res.n_steps[idx], active=sres.active)
    print("parity_sample:", p)
    assert p["worst_rel_err"] <= REL_TOL
    assert p["inner_totals_equal"] == 256 and p["step_counts_equal"] == 256
    assert p["active_set_mismatches"] == 0


@pytest.mark.parametrize("scheme", ["bdf2", "rannacher"])
@pytest.mark.parametrize("kernel", ["resident", "interleaved"])
def test_config2_bermudan_at_size(qpde, handle, oracle, scheme, kernel):
    """configs[2] at size: tutorial.cpp Bermudan puts with local volatility alpha / S on the 34-tick
    axis refined 6 times (2113 nodes), 6400 steps, 10 exercise dates, varied (alpha, r, K) drawn as
    bench.py --config 3 draws them; every node against the oracle."""
    Cn, steps = 5, 6400
    rng = np.random.default_rng(1234)
    alpha = rng.uniform(10, 30, 16384)[:Cn]; r = rng.uniform(.01, .08, 16384)[:Cn]; K = rng.uniform(80, 120, 16384)[:Cn]
    axis = K[:, None] / 100.0 * oracle.TUTORIAL_AXIS[None, :]
    spec = qpde.BatchSpec(axis=axis, refine=6, r=r, sigma=alpha, q=0., T=1., dt=1. / steps, strike=K, scheme=scheme,
                          payoff="put", vol_model="inverse_s", n_exercise=10, exercise="k_minus_s", kernel=kernel,
                          outputs=qpde.OUT_ALL)
    with qpde.Plan(handle, spec) as plan:
        assert plan.N == 2113
        plan.run()
        res = plan.fetch()
    assert np.all(res.status == 0)
    for c in range(Cn):
        o = oracle.bermudan_put(K[c], r[c], alpha[c], 1., steps, 10, 6, scheme=scheme)
        assert np.array_equal(res.S[c], o["S"])
        assert res.n_steps[c] == o["n_steps"]
        err = rel_err(res.V[c], o["V"]).max()
        assert err <= REL_TOL, "contract %d: %.3e" % (c, err)
        assert rel_err(res.value[c], oracle.interp(o["S"], o["V"], K[c])) <= REL_TOL


def test_config3_jump_at_size_varied(qpde, handle, oracle):
    """configs[3] at size: Merton jump calls, 2113 nodes x 768 BDF1 steps, n_fft 4096, four (mu, sd)
    pairs of bench.py --config 4; every node against the oracle."""
    Cn, steps = 4, 768
    rng = np.random.default_rng(1234)
    vol = rng.uniform(.1, .3, 4096)[:Cn]; K = rng.uniform(80, 120, 4096)[:Cn]
    pairs = [(-0.1 + 0.02 * (i - 4), 0.45 + 0.01 * (i - 4)) for i in (0, 3, 5, 7)]
    sp = oracle.special_axis()
    axes, jk, jx, jd, jw = [], [], [], [], []
    for c in range(Cn):
        ax = oracle.axis_union(oracle.axis_union(sp * K[c], sp * K[c]), np.array([K[c] * 1000.0]))
        x0, dx, kappa, w = qpde.jump_setup(qpde.refine_axis(ax, 6), "lognormal", pairs[c])
        axes.append(ax); jk.append(kappa); jx.append(x0); jd.append(dx); jw.append(w)
    spec = qpde.BatchSpec(axis=np.stack(axes), refine=6, r=0.05, sigma=vol, q=0., T=.25, dt=.25 / steps, strike=K,
                          scheme="bdf1", payoff="call", jump=dict(lam=0.1, kappa=jk, x0=jx, dx=jd, weights=np.stack(jw)))
    with qpde.Plan(handle, spec) as plan:
        plan.run()
        res = plan.fetch()
    for c in range(Cn):
        o = oracle.jump_call(K[c], 0.05, vol[c], 0.25, 0.1, steps, 6, params=pairs[c])
        assert res.n_steps[c] == o["n_steps"]
        err = rel_err(res.V[c], o["V"]).max()
        assert err <= REL_TOL, "contract %d: %.3e" % (c, err)
