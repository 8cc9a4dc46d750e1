"""GPU: the RESIDENT sweep at two contracts per SM (qpde_lean.cu, American sweeps on 513 < N <= 2049
nodes) against the oracle, and against k_resident_sweep (QPDE_RESIDENT_LEAN=0) on the same batch."""
import os

import numpy as np
import pytest

from helpers import REL_TOL, mask_mismatches, rel_err

pytestmark = pytest.mark.gpu


def _run(qpde, handle, spec, lean=True):
    old = os.environ.get("QPDE_RESIDENT_LEAN")
    os.environ["QPDE_RESIDENT_LEAN"] = "1" if lean else "0"
    try:
        with qpde.Plan(handle, spec) as plan:
            plan.run()
            return plan.fetch()
    finally:
        if old is None:
            del os.environ["QPDE_RESIDENT_LEAN"]
        else:
            os.environ["QPDE_RESIDENT_LEAN"] = old


def _batch(qpde, oracle, Cn, refine, steps, seed, scheme="rannacher", call=False, q=0.0, axis0=None):
    rng = np.random.default_rng(seed)
    K = rng.uniform(50, 150, Cn); vol = rng.uniform(.1, .6, Cn)
    T = rng.uniform(.25, 2., Cn); r = rng.uniform(.01, .08, Cn)
    sp = oracle.special_axis() if axis0 is None else axis0
    axis = np.stack([oracle.axis_union(sp * k, sp * k) for k in K])
    spec = qpde.BatchSpec(axis=axis, refine=refine, r=r, sigma=vol, q=q, T=T, dt=T / steps, strike=K,
                          american=True, payoff="call" if call else "put", scheme=scheme, kernel="resident",
                          outputs=qpde.OUT_ALL)
    return spec, K, vol, T, r


@pytest.mark.parametrize("refine,steps,scheme", [(5, 150, "rannacher"), (6, 120, "rannacher"),
                                                   (5, 90, "cn"), (6, 60, "bdf1")])
def test_lean_against_oracle(qpde, handle, oracle, refine, steps, scheme):
    Cn = 10
    spec, K, vol, T, r = _batch(qpde, oracle, Cn, refine, steps, seed=11 + refine, scheme=scheme)
    res = _run(qpde, handle, spec)
    assert np.all(res.status == 0)
    assert np.all(res.computed <= res.inner_total) and np.all(res.computed >= res.n_steps)
    for c in range(Cn):
        o = oracle.american_put(K[c], r[c], vol[c], T[c], steps, refine, scheme=scheme)
        assert np.array_equal(res.S[c], o["S"])                      # the closed-form nodes are refined_node's bits
        err = rel_err(res.V[c], o["V"]).max()
        assert err <= REL_TOL, "contract %d: %.3e" % (c, err)
        assert res.n_steps[c] == o["n_steps"]
        assert np.array_equal(res.inner[c, :o["n_steps"]], o["inner"])
        assert mask_mismatches(res.active[c], o["mask"], o["V"], np.maximum(K[c] - o["S"], 0.0))[0] == 0
        assert rel_err(res.value[c], oracle.interp(o["S"], o["V"], K[c])) <= REL_TOL


def test_lean_equals_the_one_cta_per_sm_sweep(qpde, handle, oracle):
    """Same batch through both RESIDENT variants: identical step and iteration counts, active sets
    identical outside the tie band, node values within 1e-12 (two elimination orders)."""
    spec, K, vol, T, r = _batch(qpde, oracle, 300, 6, 80, seed=5)
    lean = _run(qpde, handle, spec, lean=True)
    fat = _run(qpde, handle, spec, lean=False)
    assert np.array_equal(lean.S, fat.S)
    assert np.array_equal(lean.n_steps, fat.n_steps)
    assert np.array_equal(lean.inner, fat.inner) and np.array_equal(lean.inner_total, fat.inner_total)
    assert np.array_equal(lean.computed, fat.computed)
    assert rel_err(lean.V, fat.V).max() <= 1e-12
    payoff = np.maximum(K[:, None] - fat.S, 0.0)
    for c in range(300):
        assert mask_mismatches(lean.active[c], fat.active[c], fat.V[c], payoff[c])[0] == 0


def test_lean_american_call_with_dividend_yield_and_partial_grid(qpde, handle, oracle):
    """N = 1057 (the 34-tick tutorial axis refined 5 times: the last thread rows are padding) and a
    call obstacle with q > r (early exercise matters)."""
    Cn, refine, steps = 6, 5, 100
    rng = np.random.default_rng(3)
    K = rng.uniform(80, 120, Cn); vol = rng.uniform(.15, .4, Cn); T = rng.uniform(.5, 1.5, Cn)
    axis = np.stack([oracle.TUTORIAL_AXIS * (k / 100.) for k in K])
    spec = qpde.BatchSpec(axis=axis, refine=refine, r=0.03, sigma=vol, q=0.06, T=T, dt=T / steps, strike=K,
                          american=True, payoff="call", kernel="resident", outputs=qpde.OUT_ALL)
    res = _run(qpde, handle, spec)
    for c in range(Cn):
        S = oracle.refine(axis[c], refine)
        lo, di, up = oracle.bs_coeffs(S, 0.03, vol[c], 0.06)
        payoff = np.where(S < K[c], 0.0, S - K[c])
        st, n = oracle.schedule(T[c], T[c] / steps)
        o = oracle.sweep(S, lo, di, up, payoff, T[c], st, n, american=True, obstacle=payoff)
        assert np.array_equal(res.S[c], S)
        assert rel_err(res.V[c], o["V"]).max() <= REL_TOL
        assert np.array_equal(res.inner[c, :n], o["inner"])
        assert mask_mismatches(res.active[c], o["mask"], o["V"], payoff)[0] == 0
