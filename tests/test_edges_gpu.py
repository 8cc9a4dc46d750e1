"""GPU: edge cases of the batch interface — a single contract, the smallest and the
largest grids each kernel family takes, the AUTO fall-back above the RESIDENT limit,
a shared (stride 0) axis, the max_inner guard, digital payoffs and dividends."""
import numpy as np
import pytest

from helpers import REL_TOL, mask_mismatches, rel_err

pytestmark = pytest.mark.gpu


def _american(qpde, handle, oracle, K, r, vol, T, steps, refine, kernel="auto", q=0.0, call=False, **kw):
    axis = oracle.axis_union(oracle.special_axis() * K, oracle.special_axis() * K)
    spec = qpde.BatchSpec(axis=axis[None, :], refine=refine, r=r, sigma=vol, q=q, T=T, dt=T / steps,
                          strike=[K], american=True, payoff="call" if call else "put", kernel=kernel, **kw)
    with qpde.Plan(handle, spec) as plan:
        plan.run()
        return plan.kernel, plan.fetch()


def test_single_contract_smallest_grid(qpde, handle, oracle):
    o = oracle.american_put(100., .04, .2, 1., 12, 0)                 # 33 nodes, one contract
    for kernel in ("interleaved", "resident"):
        _, res = _american(qpde, handle, oracle, 100., .04, .2, 1., 12, 0, kernel=kernel)
        assert rel_err(res.V[0], o["V"]).max() <= REL_TOL
        assert np.array_equal(res.inner[0, :12], o["inner"])


def test_auto_picks_resident_up_to_2305_nodes_then_interleaved(qpde, handle, oracle):
    # 2049 nodes: RESIDENT with the tail equation; 4097 nodes (refine 7): INTERLEAVED
    o6 = oracle.american_put(100., .04, .2, 1., 40, 6)
    kernel, res = _american(qpde, handle, oracle, 100., .04, .2, 1., 40, 6)
    assert kernel == "resident" and rel_err(res.V[0], o6["V"]).max() <= REL_TOL
    assert np.array_equal(res.inner[0, :o6["n_steps"]], o6["inner"])
    o7 = oracle.american_put(100., .04, .2, 1., 20, 7)
    kernel, res = _american(qpde, handle, oracle, 100., .04, .2, 1., 20, 7)
    assert kernel == "interleaved" and len(o7["S"]) == 4097
    assert rel_err(res.V[0], o7["V"]).max() <= REL_TOL
    with pytest.raises(qpde.QpdeError) as e:
        _american(qpde, handle, oracle, 100., .04, .2, 1., 20, 7, kernel="resident")
    assert e.value.code == qpde.ERR_UNSUPPORTED


def test_largest_resident_grid_nine_rows_per_thread(qpde, handle, oracle):
    # 36 ticks refined 6x = 2241 nodes: 9 rows per thread, no tail equation
    S0 = np.concatenate([oracle.special_axis() * 100., [20000., 40000., 80000.]])
    S = oracle.refine(S0, 6)
    assert len(S) == 2241
    lo, di, up = oracle.bs_coeffs(S, .03, .25, .01)
    payoff = np.where(S < 100., 100. - S, 0.)
    st, n = oracle.schedule(1., 1. / 50)
    o = oracle.sweep(S, lo, di, up, payoff, 1., st, n, american=True, obstacle=payoff)
    spec = qpde.BatchSpec(axis=np.tile(S0, (2, 1)), refine=6, r=.03, sigma=.25, q=.01, T=1., dt=1. / 50,
                          strike=[100., 100.], american=True, payoff="put", kernel="resident")
    with qpde.Plan(handle, spec) as plan:
        plan.run(); res = plan.fetch()
    assert rel_err(res.V[1], o["V"]).max() <= REL_TOL
    assert np.array_equal(res.inner[1, :n], o["inner"])
    assert mask_mismatches(res.active[1], o["mask"], o["V"], payoff)[0] == 0


def test_shared_axis_and_digital_call_with_dividends(qpde, handle, oracle):
    axis = oracle.special_axis() * 100.
    S = oracle.refine(axis, 3)
    Cn = 5
    vol = np.linspace(.15, .45, Cn)
    spec = qpde.BatchSpec(axis=axis, refine=3, r=.05, sigma=vol, q=.03, T=.75, dt=.75 / 60, strike=np.full(Cn, 100.),
                          american=False, payoff="digital_call", scheme="rannacher", n_contracts=Cn)
    assert spec.struct.axis_stride == 0
    with qpde.Plan(handle, spec) as plan:
        plan.run(); res = plan.fetch()
    payoff = np.where(S < 100., 0., 1.)
    st, n = oracle.schedule(.75, .75 / 60)
    for c in (0, Cn - 1):
        lo, di, up = oracle.bs_coeffs(S, .05, vol[c], .03)
        o = oracle.sweep(S, lo, di, up, payoff, .75, st, n, scheme="rannacher")
        assert np.array_equal(res.S[c], S)
        assert rel_err(res.V[c], o["V"]).max() <= REL_TOL


def test_max_inner_guard_reports_status(qpde, handle, oracle):
    # the reference would loop forever; the guard stops at max_inner and says so
    K, r, vol, T, steps, refine = 100., .04, .2, 1., 12, 2
    S = oracle.vanilla_grid(K, K, refine)
    lo, di, up = oracle.bs_coeffs(S, r, vol, 0.)
    payoff = np.where(S < K, K - S, 0.)
    st, n = oracle.schedule(T, T / steps)
    o = oracle.sweep(S, lo, di, up, payoff, T, st, n, american=True, obstacle=payoff, max_inner=1)
    assert o["status"] == 1
    for kernel in ("interleaved", "resident"):
        _, res = _american(qpde, handle, oracle, K, r, vol, T, steps, refine, kernel=kernel, max_inner=1)
        assert res.status[0] == 1 and np.all(res.inner[0, :n] == 1)
        assert rel_err(res.V[0], o["V"]).max() <= REL_TOL


def test_american_call_with_dividends_both_kernels(qpde, handle, oracle):
    o = oracle.american_put(100., .03, .3, 1.5, 64, 3, q=.06, call=True)
    for kernel in ("interleaved", "resident"):
        _, res = _american(qpde, handle, oracle, 100., .03, .3, 1.5, 64, 3, kernel=kernel, q=.06, call=True)
        assert rel_err(res.V[0], o["V"]).max() <= REL_TOL
        assert np.array_equal(res.inner[0, :o["n_steps"]], o["inner"])
        payoff = np.where(o["S"] < 100., 0., o["S"] - 100.)
        assert mask_mismatches(res.active[0], o["mask"], o["V"], payoff)[0] == 0
