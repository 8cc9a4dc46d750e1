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


def test_auto_picks_resident_up_to_8193_nodes_then_interleaved(qpde, handle, oracle):
    # 2049 nodes: RESIDENT with the tail equation; 4097 / 8193 nodes (refine 7 / 8): RESIDENT on a cluster
    # of two / four CTAs per contract (tail equation in the last CTA); 16385 nodes: INTERLEAVED
    o6 = oracle.american_put(100., .04, .2, 1., 40, 6)
    kernel, res = _american(qpde, handle, oracle, 100., .04, .2, 1., 40, 6)
    assert kernel == "resident" and rel_err(res.V[0], o6["V"]).max() <= REL_TOL
    assert np.array_equal(res.inner[0, :o6["n_steps"]], o6["inner"])
    o7 = oracle.american_put(100., .04, .2, 1., 20, 7)
    for forced in ("auto", "interleaved"):
        kernel, res = _american(qpde, handle, oracle, 100., .04, .2, 1., 20, 7, kernel=forced)
        assert kernel == ("resident" if forced == "auto" else "interleaved") and len(o7["S"]) == 4097
        assert rel_err(res.V[0], o7["V"]).max() <= REL_TOL
        assert np.array_equal(res.inner[0, :o7["n_steps"]], o7["inner"])
    o8 = oracle.american_put(100., .04, .2, 1., 10, 8)
    kernel, res = _american(qpde, handle, oracle, 100., .04, .2, 1., 10, 8)
    assert kernel == "resident" and len(o8["S"]) == 8193
    assert rel_err(res.V[0], o8["V"]).max() <= REL_TOL
    assert np.array_equal(res.inner[0, :o8["n_steps"]], o8["inner"])
    o9 = oracle.american_put(100., .04, .2, 1., 20, 9)                # up to ~80 penalty iterations per step here
    kernel, res = _american(qpde, handle, oracle, 100., .04, .2, 1., 20, 9, max_inner=1000)
    assert kernel == "interleaved" and len(o9["S"]) == 16385 and res.status[0] == 0
    assert rel_err(res.V[0], o9["V"]).max() <= REL_TOL
    assert np.array_equal(res.inner[0, :20], np.minimum(o9["inner"], 255))
    with pytest.raises(qpde.QpdeError) as e:
        _american(qpde, handle, oracle, 100., .04, .2, 1., 20, 9, kernel="resident")
    assert e.value.code == qpde.ERR_UNSUPPORTED


@pytest.mark.parametrize("nodes_extra", [0, 3, 8])
@pytest.mark.parametrize("case", ["american_rannacher", "american_bdf2", "european_cn", "bermudan_bdf2", "hjb_bdf2"])
def test_cluster_of_ctas_per_contract(qpde, handle, oracle, case, nodes_extra):
    """2305 < N <= 8193: the rows of one contract are split over the two or four CTAs of a
    thread-block cluster, with (4097, 8193 nodes) and without (2433 nodes) the tail equation."""
    base = oracle.special_axis() * 100.
    if nodes_extra == 3:
        S0 = np.concatenate([base, [20000., 40000., 80000., 160000., 320000., 640000.]]); refine = 6   # 39 ticks -> 2433 nodes
    elif nodes_extra == 8:
        S0 = base; refine = 8                                                  # 8193 nodes: four CTAs + tail
    else:
        S0 = base; refine = 7                                                  # 4097 nodes: two CTAs + tail
    S = oracle.refine(S0, refine)
    K, r, vol, T, steps = 100., .04, .25, 1., 30
    Cn = 3
    common = dict(axis=np.tile(S0, (Cn, 1)), refine=refine, q=0., T=T, dt=T / steps, strike=np.full(Cn, K), kernel="resident")
    if case.startswith("american") or case == "european_cn":
        scheme = {"american_rannacher": "rannacher", "american_bdf2": "bdf2", "european_cn": "cn"}[case]
        american = case.startswith("american")
        lo, di, up = oracle.bs_coeffs(S, r, vol, 0.)
        payoff = np.where(S < K, K - S, 0.)
        st, n = oracle.schedule(T, T / steps)
        o = oracle.sweep(S, lo, di, up, payoff, T, st, n, scheme=scheme, american=american, obstacle=payoff)
        spec = qpde.BatchSpec(r=r, sigma=vol, scheme=scheme, american=american, payoff="put", **common)
    elif case == "bermudan_bdf2":
        with np.errstate(divide="ignore"):
            lo, di, up = oracle.bs_coeffs(S, r, 20. / S, 0.)
        payoff = np.maximum(K - S, 0.)
        st, n = oracle.schedule(T, T / steps, [T / 5 * e for e in range(5)])
        o = oracle.sweep(S, lo, di, up, payoff, T, st, n, scheme="bdf2", american=False, obstacle=payoff, event_obstacle=K - S)
        spec = qpde.BatchSpec(r=r, sigma=20., scheme="bdf2", american=False, payoff="put", vol_model="inverse_s",
                              n_exercise=5, exercise="k_minus_s", **common)
    else:
        rows = [oracle.bs_coeffs(S, rr, vol, 0.) for rr in (.03, .05)]
        prs = tuple(np.stack([rw[k] for rw in rows]) for k in range(3))
        payoff = np.where(S < K, K - S, S - K)
        st, n = oracle.schedule(T, T / steps)
        o = oracle.sweep(S, rows[0][0], rows[0][1], rows[0][2], payoff, T, st, n, scheme="bdf2", american=False,
                         obstacle=payoff, policy_rows=prs, policy_max=False)
        spec = qpde.BatchSpec(r=0., sigma=vol, scheme="bdf2", american=False, payoff="straddle",
                              controls=np.array([.03, .05]), policy_max=False, **common)
    with qpde.Plan(handle, spec) as plan:
        assert plan.kernel == "resident" and plan.N == len(S) and 2305 < plan.N <= 8193
        plan.run(); res = plan.fetch()
    for c in range(Cn):
        assert res.status[c] == 0 and res.n_steps[c] == n
        assert np.array_equal(res.S[c], S)
        err = rel_err(res.V[c], o["V"]).max()
        assert err <= REL_TOL, "%s: %.3e" % (case, err)
        if case.startswith("american") or case == "hjb_bdf2":
            assert np.array_equal(res.inner[c, :n], o["inner"]), case
        assert rel_err(res.value[c], oracle.lib().qpo_interp(len(S), oracle._dp(np.ascontiguousarray(S)),
                                                             oracle._dp(np.ascontiguousarray(o["V"])), K)) <= REL_TOL


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


def test_tiny_grids_stay_on_the_plain_load_sweep(qpde, handle, oracle):
    """fewer than 32 nodes: the INTERLEAVED family keeps its plain-load sweep (a TMA pass would be all
    prologue); 12 and 31 nodes against the oracle"""
    for n_nodes in (12, 31):
        S = np.concatenate([[0.], np.linspace(40., 160., n_nodes - 2), [400.]])
        K, r, vol, T, steps = 100., .04, .3, 1., 20
        lo, di, up = oracle.bs_coeffs(S, r, vol, 0.)
        payoff = np.where(S < K, K - S, 0.)
        st, n = oracle.schedule(T, T / steps)
        o = oracle.sweep(S, lo, di, up, payoff, T, st, n, american=True, obstacle=payoff)
        spec = qpde.BatchSpec(axis=np.tile(S, (3, 1)), refine=0, r=r, sigma=vol, q=0., T=T, dt=T / steps,
                              strike=np.full(3, K), american=True, payoff="put", kernel="interleaved")
        with qpde.Plan(handle, spec) as plan:
            assert plan.kernel == "interleaved" and not plan.tma_staged
            plan.run()
            res = plan.fetch()
        assert rel_err(res.V[2], o["V"]).max() <= REL_TOL
        assert np.array_equal(res.inner[2, :n], o["inner"])
