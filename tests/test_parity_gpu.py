"""GPU: the CUDA path, called through the C ABI, against (i) the committed
outputs of the reference itself and (ii) the oracle on seeded inputs.

Bar (BASELINE.json north_star): every node value within 1e-9 relative
(|a-b| / max(1,|a|,|b|), the reference's own relativeError with scale 1),
identical per-step inner iteration counts, identical final active sets."""
import numpy as np
import pytest

from helpers import (REL_TOL, case_args, case_obstacle, mask_mismatches, rel_err,
                     run_oracle_case, spec_for_case)

pytestmark = pytest.mark.gpu

KERNELS = ["interleaved", "resident"]


def _supported(qpde, handle, spec):
    try:
        return qpde.Plan(handle, spec)
    except qpde.QpdeError as e:
        if e.code == qpde.ERR_UNSUPPORTED:
            return None
        raise


@pytest.mark.parametrize("kernel", KERNELS)
def test_golden_cases(qpde, handle, oracle, golden, kernel):
    ran = 0
    for name in [str(n) for n in golden["names"]]:
        mode, kw = case_args(golden, name)
        if mode == "jump" and kernel == "interleaved":
            continue                 # the jump-diffusion sweep exists in the RESIDENT family only
        spec = spec_for_case(qpde, oracle, mode, kw, kernel=kernel, copies=3)
        plan = _supported(qpde, handle, spec)
        if plan is None:
            continue
        with plan:
            plan.run()
            res = plan.fetch()
        ran += 1
        V_ref, S_ref = golden[name + "/V"], golden[name + "/S"]
        for c in range(3):
            assert np.array_equal(res.S[c], S_ref), name                 # grid: bit-exact
            err = rel_err(res.V[c], V_ref).max()
            assert err <= REL_TOL, "%s: node error %.3e" % (name, err)
            assert res.n_steps[c] == int(golden[name + "/steps"]), name
            assert res.status[c] == 0
            if name + "/inner" in golden.files and len(golden[name + "/inner"]):
                n = res.n_steps[c]
                assert np.array_equal(res.inner[c, :n], golden[name + "/inner"]), name
                assert res.inner_total[c] == golden[name + "/inner"].sum()
            if name + "/mask" in golden.files and len(golden[name + "/mask"]):
                hard, _ = mask_mismatches(res.active[c], golden[name + "/mask"], V_ref,
                                          case_obstacle(kw, S_ref))
                assert hard == 0, name
            assert rel_err(res.value[c], float(golden[name + "/value"])) <= REL_TOL
    assert ran >= (40 if kernel == "interleaved" else 42)


@pytest.mark.parametrize("kernel", KERNELS)
def test_american_batch_vs_oracle(qpde, handle, oracle, kernel):
    """config-2 style: varied strike / vol / T / r, seeded as SURVEY.md §8(d)."""
    rng = np.random.default_rng(1234)
    Cn, refine, steps = 24, 4, 200        # N = 513
    K = rng.uniform(50, 150, Cn); vol = rng.uniform(.1, .6, Cn)
    T = rng.uniform(.25, 2., Cn); r = rng.uniform(.01, .08, Cn)
    sp = oracle.special_axis()
    axis = np.stack([oracle.axis_union(sp * k, sp * k) for k in K])
    spec = qpde.BatchSpec(axis=axis, refine=refine, r=r, sigma=vol, q=0.0, T=T, dt=T / steps,
                          strike=K, american=True, payoff="put", kernel=kernel)
    plan = _supported(qpde, handle, spec)
    if plan is None:
        pytest.skip("kernel does not support this size")
    with plan:
        plan.run()
        res = plan.fetch()
    worst, iters_equal, hard_total, ties = 0.0, 0, 0, 0
    for c in range(Cn):
        o = oracle.american_put(K[c], r[c], vol[c], T[c], steps, refine)
        assert np.array_equal(res.S[c], o["S"])
        worst = max(worst, rel_err(res.V[c], o["V"]).max())
        assert res.n_steps[c] == o["n_steps"]
        iters_equal += int(np.array_equal(res.inner[c, :o["n_steps"]], o["inner"]))
        hard, tie = mask_mismatches(res.active[c], o["mask"], o["V"], np.maximum(K[c] - o["S"], 0.0))
        hard_total += hard; ties += tie
        if o["n_steps"] == steps:
            # no ~1e-16-long extra last step: the penalty gap V - obstacle is O(1e-8), no ties
            assert tie == 0
        assert rel_err(res.value[c], oracle.interp(o["S"], o["V"], K[c])) <= REL_TOL
    assert worst <= REL_TOL, "worst node error %.3e" % worst
    assert iters_equal == Cn and hard_total == 0
    print("active-set ties (within 64 ulp of the obstacle, see DESIGN.md): %d nodes" % ties)


@pytest.mark.parametrize("kernel", KERNELS)
def test_full_size_properties(qpde, handle, oracle, kernel):
    """N = 2049 x 1000 steps (BASELINE.json configs[1] shape, small batch):
    one contract against the oracle, plus size-independent properties."""
    rng = np.random.default_rng(7)
    Cn, refine, steps = 8, 6, 1000
    K = rng.uniform(50, 150, Cn); vol = rng.uniform(.1, .6, Cn)
    T = rng.uniform(.25, 2., Cn); r = rng.uniform(.01, .08, Cn)
    K[1], vol[1], T[1], r[1] = K[0], vol[0], T[0], r[0]       # duplicate contract
    sp = oracle.special_axis()
    axis = np.stack([oracle.axis_union(sp * k, sp * k) for k in K])
    spec = qpde.BatchSpec(axis=axis, refine=refine, r=r, sigma=vol, q=0.0, T=T, dt=T / steps,
                          strike=K, american=True, payoff="put", kernel=kernel)
    plan = _supported(qpde, handle, spec)
    if plan is None:
        pytest.skip("kernel does not support this size")
    with plan:
        plan.run()
        res = plan.fetch()
        plan.run()                       # idempotence: a second run gives the same bits
        res2 = plan.fetch()
    assert np.array_equal(res.V, res2.V)
    assert np.array_equal(res.V[0], res.V[1])                   # determinism across the batch
    assert np.all(res.status == 0)
    payoff = np.maximum(K[:, None] - res.S, 0.0)
    assert np.all(res.V >= payoff - 1e-6 * np.maximum(1.0, payoff))   # V >= obstacle - penalty slack
    # Without the reference's extra ~1e-16-long last step the active set of a put is a
    # prefix in S; with it, V - obstacle collapses to rounding level and the set is noise.
    act = res.active.astype(np.int8)[res.n_steps == steps]
    assert np.all(np.diff(act, axis=1) <= 0) and np.all(act[:, 0] == 1) and np.all(act[:, -1] == 0)
    assert np.all(res.n_steps >= steps) and np.all(res.n_steps <= steps + 1)
    for c in (0, 5):
        o = oracle.american_put(K[c], r[c], vol[c], T[c], steps, refine)
        err = rel_err(res.V[c], o["V"]).max()
        assert err <= REL_TOL, "contract %d: %.3e" % (c, err)
        assert np.array_equal(res.inner[c, :o["n_steps"]], o["inner"])
        assert mask_mismatches(res.active[c], o["mask"], o["V"], np.maximum(K[c] - o["S"], 0.0))[0] == 0


@pytest.mark.parametrize("kernel", KERNELS)
def test_policy_batch_vs_oracle(qpde, handle, oracle, kernel):
    """unequal borrowing / lending rates (PolicyIteration over two interest rates),
    varied contracts, both positions."""
    rng = np.random.default_rng(99)
    Cn, refine, steps = 12, 3, 96
    K = rng.uniform(60, 140, Cn); vol = rng.uniform(.15, .5, Cn); T = rng.uniform(.3, 1.5, Cn)
    r_l = rng.uniform(.01, .04, Cn); r_b = r_l + rng.uniform(.005, .04, Cn)
    sp = oracle.special_axis()
    axis = np.stack([oracle.axis_union(sp * k, sp * k) for k in K])
    for long_position in (False, True):
        spec = qpde.BatchSpec(axis=axis, refine=refine, r=0.0, sigma=vol, q=0.0, T=T, dt=T / steps,
                              strike=K, scheme="bdf2", payoff="straddle",
                              controls=np.stack([r_l, r_b], axis=1), policy_max=long_position,
                              kernel=kernel)
        plan = _supported(qpde, handle, spec)
        if plan is None:
            pytest.skip("kernel does not support this configuration")
        with plan:
            plan.run()
            res = plan.fetch()
        for c in range(Cn):
            o = oracle.hjb_straddle(K[c], r_l[c], r_b[c], vol[c], T[c], steps, refine,
                                    long_position=long_position)
            assert res.n_steps[c] == o["n_steps"]
            err = rel_err(res.V[c], o["V"]).max()
            assert err <= REL_TOL, "contract %d: %.3e" % (c, err)
            assert np.array_equal(res.inner[c, :o["n_steps"]], o["inner"])
            assert res.status[c] == 0


def test_jump_batch_full_size(qpde, handle, oracle):
    """config-4 shape: Merton jump European calls on the 2113-node grid (34 ticks
    refined 6x), n_fft = 4096, 768 BDF1 steps; one contract against the oracle."""
    rng = np.random.default_rng(4)
    Cn, refine, steps = 6, 6, 768
    K = rng.uniform(80, 120, Cn); vol = rng.uniform(.1, .3, Cn)
    sp = oracle.special_axis()
    axes, jl, jk, jx, jd, jw = [], [], [], [], [], []
    for c in range(Cn):
        axis = oracle.axis_union(oracle.axis_union(sp * K[c], sp * K[c]), np.array([K[c] * 1000.0]))
        S = qpde.refine_axis(axis, refine)
        x0, dx, kappa, w = qpde.jump_setup(S, "lognormal", (-0.1, 0.45))
        axes.append(axis); jk.append(kappa); jx.append(x0); jd.append(dx); jw.append(w)
    spec = qpde.BatchSpec(axis=np.stack(axes), refine=refine, r=0.05, sigma=vol, q=0.0, T=0.25,
                          dt=0.25 / steps, strike=K, scheme="bdf1", payoff="call",
                          jump=dict(lam=0.1, kappa=jk, x0=jx, dx=jd, weights=np.stack(jw)))
    with qpde.Plan(handle, spec) as plan:
        assert plan.N == 2113
        plan.run()
        res = plan.fetch()
    assert np.all(res.status == 0) and np.all(res.n_steps == 768)
    for c in (0, Cn - 1):
        o = oracle.jump_call(K[c], 0.05, vol[c], 0.25, 0.1, steps, refine)
        assert np.array_equal(res.S[c], o["S"])
        err = rel_err(res.V[c], o["V"]).max()
        assert err <= REL_TOL, "contract %d: %.3e" % (c, err)
        assert rel_err(res.value[c], oracle.interp(o["S"], o["V"], K[c])) <= REL_TOL


def test_one_shot_sweep_host_buffers(qpde, handle, oracle):
    """qpde_bs1d_sweep: the reference-facing call (host pointers in and out)."""
    o = oracle.american_put(100., .04, .2, 1., 48, 1)
    axis = oracle.axis_union(oracle.special_axis() * 100., oracle.special_axis() * 100.)
    spec = qpde.BatchSpec(axis=axis[None, :], refine=1, r=.04, sigma=.2, q=0., T=1., dt=1. / 48,
                          strike=[100.], american=True, payoff="put")
    res = qpde.sweep(handle, spec, outputs=qpde.OUT_V | qpde.OUT_VALUE | qpde.OUT_INNER_TOTAL)
    assert rel_err(res.V[0], o["V"]).max() <= REL_TOL
    assert res.inner_total[0] == o["inner"].sum()
    assert res.S is None and res.active is None


def test_one_shot_sweep_of_a_large_batch_matches_the_plan(qpde, handle, oracle):
    """qpde_bs1d_sweep on a batch of many waves (5,000 contracts on 148 SMs): every output equals
    the plan path bit for bit, also into strided host buffers."""
    Cn = 5000
    rng = np.random.default_rng(7)
    K = rng.uniform(50., 150., Cn); vol = rng.uniform(.1, .6, Cn); T = rng.uniform(.25, 2., Cn)
    r = rng.uniform(.01, .08, Cn)
    special = oracle.special_axis()
    spec = qpde.BatchSpec(axis=K[:, None] * special[None, :], refine=1, r=r, sigma=vol, q=0., T=T,
                          dt=T / 20, strike=K, american=True, payoff="put", kernel="resident")
    with qpde.Plan(handle, spec) as plan:
        plan.run()
        ref = plan.fetch()
    res = qpde.sweep(handle, spec)
    for name in ("V", "S", "value", "n_steps", "inner_total", "active", "status"):
        assert np.array_equal(getattr(res, name), getattr(ref, name)), name
    w = min(res.inner.shape[1], ref.inner.shape[1])
    assert np.array_equal(res.inner[:, :w], ref.inner[:, :w]) and np.all(ref.n_steps <= w)
    # strided host buffers (a wider V row than N)
    wide = qpde.ResultBuffers(Cn, spec.N + 5, qpde.sweep_max_steps(spec), qpde.OUT_V | qpde.OUT_VALUE)
    wide.struct.V_stride = spec.N + 5
    qpde.sweep(handle, spec, result=wide)
    assert np.array_equal(wide.V[:, :spec.N], ref.V) and np.array_equal(wide.value, ref.value)


def test_payoff_and_sigma_arrays(qpde, handle, oracle):
    """explicit [C][N] initial condition / obstacle / per-node volatility."""
    K, refine, steps = 100., 2, 80
    S = oracle.refine(oracle.TUTORIAL_AXIS, refine)
    with np.errstate(divide="ignore"):
        vol = 20. / S
    vol[0] = 0.3                                           # node 0 is a boundary row; value unused
    lo, di, up = oracle.bs_coeffs(S, .04, vol, 0.)
    payoff = np.abs(S - K)                                 # straddle as an array
    st, n = oracle.schedule(1., 1. / steps)
    o = oracle.sweep(S, lo, di, up, payoff, 1., st, n, american=True, obstacle=payoff)
    spec = qpde.BatchSpec(axis=np.tile(oracle.TUTORIAL_AXIS, (2, 1)), refine=refine, r=.04, sigma=0.,
                          q=0., T=1., dt=1. / steps, strike=[K, K], american=True,
                          payoff_array=np.tile(payoff, (2, 1)), obstacle_array=np.tile(payoff, (2, 1)),
                          vol_model="nodes", sigma_nodes=np.tile(vol, (2, 1)))
    with qpde.Plan(handle, spec) as plan:
        plan.run()
        res = plan.fetch()
    assert rel_err(res.V[1], o["V"]).max() <= REL_TOL
    assert np.array_equal(res.inner[1, :n], o["inner"])
    assert mask_mismatches(res.active[1], o["mask"], o["V"], payoff)[0] == 0


def test_invalid_arguments(qpde, handle):
    axis = np.array([[0., 1., 2., 3.]])
    good = dict(axis=axis, refine=1, r=.04, sigma=.2, q=0., T=1., dt=.1, strike=[1.], american=True)
    with pytest.raises(qpde.QpdeError) as e:
        qpde.Plan(handle, qpde.BatchSpec(**{**good, "dt": 2.0}))
    assert e.value.code == qpde.ERR_INVALID
    with pytest.raises(qpde.QpdeError):
        qpde.Plan(handle, qpde.BatchSpec(**{**good, "max_inner": 0}))
    spec = qpde.BatchSpec(**good)
    spec.struct.abi_version = 99
    with pytest.raises(qpde.QpdeError):
        qpde.Plan(handle, spec)


def test_policy_iteration_under_theta_schemes(qpde, handle, oracle, golden):
    """PolicyIteration under ReverseRannacher / ReverseCrankNicolson: the explicit half of the step is the policy
    node's A(t0) under the controls of the running inner iteration (Rannacher.hpp:60-73, CrankNicolson.hpp:64-78,
    PolicyIteration.hpp:107-115), so the right-hand side follows every control update.  INTERLEAVED family (AUTO
    falls to it); the four fixtures are the reference's own outputs, per-step inner iteration counts included."""
    names = ["hjb_short_rannacher_k1", "hjb_short_cn_k1", "hjb_long_rannacher_k2", "hjb_long_cn_k2"]
    for name in names:
        mode, kw = case_args(golden, name)
        for kernel in ("interleaved", "auto"):
            spec = spec_for_case(qpde, oracle, mode, kw, kernel=kernel, copies=2)
            with qpde.Plan(handle, spec) as plan:
                assert plan.kernel == "interleaved"
                plan.run()
                res = plan.fetch()
            n = res.n_steps[0]
            assert n == int(golden[name + "/steps"]) and res.status[0] == 0, name
            assert rel_err(res.V[0], golden[name + "/V"]).max() <= REL_TOL, name
            assert np.array_equal(res.inner[0, :n], golden[name + "/inner"]), name
        with pytest.raises(qpde.QpdeError):
            qpde.Plan(handle, spec_for_case(qpde, oracle, mode, kw, kernel="resident", copies=1))
