"""CPU: the oracle (oracle/qpde_oracle.c) against the committed outputs of the
reference itself (tests/golden/reference_cases.npz, made by
tests/golden/make_golden.py from the unmodified reference headers), against the
live reference binary when it is present, and against closed-form anchors."""
import math

import numpy as np
import pytest

from helpers import case_args, run_oracle_case


def test_golden_cases_bit_exact(oracle, golden):
    names = [str(n) for n in golden["names"]]
    assert len(names) >= 76
    for name in names:
        mode, kw = case_args(golden, name)
        out = run_oracle_case(oracle, mode, kw)
        assert out["n_steps"] == int(golden[name + "/steps"]), name
        assert np.array_equal(out["S"], golden[name + "/S"]), name
        # bit-exact: same operation order as the reference + the LU stand-in
        assert np.array_equal(out["V"], golden[name + "/V"]), name
        if name + "/inner" in golden.files and len(golden[name + "/inner"]):
            assert np.array_equal(out["inner"], golden[name + "/inner"]), name
        if name + "/mask" in golden.files and len(golden[name + "/mask"]):
            assert np.array_equal(out["mask"], golden[name + "/mask"]), name


def test_live_reference_when_present(oracle):
    if not oracle.have_ref():
        pytest.skip("oracle/_ref/qpde_ref not built")
    kw = dict(american=1, steps=60, refine=2, K=87.5, r=0.03, vol=0.45, T=0.7)
    ref = oracle.run_ref("vanilla", **kw)
    out = oracle.american_put(kw["K"], kw["r"], kw["vol"], kw["T"], kw["steps"], kw["refine"])
    assert np.array_equal(out["V"], ref["V"])
    assert np.array_equal(out["inner"], ref["inner"])
    assert np.array_equal(out["mask"], ref["mask"])


def test_reference_known_answer_table(oracle, golden):
    """The one known-answer table the reference records (examples/jump_diffusion.cpp:179-190,
    Kou double-exponential jumps, "Tested February 18, 2016"): value and step count per level."""
    table = {3: (265, 97, 3.970527), 4: (529, 192, 3.971451), 5: (1057, 385, 3.972397), 6: (2113, 768, 3.972930)}
    for k, (nodes, steps, value) in table.items():
        name = "jump_kou_k%d" % k
        assert len(golden[name + "/S"]) == nodes
        assert int(golden[name + "/steps"]) == steps
        assert round(float(golden[name + "/value"]), 6) == value     # the reference, compiled here
        o = oracle.jump_call(100., .05, .15, .25, .1, 12 * 2 ** k, k, density="kou",
                             params=(0.3445, 3.0465, 3.0775))
        assert o["n_steps"] == steps and round(oracle.interp(o["S"], o["V"], 100.), 6) == value


def test_axis_special_and_refine(oracle):
    sp = oracle.special_axis()
    assert len(sp) == 33 and sp[0] == 0.0 and sp[-1] == 100.0 and sp[15] == 1.0
    fine = oracle.refine(sp * 100.0, 6)
    assert len(fine) == 2049                       # 2^6 * 32 + 1  
    assert np.all(np.diff(fine) > 0)
    assert np.array_equal(fine[::64], sp * 100.0)  # coarse ticks survive refinement
    assert np.array_equal(oracle.refine(sp, 0), sp)
    u = oracle.axis_union(np.array([0., 1., 2.]), np.array([1., 1.5, 3.]))
    assert np.array_equal(u, np.array([0., 1., 1.5, 2., 3.]))


def test_schedule_artifacts(oracle):
    # constant steps that divide T exactly in floating point
    st, n = oracle.schedule(1.0, 0.25)
    assert n == 4 and st[n - 1].t == 0.0 and st[n - 1].event_after == 1
    # accumulated t -= dt can miss 0 by more than eps: one extra tiny step
    # (the 97/385/1537/6145 step counts of examples/jump_diffusion.cpp:184-190)
    counts = [oracle.schedule(0.25, 0.25 / (12 * 2 ** k * 8))[1] for k in range(7)]
    assert all(c in (96 * 2 ** k, 96 * 2 ** k + 1) for k, c in enumerate(counts))
    # events: each of the E bodies ends with event_after, user events are counted
    T, E = 1.0, 10
    st, n = oracle.schedule(T, 0.01, [T / E * e for e in range(E)])
    ends = [i for i in range(n) if st[i].event_after]
    assert len(ends) == E and sum(st[i].user_events for i in ends) == E
    assert st[n - 1].t == 0.0 and st[n - 1].user_events == 1  # event at t = 0 fires with the NullEvent


def _bs_put(S, K, r, vol, T):
    d1 = (math.log(S / K) + (r + vol * vol / 2) * T) / (vol * math.sqrt(T))
    d2 = d1 - vol * math.sqrt(T)
    N = lambda x: 0.5 * (1 + math.erf(x / math.sqrt(2)))
    return K * math.exp(-r * T) * N(-d2) - S * N(-d1)


def test_european_converges_to_black_scholes(oracle):
    exact = _bs_put(100., 100., .04, .2, 1.)
    errs = []
    for k in range(4):
        out = oracle.american_put(100., .04, .2, 1., 12 * 2 ** k, k, american=False)
        errs.append(abs(oracle.interp(out["S"], out["V"], 100.) - exact))
    assert errs[-1] < 2e-3
    ratios = [errs[i] / errs[i + 1] for i in range(len(errs) - 1)]
    assert all(r > 3.0 for r in ratios)            # second order: ratio -> 4


@pytest.mark.parametrize("scheme", ["bdf3", "bdf4", "bdf5", "bdf6"])
def test_high_order_bdf_converges_to_black_scholes(oracle, scheme):
    """closed-form anchor for the BDF3 .. BDF6 restatement (the spatial error, second order, dominates)"""
    exact = _bs_put(100., 100., .04, .2, 1.)
    errs = []
    for k in range(4):
        out = oracle.american_put(100., .04, .2, 1., 12 * 2 ** k, k, american=False, scheme=scheme)
        errs.append(abs(oracle.interp(out["S"], out["V"], 100.) - exact))
    assert errs[-1] < 2e-3
    assert all(errs[i] / errs[i + 1] > 3.0 for i in range(len(errs) - 1))


def test_american_table_ratio_and_properties(oracle):
    vals, its = [], []
    for k in range(4):
        out = oracle.american_put(100., .04, .2, 1., 12 * 4 ** k, k)
        vals.append(oracle.interp(out["S"], out["V"], 100.))
        its.append(out["inner"].mean())
        payoff = np.maximum(100. - out["S"], 0.)
        assert np.all(out["V"] >= payoff - 1e-6 * np.maximum(1., payoff))
        m = out["mask"].astype(bool)
        assert m[0] and not m[-1]
        assert np.all(np.diff(m.astype(int)) <= 0)  # active set is a prefix in S for a put
        assert out["status"] == 0 and out["inner"].min() >= 1
    changes = np.diff(vals)
    assert all(c > 0 for c in changes)
    assert 2.0 < changes[0] / changes[1] < 8.0 and 2.0 < changes[1] / changes[2] < 8.0
    assert all(1.5 < x < 3.0 for x in its)


def test_gmwb_reference_fixture(oracle):
    """Fixtures of the reference's GMWB example (examples/hjbqvi/gmwb.cpp) for the one
    §8(a) row the device path does not cover yet.  Checks what can be checked without a
    restatement: the convergence table's first two values agree with the no-fee GMWB
    value of the literature (~107.7 at w = a = 100), the solution dominates the exit
    payoff, controls take grid values; and — when the reference binary is present —
    that it reproduces the fixture bit for bit."""
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "gmwb_reference.npz"))
    v0, v1 = float(g["gmwb_k0/value"]), float(g["gmwb_k1/value"])
    assert 107.5 < v0 < v1 < 107.8
    for name in ("gmwb_small", "gmwb_k0", "gmwb_k1"):
        W, A, V = g[name + "/W"], g[name + "/A"], g[name + "/V"].reshape(len(g[name + "/A"]), -1)
        assert V.shape == (len(A), len(W))                     # W fastest (Domain.hpp:1416-1425)
        payoff = np.maximum(W[None, :], 0.9 * A[:, None])      # exit function max(w, (1 - k) a - c)
        assert np.all(V >= payoff - 1e-6 * np.maximum(1., payoff))
        q = g[name + "/Q"]; q = q[np.isfinite(q)]              # entries no candidate claimed stay unset
        assert set(np.unique(q)) <= {0.0, 10.0}
        z = g[name + "/Z"]; z = z[np.isfinite(z)]
        assert len(z) == 0 or (z.min() >= 0.0 and z.max() <= 1.0)
    if oracle.have_ref():
        ref = oracle.run_ref("gmwb", refine=0, a_points=11, timesteps=8, control_points=3)
        assert np.array_equal(ref["V"], g["gmwb_small/V"])


def test_gmwb_oracle_against_reference_fixtures(oracle):
    """oracle/qpde_oracle_gmwb.c against the reference's own GMWB outputs: every node within
    1e-9 relative (measured: 3e-14), identical step and mean inner-iteration counts."""
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "gmwb_reference.npz"))
    for name, kw in (("gmwb_small", dict(refine_times=0, a_points=11, timesteps=8)),
                     ("gmwb_k0", dict(refine_times=0)), ("gmwb_k1", dict(refine_times=1))):
        o = oracle.gmwb(**kw)
        ref = g[name + "/V"].reshape(o["V"].shape)
        assert o["status"] == 0 and o["n_steps"] == int(g[name + "/steps"])
        assert o["mean_inner"] == float(g[name + "/mean_inner"])
        err = np.abs(o["V"] - ref) / np.maximum(1., np.abs(ref))
        assert err.max() <= 1e-9, "%s: %.3e" % (name, err.max())
        assert np.array_equal(o["W"], g[name + "/W"]) and np.array_equal(o["A"], g[name + "/A"])


def _er_kwargs(args):
    """`qpde_ref exchange` arguments -> keyword arguments of the model (oracle.exchange_rate / capi)."""
    kw = {("lam" if k == "lambda" else k): v for k, v in args.items()
          if k not in ("refine", "q_points", "z_points", "points", "timesteps")}
    return dict(refine_times=args.get("refine", 0), timesteps=args.get("timesteps", 16), **kw)


def test_hjbqvi1d_oracle_against_reference_fixtures(oracle):
    """oracle/qpde_oracle_hjbqvi1d.c against the reference's own outputs of examples/hjbqvi/exchange_rate.cpp
    (tests/golden/hjbqvi1d_reference.npz): grids bit-exact, every node within 1e-9 relative (measured: 1e-13),
    identical "Mean Policy Iterations", identical control vectors and impulse regions."""
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "hjbqvi1d_reference.npz"))
    for name in g["names"]:
        name = str(name)
        args = eval(str(g[name + "/args"]))
        o = oracle.exchange_rate(g[name + "/X0"], g[name + "/Q0"], g[name + "/Z0"], **_er_kwargs(args))
        assert o["status"] == 0 and o["n_steps"] == int(g[name + "/steps"]), name
        for key in ("X", "QA", "ZA"):
            assert np.array_equal(o[key], g[name + "/" + key]), (name, key)
        ref = g[name + "/V"]
        err = np.abs(o["V"] - ref) / np.maximum(1., np.abs(ref))
        assert err.max() <= 1e-9, "%s: %.3e" % (name, err.max())
        assert o["mean_inner"] == float(g[name + "/mean_inner"]), name
        for key in ("Q", "Z"):
            assert np.array_equal(o[key], g[name + "/" + key], equal_nan=True), (name, key)
        assert np.array_equal(o["mask"] != 0, np.isnan(g[name + "/Q"])), name
    # the convergence table of the example (the value at the parity m): -1.5954, -1.6018, -1.6001, -1.5988, -1.5980
    vals = [float(g["er_k%d/value" % k]) for k in range(5)]
    assert all(-1.61 < v < -1.59 for v in vals)
    if oracle.have_ref():
        ref = oracle.run_ref("exchange", refine=1)
        assert np.array_equal(ref["V"], g["er_k1/V"])


def test_hjbqvi2d_oracle_against_reference_fixtures(oracle):
    """oracle/qpde_oracle_hjbqvi2d.c against the reference's own outputs of examples/hjbqvi/optimal_consumption.cpp
    (tests/golden/hjbqvi2d_reference.npz; refinements 0 - 2 and three parameter / grid variations): every node
    within 1e-9 relative (measured: 2e-14), identical "Mean Policy Iterations", control vectors and transaction
    regions."""
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "hjbqvi2d_reference.npz"))
    for name in g["names"]:
        name = str(name)
        args = eval(str(g[name + "/args"]))
        kw = {("lam" if k == "lambda" else k): v for k, v in args.items() if k != "refine"}
        o = oracle.optimal_consumption(refine_times=args.get("refine", 0), **kw)
        assert o["status"] == 0 and o["n_steps"] == int(g[name + "/steps"]), name
        assert np.array_equal(o["W"], g[name + "/W"]) and np.array_equal(o["B"], g[name + "/B"]), name
        ref = g[name + "/V"].reshape(o["V"].shape)
        err = np.abs(o["V"] - ref) / np.maximum(1., np.abs(ref))
        assert err.max() <= 1e-9, "%s: %.3e" % (name, err.max())
        assert o["mean_inner"] == float(g[name + "/mean_inner"]), name
        for key in ("Q", "Z"):
            assert np.array_equal(o[key], g[name + "/" + key].reshape(ref.shape), equal_nan=True), (name, key)
    # the example's convergence table (value at w = b = 45.2): 56.058, 58.739, 59.420, 59.658
    vals = [float(g["oc_k%d/value" % k]) for k in range(4)]
    assert vals[0] < vals[1] < vals[2] < vals[3] < 60. and vals[0] > 56.
