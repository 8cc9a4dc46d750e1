"""Shared helpers: run one golden case on the oracle / through the C ABI."""
import ast

import numpy as np

# The parity bar of BASELINE.json's north_star: "every node value matches the
# reference's CPU path ... to within 1e-9 relative".  Relative is measured the
# way the reference itself measures it (relativeError, scale = 1:
# Core/IterativeMethod.hpp:1125-1137): |a-b| / max(1, |a|, |b|).
REL_TOL = 1e-9


def rel_err(a, b):
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    return np.abs(a - b) / np.maximum(1.0, np.maximum(np.abs(a), np.abs(b)))


def mask_mismatches(mask, mask_ref, V_ref, obstacle):
    """Active-set comparison with the tie policy of DESIGN.md §6: the sets must be
    identical except at nodes where the reference's own V - obstacle is within a
    few ulp of zero.  There the sign is decided by the last rounding of the
    elimination (Thomas / partition / pivoted LU all differ) — typically the
    free-boundary node after the reference's extra ~1e-16-long last time step
    (SURVEY.md §8(a) row 3), where V - obstacle = O(1e-14).
    Returns (hard_mismatches, ties)."""
    mask = np.asarray(mask).astype(bool); mask_ref = np.asarray(mask_ref).astype(bool)
    diff = mask != mask_ref
    gap = np.abs(np.asarray(V_ref) - np.asarray(obstacle))
    tie = gap <= 64 * np.finfo(np.float64).eps * np.maximum(1.0, np.abs(obstacle))
    return int(np.sum(diff & ~tie)), int(np.sum(diff & tie))


def case_obstacle(kw, S):
    K = kw["K"]
    return np.where(S < K, 0.0, S - K) if kw.get("call", 0) else np.where(S < K, K - S, 0.0)


def jump_params(kw):
    return (kw["p_up"], kw["eta1"], kw["eta2"]) if kw["density"] == "kou" else (kw["mu"], kw["sd"])


def case_args(golden, name):
    return str(golden[name + "/mode"]), ast.literal_eval(str(golden[name + "/args"]))


def run_oracle_case(po, mode, kw):
    if mode == "vanilla":
        return po.american_put(kw["K"], kw["r"], kw["vol"], kw["T"], kw["steps"], kw["refine"],
                               q=kw.get("q", 0.0), call=bool(kw.get("call", 0)),
                               american=bool(kw["american"]), scheme=kw.get("scheme", "rannacher"),
                               variable_target=kw.get("variable_target"), E=kw.get("E", 0),
                               dividend=kw.get("dividend"), dividend_kind=kw.get("dividend_kind", "cash"),
                               dividend_first=bool(kw.get("dividend_first", 0)), exercise=bool(kw.get("exercise", 0)),
                               event=po.glwb_event(kw["E"]) if kw.get("event") == "glwb" else None)
    if mode == "glwb":
        return po.glwb(kw["years"], kw["steps"], kw["refine"], scheme=kw.get("scheme", "bdf2"))
    if mode == "forward":
        return po.forward_option(kw["K"], kw["r"], kw["vol"], kw["T"], kw["steps"], kw["refine"], call=bool(kw.get("call", 0)),
                                 american=bool(kw["american"]), scheme=kw["scheme"], alpha=kw.get("alpha", 0.0),
                                 E=kw.get("E", 0), dividend=kw.get("dividend"), dividend_kind=kw.get("dividend_kind", "cash"),
                                 dividend_first=bool(kw.get("dividend_first", 0)), exercise=bool(kw.get("exercise", 0)))
    if mode == "jump":
        return po.jump_call(kw["K"], kw["r"], kw["vol"], kw["T"], kw["lambda"], kw["steps"], kw["refine"],
                            density=kw["density"], params=jump_params(kw), scheme=kw.get("scheme", "bdf1"))
    if mode == "hjb":
        return po.hjb_straddle(kw["K"], kw["r_l"], kw["r_b"], kw["vol"], kw["T"], kw["steps"], kw["refine"],
                               long_position=bool(kw["long"]), scheme=kw["scheme"])
    return po.bermudan_put(kw["K"], kw["r"], kw["alpha"], kw["T"], kw["steps"], kw["E"],
                           kw["refine"], scheme=kw["scheme"], dividend=kw.get("dividend"),
                           dividend_kind=kw.get("dividend_kind", "cash"),
                           dividend_first=bool(kw.get("dividend_first", 0)), exercise=bool(kw.get("exercise", 1)))


def spec_for_case(qpde, po, mode, kw, kernel="auto", copies=1):
    """BatchSpec reproducing a golden case (optionally replicated `copies` times)."""
    if mode == "vanilla":
        K = kw["K"]
        axis = po.axis_union(po.special_axis() * K, po.special_axis() * K)
        call = bool(kw.get("call", 0))
        return qpde.BatchSpec(
            axis=np.tile(axis, (copies, 1)), refine=kw["refine"], r=kw["r"], sigma=kw["vol"],
            q=kw.get("q", 0.0), T=kw["T"], dt=kw["T"] / kw["steps"], strike=np.full(copies, K),
            scheme=kw.get("scheme", "rannacher"), american=bool(kw["american"]),
            payoff="call" if call else "put", kernel=kernel,
            variable_target=kw.get("variable_target"), max_steps=4000 if "variable_target" in kw else 0,
            n_exercise=kw.get("E", 0), exercise="k_minus_s" if kw.get("exercise", 0) else "none",
            dividend=None if "dividend" not in kw else dict(kind=kw["dividend_kind"], amount=kw["dividend"],
                                                          first=bool(kw.get("dividend_first", 0))),
            event=po.glwb_event(kw["E"]) if kw.get("event") == "glwb" else None)
    if mode == "glwb":
        years, S0 = kw["years"], 100.0
        _, source, params = po.glwb_tables(years)
        ev = po.glwb_event(years, bonus=.06, GQ=.05 * S0)
        axis = po.special_axis() * S0
        N = qpde.refined_size(len(axis), kw["refine"])
        return qpde.BatchSpec(
            axis=np.tile(axis, (copies, 1)), refine=kw["refine"], r=.04, sigma=.2, q=.015, T=float(years),
            dt=float(years) / kw["steps"], strike=np.full(copies, S0), scheme=kw.get("scheme", "bdf2"), american=False,
            payoff_array=np.zeros((copies, N)), n_exercise=years, exercise="none", kernel=kernel,
            event=dict(program=ev["program"], consts=ev["consts"], params=params),
            event_times=[float(n) for n in range(1, years + 1)], source=source)
    if mode == "forward":
        K = kw["K"]
        axis = po.axis_union(po.special_axis() * K, po.special_axis() * K)
        local = kw.get("alpha", 0.0) > 0
        return qpde.BatchSpec(
            axis=np.tile(axis, (copies, 1)), refine=kw["refine"], r=kw["r"], sigma=kw["alpha"] if local else kw["vol"],
            vol_model="inverse_s" if local else "const", q=0.0, T=kw["T"], dt=kw["T"] / kw["steps"],
            strike=np.full(copies, K), scheme=kw["scheme"], american=bool(kw["american"]),
            payoff="call" if kw.get("call", 0) else "put", kernel=kernel, forward=True,
            n_exercise=kw.get("E", 0), exercise="k_minus_s" if kw.get("exercise", 0) else "none",
            dividend=None if "dividend" not in kw else dict(kind=kw["dividend_kind"], amount=kw["dividend"],
                                                          first=bool(kw.get("dividend_first", 0))))
    if mode == "jump":
        K = kw["K"]
        sp = po.special_axis()
        axis = po.axis_union(po.axis_union(sp * K, sp * K), np.array([K * 1000.0]))
        S = qpde.refine_axis(axis, kw["refine"])
        x0, dx, kappa, w = qpde.jump_setup(S, kw["density"], jump_params(kw))   # product host helper
        return qpde.BatchSpec(
            axis=np.tile(axis, (copies, 1)), refine=kw["refine"], r=kw["r"], sigma=kw["vol"], q=0.0,
            T=kw["T"], dt=kw["T"] / kw["steps"], strike=np.full(copies, K), scheme=kw.get("scheme", "bdf1"),
            american=False, payoff="call", kernel=kernel,
            jump=dict(lam=kw["lambda"], kappa=kappa, x0=x0, dx=dx, weights=w))
    if mode == "hjb":
        K = kw["K"]
        axis = po.axis_union(po.special_axis() * K, po.special_axis() * K)
        return qpde.BatchSpec(
            axis=np.tile(axis, (copies, 1)), refine=kw["refine"], r=0.0, sigma=kw["vol"], q=0.0,
            T=kw["T"], dt=kw["T"] / kw["steps"], strike=np.full(copies, K), scheme=kw["scheme"],
            american=False, payoff="straddle", controls=np.array([kw["r_l"], kw["r_b"]]),
            policy_max=bool(kw["long"]), kernel=kernel)
    K = kw["K"]
    axis = po.TUTORIAL_AXIS * (K / 100.0)
    return qpde.BatchSpec(
        axis=np.tile(axis, (copies, 1)), refine=kw["refine"], r=kw["r"], sigma=kw["alpha"], q=0.0,
        T=kw["T"], dt=kw["T"] / kw["steps"], strike=np.full(copies, K), scheme=kw["scheme"],
        american=False, payoff="put", vol_model="inverse_s", n_exercise=kw["E"],
        exercise="k_minus_s" if kw.get("exercise", 1) else "none", kernel=kernel,
        dividend=None if "dividend" not in kw else dict(kind=kw["dividend_kind"], amount=kw["dividend"],
                                                      first=bool(kw.get("dividend_first", 0))))
