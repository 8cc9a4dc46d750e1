"""Throughput of the other BASELINE.json configs (3: Bermudan local-vol, 4: Merton
jump European call, plus the HJB policy-iteration example) on one GPU, with the
reference's CPU path timed beside each on a small sample.  bench.py stays the
contract for configs[1]; this tool feeds profiles/ and DESIGN.md.

usage: python tools/bench_configs.py [--scale 1.0] [--cpu-sample 4] [--runs 3]
"""
import argparse
import json
import os
import sys
import time
from concurrent.futures import ProcessPoolExecutor

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from quantpde_b200 import capi  # noqa: E402

SPECIAL = np.array([0., .1, .2, .3, .4, .5, .6, .7, .80, .84, .88, .92, .94, .96, .98, 1.,
                    1.02, 1.04, 1.06, 1.08, 1.10, 1.14, 1.18, 1.23, 1.30, 1.40, 1.50, 1.75,
                    2.25, 3.00, 7.50, 20., 100.])
TUTORIAL = np.array([0., 10., 20., 30., 40., 50., 60., 70., 75., 80., 84., 88., 92., 94., 96.,
                     98., 100., 102., 104., 106., 108., 110., 114., 118., 123., 130., 140.,
                     150., 175., 225., 300., 750., 2000., 10000.])


def _ref(args):
    import bench                                   # the cpu_baseline leg lives in bench.py
    return bench.reference_seconds(args)


def _jump_setup(args):
    S, params = args
    return capi.jump_setup(S, "lognormal", params)


def cpu_reference(mode, kws):
    cores = os.cpu_count() or 1
    t0 = time.perf_counter()
    with ProcessPoolExecutor(max_workers=cores) as ex:
        secs = list(ex.map(_ref, [(mode, kw) for kw in kws]))
    wall = time.perf_counter() - t0
    return {"solves_per_s_all_cores": len(kws) / wall, "cores": cores,
            "seconds_per_contract_1core": float(np.mean(secs)), "sample": len(kws)}


def time_plan(handle, spec, runs):
    with capi.Plan(handle, spec) as plan:
        plan.run(); handle.synchronize()
        ts = []
        for _ in range(runs):
            t0 = time.perf_counter()
            plan.run(); handle.synchronize()
            ts.append(time.perf_counter() - t0)
        res = plan.fetch(capi.OUT_STATUS | capi.OUT_STEPS | capi.OUT_INNER_TOTAL | capi.OUT_VALUE)
        return float(np.min(ts)), plan, res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scale", type=float, default=1.0, help="fraction of the BASELINE batch sizes")
    ap.add_argument("--cpu-sample", type=int, default=4)
    ap.add_argument("--runs", type=int, default=3)
    a = ap.parse_args()
    out = []
    with capi.Handle(0) as h:
        # ---- config 3: Bermudan puts, local volatility alpha / S, 10 exercise dates
        Cn = max(1, int(16384 * a.scale))
        rng = np.random.default_rng(1234)
        alpha = rng.uniform(10, 30, Cn); r = rng.uniform(.01, .08, Cn); K = rng.uniform(80, 120, Cn)
        axis = K[:, None] / 100.0 * TUTORIAL[None, :]
        for scheme, steps in (("bdf2", 6400), ("rannacher", 6400)):
            spec = capi.BatchSpec(axis=axis, refine=6, r=r, sigma=alpha, q=0., T=1., dt=1. / steps, strike=K,
                                  scheme=scheme, payoff="put", vol_model="inverse_s", n_exercise=10,
                                  exercise="k_minus_s", outputs=capi.OUT_VALUE | capi.OUT_STATUS | capi.OUT_STEPS | capi.OUT_INNER_TOTAL)
            secs, plan, res = time_plan(h, spec, a.runs)
            kws = [dict(steps=steps, refine=6, scheme=scheme, K=float(K[c]), r=float(r[c]), alpha=float(alpha[c]),
                        T=1.0, E=10) for c in range(a.cpu_sample)]
            out.append({"config": "3: Bermudan put, local vol, %s, 2113 nodes x %d steps, 10 exercise dates" % (scheme, steps),
                        "contracts": Cn, "seconds": secs, "solves_per_s": Cn / secs, "kernel": plan.kernel,
                        "steps_taken": int(res.n_steps[0]), "cpu_reference": cpu_reference("bermudan", kws)})
        # ---- config 4: Merton jump European calls, 8 (mu, sd) pairs, n_fft = 4096
        Cn = max(1, int(4096 * a.scale))
        rng = np.random.default_rng(1234)
        vol = rng.uniform(.1, .3, Cn); K = rng.uniform(80, 120, Cn)
        pairs = [(-0.1 + 0.02 * (i - 4), 0.45 + 0.01 * (i - 4)) for i in range(8)]
        which = rng.integers(0, 8, Cn)
        axes, jk, jx, jd, jw = [], [], [], [], []
        t0 = time.perf_counter()
        # The set-up (kappa and 4096 density weights by adaptive quadrature) is host
        # work done once per operator, as in the reference.  It depends on the contract
        # only through (density parameters, dx); identical keys share one computation.
        keys, grids = {}, []
        for c in range(Cn):
            axis_c = np.union1d(SPECIAL * K[c], [K[c] * 1000.0])
            S = capi.refine_axis(axis_c, 6)
            x0 = np.log(S[1]); dx = (np.log(S[-2]) - x0) / (4096 - 1)
            axes.append(axis_c); grids.append(S)
            keys.setdefault((int(which[c]), float(dx)), c)
        reps = sorted(keys.values())
        with ProcessPoolExecutor(max_workers=os.cpu_count() or 1) as ex:
            done = dict(zip(reps, ex.map(_jump_setup, [(grids[c], pairs[which[c]]) for c in reps])))
        for c in range(Cn):
            x0 = np.log(grids[c][1]); dx = (np.log(grids[c][-2]) - x0) / (4096 - 1)
            _, _, kappa, w = done[keys[(int(which[c]), float(dx))]]
            jk.append(kappa); jx.append(float(x0)); jd.append(float(dx)); jw.append(w)
        setup_s = time.perf_counter() - t0
        spec = capi.BatchSpec(axis=np.stack(axes), refine=6, r=0.05, sigma=vol, q=0., T=.25, dt=.25 / 768, strike=K,
                              scheme="bdf1", payoff="call",
                              jump=dict(lam=0.1, kappa=jk, x0=jx, dx=jd, weights=np.stack(jw)),
                              outputs=capi.OUT_VALUE | capi.OUT_STATUS | capi.OUT_STEPS | capi.OUT_INNER_TOTAL)
        secs, plan, res = time_plan(h, spec, a.runs)
        kws = [dict(steps=768, refine=6, density="lognormal", K=float(K[c]), r=0.05, vol=float(vol[c]), T=0.25,
                    mu=pairs[which[c]][0], sd=pairs[which[c]][1], **{"lambda": 0.1}) for c in range(a.cpu_sample)]
        out.append({"config": "4: Merton jump European call, BDF1, 2113 nodes x 768 steps, n_fft 4096",
                    "contracts": Cn, "seconds": secs, "solves_per_s": Cn / secs, "kernel": "resident (jump)",
                    "host_setup_seconds": setup_s, "distinct_weight_vectors": len(reps),
                    "steps_taken": int(res.n_steps[0]),
                    "cpu_reference": cpu_reference("jump", kws)})
        # ---- HJB: unequal borrowing / lending, policy iteration, BDF2
        Cn = max(1, int(8192 * a.scale))
        rng = np.random.default_rng(1234)
        K = rng.uniform(60, 140, Cn); vol = rng.uniform(.15, .5, Cn)
        r_l = rng.uniform(.01, .04, Cn); r_b = r_l + rng.uniform(.005, .04, Cn)
        spec = capi.BatchSpec(axis=K[:, None] * SPECIAL[None, :], refine=6, r=0., sigma=vol, q=0., T=1., dt=1. / 768,
                              strike=K, scheme="bdf2", payoff="straddle", controls=np.stack([r_l, r_b], axis=1),
                              outputs=capi.OUT_VALUE | capi.OUT_STATUS | capi.OUT_STEPS | capi.OUT_INNER_TOTAL)
        secs, plan, res = time_plan(h, spec, a.runs)
        kws = [dict(steps=768, refine=6, long=0, scheme="bdf2", K=float(K[c]), r_l=float(r_l[c]), r_b=float(r_b[c]),
                    vol=float(vol[c]), T=1.0) for c in range(a.cpu_sample)]
        out.append({"config": "HJB unequal borrowing/lending, policy iteration, BDF2, 2049 nodes x 768 steps",
                    "contracts": Cn, "seconds": secs, "solves_per_s": Cn / secs, "kernel": plan.kernel,
                    "mean_inner": float(res.inner_total.mean() / res.n_steps.mean()),
                    "cpu_reference": cpu_reference("hjb", kws)})
    for o in out:
        o["speedup_vs_all_cores"] = o["solves_per_s"] / o["cpu_reference"]["solves_per_s_all_cores"]
        print(json.dumps(o))


if __name__ == "__main__":
    main()
