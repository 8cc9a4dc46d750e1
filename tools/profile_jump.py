"""BASELINE configs[3] shape (Merton jump European calls, 2113 nodes x 768 BDF1 steps, n_fft 4096) on a small
batch, for ncu / quick timing.  usage: python tools/profile_jump.py [contracts] [runs]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from quantpde_b200 import capi  # noqa: E402
from tools.bench_configs import SPECIAL  # noqa: E402

Cn = int(sys.argv[1]) if len(sys.argv) > 1 else 148
runs = int(sys.argv[2]) if len(sys.argv) > 2 else 2
rng = np.random.default_rng(1234)
vol = rng.uniform(.1, .3, Cn); K = np.full(Cn, 100.)
axis = np.union1d(SPECIAL * 100., [100. * 1000.])
S = capi.refine_axis(axis, 6)
x0, dx, kappa, w = capi.jump_setup(S, "lognormal", (-0.1, 0.45))
with capi.Handle(0) as h:
    spec = capi.BatchSpec(axis=np.tile(axis, (Cn, 1)), refine=6, r=0.05, sigma=vol, q=0., T=.25, dt=.25 / 768, strike=K,
                          scheme="bdf1", payoff="call", jump=dict(lam=0.1, kappa=kappa, x0=x0, dx=dx, weights=w),
                          outputs=capi.OUT_VALUE | capi.OUT_STATUS | capi.OUT_STEPS)
    with capi.Plan(h, spec) as plan:
        for i in range(runs):
            t0 = time.perf_counter()
            plan.run(); h.synchronize()
            dt = time.perf_counter() - t0
            print("run %d: %d contracts in %.2f ms -> %.0f solves/s, %.1f us per step per contract-SM" % (
                i, Cn, dt * 1e3, Cn / dt, dt * 1e6 / 768 / max(1, Cn / 148)))
        res = plan.fetch(capi.OUT_VALUE | capi.OUT_STATUS | capi.OUT_STEPS)
        print("value[0] %.9f steps %d status %d" % (res.value[0], res.n_steps[0], res.status.max()))
