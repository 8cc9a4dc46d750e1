"""Small Bermudan (config 3) run for ncu / timing. usage: profile_bermudan.py [contracts] [scheme] [steps] [runs]"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from quantpde_b200 import capi
from tools.bench_configs import TUTORIAL
Cn = int(sys.argv[1]) if len(sys.argv) > 1 else 296
scheme = sys.argv[2] if len(sys.argv) > 2 else "bdf2"
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 800
runs = int(sys.argv[4]) if len(sys.argv) > 4 else 2
rng = np.random.default_rng(1234)
alpha = rng.uniform(10, 30, Cn); r = rng.uniform(.01, .08, Cn); K = rng.uniform(80, 120, Cn)
with capi.Handle(0) as h:
    spec = capi.BatchSpec(axis=K[:, None] / 100.0 * TUTORIAL[None, :], refine=6, r=r, sigma=alpha, q=0., T=1.,
                          dt=1. / steps, strike=K, scheme=scheme, payoff="put", vol_model="inverse_s",
                          n_exercise=10, exercise="k_minus_s", outputs=capi.OUT_VALUE | capi.OUT_STEPS)
    with capi.Plan(h, spec) as plan:
        for i in range(runs):
            t0 = time.perf_counter(); plan.run(); h.synchronize(); dt = time.perf_counter() - t0
            print("run %d: %d contracts x %d steps (%s) in %.2f ms" % (i, Cn, steps, scheme, dt * 1e3))
