"""Runs the BASELINE configs[1] workload on a small batch a few times (for ncu /
quick timing).  usage: python tools/profile_case.py [contracts] [kernel] [runs]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from quantpde_b200 import capi  # noqa: E402

Cn = int(sys.argv[1]) if len(sys.argv) > 1 else 296
kernel = sys.argv[2] if len(sys.argv) > 2 else "resident"
runs = int(sys.argv[3]) if len(sys.argv) > 3 else 2
K, vol, T, r = bench.contract_parameters(Cn)
axis = np.ascontiguousarray(K[:, None] * bench.SPECIAL[None, :])
with capi.Handle(0) as h:
    spec = capi.BatchSpec(axis=axis, refine=bench.N_NODES_REFINE, r=r, sigma=vol, q=0.0, T=T,
                          dt=T / bench.N_STEPS, strike=K, american=True, payoff="put", kernel=kernel,
                          outputs=capi.OUT_V | capi.OUT_VALUE | capi.OUT_INNER_TOTAL | capi.OUT_STATUS)
    with capi.Plan(h, spec) as plan:
        for i in range(runs):
            t0 = time.perf_counter()
            plan.run()
            h.synchronize()
            dt = time.perf_counter() - t0
            print("run %d: %d contracts in %.2f ms -> %.0f solves/s (%s)" % (i, Cn, dt * 1e3, Cn / dt, plan.kernel))
        res = plan.fetch(capi.OUT_INNER_TOTAL | capi.OUT_STATUS)
        print("mean inner per step %.3f, not converged %d" % (res.inner_total.mean() / bench.N_STEPS, (res.status != 0).sum()))
