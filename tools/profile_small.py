"""RESIDENT family on small grids (American puts, 1000 steps). usage: python tools/profile_small.py"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from quantpde_b200 import capi  # noqa: E402

with capi.Handle(0) as h:
    for refine, Cn in ((0, 131072), (1, 131072), (2, 131072), (3, 65536), (4, 37888)):
        K, vol, T, r = bench.contract_parameters(Cn)
        axis = np.ascontiguousarray(K[:, None] * bench.SPECIAL[None, :])
        spec = capi.BatchSpec(axis=axis, refine=refine, r=r, sigma=vol, q=0.0, T=T, dt=T / 1000, strike=K, american=True,
                              payoff="put", kernel="resident", outputs=capi.OUT_VALUE | capi.OUT_INNER_TOTAL | capi.OUT_STATUS)
        with capi.Plan(h, spec) as plan:
            plan.run(); h.synchronize()
            t0 = time.perf_counter(); plan.run(); h.synchronize(); dt = time.perf_counter() - t0
            res = plan.fetch(capi.OUT_INNER_TOTAL | capi.OUT_STATUS)
            solves = float(res.inner_total.sum())
            print("N %5d  %-9s %6d contracts x 1000 steps: %9.0f contracts/s, %7.1f M inner solves/s, algorithmic %6.0f GB/s, status %d"
                  % (plan.N, plan.kernel, Cn, Cn / dt, solves / dt / 1e6, 56. * plan.N * solves / dt / 1e9, res.status.max()))
