"""American put batch at other grid sizes / forced kernel family (throughput cliff check).
usage: python tools/profile_sizes.py"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from quantpde_b200 import capi  # noqa: E402

with capi.Handle(0) as h:
    for refine, kernel, Cn, steps in ((6, "resident", 4736, 1000), (6, "interleaved", 16384, 1000), (7, "auto", 4736, 1000), (7, "interleaved", 16384, 1000),
                                      (4, "resident", 37888, 1000), (4, "interleaved", 65536, 1000), (2, "auto", 65536, 1000)):
        K, vol, T, r = bench.contract_parameters(Cn)
        axis = np.ascontiguousarray(K[:, None] * bench.SPECIAL[None, :])
        spec = capi.BatchSpec(axis=axis, refine=refine, r=r, sigma=vol, q=0.0, T=T, dt=T / steps, strike=K, american=True,
                              payoff="put", kernel=kernel, outputs=capi.OUT_VALUE | capi.OUT_INNER_TOTAL | capi.OUT_STATUS)
        with capi.Plan(h, spec) as plan:
            plan.run(); h.synchronize()
            t0 = time.perf_counter(); plan.run(); h.synchronize(); dt = time.perf_counter() - t0
            res = plan.fetch(capi.OUT_INNER_TOTAL | capi.OUT_STATUS)
            solves = float(res.inner_total.sum())
            print("N %5d  %-11s %6d contracts x %d steps: %8.0f contracts/s, %6.1f M inner solves/s, algorithmic %6.0f GB/s, status %d"
                  % (plan.N, plan.kernel, Cn, steps, Cn / dt, solves / dt / 1e6, 56. * plan.N * solves / dt / 1e9, res.status.max()))
