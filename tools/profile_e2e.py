"""Where the end-to-end call spends its time: plan creation (allocation + H2D), the sweep, the fetch (D2H)
into page-locked and into pageable host buffers.  usage: python tools/profile_e2e.py [contracts]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from quantpde_b200 import capi  # noqa: E402

Cn = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
K, vol, T, r = bench.contract_parameters(Cn)
axis = np.ascontiguousarray(K[:, None] * bench.SPECIAL[None, :])
out = capi.OUT_V | capi.OUT_VALUE | capi.OUT_INNER_TOTAL | capi.OUT_STATUS
with capi.Handle(0) as h:
    spec = capi.BatchSpec(axis=axis, refine=6, r=r, sigma=vol, q=0.0, T=T, dt=T / 1000, strike=K, american=True,
                          payoff="put", outputs=out)
    for rep in range(3):
        t0 = time.perf_counter()
        plan = capi.Plan(h, spec); h.synchronize()
        t1 = time.perf_counter()
        plan.run(); h.synchronize()
        t2 = time.perf_counter()
        res = capi.ResultBuffers(Cn, plan.N, plan.max_steps, out, pinned=True)
        t3 = time.perf_counter()
        capi.check(capi.lib().qpde_bs1d_plan_fetch(plan.ptr, capi.C.byref(res.struct)), h.ptr)
        t4 = time.perf_counter()
        res2 = capi.ResultBuffers(Cn, plan.N, plan.max_steps, out, pinned=False)
        t5 = time.perf_counter()
        capi.check(capi.lib().qpde_bs1d_plan_fetch(plan.ptr, capi.C.byref(res2.struct)), h.ptr)
        t6 = time.perf_counter()
        plan.close(); h.synchronize()
        t7 = time.perf_counter()
        res.close()
        print("create %.1f ms  run %.1f ms  [pinned alloc %.1f ms]  fetch(pinned) %.1f ms = %.1f GB/s  fetch(pageable) %.1f ms  destroy %.1f ms"
              % (1e3 * (t1 - t0), 1e3 * (t2 - t1), 1e3 * (t3 - t2), 1e3 * (t4 - t3), res.V.nbytes / (t4 - t3) / 1e9 if False else Cn * plan.N * 8 / (t4 - t3) / 1e9,
                 1e3 * (t6 - t5), 1e3 * (t7 - t6)))
