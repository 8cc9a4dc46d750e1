"""INTERLEAVED family on many small systems ("systems that fit a thread").
usage: profile_interleaved.py [contracts] [refine] [steps] [runs] [kernel]"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from quantpde_b200 import capi
Cn = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
refine = int(sys.argv[2]) if len(sys.argv) > 2 else 2
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 100
runs = int(sys.argv[4]) if len(sys.argv) > 4 else 3
kernel = sys.argv[5] if len(sys.argv) > 5 else "interleaved"
K, vol, T, r = bench.contract_parameters(Cn)
with capi.Handle(0) as h:
    spec = capi.BatchSpec(axis=np.ascontiguousarray(K[:, None] * bench.SPECIAL[None, :]), refine=refine, r=r, sigma=vol,
                          q=0., T=T, dt=T / steps, strike=K, american=True, payoff="put", kernel=kernel,
                          outputs=capi.OUT_VALUE | capi.OUT_INNER_TOTAL | capi.OUT_STATUS)
    with capi.Plan(h, spec) as plan:
        N = plan.N
        for i in range(runs):
            t0 = time.perf_counter(); plan.run(); h.synchronize(); dt = time.perf_counter() - t0
            res = plan.fetch(capi.OUT_INNER_TOTAL)
            inner = int(res.inner_total.astype(np.int64).sum())
            print("run %d: %d contracts, N=%d, %d steps (%s): %.1f ms, %.0f solves/s, algorithmic %.0f GB/s (56 B), real-model %.0f GB/s (96 B)"
                  % (i, Cn, N, steps, plan.kernel, dt * 1e3, Cn / dt, 56. * N * inner / dt / 1e9, 96. * N * inner / dt / 1e9))
