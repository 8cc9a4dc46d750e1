"""Optimal consumption (2-D HJBQVI, impulses both ways) timing.  usage: profile_hjbqvi2d.py [refine ...]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from quantpde_b200 import capi
levels = [int(a) for a in sys.argv[1:]] or [2, 3]
with capi.Handle(0) as h:
    capi.optimal_consumption_solve(h, 0)       # warm-up
    for k in levels:
        l0 = h.launches; t0 = time.perf_counter()
        res = capi.optimal_consumption_solve(h, k)
        dt = time.perf_counter() - t0
        iw, ib = abs(res["W"] - 45.2).argmin(), abs(res["B"] - 45.2).argmin()
        print("refine %d: %d x %d nodes, %d steps, mean policy iterations %.4f, %.1f line sweeps per iteration, status %d, V(~45.2, ~45.2) = %.6f, %.2f s, %d launches"
              % (k, len(res["W"]), len(res["B"]), res["n_steps"], res["mean_inner"], res["mean_sweeps"], res["status"], res["V"][ib, iw], dt, h.launches - l0))
