"""Several GMWB problems at once on one GPU: K host threads, one handle (= one stream set) each.
The line pass of one problem is a latency-bound chain on a cluster of 8 SMs; the other
140 SMs are only busy during its control search, so independent problems fill each other's gaps.
usage: profile_gmwb_concurrent.py [refine] [threads ...]     prints one JSON line per thread count"""
import json, os, sys, threading, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from quantpde_b200 import capi
# the 65-tick W axis of gmwb.cpp:92-105
W = np.array([0., 5., 10., 15., 20., 25., 30., 35., 40., 45., 50., 55., 60., 65., 70., 72.5, 75., 77.5, 80., 82., 84.,
    86., 88., 90., 91., 92., 93., 94., 95., 96., 97., 98., 99., 100., 101., 102., 103., 104., 105., 106.,
    107., 108., 109., 110., 112., 114., 116., 118., 120., 123., 126., 130., 135., 140., 145., 150., 160.,
    175., 200., 225., 250., 300., 500., 750., 1000.])

refine = int(sys.argv[1]) if len(sys.argv) > 1 else 2
counts = [int(a) for a in sys.argv[2:]] or [1, 2, 4, 8]
A = np.linspace(0., 100., 17)          # BASELINE.json configs[4]: uniform(0, 100, 17)
Z = np.array([0., .5, 1.])
steps = 32 * 2 ** refine


def solve(h, sigma, out, k):
    res = capi.gmwb_solve(h, W, A, [0., 10.], Z, refine, steps, sigma=sigma)
    out[k] = (res["status"], float(res["V"][np.searchsorted(res["A"], 100.), np.searchsorted(res["W"], 100.)]), res["V"])


with capi.Handle(0) as h0:
    capi.gmwb_solve(h0, W, A, [0., 10.], Z, 1, 64)        # warm-up: module load, clocks
    ref = {}
    solve(h0, .3, ref, 0)
base = None
for K in counts:
    handles = [capi.Handle(0) for _ in range(K)]
    sig = [.3 if k == 0 else .2 + .02 * k for k in range(K)]
    out = {}
    th = [threading.Thread(target=solve, args=(handles[k], sig[k], out, k)) for k in range(K)]
    t0 = time.perf_counter()
    for t in th: t.start()
    for t in th: t.join()
    dt = time.perf_counter() - t0
    for hh in handles: hh.close()
    same = bool(np.array_equal(out[0][2], ref[0][2]))     # problem 0 is the single-problem run bit for bit
    base = base or dt
    print(json.dumps({"tool": "profile_gmwb_concurrent", "grid": "%d x %d" % ((len(W) - 1) * 2 ** refine + 1, 16 * 2 ** refine + 1),
                      "steps": steps, "problems": K, "seconds": round(dt, 3), "problems_per_s": round(K / dt, 3),
                      "speedup_vs_sequential": round(K * base / dt / counts[0], 3) if counts[0] == 1 else None,
                      "status": [out[k][0] for k in range(K)], "problem0_bits_equal_single_run": same}))
