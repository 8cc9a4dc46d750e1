"""GMWB timing. usage: profile_gmwb.py [refine] [a_points] [timesteps0]"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from quantpde_b200 import capi
W = np.array([0., 5., 10., 15., 20., 25., 30., 35., 40., 45., 50., 55., 60., 65., 70., 72.5, 75., 77.5, 80., 82., 84.,
    86., 88., 90., 91., 92., 93., 94., 95., 96., 97., 98., 99., 100., 101., 102., 103., 104., 105., 106.,
    107., 108., 109., 110., 112., 114., 116., 118., 120., 123., 126., 130., 135., 140., 145., 150., 160.,
    175., 200., 225., 250., 300., 500., 750., 1000.])
refine = int(sys.argv[1]) if len(sys.argv) > 1 else 2
a_points = int(sys.argv[2]) if len(sys.argv) > 2 else 51
ts0 = int(sys.argv[3]) if len(sys.argv) > 3 else 32
dx = 100. / (a_points - 1)
A = np.array([0. + dx * i for i in range(a_points)])
Z = np.array([0., .5, 1.])
with capi.Handle(0) as h:
    capi.gmwb_solve(h, W, A, [0., 10.], Z, min(refine, 1), ts0)       # warm-up: module load, clocks
    dt = 1e30
    for _ in range(2):
        l0 = h.launches
        t0 = time.perf_counter()
        res = capi.gmwb_solve(h, W, A, [0., 10.], Z, refine, ts0 * 2 ** refine)
        dt = min(dt, time.perf_counter() - t0)
        launches = h.launches - l0
    iw, ia = np.searchsorted(res["W"], 100.), np.searchsorted(res["A"], 100.)
    print("refine %d: %d x %d nodes, %d steps, mean inner %.4f, value %.9f, status %d, %.2f s, launches %d"
          % (refine, len(res["W"]), len(res["A"]), res["n_steps"], res["mean_inner"], res["V"][ia, iw], res["status"], dt, launches))
