"""Small HJBQVI cases for compute-sanitizer (memcheck / racecheck): the 1-D kernel at two geometries and the 2-D driver."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from quantpde_b200 import capi
g = np.load(os.path.join(ROOT, "tests", "golden", "hjbqvi1d_reference.npz"))
with capi.Handle(0) as h:
    par = [[.02, .3, .25, 3., 0., 1., .1], [.02, .2, .5, 1., 0., .5, .05]]
    r = capi.exchange_rate_solve(h, g["er_k0/X0"], g["er_k0/Q0"], g["er_k0/Z0"], par, 1, 4)      # M = 2
    print("1-D, 65 nodes:", r["status"], r["mean_inner"])
    r = capi.exchange_rate_solve(h, g["er_k0/X0"], g["er_k0/Q0"], g["er_k0/Z0"], par[:1], 5, 2)   # M = 5
    print("1-D, 1025 nodes:", r["status"], r["mean_inner"])
    r = capi.optimal_consumption_solve(h, 0, timesteps=2)
    print("2-D, 20 x 20 nodes:", r["status"], r["mean_inner"])
