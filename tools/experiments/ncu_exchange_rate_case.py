"""One wave of exchange-rate problems (2 per SM) at refinement 4 with a shortened time loop, for ncu."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from quantpde_b200 import capi
g = np.load(os.path.join(ROOT, "tests", "golden", "hjbqvi1d_reference.npz"))
with capi.Handle(0) as h:
    par = [[.02, .3, .25, 3., 0., 1., .1]] * 296
    capi.exchange_rate_solve(h, g["er_k0/X0"], g["er_k0/Q0"], g["er_k0/Z0"], par, 4, 16)   # 513 nodes, 16 steps
