import sys, numpy as np
sys.path.insert(0, '/root/repo')
from quantpde_b200 import capi
g = np.load('/root/repo/tests/golden/hjbqvi1d_reference.npz')
with capi.Handle(0) as h:
    par = [[.02, .3, .25, 3., 0., 1., .1]] * 148
    capi.exchange_rate_solve(h, g["er_k0/X0"], g["er_k0/Q0"], g["er_k0/Z0"], par, 4, 16)   # 513 nodes, 16 of the 256 steps
