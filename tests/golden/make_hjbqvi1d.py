"""Fixtures of the reference for the 1-D impulse-control HJBQVI path (examples/hjbqvi/exchange_rate.cpp):
runs oracle/_ref/qpde_ref exchange (the UNMODIFIED reference headers on oracle/shim, wired as the example)
in this container and stores its dumps in tests/golden/hjbqvi1d_reference.npz.

usage: python tests/golden/make_hjbqvi1d.py      (needs /root/reference; ~10 s)"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import pyoracle as po  # noqa: E402

# name -> arguments of `qpde_ref exchange` (defaults = the example's parameters, exchange_rate.cpp:36-76)
CASES = (
    ("er_k0", dict(refine=0)), ("er_k1", dict(refine=1)), ("er_k2", dict(refine=2)),
    ("er_k3", dict(refine=3)), ("er_k4", dict(refine=4)),
    ("er_cheap_intervention", dict(refine=2, sigma=.2, c=.05, **{"lambda": .5})),
    ("er_strong_rates", dict(refine=2, a=.5, b=1., rho=.05, T=5.)),
    ("er_parity_off_centre", dict(refine=1, m=.25)),
    ("er_coarse_controls", dict(refine=3, q_points=5, z_points=8, timesteps=8)),
)

out = {"names": np.array([n for n, _ in CASES])}
for name, kw in CASES:
    ref = po.run_ref("exchange", **kw)
    out[name + "/args"] = np.array(repr(kw))
    for key in ("X0", "Q0", "Z0", "X", "QA", "ZA", "V", "Q", "Z", "value", "steps", "mean_inner", "mean_solver"):
        out[name + "/" + key] = np.asarray(ref[key])
    print(name, len(ref["X"]), ref["steps"], ref["value"], ref["mean_inner"], ref["mean_solver"])
np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "hjbqvi1d_reference.npz"), **out)
