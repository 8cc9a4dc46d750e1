"""Fixtures of the reference for the 2-D impulse-control HJBQVI path with impulses both ways
(examples/hjbqvi/optimal_consumption.cpp): runs oracle/_ref/qpde_ref consumption (the UNMODIFIED reference
headers on oracle/shim, wired as the example) in this container and stores its dumps in
tests/golden/hjbqvi2d_reference.npz.

usage: python tests/golden/make_hjbqvi2d.py [--k2-dump FILE] [--k3-dump FILE]
  refinement 2 (5929 nodes) takes the reference 45 s and refinement 3 (23,409 nodes, 256 steps) 11 minutes on
  one core: with --k2-dump / --k3-dump the saved stdout of `qpde_ref consumption refine=2|3` is used instead of
  a new run; refinement 3 is kept at every 4th node in both directions; without the flags an existing fixture
  keeps its k2 / k3 entries."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import pyoracle as po  # noqa: E402

PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "hjbqvi2d_reference.npz")
CASES = (
    ("oc_k0", dict(refine=0)), ("oc_k1", dict(refine=1)),
    ("oc_cheap_transfers", dict(refine=0, sigma=.2, c=.1, **{"lambda": .05})),
    ("oc_coarse", dict(refine=1, spatial_points=12, control_points=8, timesteps=16)),
    ("oc_risky", dict(refine=1, spatial_points=10, control_points=6, mu=.15, sigma=.4, gamma=.5, rho=.15)),
)
KEYS = ("W", "B", "QA", "ZA", "V", "Q", "Z", "value", "steps", "mean_inner", "mean_solver")


def dump_arg(flag):
    return po.parse_ref_dump(open(sys.argv[sys.argv.index(flag) + 1]).read()) if flag in sys.argv else None


out = {}
names = []
for name, kw in CASES:
    ref = po.run_ref("consumption", **kw)
    names.append(name)
    out[name + "/args"] = np.array(repr(kw))
    for key in KEYS:
        out[name + "/" + key] = np.asarray(ref[key])
    print(name, len(ref["W"]), len(ref["B"]), ref["steps"], ref["value"], ref["mean_inner"], ref["mean_solver"])
old = np.load(PATH, allow_pickle=False) if os.path.exists(PATH) else None
k2 = dump_arg("--k2-dump")
if k2 is not None:
    names.append("oc_k2")
    out["oc_k2/args"] = np.array(repr(dict(refine=2)))
    for key in KEYS:
        out["oc_k2/" + key] = np.asarray(k2[key])
    print("oc_k2", len(k2["W"]), k2["steps"], k2["value"], k2["mean_inner"])
k3 = dump_arg("--k3-dump")
if k3 is not None:
    nw3, nb3 = len(k3["W"]), len(k3["B"])
    out["oc_k3/args"] = np.array(repr(dict(refine=3)))
    out["oc_k3/stride"] = np.array(4)
    out["oc_k3/V_strided"] = k3["V"].reshape(nb3, nw3)[::4, ::4].copy()
    out["oc_k3/mask_strided"] = np.isnan(k3["Q"]).reshape(nb3, nw3)[::4, ::4].copy()
    for key in ("W", "B", "value", "steps", "mean_inner", "mean_solver"):
        out["oc_k3/" + key] = np.asarray(k3[key])
    print("oc_k3", nw3, nb3, k3["steps"], k3["value"], k3["mean_inner"])
if old is not None:
    for key in old.files:
        if (key.startswith("oc_k2/") and k2 is None) or (key.startswith("oc_k3/") and k3 is None):
            out[key] = old[key]
    if k2 is None and "oc_k2/V" in old.files:
        names.append("oc_k2")
out["names"] = np.array(names)
np.savez_compressed(PATH, **out)
