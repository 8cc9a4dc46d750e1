"""Generates tests/golden/reference_cases.npz by running oracle/_ref/qpde_ref —
the UNMODIFIED reference headers under /root/reference compiled against
oracle/shim (see oracle/Makefile).  Run in the build container only (the GPU
box has no /root/reference and uses the committed .npz):

    make -C oracle && python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import pyoracle as po  # noqa: E402

CASES = []
# config 1: examples/vanilla_options.cpp American put convergence table, k = 0..3
for k in range(4):
    CASES.append(("american_table_k%d" % k, "vanilla",
                  dict(american=1, steps=12 * 4 ** k, refine=k, K=100.0, r=0.04, vol=0.2, T=1.0)))
# config 2 style: varied strike / vol / T / r (seeded), 100 steps, 257 nodes
rng = np.random.default_rng(1234)
for j in range(4):
    K = float(rng.uniform(50, 150)); vol = float(rng.uniform(.1, .6))
    T = float(rng.uniform(.25, 2.)); r = float(rng.uniform(.01, .08))
    CASES.append(("american_varied_%d" % j, "vanilla",
                  dict(american=1, steps=100, refine=3, K=K, r=r, vol=vol, T=T)))
# American call with dividends (early exercise matters)
CASES.append(("american_call_div", "vanilla",
              dict(american=1, call=1, steps=64, refine=2, K=100.0, r=0.03, vol=0.3, T=1.5, q=0.06)))
# European, every scheme
for sch in ("rannacher", "cn", "bdf1", "bdf2"):
    CASES.append(("european_%s" % sch, "vanilla",
                  dict(american=0, steps=100, refine=2, K=100.0, r=0.04, vol=0.2, T=1.0, scheme=sch)))
# config 3: examples/tutorial.cpp Bermudan put, local volatility, 10 exercise dates
for sch in ("bdf2", "rannacher", "bdf1", "cn"):
    for k in (0, 1):
        CASES.append(("bermudan_%s_k%d" % (sch, k), "bermudan",
                      dict(steps=100 * 2 ** k, refine=k, scheme=sch, K=100.0, r=0.04, alpha=20.0, T=1.0, E=10)))

# ReverseVariableStepper (Core/Stepper.hpp:60-126) as examples/vanilla_options.cpp:52-58 builds it:
# dt0 = T / steps, target = 1 / factor
for name, kw in (("variable_american_k1", dict(american=1, steps=48, refine=1, variable_target=0.25)),
                 ("variable_american_k2", dict(american=1, steps=192, refine=2, variable_target=1.0 / 16)),
                 ("variable_american_k3", dict(american=1, steps=768, refine=3, variable_target=1.0 / 64)),
                 ("variable_european_k2", dict(american=0, steps=48, refine=2, variable_target=0.25)),
                 ("variable_american_bdf2", dict(american=1, steps=40, refine=1, variable_target=0.3, scheme="bdf2")),
                 ("variable_european_cn", dict(american=0, steps=40, refine=1, variable_target=0.3, scheme="cn")),
                 ("variable_american_bdf1", dict(american=1, steps=40, refine=2, variable_target=0.2, scheme="bdf1"))):
    kw.update(K=100.0, r=0.04, vol=0.2, T=1.0)
    CASES.append((name, "vanilla", kw))

# README.md:357-375: discrete dividends at the event times (general interpolating events,
# Event.hpp:100-112); dividend_first = which transform the reverse sweep applies first
CASES.append(("dividend_cash_bdf2_k1", "bermudan",
              dict(steps=200, refine=1, scheme="bdf2", K=100.0, r=0.04, alpha=20.0, T=1.0, E=10,
                   dividend=1.0, dividend_kind="cash", dividend_first=0)))
CASES.append(("dividend_cash_first_rannacher_k0", "bermudan",
              dict(steps=100, refine=0, scheme="rannacher", K=100.0, r=0.04, alpha=20.0, T=1.0, E=10,
                   dividend=1.0, dividend_kind="cash", dividend_first=1)))
CASES.append(("dividend_proportional_bdf2_k0", "bermudan",
              dict(steps=100, refine=0, scheme="bdf2", K=110.0, r=0.03, alpha=25.0, T=0.75, E=4,
                   dividend=0.02, dividend_kind="proportional", dividend_first=0)))
CASES.append(("dividend_cash_european_cn_k2", "bermudan",
              dict(steps=80, refine=2, scheme="cn", K=100.0, r=0.04, alpha=20.0, T=1.0, E=10,
                   dividend=2.5, dividend_kind="cash", dividend_first=0, exercise=0)))
CASES.append(("dividend_cash_bdf1_k1", "bermudan",
              dict(steps=60, refine=1, scheme="bdf1", K=95.0, r=0.05, alpha=18.0, T=0.5, E=3,
                   dividend=0.7, dividend_kind="cash", dividend_first=1)))

# examples/unequal_borrowing_lending_rates.cpp: HJB with the interest rate as a control
for lp in (0, 1):
    for k in (0, 1, 2):
        CASES.append(("hjb_%s_bdf2_k%d" % ("long" if lp else "short", k), "hjb",
                      dict(steps=12 * 2 ** k, refine=k, long=lp, scheme="bdf2", K=100.0, r_l=0.03, r_b=0.05,
                           vol=0.3, T=1.0)))
CASES.append(("hjb_short_bdf1_k1", "hjb",
              dict(steps=40, refine=1, long=0, scheme="bdf1", K=90.0, r_l=0.02, r_b=0.07, vol=0.25, T=0.8)))

# examples/jump_diffusion.cpp: European call, BDF1 + explicit correlation-integral jump term.
# The Kou cases k = 3, 4 are rows of the reference's own recorded table
# (examples/jump_diffusion.cpp:184-185: 3.970527 at 265 x 97, 3.971451 at 529 x 192).
for k in (0, 3, 4):
    CASES.append(("jump_kou_k%d" % k, "jump",
                  dict(steps=12 * 2 ** k, refine=k, density="kou", K=100.0, r=0.05, vol=0.15, T=0.25,
                       p_up=0.3445, eta1=3.0465, eta2=3.0775, **{"lambda": 0.1})))
for k in (1, 3):
    CASES.append(("jump_merton_k%d" % k, "jump",
                  dict(steps=12 * 2 ** k, refine=k, density="lognormal", K=100.0, r=0.05, vol=0.15, T=0.25,
                       mu=-0.1, sd=0.45, **{"lambda": 0.1})))
CASES.append(("jump_merton_bdf2", "jump",
              dict(steps=40, refine=2, density="lognormal", K=90.0, r=0.03, vol=0.25, T=0.5,
                   mu=-0.2, sd=0.3, scheme="bdf2", **{"lambda": 0.4})))

# ReverseBDFThree .. ReverseBDFSix (Core/BDF.hpp:139-470,595-942): start-up through the lower
# orders, restart after every event, variable-step coefficients when the last step is clipped
for order in (3, 4, 5, 6):
    sch = "bdf%d" % order
    CASES.append(("european_%s" % sch, "vanilla",
                  dict(american=0, steps=100, refine=2, K=100.0, r=0.04, vol=0.2, T=1.0, scheme=sch)))
    CASES.append(("bermudan_%s_k1" % sch, "bermudan",
                  dict(steps=200, refine=1, scheme=sch, K=100.0, r=0.04, alpha=20.0, T=1.0, E=10)))
CASES.append(("american_bdf3", "vanilla",
              dict(american=1, steps=60, refine=2, K=100.0, r=0.04, vol=0.2, T=1.0, scheme="bdf3")))
CASES.append(("american_bdf6_varied", "vanilla",
              dict(american=1, steps=90, refine=3, K=83.0, r=0.06, vol=0.45, T=1.7, scheme="bdf6")))
CASES.append(("dividend_cash_bdf4_k1", "bermudan",
              dict(steps=120, refine=1, scheme="bdf4", K=100.0, r=0.04, alpha=20.0, T=1.0, E=6,
                   dividend=1.0, dividend_kind="cash", dividend_first=0)))

CASES.append(("hjb_short_bdf4_k2", "hjb",
              dict(steps=48, refine=2, long=0, scheme="bdf4", K=100.0, r_l=0.03, r_b=0.05, vol=0.3, T=1.0)))
# policy iteration under the theta schemes: the explicit half follows the controls of the running inner iteration
# (Rannacher.hpp:60-73 / CrankNicolson.hpp:64-78 call the policy node's A(t0))
for lp, k in ((0, 1), (1, 2)):
    for sch in ("rannacher", "cn"):
        CASES.append(("hjb_%s_%s_k%d" % ("long" if lp else "short", sch, k), "hjb",
                      dict(steps=12 * 2 ** k, refine=k, long=lp, scheme=sch, K=100.0, r_l=0.03, r_b=0.05, vol=0.3, T=1.0)))
CASES.append(("variable_american_bdf4", "vanilla",
              dict(american=1, steps=40, refine=2, variable_target=0.3, scheme="bdf4", K=100.0, r=0.04, vol=0.2, T=1.0)))

# ---- round 2: the BASELINE.json config sizes themselves -----------------------------------
# configs[0]: the rest of the reference's default table (examples/vanilla_options.cpp:36-42,139:
# maximum_refinement 5 -> k = 0 .. 5; k = 6 = 2049 x 49,152 is checked against the oracle live)
for k in (4, 5):
    CASES.append(("american_table_k%d" % k, "vanilla",
                  dict(american=1, steps=12 * 4 ** k, refine=k, K=100.0, r=0.04, vol=0.2, T=1.0)))
# configs[2] at size: tutorial.cpp Bermudan put, local volatility, 2113 nodes x 6400 steps x 10 exercise dates
for sch in ("bdf2", "rannacher"):
    CASES.append(("bermudan_%s_k6_6400" % sch, "bermudan",
                  dict(steps=6400, refine=6, scheme=sch, K=100.0, r=0.04, alpha=20.0, T=1.0, E=10)))
# configs[3] grid: the rows 1057 x 385 and 2113 x 768 of the reference's recorded Kou table
# (examples/jump_diffusion.cpp:186-187: 3.972397, 3.972930)
for k in (5, 6):
    CASES.append(("jump_kou_k%d" % k, "jump",
                  dict(steps=12 * 2 ** k, refine=k, density="kou", K=100.0, r=0.05, vol=0.15, T=0.25,
                       p_up=0.3445, eta1=3.0465, eta2=3.0775, **{"lambda": 0.1})))
CASES.append(("jump_merton_k6", "jump",
              dict(steps=768, refine=6, density="lognormal", K=100.0, r=0.05, vol=0.15, T=0.25,
                   mu=-0.1, sd=0.45, **{"lambda": 0.1})))
# penalty method + events: the README's discrete dividends (README.md:357-375) on an American put
# (examples/vanilla_options.cpp wiring + stepper.add), also with an exercise transform, up to the
# BASELINE grid size
CASES.append(("american_dividend_cash_rannacher_k2", "vanilla",
              dict(american=1, steps=100, refine=2, K=100.0, r=0.04, vol=0.2, T=1.0, E=4,
                   dividend=1.0, dividend_kind="cash", dividend_first=0)))
CASES.append(("american_dividend_prop_bdf2_k3", "vanilla",
              dict(american=1, steps=150, refine=3, K=95.0, r=0.03, vol=0.3, T=1.5, E=5, scheme="bdf2",
                   dividend=0.02, dividend_kind="proportional", dividend_first=1)))
CASES.append(("american_exercise_dividend_bdf1_k2", "vanilla",
              dict(american=1, steps=80, refine=2, K=100.0, r=0.04, vol=0.25, T=1.0, E=4, scheme="bdf1",
                   dividend=0.7, dividend_kind="cash", dividend_first=0, exercise=1)))
CASES.append(("american_dividend_cash_rannacher_k6", "vanilla",
              dict(american=1, steps=200, refine=6, K=100.0, r=0.04, vol=0.2, T=1.0, E=4,
                   dividend=1.0, dividend_kind="cash", dividend_first=0)))
# forward-time classes (TimeIteration<true>, ForwardConstantStepper, ForwardRannacher / CrankNicolson /
# BDFk: IterativeMethod.hpp:1494-1495, Stepper.hpp:41): no example of the reference uses them, so these
# run the classes themselves (qpde_ref forward) — European and American, every scheme family, events
for name, kw in (
        ("forward_european_rannacher_k2", dict(steps=50, refine=2, american=0, scheme="rannacher")),
        ("forward_european_cn_k2", dict(steps=50, refine=2, american=0, scheme="cn", call=1)),
        ("forward_european_bdf2_k2", dict(steps=50, refine=2, american=0, scheme="bdf2")),
        ("forward_european_bdf4_events_k2", dict(steps=50, refine=2, american=0, scheme="bdf4", E=3, exercise=1)),
        ("forward_american_rannacher_k3", dict(steps=100, refine=3, american=1, scheme="rannacher")),
        ("forward_american_bdf1_k1", dict(steps=37, refine=1, american=1, scheme="bdf1", T=0.7, K=93.0, vol=0.35, r=0.02)),
        ("forward_american_cn_dividend_k2", dict(steps=64, refine=2, american=1, scheme="cn", E=4, dividend=1.0,
                                                 dividend_kind="cash")),
        ("forward_localvol_events_k2", dict(steps=64, refine=2, american=0, scheme="rannacher", alpha=20.0, E=5,
                                            dividend=0.02, dividend_kind="proportional", dividend_first=1, exercise=1)),
        ("forward_american_rannacher_k6", dict(steps=200, refine=6, american=1, scheme="rannacher"))):
    base = dict(K=100.0, r=0.04, vol=0.2, T=1.0)
    base.update(kw)
    CASES.append((name, "forward", base))
# general Event<1> transforms: the three-way choice of examples/experimental/glwb.cpp:84-121 (bonus /
# withdrawal / surrender, per-event survival and penalty rates) as the event of a Black-Scholes sweep
# (qpde_ref vanilla event=glwb); the device evaluates it as a QPDE_EV_* program
for name, kw in (("event_glwb_bdf2_k2", dict(american=0, steps=50, refine=2, E=5, scheme="bdf2")),
                 ("event_glwb_american_rannacher_k3", dict(american=1, steps=80, refine=3, E=4, scheme="rannacher")),
                 ("event_glwb_cn_k2", dict(american=0, steps=64, refine=2, E=8, scheme="cn", vol=0.3, K=90.0)),
                 ("event_glwb_bdf2_k6", dict(american=0, steps=200, refine=6, E=10, scheme="bdf2"))):
    base = dict(K=100.0, r=0.04, vol=0.2, T=1.0, event="glwb")
    base.update(kw)
    CASES.append((name, "vanilla", base))
# examples/experimental/glwb.cpp run(k) itself (qpde_ref glwb: GLWBOperator with its source term b(t), an
# event at every year incl. one AT the initial time, zero initial condition; short synthetic tables)
for name, kw in (("glwb_bdf2_k1", dict(years=4, steps=24, refine=1)), ("glwb_bdf2_k2", dict(years=6, steps=72, refine=2)),
                 ("glwb_bdf1_k2", dict(years=5, steps=40, refine=2, scheme="bdf1")),
                 ("glwb_bdf2_k4", dict(years=8, steps=192, refine=4)), ("glwb_bdf2_k6", dict(years=6, steps=144, refine=6))):
    CASES.append((name, "glwb", dict(kw)))
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_cases.npz")
# cases already in the file with the same arguments are kept (pass --all to re-run everything)
have = np.load(GOLDEN, allow_pickle=False) if os.path.exists(GOLDEN) and "--all" not in sys.argv else None
out = {}
names = []
for name, mode, kw in CASES:
    if have is not None and name + "/args" in have.files and str(have[name + "/args"]) == repr(kw):
        names.append(name)
        for key in have.files:
            if key.startswith(name + "/"):
                out[key] = have[key]
        continue
    ref = po.run_ref(mode, **kw)
    names.append(name)
    out[name + "/mode"] = np.array(mode)
    out[name + "/args"] = np.array(repr(kw))
    for key in ("S", "V", "inner", "mask"):
        if key in ref:
            out[name + "/" + key] = ref[key]
    out[name + "/steps"] = np.array(ref["steps"])
    out[name + "/value"] = np.array(ref["value"])
    print(name, ref["steps"], ref["value"])
out["names"] = np.array(names)
np.savez_compressed(GOLDEN, **out)
if "--no-gmwb" in sys.argv:
    sys.exit(0)

# ---------------------------------------------------------------------------
# GMWB (examples/hjbqvi/gmwb.cpp): fixtures of the reference for the §8(a) row 17 path
# (qpde_gmwb2d_solve, oracle/qpde_oracle_gmwb.c).
# ---------------------------------------------------------------------------
gm = {}
for name, kw in (("gmwb_small", dict(refine=0, a_points=11, timesteps=8, control_points=3)),
                 ("gmwb_k0", dict(refine=0)), ("gmwb_k1", dict(refine=1))):
    ref = po.run_ref("gmwb", **kw)
    gm[name + "/args"] = np.array(repr(kw))
    for key in ("W", "A", "V", "Q", "Z", "value", "steps", "mean_inner", "mean_solver"):
        gm[name + "/" + key] = np.asarray(ref[key])
    print(name, len(ref["W"]), len(ref["A"]), ref["steps"], ref["value"], ref["mean_inner"], ref["mean_solver"])
# refinement 3 (513 x 401 nodes, 256 steps): 17 minutes of the reference on one core, so it is run
# only on request (`--gmwb-k3`, or `--gmwb-k3-dump FILE` with the saved stdout of
# `oracle/_ref/qpde_ref gmwb refine=3`); every 4th node in both directions is kept (129 x 101 values)
GM_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "gmwb_reference.npz")
ref3 = None
if "--gmwb-k3-dump" in sys.argv:
    ref3 = po.parse_ref_dump(open(sys.argv[sys.argv.index("--gmwb-k3-dump") + 1]).read())
elif "--gmwb-k3" in sys.argv:
    ref3 = po.run_ref("gmwb", refine=3)
if ref3 is not None:
    nw3, na3 = len(ref3["W"]), len(ref3["A"])
    gm["gmwb_k3/args"] = np.array(repr(dict(refine=3)))
    gm["gmwb_k3/stride"] = np.array(4)
    gm["gmwb_k3/V_strided"] = ref3["V"].reshape(na3, nw3)[::4, ::4].copy()
    for key in ("W", "A", "value", "steps", "mean_inner", "mean_solver"):
        gm["gmwb_k3/" + key] = np.asarray(ref3[key])
    print("gmwb_k3", nw3, na3, ref3["steps"], ref3["value"], ref3["mean_inner"])
elif os.path.exists(GM_PATH):
    old = np.load(GM_PATH, allow_pickle=False)
    for key in old.files:
        if key.startswith("gmwb_k3/"):
            gm[key] = old[key]
np.savez_compressed(GM_PATH, **gm)
