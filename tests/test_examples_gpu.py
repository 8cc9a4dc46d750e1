"""GPU: the example programs written against the header-only host API
(quantpde_b200/include/QuantPDE_B200/Core.hpp) — the reference-facing boundary a
C++ user of examples/tutorial.cpp / vanilla_options.cpp would switch to — against
the committed outputs of the reference itself."""
import os
import subprocess

import numpy as np
import pytest

from helpers import REL_TOL, rel_err

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def examples():
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "examples")], check=True)
    return os.path.join(ROOT, "examples", "bin")


def _table(text):
    return np.array([[float(t) for t in ln.split()] for ln in text.strip().splitlines()])


@pytest.mark.parametrize("case,args", [
    ("bermudan_bdf2_k1", ["dt=0.005", "refine=1"]),
    ("dividend_cash_bdf2_k1", ["dt=0.005", "refine=1", "D=1"]),
])
def test_tutorial_example(examples, golden, oracle, case, args):
    out = subprocess.run([os.path.join(examples, "tutorial_b200")] + args, check=True, capture_output=True, text=True)
    tab = _table(out.stdout)
    S_ref, V_ref = golden[case + "/S"], golden[case + "/V"]
    for S, V in tab:
        ref = oracle.lib().qpo_interp(len(S_ref), oracle._dp(np.ascontiguousarray(S_ref)),
                                      oracle._dp(np.ascontiguousarray(V_ref)), float(S))
        assert rel_err(V, ref) <= 1e-9 + 1e-11, (case, S, V, ref)     # 12 printed digits


def test_vanilla_options_example_table(examples, golden):
    out = subprocess.run([os.path.join(examples, "vanilla_options_b200"), "kn=2"], check=True,
                         capture_output=True, text=True)
    rows = out.stdout.strip().splitlines()[1:]
    assert len(rows) == 3
    for k, row in enumerate(rows):
        cols = row.split()
        name = "american_table_k%d" % k      # the reference's own convergence table (vanilla_options.cpp)
        assert int(float(cols[0])) == len(golden[name + "/S"])
        assert int(float(cols[1])) == int(golden[name + "/steps"])
        assert rel_err(float(cols[3]), float(golden[name + "/value"])) <= 1e-6   # printed with 6 digits


def test_vanilla_options_example_variable_timestepping(examples, golden):
    """use_variable_timestepping of vanilla_options.cpp:52-58 through ReverseVariableStepper."""
    out = subprocess.run([os.path.join(examples, "vanilla_options_b200"), "k0=1", "kn=3", "variable=1"], check=True,
                         capture_output=True, text=True)
    rows = out.stdout.strip().splitlines()[1:]
    assert len(rows) == 3
    for k, row in zip((1, 2, 3), rows):
        cols = row.split()
        name = "variable_american_k%d" % k
        assert int(float(cols[0])) == len(golden[name + "/S"])
        assert int(float(cols[1])) == int(golden[name + "/steps"])
        assert rel_err(float(cols[3]), float(golden[name + "/value"])) <= 1e-6   # printed with 6 digits
