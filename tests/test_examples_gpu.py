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


def _rows(text):
    """The rows of the convergence table: what follows the header line up to the first empty line (the
    solution on the print grid comes after it, Results.hpp:148-164)."""
    rows = []
    for ln in text.splitlines()[1:]:
        if not ln.strip():
            break
        rows.append(ln)
    return rows


@pytest.mark.parametrize("case,args", [
    ("bermudan_bdf2_k1", ["dt=0.005", "refine=1"]),
    ("dividend_cash_bdf2_k1", ["dt=0.005", "refine=1", "D=1"]),
])
def test_tutorial_example(examples, golden, oracle, case, args):
    out = subprocess.run([os.path.join(examples, "tutorial_b200")] + args
                         + ["precision=12", "print_min=80", "print_step=5", "print_max=120"],
                         check=True, capture_output=True, text=True)
    tab = _table(out.stdout)
    S_ref, V_ref = golden[case + "/S"], golden[case + "/V"]
    for S, V in tab:
        ref = oracle.lib().qpo_interp(len(S_ref), oracle._dp(np.ascontiguousarray(S_ref)),
                                      oracle._dp(np.ascontiguousarray(V_ref)), float(S))
        assert rel_err(V, ref) <= 1e-9 + 1e-11, (case, S, V, ref)     # 12 printed digits


def test_vanilla_options_example_table(examples, golden):
    out = subprocess.run([os.path.join(examples, "vanilla_options_b200"), "is_american=true", "is_call=false",
                          "maximum_refinement=2"], check=True, capture_output=True, text=True)
    rows = _rows(out.stdout)
    assert len(rows) == 3
    for k, row in enumerate(rows):
        cols = row.split()
        name = "american_table_k%d" % k      # the reference's own convergence table (vanilla_options.cpp)
        assert int(float(cols[0])) == len(golden[name + "/S"])
        assert int(float(cols[1])) == int(golden[name + "/steps"])
        assert rel_err(float(cols[3]), float(golden[name + "/value"])) <= 1e-6   # printed with 6 digits


def test_vanilla_options_example_variable_timestepping(examples, golden):
    """use_variable_timestepping of vanilla_options.cpp:52-58 through ReverseVariableStepper."""
    out = subprocess.run([os.path.join(examples, "vanilla_options_b200"), "is_american=true", "is_call=false",
                          "minimum_refinement=1", "maximum_refinement=3", "use_variable_timestepping=true"], check=True,
                         capture_output=True, text=True)
    rows = _rows(out.stdout)
    assert len(rows) == 3
    for k, row in zip((1, 2, 3), rows):
        cols = row.split()
        name = "variable_american_k%d" % k
        assert int(float(cols[0])) == len(golden[name + "/S"])
        assert int(float(cols[1])) == int(golden[name + "/steps"])
        assert rel_err(float(cols[3]), float(golden[name + "/value"])) <= 1e-6   # printed with 6 digits


def test_jump_diffusion_example_reproduces_the_reference_table(examples, golden):
    """examples/jump_diffusion.cpp:179-190 records this table for the Kou model (the only known-answer
    table in the reference): 3.970527 at 265 nodes x 97 steps, 3.971451 at 529 x 192, 3.972397 at
    1057 x 385, 3.972930 at 2113 x 768 (the grid of BASELINE.json configs[3])."""
    out = subprocess.run([os.path.join(examples, "jump_diffusion_b200"), "jump_type=double_exponential",
                          "minimum_refinement=3", "maximum_refinement=6"], check=True, capture_output=True, text=True)
    rows = [r.split() for r in _rows(out.stdout)]
    assert len(rows) == 4
    for row, (nodes, steps, value, k) in zip(rows, ((265, 97, 3.970527, 3), (529, 192, 3.971451, 4),
                                                    (1057, 385, 3.972397, 5), (2113, 768, 3.972930, 6))):
        assert int(float(row[0])) == nodes and int(float(row[1])) == steps
        assert abs(float(row[2]) - value) <= 1e-6                                  # the recorded 7 digits
        assert rel_err(float(row[2]), float(golden["jump_kou_k%d/value" % k])) <= 1e-6   # printed with 6 digits
    assert abs(float(rows[1][3]) - 9.242709e-04) <= 2e-9                           # the recorded "Change"
    assert abs(float(rows[2][3]) - 9.454750e-04) <= 2e-9 and abs(float(rows[3][3]) - 5.331950e-04) <= 2e-9
    assert abs(float(rows[3][4]) - 1.773226e+00) <= 2e-5                           # the recorded "Ratio"
    out = subprocess.run([os.path.join(examples, "jump_diffusion_b200"), "minimum_refinement=1",
                          "maximum_refinement=3"], check=True, capture_output=True, text=True)   # Merton (the default)
    rows = [r.split() for r in _rows(out.stdout)]
    for row, k in ((rows[0], 1), (rows[2], 3)):
        assert int(float(row[1])) == int(golden["jump_merton_k%d/steps" % k])
        assert rel_err(float(row[2]), float(golden["jump_merton_k%d/value" % k])) <= 1e-6


@pytest.mark.parametrize("position", ["short", "long"])
def test_unequal_borrowing_lending_rates_example(examples, golden, position):
    out = subprocess.run([os.path.join(examples, "unequal_borrowing_lending_rates_b200"), "maximum_refinement=2",
                          "long_position=%s" % ("true" if position == "long" else "false")], check=True,
                         capture_output=True, text=True)
    rows = [r.split() for r in _rows(out.stdout)]
    assert len(rows) == 3
    for k, row in enumerate(rows):
        name = "hjb_%s_bdf2_k%d" % (position, k)       # outputs of the reference's own example wiring
        inner = golden[name + "/inner"]
        assert int(float(row[0])) == len(golden[name + "/S"]) and int(float(row[1])) == int(golden[name + "/steps"])
        assert abs(float(row[2]) - inner.mean()) <= 1e-6 * max(1., inner.mean())   # "Mean Policy Iterations"
        assert rel_err(float(row[3]), float(golden[name + "/value"])) <= 1e-6


def test_gmwb_example_table(examples):
    """examples/hjbqvi/gmwb.cpp through QuantPDE_B200::GMWB: values and "Mean Policy Iterations" of the
    reference's own runs (tests/golden/gmwb_reference.npz) at refinements 0 and 1."""
    g = np.load(os.path.join(ROOT, "tests", "golden", "gmwb_reference.npz"), allow_pickle=False)
    out = subprocess.run([os.path.join(examples, "gmwb_b200"), "max_refinement=1"], check=True,
                         capture_output=True, text=True)
    rows = [r.split() for r in out.stdout.strip().splitlines()[1:]]
    assert len(rows) == 2
    for row, name in zip(rows, ("gmwb_k0", "gmwb_k1")):
        assert int(row[0]) == g[name + "/V"].size and int(row[3]) == int(g[name + "/steps"])
        assert abs(float(row[6]) - float(g[name + "/mean_inner"])) <= 1e-12       # identical policy iteration counts
        assert rel_err(float(row[8]), float(g[name + "/value"])) <= REL_TOL


def test_batch_example_many_contracts_in_one_sweep(examples, oracle):
    """QuantPDE_B200::Batch from C++: 200 American puts in one device sweep; the printed contracts
    against the oracle"""
    out = subprocess.run([os.path.join(examples, "batch_american_b200"), "contracts=200", "refine=3", "steps=100"],
                         check=True, capture_output=True, text=True)
    lines = out.stdout.strip().splitlines()
    head = lines[0].split()
    assert int(head[1]) == 200 and int(head[3]) == 257
    assert 1.5 < float(head[9]) < 4.0                                   # mean penalty iterations per step
    assert len(lines) == 4
    for ln in lines[1:]:
        t = ln.split()
        K, vol, T, r, V = (float(t[i]) for i in (3, 5, 7, 9, 11))
        o = oracle.american_put(K, r, vol, T, 100, 3)
        assert rel_err(V, oracle.interp(o["S"], o["V"], K)) <= REL_TOL
