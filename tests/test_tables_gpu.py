"""GPU: the example programs reproduce the stdout of the reference's own example programs byte for
byte ("same convergence-table order"): same invocation (`prog CONF_FILE`, the JSON
keys of the reference's examples), same table (Modules/Utilities/Results.hpp:79-168), same solution
on the print grid.  Fixtures: tests/golden/tables/*.txt, recorded by tests/golden/make_tables.py from
the unmodified examples/*.cpp of the reference compiled against oracle/shim.  Only the timing
column (the last 23 characters of a table row) is masked."""
import os
import subprocess

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TABLES = os.path.join(ROOT, "tests", "golden", "tables")

PROGRAM = {"vanilla": "vanilla_options_b200", "jump": "jump_diffusion_b200", "glwb": "glwb_b200",
           "hjb": "unequal_borrowing_lending_rates_b200", "tutorial": "tutorial_b200"}
CASES = sorted(f[:-4] for f in os.listdir(TABLES) if f.endswith(".txt"))


@pytest.fixture(scope="module")
def examples():
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "examples")], check=True)
    return os.path.join(ROOT, "examples", "bin")


def mask_timing(text):
    """Blanks the "Timing (Seconds)" column of the table rows (the block before the first empty line)."""
    out, table = [], True
    for k, ln in enumerate(text.splitlines()):
        if ln.strip() == "":
            table = False
        if table and k > 0 and "Timing (Seconds)" in text.splitlines()[0]:
            ln = ln[:-23] + " " * 23
        out.append(ln)
    return "\n".join(out)


def assert_same_stdout(got, want):
    """Byte for byte, the timing column apart.  One allowance: "Change" and "Ratio" are differences and
    quotients of values that agree with the reference to ~1e-12 relative, so their 7th printed digit may
    differ by one unit — those two columns (the last two before the timing) may differ by 2e-6 relative, in
    place, with every other character of the line equal."""
    got, want = mask_timing(got).splitlines(), mask_timing(want).splitlines()
    assert len(got) == len(want)
    derived = "Change" in want[0]
    for k, (g, w) in enumerate(zip(got, want)):
        if g == w:
            continue
        assert derived and k > 0 and len(g) == len(w), (g, w)
        n = len(w.rstrip())
        assert g[:n - 46] == w[:n - 46], (g, w)                  # everything up to the Change column
        for a, b in zip(g[n - 46:].split(), w[n - 46:].split()):
            assert a == b or abs(float(a) - float(b)) <= 2e-6 * abs(float(b)), (g, w)


@pytest.mark.parametrize("case", CASES)
def test_stdout_equals_the_reference_program(examples, case):
    prog = PROGRAM[case.split("_")[0]]
    conf = os.path.join(TABLES, case + ".json")
    args = [os.path.join(examples, prog)] + ([conf] if os.path.exists(conf) else [])
    got = subprocess.run(args, check=True, capture_output=True, text=True).stdout
    want = open(os.path.join(TABLES, case + ".txt")).read()
    assert_same_stdout(got, want)


def test_interactive_mode_reads_the_configuration_from_stdin(examples):
    """`prog -i < conf` (Configuration.hpp:36-76)"""
    conf = open(os.path.join(TABLES, "vanilla_digital_put.json")).read()
    got = subprocess.run([os.path.join(examples, "vanilla_options_b200"), "-i"], input=conf, check=True,
                         capture_output=True, text=True).stdout
    assert_same_stdout(got, open(os.path.join(TABLES, "vanilla_digital_put.txt")).read())
