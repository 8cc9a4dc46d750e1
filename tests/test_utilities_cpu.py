"""CPU: QuantPDE_B200/Utilities.hpp — the JSON configuration reader and the convergence-table writer
that stand in for Modules/Utilities/Configuration.hpp and Results.hpp of the reference."""
import json
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TABLES = os.path.join(ROOT, "tests", "golden", "tables")
INC = ["-I" + os.path.join(ROOT, "include"), "-I" + os.path.join(ROOT, "quantpde_b200", "include")]
LIB = ["-L" + os.path.join(ROOT, "quantpde_b200", "lib"), "-lqpde_b200",
       "-Wl,-rpath," + os.path.join(ROOT, "quantpde_b200", "lib")]


def _build(tmp_path, name, source):
    exe = str(tmp_path / name)
    subprocess.run(["g++", "-std=c++11", "-O1", "-Wall", "-Werror"] + INC + [source, "-o", exe] + LIB, check=True)
    return exe


def _mask(text):
    lines = text.splitlines()
    out, table = [], True
    for k, ln in enumerate(lines):
        if ln.strip() == "":
            table = False
        out.append(ln[:-23] + " " * 23 if table and k > 0 else ln)
    return "\n".join(out)


def test_results_buffer_reproduces_the_reference_table_byte_for_byte(tmp_path, golden):
    """The recorded solutions of the reference (tests/golden/reference_cases.npz: american_table_k0..k5,
    examples/vanilla_options.cpp with is_american, is_call = false) through ResultsBuffer1 + Solution:
    stdout must equal the reference program's own (tests/golden/tables/vanilla_american_put.txt) except
    for the timing column — headers, fixed / scientific notation per column, Change / Ratio, the
    solution on the print grid."""
    exe = _build(tmp_path, "replay_table", os.path.join(ROOT, "tests", "cpp", "replay_table.cpp"))
    levels = 6
    lines = [float(100).hex(), float(0).hex(), float(10).hex(), float(200).hex(), "3", "Nodes", "Steps",
             "Mean Policy Iterations", str(levels)]
    for k in range(levels):
        name = "american_table_k%d" % k
        S, V, inner = golden[name + "/S"], golden[name + "/V"], golden[name + "/inner"]
        steps = int(golden[name + "/steps"])
        mean = float(inner[:steps].astype(np.float64).sum()) / steps
        lines.append("3 %s %s %s" % (float(len(S)).hex(), float(steps).hex(), mean.hex()))
        lines.append(str(len(S)))
        lines.append(" ".join(float(s).hex() for s in S))
        lines.append(" ".join(float(v).hex() for v in V))
    replay = tmp_path / "replay.txt"
    replay.write_text("\n".join(lines) + "\n")
    got = subprocess.run([exe, str(replay)], check=True, capture_output=True, text=True).stdout
    want = open(os.path.join(TABLES, "vanilla_american_put.txt")).read()
    assert _mask(got) == _mask(want)


CONFIG_PROGRAM = r"""
#include <QuantPDE_B200/Core.hpp>
#include <QuantPDE_B200/Utilities.hpp>
using namespace QuantPDE_B200;
int main(int argc, char **argv) {
	Configuration c = getConfiguration(argc, argv);
	const int kn = getInt(c, "maximum_refinement", 5);
	const Real vol = getReal(c, "volatility", .2);
	const bool am = getBool(c, "is_american", false);
	const std::string jt = getString(c, "jump_type", "lognormal");
	const Real fromInt = getReal(c, "strike_price", 100.);       // an integer literal read as a Real
	RectilinearGrid1 g = getGrid(c, "initial_grid", RectilinearGrid1(Axis { 0., 1., 2. }));
	RectilinearGrid1 d = getGrid(c, "other_grid", RectilinearGrid1(Axis { 0., .5 }));
	std::printf("%d %.17g %d %s %.17g %d %.17g %d\n", kn, vol, (int) am, jt.c_str(), fromInt, g.size(), g.nodes()[2], d.size());
	std::cout << c << std::endl;
	return 0;
}
"""


def test_configuration_reads_the_reference_keys_and_writes_back_the_defaults(tmp_path):
    """Configuration.hpp:17-34,78-134: get*(configuration, key, default) returns the entry or the default
    and stores what was used; the printed configuration is itself valid JSON holding every parameter."""
    src = tmp_path / "conf.cpp"
    src.write_text(CONFIG_PROGRAM)
    exe = _build(tmp_path, "conf", str(src))
    conf = tmp_path / "c.json"
    conf.write_text('// a comment, as JsonCpp tolerates\n{"maximum_refinement": 3, "volatility": 2.5e-1, /* inline */ '
                    '"is_american": true, "strike_price": 95, "initial_grid": [[0, 10.5, 20, 1e3]], "note": "a \\"q\\""}')
    out = subprocess.run([exe, str(conf)], check=True, capture_output=True, text=True).stdout
    first, rest = out.split("\n", 1)
    assert first == "3 0.25 1 lognormal 95 4 20 2"
    doc = json.loads(rest)
    assert doc["maximum_refinement"] == 3 and doc["volatility"] == 0.25 and doc["is_american"] is True
    assert doc["jump_type"] == "lognormal" and doc["strike_price"] == 95.0 and doc["note"] == 'a "q"'
    assert doc["initial_grid"] == [[0, 10.5, 20, 1000.0]] and doc["other_grid"] == [[0.0, 0.5]]
    # -i reads stdin; key=value overrides one entry; no arguments = all defaults
    out = subprocess.run([exe, "-i", "volatility=0.5", "jump_type=double_exponential"], input='{"maximum_refinement": 1}',
                         check=True, capture_output=True, text=True).stdout
    assert out.split("\n", 1)[0] == "1 0.5 0 double_exponential 100 3 2 2"
    out = subprocess.run([exe], check=True, capture_output=True, text=True).stdout
    assert out.split("\n", 1)[0] == "5 0.20000000000000001 0 lognormal 100 3 2 2"
    bad = subprocess.run([exe, "-i"], input='{"maximum_refinement": ', capture_output=True, text=True)
    assert bad.returncode != 0


def test_help_exits_with_2(tmp_path):
    src = tmp_path / "conf.cpp"
    src.write_text(CONFIG_PROGRAM)
    exe = _build(tmp_path, "conf", str(src))
    r = subprocess.run([exe, "-h"], capture_output=True, text=True)
    assert r.returncode == 2 and "CONF_FILE" in r.stderr
