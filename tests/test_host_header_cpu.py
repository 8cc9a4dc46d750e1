"""CPU: the header-only host API (quantpde_b200/include/QuantPDE_B200/Core.hpp) is inline-clean — two
translation units that include it link into one program (the reference's headers cannot: global
definitions, one TU only) — and a program built on it fails loudly without a B200
(no CPU fallback)."""
import os
import subprocess
import textwrap

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "quantpde_b200", "lib")


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def test_two_translation_units_link_and_fail_loudly_without_a_gpu(tmp_path, qpde):
    a = tmp_path / "a.cpp"
    b = tmp_path / "b.cpp"
    a.write_text(textwrap.dedent("""
        #include <QuantPDE_B200/Core.hpp>
        using namespace QuantPDE_B200;
        int nodes_of_other_unit(int refine);
        int main() {
            RectilinearGrid1 grid(100. * Axis::special());
            if(grid.refined(2).size() != nodes_of_other_unit(2)) return 3;
            GMWB gmwb;                                  // parameters only: no device needed yet
            if(gmwb.wAxis.size() != 65 || gmwb.aAxis.size() != 51) return 4;
            try {
                B200Solver solver(0);                   // needs an sm_100 device
                BlackScholesJumpDiffusion1 bs(grid, .05, .15, 0., .1, doubleExponential(.3445, 3.0465, 3.0775));
                ReverseConstantStepper stepper(0., .25, .25 / 12);
                bs.setIteration(stepper);
                ReverseBDFThree discretization(grid, bs);
                discretization.setIteration(stepper);
                return 0;
            } catch(const std::exception &e) {
                std::fprintf(stderr, "%s\\n", e.what());
                return 7;
            }
        }
    """))
    b.write_text(textwrap.dedent("""
        #include <QuantPDE_B200/Core.hpp>
        int nodes_of_other_unit(int refine) {
            QuantPDE_B200::RectilinearGrid1 grid(100. * QuantPDE_B200::Axis::special());
            return grid.refined(refine).size();
        }
    """))
    exe = tmp_path / "two_units"
    subprocess.run(["g++", "-std=c++11", "-Wall", "-Werror", "-I" + os.path.join(ROOT, "include"),
                    "-I" + os.path.join(ROOT, "quantpde_b200", "include"), str(a), str(b), "-o", str(exe),
                    "-L" + LIB, "-lqpde_b200", "-Wl,-rpath," + LIB], check=True)
    run = subprocess.run([str(exe)], capture_output=True, text=True)
    if _has_gpu():
        assert run.returncode == 0, run.stderr
    else:
        assert run.returncode == 7 and run.stderr.strip(), (run.returncode, run.stderr)   # an error message, not a fallback


EVENT_PROGRAM = r"""
#include <QuantPDE_B200/Core.hpp>
#include <cstdio>
#include <limits>
using namespace QuantPDE_B200;
int main() {
	// examples/experimental/glwb.cpp:84-121 as an expression
	const Real bonus = .06, GQ = 5.;
	EventExpression S = EventExpression::S(), R = EventExpression::param(0), pen = EventExpression::param(1);
	EventExpression best = max(max(c(1. + bonus) * V(S / c(1. + bonus)), V(max(S - c(GQ), c(0.))) + R * c(GQ)),
		select(S > c(GQ), R * (c(GQ) + (c(1.) - pen) * (S - c(GQ))), c(-std::numeric_limits<Real>::infinity())));
	std::vector<int32_t> words; std::vector<Real> consts;
	best.lower(words, consts);
	for(int32_t w : words) std::printf("%d ", (int) w);
	std::printf("\n");
	for(Real k : consts) std::printf("%a ", k);
	std::printf("\n");
	// the stepper keeps it with its parameter table; forward steppers and Forward* discretisations exist
	ReverseConstantStepper stepper(0., 1., .01);
	stepper.addEvents(2, best, {{.99, .03}, {.98, .02}});
	ForwardConstantStepper forward(0., 1., .01);
	RectilinearGrid1 grid(Axis::special());
	BlackScholes1 bs(grid, .04, .2, 0.);
	ForwardRannacher fr(grid, bs); ForwardBDFTwo fb(grid, bs);
	return stepper.eventDates() == 2 && forward.isForward() && fr.forward && fb.forward ? 0 : 1;
}
"""


def test_event_expression_lowers_to_a_program_with_the_lambdas_value(tmp_path):
    """EventExpression (Core.hpp): the GLWB three-way choice written as a C++ expression lowers to QPDE_EV_*
    words whose value — interpreted here — equals the lambda of examples/experimental/glwb.cpp:84-121 bit for bit."""
    import numpy as np
    src = tmp_path / "ev.cpp"
    src.write_text(EVENT_PROGRAM)
    exe = str(tmp_path / "ev")
    subprocess.run(["g++", "-std=c++11", "-Wall", "-Werror", "-I" + os.path.join(ROOT, "include"),
                    "-I" + os.path.join(ROOT, "quantpde_b200", "include"), str(src), "-o", exe,
                    "-L" + os.path.join(ROOT, "quantpde_b200", "lib"), "-lqpde_b200",
                    "-Wl,-rpath," + os.path.join(ROOT, "quantpde_b200", "lib")], check=True)
    out = subprocess.run([exe], check=True, capture_output=True, text=True).stdout.splitlines()
    words = [int(t) for t in out[0].split()]
    consts = [float.fromhex(t) for t in out[1].split()]
    xs = np.linspace(0., 200., 41); vs = np.maximum(100. - xs, 0.) + .01 * xs

    def V(at):
        return float(np.interp(at, xs, vs))

    def run(S, params):
        st, pc = [], 0
        while True:
            op = words[pc]; pc += 1
            if op == 0:
                assert len(st) == 1
                return st[0]
            if op == 1: st.append(S)
            elif op == 2: st.append(consts[words[pc]]); pc += 1
            elif op == 3: st.append(params[words[pc]]); pc += 1
            elif op == 10: st[-1] = -st[-1]
            elif op == 11: st[-1] = V(st[-1])
            elif op == 13:
                f = st.pop(); t = st.pop(); st[-1] = t if st[-1] != 0. else f
            else:
                b = st.pop(); a = st[-1]
                st[-1] = {4: a + b, 5: a - b, 6: a * b, 7: a / b if b else float("inf"), 8: b if b > a else a,
                          9: b if b < a else a, 12: 1. if a > b else 0.}[op]
            assert len(st) <= 8

    bonus, GQ = .06, 5.
    for S in (0., 3., 5., 5.0001, 60., 100., 140., 400.):
        for Rn, pen in ((.99, .03), (.5, 0.)):
            best = (1. + bonus) * V(S / (1. + bonus))
            tmp = V(max(S - GQ, 0.)) + Rn * GQ
            best = tmp if tmp > best else best
            if S > GQ:
                tmp = Rn * (GQ + (1. - pen) * (S - GQ))
                best = tmp if tmp > best else best
            assert run(S, (Rn, pen)) == best


def test_axis_cluster_reproduces_the_reference_axes(tmp_path):
    """Axis::cluster (Core/Axis.hpp:51-90,238-257) on the host gives the ticks the reference's own program dumped
    (tests/golden/hjbqvi1d_reference.npz: X0 = cluster(-2, 0, 2, 32, 10), Z0 = cluster(-2, 0, 2, 16, 10)) bit for
    bit, and the ExchangeRate / OptimalConsumption classes carry the examples' default parameters."""
    import numpy as np
    src = tmp_path / "cluster.cpp"
    src.write_text(textwrap.dedent("""
        #include <QuantPDE_B200/Core.hpp>
        #include <cstdio>
        using namespace QuantPDE_B200;
        int main() {
            ExchangeRate e;
            OptimalConsumption o;
            const Axis x = e.xAxis(), z = e.impulseAxis(), q = e.controlAxis();
            std::printf("%d %d %d %d\\n", (int) x.size(), (int) z.size(), (int) q.size(), o.spatialPoints);
            for(Index i = 0; i < x.size(); ++i) std::printf("%a ", x[i]);
            std::printf("\\n");
            for(Index i = 0; i < z.size(); ++i) std::printf("%a ", z[i]);
            std::printf("\\n");
            for(Index i = 0; i < q.size(); ++i) std::printf("%a ", q[i]);
            std::printf("\\n");
            ExchangeRate off; off.m = .25;                 // the flipped branch of nonsymmetricCluster on one side
            const Axis xo = off.xAxis();
            for(Index i = 0; i < xo.size(); ++i) std::printf("%a ", xo[i]);
            std::printf("\\n");
            return e.rho == .02 && e.lambda == 1. && o.gamma == .3 && o.boundary == 200. ? 0 : 1;
        }
    """))
    exe = tmp_path / "cluster"
    subprocess.run(["g++", "-std=c++11", "-Wall", "-Werror", "-I" + os.path.join(ROOT, "include"),
                    "-I" + os.path.join(ROOT, "quantpde_b200", "include"), str(src), "-o", str(exe),
                    "-L" + LIB, "-lqpde_b200", "-Wl,-rpath," + LIB], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.splitlines()
    g = np.load(os.path.join(ROOT, "tests", "golden", "hjbqvi1d_reference.npz"))
    assert out[0].split() == ["33", "17", "9", "20"]
    for line, key in zip(out[1:5], ("er_k0/X0", "er_k0/Z0", "er_k0/Q0", "er_parity_off_centre/X0")):
        assert np.array_equal(np.array([float.fromhex(t) for t in line.split()]), g[key]), key
