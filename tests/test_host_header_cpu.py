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
