"""CPU: the C-ABI library loads, exports every symbol include/qpde_b200.h
declares, its host-only helpers agree with the oracle, and it fails loudly
without a GPU (no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_functions():
    src = open(os.path.join(ROOT, "include", "qpde_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(qpde_[a-z0-9_]+)\s*\(", src)))


def test_exports_match_header(qpde):
    declared = _declared_functions()
    assert len(declared) >= 20
    L = qpde.lib()
    for name in declared:
        assert hasattr(L, name), "libqpde_b200.so does not export %s" % name
    assert sorted(qpde.EXPORTS) == declared
    assert L.qpde_abi_version() == qpde.ABI_VERSION


def test_struct_layout_matches_header(qpde):
    # compile a tiny C program against the header and compare sizeof/offsetof
    import subprocess, tempfile
    prog = r'''
#include <stdio.h>
#include <stddef.h>
#include "qpde_b200.h"
int main(void) {
  printf("%zu %zu %zu %zu %zu %zu %zu %zu\n", sizeof(qpde_bs1d_batch), sizeof(qpde_bs1d_result),
    sizeof(qpde_step), offsetof(qpde_bs1d_batch, T), offsetof(qpde_bs1d_batch, strike),
    offsetof(qpde_bs1d_batch, tolerance), offsetof(qpde_bs1d_batch, kernel),
    offsetof(qpde_bs1d_result, active));
  printf("%zu %zu %zu %zu %zu %zu\n", sizeof(qpde_hjbqvi1d_batch), sizeof(qpde_hjbqvi1d_result),
    offsetof(qpde_hjbqvi1d_batch, params_stride), offsetof(qpde_hjbqvi1d_batch, T),
    offsetof(qpde_hjbqvi1d_batch, max_sweeps), offsetof(qpde_hjbqvi1d_result, status));
  return 0; }'''
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "t.c"), "w").write(prog)
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), os.path.join(d, "t.c"),
                               "-o", os.path.join(d, "t")])
        got = [int(x) for x in subprocess.check_output([os.path.join(d, "t")]).split()]
    B, R = qpde.Batch, qpde.Result
    want = [C.sizeof(B), C.sizeof(R), C.sizeof(qpde.Step), B.T.offset, B.strike.offset,
            B.tolerance.offset, B.kernel.offset, R.active.offset]
    HB, HR = qpde.Hjbqvi1dBatch, qpde.Hjbqvi1dResult
    want += [C.sizeof(HB), C.sizeof(HR), HB.params_stride.offset, HB.T.offset, HB.max_sweeps.offset, HR.status.offset]
    assert got == want


def test_schedule_helper_matches_oracle(qpde, oracle):
    cases = [(1.0, 1.0 / 12, ()), (0.7316, 0.7316 / 1000, ()), (2.0, 0.013, ()),
             (1.0, 0.01, tuple(1.0 / 10 * e for e in range(10))),
             (1.3, 1.3 / 77, (0.0, 0.4, 0.4, 1.1))]
    for T, dt, ev in cases:
        mine = qpde.make_schedule(T, dt, ev)
        theirs, n = oracle.schedule(T, dt, ev)
        assert len(mine) == n
        for i in range(n):
            a, b = mine[i], theirs[i]
            assert (a.t, a.dt, a.same_dt, a.event_after, a.user_events) == \
                   (b.t, b.dt, b.same_dt, b.event_after, b.user_events)


def test_forward_schedule_helper_matches_oracle(qpde, oracle):
    """ForwardConstantStepper's loop (TimeIteration<true>): time runs up from 0, smallest event first,
    the terminal NullEvent at T; the oracle's replay is pinned bit for bit by the forward_* fixtures."""
    cases = [(1.0, 1.0 / 12, ()), (0.7316, 0.7316 / 1000, ()), (2.0, 0.013, ()),
             (1.0, 0.01, tuple(1.0 / 10 * e for e in range(1, 11))),
             (1.3, 1.3 / 77, (0.4, 0.4, 1.1, 1.3)), (0.9, 0.9 / 64, tuple(0.9 / 7 * e for e in range(1, 8)))]
    for T, dt, ev in cases:
        mine = qpde.make_schedule(T, dt, ev, forward=True)
        theirs, n = oracle.schedule(T, dt, ev, forward=True)
        assert len(mine) == n
        for i in range(n):
            a, b = mine[i], theirs[i]
            assert (a.t, a.dt, a.same_dt, a.event_after, a.user_events) == \
                   (b.t, b.dt, b.same_dt, b.event_after, b.user_events)
        assert mine[n - 1].t == T and mine[n - 1].event_after == 1


def test_jump_setup_matches_oracle(qpde, oracle):
    """qpde_jump_setup (product host code) vs the oracle's restatement of the
    reference's construction-time quadrature: bit-identical x0, dx, kappa, weights."""
    sp = oracle.special_axis()
    axis = oracle.axis_union(oracle.axis_union(sp * 100., sp * 100.), np.array([100000.]))
    for refine, density, params in ((0, "lognormal", (-0.1, 0.45)), (2, "kou", (0.3445, 3.0465, 3.0775)),
                                    (3, "lognormal", (-0.25, 0.3))):
        S = qpde.refine_axis(axis, refine)
        assert np.array_equal(S, oracle.refine(axis, refine))
        x0, dx, kappa, w = qpde.jump_setup(S, density, params)
        ox0, odx, ow, okappa = oracle.jump_setup(S, density, params)
        assert (x0, dx, kappa) == (ox0, odx, okappa)
        assert np.array_equal(w, ow)
        assert abs(w.sum() - 1.0) < 1e-3          # the weights are a probability mass function
    # closed form for the lognormal density: kappa = exp(mu + sigma^2 / 2) - 1
    assert abs(qpde.jump_setup(S, "lognormal", (-0.1, 0.45))[2] - (np.exp(-0.1 + 0.45 ** 2 / 2) - 1)) < 1e-6


def test_refined_size(qpde):
    assert qpde.refined_size(33, 6) == 2049
    assert qpde.refined_size(34, 6) == 2113
    assert qpde.refined_size(33, 0) == 33


def test_no_cpu_fallback(qpde):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(qpde.QpdeError) as e:
        qpde.Handle(0)
    assert e.value.code == qpde.ERR_NO_DEVICE


def test_product_never_touches_oracle():
    # the product path may not import, link or execute anything under oracle/
    pkg = os.path.join(ROOT, "quantpde_b200")
    for base, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".hpp", ".h", ".cpp", "Makefile")):
                txt = open(os.path.join(base, f), errors="ignore").read()
                assert "oracle" not in txt.lower() or f == "__init__.py" and "oracle" not in txt.lower(), \
                    "%s mentions the oracle" % os.path.join(base, f)
