"""GPU: the 2-D GMWB impulse-control HJBQVI (BASELINE.json configs[4] path) through the
C ABI against the reference's own outputs (tests/golden/gmwb_reference.npz, made from
examples/hjbqvi/gmwb.cpp on the unmodified reference headers) and against the oracle."""
import os

import numpy as np
import pytest

from helpers import REL_TOL

pytestmark = pytest.mark.gpu


def _solve(qpde, handle, oracle, refine, a_points=51, timesteps=32, **kw):
    return qpde.gmwb_solve(handle, oracle.GMWB_W_AXIS, oracle.uniform_axis(0., 100., a_points), [0., 10.],
                           oracle.uniform_axis(0., 1., 3), refine, timesteps * 2 ** refine, **kw)


def test_gmwb_against_reference_fixtures(qpde, handle, oracle):
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "gmwb_reference.npz"))
    for name, kw in (("gmwb_small", dict(refine=0, a_points=11, timesteps=8)),
                     ("gmwb_k0", dict(refine=0)), ("gmwb_k1", dict(refine=1))):
        res = _solve(qpde, handle, oracle, **kw)
        ref = g[name + "/V"].reshape(res["V"].shape)
        assert res["status"] == 0 and res["n_steps"] == int(g[name + "/steps"])
        assert np.array_equal(res["W"], g[name + "/W"]) and np.array_equal(res["A"], g[name + "/A"])
        err = np.abs(res["V"] - ref) / np.maximum(1., np.abs(ref))
        assert err.max() <= REL_TOL, "%s: %.3e" % (name, err.max())
        assert res["mean_inner"] == float(g[name + "/mean_inner"]), name     # "Mean Policy Iterations"


def test_gmwb_refinement_two_against_oracle(qpde, handle, oracle):
    """257 x 201 nodes, 128 steps: the W-line solve uses the tail equation (257 = 32 * 8 + 1)."""
    res = _solve(qpde, handle, oracle, refine=2)
    o = oracle.gmwb(refine_times=2)
    assert res["status"] == 0 and o["status"] == 0 and res["n_steps"] == o["n_steps"]
    err = np.abs(res["V"] - o["V"]) / np.maximum(1., np.abs(o["V"]))
    assert err.max() <= REL_TOL, "%.3e" % err.max()
    assert res["mean_inner"] == o["mean_inner"]
    # convergence table value at (w, a) = (100, 100): the no-fee GMWB value of the literature
    iw, ia = np.searchsorted(res["W"], 100.), np.searchsorted(res["A"], 100.)
    assert res["W"][iw] == 100. and res["A"][ia] == 100. and 107.6 < res["V"][ia, iw] < 107.8


def test_gmwb_refinement_three_against_the_reference(qpde, handle, oracle):
    """513 x 401 nodes, 256 steps against the reference's own output (17 minutes of qpde_ref on one
    core; tests/golden/gmwb_reference.npz keeps every 4th node in both directions)."""
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "gmwb_reference.npz"))
    res = _solve(qpde, handle, oracle, refine=3)
    stride = int(g["gmwb_k3/stride"])
    ref = g["gmwb_k3/V_strided"]
    assert res["status"] == 0 and res["n_steps"] == int(g["gmwb_k3/steps"]) == 256
    assert np.array_equal(res["W"], g["gmwb_k3/W"]) and np.array_equal(res["A"], g["gmwb_k3/A"])
    got = res["V"][::stride, ::stride]
    assert got.shape == ref.shape == (101, 129)
    err = np.abs(got - ref) / np.maximum(1., np.abs(ref))
    assert err.max() <= REL_TOL, "%.3e" % err.max()
    assert res["mean_inner"] == float(g["gmwb_k3/mean_inner"])              # "Mean Policy Iterations"
    iw, ia = np.searchsorted(res["W"], 100.), np.searchsorted(res["A"], 100.)
    assert abs(res["V"][ia, iw] - float(g["gmwb_k3/value"])) <= REL_TOL * 108.


def test_gmwb_driver_depths_give_the_same_bits(qpde, handle, oracle):
    """QPDE_GMWB_DEPTH 0 (iterations one after the other), 1 (control search overlapped with the line
    pass), 2 (+ speculative trailing pass): identical values and iteration counts."""
    runs = []
    for depth in ("0", "1", "2"):
        os.environ["QPDE_GMWB_DEPTH"] = depth
        try:
            runs.append(_solve(qpde, handle, oracle, refine=2, a_points=65))     # 257 lines: two chunks
        finally:
            del os.environ["QPDE_GMWB_DEPTH"]
    for r in runs[1:]:
        assert np.array_equal(r["V"], runs[0]["V"]) and np.array_equal(r["inner"], runs[0]["inner"])
        assert r["status"] == 0 and r["mean_inner"] == runs[0]["mean_inner"]


def test_gmwb_properties_and_errors(qpde, handle, oracle):
    res = _solve(qpde, handle, oracle, refine=1, sigma=0.3)
    payoff = np.maximum(res["W"][None, :], 0.9 * res["A"][:, None])
    assert np.all(res["V"] >= payoff - 1e-6 * np.maximum(1., payoff))      # no fee: V dominates the exit payoff
    assert np.all(np.diff(res["V"], axis=1) >= -1e-9)                      # increasing in the account value
    fee = _solve(qpde, handle, oracle, refine=1, sigma=0.3, eta=0.01)
    assert np.all(fee["V"] <= res["V"] + 1e-9)                             # a fee can only lower the value
    with pytest.raises(qpde.QpdeError):
        qpde.gmwb_solve(handle, oracle.GMWB_W_AXIS + 1.0, oracle.uniform_axis(0., 100., 11), [0., 10.],
                        oracle.uniform_axis(0., 1., 3), 0, 8)


class _TwoShards:
    """`count` A-blocks on one GPU, swept in ascending A as that many ranks would."""

    def __init__(self, handle, problem, count=2):
        from quantpde_b200 import gmwb_sharded
        cuts = [problem.na * k // count for k in range(count + 1)]
        self.parts = [gmwb_sharded.CudaBackend(handle, problem, cuts[k], cuts[k + 1]) for k in range(count)]
        self.stream = self.parts[0].stream
        self.empty, self.flag = self.parts[0].empty, self.parts[0].flag

    def exit_values(self, x): [p.exit_values(x) for p in self.parts]
    def policy(self, x, v0, h0): [p.policy(x, v0, h0) for p in self.parts]
    def sweep(self, x, v0, h0, moved): [p.sweep(x, v0, h0, moved) for p in self.parts]
    def relerr(self, x, xprev, notdone): [p.relerr(x, xprev, notdone) for p in self.parts]
    def close(self): [p.close() for p in self.parts]


def test_gmwb_shard_api_two_blocks_equal_whole(qpde, handle, oracle):
    """qpde_gmwb2d_shard_*: the grid split in two A-blocks and driven by gmwb_sharded.solve gives
    the bits of qpde_gmwb2d_solve (same sweep order), for one block and for two."""
    from quantpde_b200 import gmwb_sharded
    args = (oracle.GMWB_W_AXIS, oracle.uniform_axis(0., 100., 51), [0., 10.], oracle.uniform_axis(0., 1., 3), 1, 64)
    whole = qpde.gmwb_solve(handle, *args)
    prob = qpde.GmwbProblem(*args)
    one = gmwb_sharded.CudaBackend(handle, prob, 0, prob.na)
    two = _TwoShards(handle, prob)
    for backend in (one, two):
        res = gmwb_sharded.solve(prob, backend, 0, 1)
        backend.close()
        assert res["status"] == 0 and res["n_steps"] == whole["n_steps"] and res["mean_inner"] == whole["mean_inner"]
        assert np.array_equal(res["V"], whole["V"])


def test_gmwb_config_size_shard_counts_give_the_same_bits(qpde, handle, oracle):
    """BASELINE.json configs[4] grid (4097 x 1025 nodes, 129 impulse candidates), 64 time steps: the
    single-GPU solve and the sharded driver with 1, 2 and 4 A-blocks agree bit for bit."""
    from quantpde_b200 import gmwb_sharded
    args = (oracle.GMWB_W_AXIS, oracle.uniform_axis(0., 100., 17), [0., 10.], oracle.uniform_axis(0., 1., 3), 6, 64)
    whole = qpde.gmwb_solve(handle, *args)
    assert whole["V"].shape == (1025, 4097) and whole["status"] == 0
    prob = qpde.GmwbProblem(*args)
    for count in (1, 2, 4):
        backend = _TwoShards(handle, prob, count)
        res = gmwb_sharded.solve(prob, backend, 0, 1)
        backend.close()
        assert res["status"] == 0 and res["mean_inner"] == whole["mean_inner"], count
        assert np.array_equal(res["V"], whole["V"]), count


def test_gmwb_two_ranks_nccl(qpde, oracle, tmp_path):
    """World size 2 over NCCL (needs two GPUs; the driver's scaling run has them)."""
    import subprocess
    import sys
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("one GPU visible")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = str(tmp_path / "sharded.npz")
    subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                    "--master-addr", "127.0.0.1", "--master-port", "29631",
                    os.path.join(root, "tools", "gmwb_sharded_run.py"), "--refine", "1", "--out", out],
                   check=True, cwd=root, timeout=900)
    got = np.load(out)
    with qpde.Handle(0) as h:
        whole = _solve(qpde, h, oracle, refine=1)
    assert np.array_equal(got["V"], whole["V"]) and float(got["mean_inner"]) == whole["mean_inner"]


def test_gmwb_library_sharded_two_ranks(qpde, oracle, tmp_path):
    """qpde_comm_init + qpde_gmwb2d_solve_sharded at world size 2: NCCL inside the library, torch only
    ships the communicator id.  Bit-identical to the single-GPU solve (needs two GPUs)."""
    import subprocess
    import sys
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("one GPU visible")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = str(tmp_path / "sharded_library.npz")
    subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                    "--master-addr", "127.0.0.1", "--master-port", "29633",
                    os.path.join(root, "tools", "gmwb_sharded_library.py"), "--refine", "2", "--out", out],
                   check=True, cwd=root, timeout=900)
    got = np.load(out)
    with qpde.Handle(0) as h:
        whole = _solve(qpde, h, oracle, refine=2)
    assert np.array_equal(got["V"], whole["V"]) and float(got["mean_inner"]) == whole["mean_inner"]


def test_sharded_solve_needs_a_communicator(qpde, handle, oracle):
    with pytest.raises(qpde.QpdeError):
        qpde.gmwb_solve(handle, oracle.GMWB_W_AXIS, oracle.uniform_axis(0., 100., 11), [0., 10.], [0., .5, 1.], 0, 8,
                        sharded=True)


@pytest.mark.parametrize("depth,partition", [("1", "0"), ("2", "1"), ("2", "0")])
def test_split_search_and_assembly_equals_the_fused_kernel(qpde, handle, oracle, monkeypatch, depth, partition):
    """The sharded code path on one rank (QPDE_GMWB_SPLIT=1): control search -> one decision word per node ->
    row assembly from the words, with and without the trailing pass and the SM partition (green contexts).
    Bit-identical to the fused search + assembly kernel."""
    whole = _solve(qpde, handle, oracle, refine=2)
    monkeypatch.setenv("QPDE_GMWB_SPLIT", "1")
    monkeypatch.setenv("QPDE_GMWB_DEPTH", depth)
    monkeypatch.setenv("QPDE_GMWB_PARTITION", partition)
    res = _solve(qpde, handle, oracle, refine=2)
    assert res["status"] == 0 and res["mean_inner"] == whole["mean_inner"]
    assert np.array_equal(res["V"], whole["V"])


def test_gmwb_problems_of_several_threads_overlap_and_keep_their_bits(qpde, handle, oracle):
    """Several GMWB problems at once: one handle (its own streams) per host thread.  The line pass of a
    problem occupies one cluster of 8 SMs, so independent problems fill the rest of the GPU
    (tools/profile_gmwb_concurrent.py: 3.7x the throughput at 8 problems); every problem returns the bits
    of its stand-alone run."""
    import threading
    sig = [.2, .25, .3, .35]
    alone = [_solve(qpde, handle, oracle, refine=1, sigma=s) for s in sig]
    handles = [qpde.Handle(0) for _ in sig]
    out = [None] * len(sig)

    def run(k):
        out[k] = _solve(qpde, handles[k], oracle, refine=1, sigma=sig[k])
    th = [threading.Thread(target=run, args=(k,)) for k in range(len(sig))]
    for t in th: t.start()
    for t in th: t.join()
    for hh in handles: hh.close()
    for k in range(len(sig)):
        assert out[k] is not None and out[k]["status"] == 0
        assert np.array_equal(out[k]["V"], alone[k]["V"]) and out[k]["mean_inner"] == alone[k]["mean_inner"]
    assert not np.array_equal(out[0]["V"], out[1]["V"])
