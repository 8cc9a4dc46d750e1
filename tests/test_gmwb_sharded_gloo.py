"""CPU, world_size 2, gloo: the A-axis sharded GMWB driver (quantpde_b200/gmwb_sharded.py)
— block ownership, the ascending-A broadcast chain and the flag reductions — with the
oracle's per-line-range pieces standing in for the device kernels.  The sharded sweep
order equals the monolithic one, so the result must equal oracle.gmwb() bit for bit."""
import ctypes as C
import os
import socket

import numpy as np
import torch
import torch.multiprocessing as mp

from quantpde_b200 import gmwb_sharded
from quantpde_b200.shard import contract_slice


class OracleBackend:
    def __init__(self, po, gm, keep, a_begin, a_end):
        self.po, self.gm, self.keep, self.lo, self.hi = po, gm, keep, a_begin, a_end
        L = po.lib()
        L.qpo_gmwb_state_new.restype = C.c_void_p
        L.qpo_gmwb_sweep.restype = C.c_double
        L.qpo_gmwb_relerr.restype = C.c_double
        self.L = L
        self.state = C.c_void_p(L.qpo_gmwb_state_new(C.byref(gm)))

    @staticmethod
    def _p(t):
        return C.cast(t.data_ptr(), C.POINTER(C.c_double))

    def empty(self, n): return torch.zeros(n, dtype=torch.float64)
    def flag(self): return torch.zeros(1, dtype=torch.int32)
    def exit_values(self, x): self.L.qpo_gmwb_exit(C.byref(self.gm), self.lo, self.hi, self._p(x))
    def policy(self, x, v0, h0): self.L.qpo_gmwb_policy(self.state, self._p(x), self.lo, self.hi)

    def sweep(self, x, v0, h0, moved):
        worst = self.L.qpo_gmwb_sweep(self.state, self._p(x), self._p(v0), C.c_double(h0), self.lo, self.hi)
        if worst > 1e-15:
            moved.fill_(1)

    def relerr(self, x, xprev, notdone):
        if self.L.qpo_gmwb_relerr(C.byref(self.gm), self._p(x), self._p(xprev), self.lo, self.hi) > self.gm.tolerance:
            notdone.fill_(1)


class _Problem:
    pass


def _setup(po):
    W = po.GMWB_W_AXIS.copy(); A = po.uniform_axis(0., 100., 11); z = po.uniform_axis(0., 1., 3); q = np.array([0., 10.])
    gm = po.Gmwb()
    gm.n_w, gm.n_a, gm.W, gm.A = len(W), len(A), po._dp(W), po._dp(A)
    gm.n_q, gm.q, gm.n_z, gm.zfrac = 2, po._dp(q), len(z), po._dp(z)
    gm.T, gm.timesteps = 10., 8
    gm.r, gm.eta, gm.sigma, gm.kfee, gm.c = .05, 0., .2, .1, 0.
    gm.scaling_factor, gm.tolerance, gm.max_inner = 1e-2, 1e-6, 1000
    prob = _Problem(); prob.nw, prob.na = len(W), len(A)
    prob.struct = _Problem(); prob.struct.T, prob.struct.timesteps = 10., 8
    return gm, (W, A, z, q), prob


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, out_dir):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import pyoracle as po
    gm, keep, prob = _setup(po)
    lo, hi = contract_slice(prob.na, rank, world)
    res = gmwb_sharded.solve(prob, OracleBackend(po, gm, keep, lo, hi), rank, world)
    np.save(os.path.join(out_dir, "rank%d.npy" % rank), res["V"])
    np.save(os.path.join(out_dir, "meta%d.npy" % rank), np.array([res["n_steps"], res["mean_inner"], res["status"]]))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharded_equals_monolithic(tmp_path):
    mp.spawn(_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    from oracle import pyoracle as po
    o = po.gmwb(refine_times=0, a_points=11, timesteps=8)
    for r in range(2):
        V = np.load(tmp_path / ("rank%d.npy" % r))
        meta = np.load(tmp_path / ("meta%d.npy" % r))
        assert np.array_equal(V, o["V"])                       # same sweep order => same bits
        assert meta[0] == o["n_steps"] and meta[1] == o["mean_inner"] and meta[2] == 0
