"""GMWB sharded by the A-axis: one rank (one GPU) per contiguous block of A-lines
(BASELINE.json configs[4]).

Every rank keeps the full-grid iterate (33.6 MB at 4097 x 1025) and owns the rows
of its lines.  Per inner iteration:
  * control searches + row assembly of the owned lines      (parallel over ranks;
    reads the whole previous iterate: the impulse targets (w - z, a - z) live on
    arbitrary lower A-lines)
  * the line solves in ascending A: rank k solves its lines after ranks < k have
    published theirs, then broadcasts its block — the "halo" for the upwinded a-1
    coupling and the all-gather for the impulse step in one exchange.  The system
    is block lower-triangular by A-line, so one round is a direct solve (a backend
    without that guarantee, `exact_pass = False`, is swept until nothing moves)
  * MAX-reduction of the stopping flag (relativeError > tolerance).
The collectives are torch.distributed's (NCCL on GPUs, gloo in the CPU test); the
per-rank work goes through the C ABI (qpde_gmwb2d_shard_*).  The `backend` object
supplies the per-rank work, so the same driver runs the CPU test with a stand-in.
"""
import numpy as np

from .shard import contract_slice


class CudaBackend:
    """The product path: libqpde_b200 through the C ABI, tensors on the rank's GPU."""

    exact_pass = True       # qpde_gmwb2d_shard_sweep is a direct solve of the owned lines

    def __init__(self, handle, problem, a_begin, a_end):
        import torch
        from . import capi
        self.torch, self.capi = torch, capi
        self.device = torch.device("cuda", torch.cuda.current_device())
        self.shard = capi.GmwbShard(handle, problem, a_begin, a_end)
        self.handle = handle
        self.stream = torch.cuda.ExternalStream(handle.stream, device=self.device)

    def empty(self, n):
        return self.torch.zeros(n, dtype=self.torch.float64, device=self.device)

    def flag(self):
        return self.torch.zeros(1, dtype=self.torch.int32, device=self.device)

    def exit_values(self, x): self.shard.exit_values(x)
    def policy(self, x, v0, h0): self.shard.policy(x, h0)
    def sweep(self, x, v0, h0, status): self.shard.sweep(x, v0, status)
    def relerr(self, x, xprev, notdone): self.shard.relerr(x, xprev, notdone)
    def close(self): self.shard.close()


def solve(problem, backend, rank, world, max_inner=200, max_sweeps=500, group=None):
    """Runs the whole backward sweep; every rank returns the full solution
    dict(V [n_a][n_w], n_steps, mean_inner, status).  `problem` needs .nw, .na, .struct
    (T, timesteps) — quantpde_b200.capi.GmwbProblem or an equivalent."""
    import torch
    import torch.distributed as dist
    from . import capi

    nw, na = problem.nw, problem.na
    T, timesteps = problem.struct.T, problem.struct.timesteps
    blocks = [contract_slice(na, r, world) for r in range(world)]          # contiguous A-lines per rank
    ctx = torch.cuda.stream(backend.stream) if hasattr(backend, "stream") else _Null()
    with ctx:
        x, xprev, v0 = backend.empty(nw * na), backend.empty(nw * na), backend.empty(nw * na)
        moved, notdone, failed = backend.flag(), backend.flag(), backend.flag()
        exact = getattr(backend, "exact_pass", False)

        def publish(r):
            lo, hi = blocks[r]
            if world > 1 and hi > lo:
                dist.broadcast(x[lo * nw:hi * nw], src=r, group=group)

        backend.exit_values(x)
        for r in range(world):
            publish(r)
        steps = capi.make_schedule(T, T / timesteps)
        time1, inner_sum, status = T, 0, 0
        for st in steps:
            h0 = time1 - st.t
            v0.copy_(x)
            its = 0
            while True:
                xprev.copy_(x)
                backend.policy(xprev, v0, h0)
                for sweep in range(max_sweeps):
                    moved.zero_()
                    for r in range(world):                 # ascending A across the ranks
                        if r == rank:
                            backend.sweep(x, v0, h0, failed if exact else moved)
                        publish(r)
                    if exact:
                        break
                    if world > 1:
                        dist.all_reduce(moved, op=dist.ReduceOp.MAX, group=group)
                    if int(moved.item()) == 0:
                        break
                else:
                    status = 2
                notdone.zero_()
                backend.relerr(x, xprev, notdone)
                if world > 1:
                    dist.all_reduce(notdone, op=dist.ReduceOp.MAX, group=group)
                its += 1
                if int(notdone.item()) == 0:
                    break
                if its >= max_inner:
                    status = status or 1
                    break
            inner_sum += its
            time1 = st.t
        if exact:
            if world > 1:
                dist.all_reduce(failed, op=dist.ReduceOp.MAX, group=group)
            status = status or int(failed.item())
        V = x.cpu().numpy().reshape(na, nw).copy()
    return dict(V=V, n_steps=len(steps), mean_inner=inner_sum / len(steps), status=status)


class _Null:
    def __enter__(self): return self
    def __exit__(self, *a): return False
