"""GMWB sharded by the A-axis over the GPUs of one node.

usage: python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
           --master-port P tools/gmwb_sharded_run.py [--refine K] [--a-points 51] [--timesteps0 32] [--out f.npz]
       (or plain `python tools/gmwb_sharded_run.py` for one GPU)
Rank 0 prints one JSON line: grid, steps, mean inner iterations, seconds (device, max over ranks), value at (100, 100).
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from quantpde_b200 import capi, gmwb_sharded  # noqa: E402
from quantpde_b200.shard import contract_slice  # noqa: E402

W = np.array([0., 5., 10., 15., 20., 25., 30., 35., 40., 45., 50., 55., 60., 65., 70., 72.5, 75., 77.5, 80., 82., 84.,
              86., 88., 90., 91., 92., 93., 94., 95., 96., 97., 98., 99., 100., 101., 102., 103., 104., 105., 106.,
              107., 108., 109., 110., 112., 114., 116., 118., 120., 123., 126., 130., 135., 140., 145., 150., 160.,
              175., 200., 225., 250., 300., 500., 750., 1000.])


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--refine", type=int, default=2)
    ap.add_argument("--a-points", type=int, default=51)
    ap.add_argument("--timesteps0", type=int, default=32)
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dxa = 100. / (a.a_points - 1)
    A = np.array([dxa * i for i in range(a.a_points)])
    prob = capi.GmwbProblem(W, A, [0., 10.], [0., .5, 1.], a.refine, a.timesteps0 * 2 ** a.refine)
    with capi.Handle(local) as h:
        lo, hi = contract_slice(prob.na, rank, world)
        backend = gmwb_sharded.CudaBackend(h, prob, lo, hi)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(backend.stream)
        res = gmwb_sharded.solve(prob, backend, rank, world)
        e1.record(backend.stream)
        torch.cuda.synchronize()
        secs = torch.tensor([e0.elapsed_time(e1) / 1e3], device="cuda")
        if world > 1:
            dist.all_reduce(secs, op=dist.ReduceOp.MAX)
        launches = h.launches
        backend.close()
    if rank == 0:
        iw, ia = np.searchsorted(prob.W, 100.), np.searchsorted(prob.A, 100.)
        print(json.dumps({"workload": "GMWB HJBQVI %d x %d nodes, %d steps" % (prob.nw, prob.na, res["n_steps"]),
                          "n_gpus": world, "seconds": float(secs.item()), "mean_inner": res["mean_inner"],
                          "status": res["status"], "value_100_100": float(res["V"][ia, iw]),
                          "launches_rank0": launches}))
        if a.out:
            np.savez(a.out, V=res["V"], mean_inner=res["mean_inner"])
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
