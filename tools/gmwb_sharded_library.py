"""GMWB with the control search sharded by the A-axis over the GPUs of one node, driven by the LIBRARY
(qpde_comm_init + qpde_gmwb2d_solve_sharded: NCCL inside libqpde_b200, no torch on the data path).

usage: python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
           --master-port P tools/gmwb_sharded_library.py [--refine K] [--a-points 51] [--timesteps0 32] [--out f.npz]
torch.distributed (gloo) only ships the 128-byte communicator id and reduces the wall time.  Rank 0
prints one JSON line; --check also solves on one GPU (rank 0) and compares bit for bit.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from quantpde_b200 import capi  # noqa: E402

W = np.array([0., 5., 10., 15., 20., 25., 30., 35., 40., 45., 50., 55., 60., 65., 70., 72.5, 75., 77.5, 80., 82., 84.,
              86., 88., 90., 91., 92., 93., 94., 95., 96., 97., 98., 99., 100., 101., 102., 103., 104., 105., 106.,
              107., 108., 109., 110., 112., 114., 116., 118., 120., 123., 126., 130., 135., 140., 145., 150., 160.,
              175., 200., 225., 250., 300., 500., 750., 1000.])


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--refine", type=int, default=2)
    ap.add_argument("--a-points", type=int, default=51)
    ap.add_argument("--timesteps0", type=int, default=32)
    ap.add_argument("--runs", type=int, default=1)
    ap.add_argument("--check", action="store_true")
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    if world > 1:
        dist.init_process_group("gloo")
    dxa = 100. / (a.a_points - 1)
    A = np.array([dxa * i for i in range(a.a_points)])
    Z = np.array([0., .5, 1.])
    steps = a.timesteps0 * 2 ** a.refine
    with capi.Handle(local) as h:
        if world > 1:
            ident = torch.zeros(capi.COMM_ID_BYTES, dtype=torch.uint8)
            if rank == 0:
                ident = torch.frombuffer(bytearray(capi.comm_unique_id()), dtype=torch.uint8).clone()
            dist.broadcast(ident, 0)
            h.comm_init(ident.numpy().tobytes(), rank, world)
        capi.gmwb_solve(h, W, A, [0., 10.], Z, min(a.refine, 1), 4, sharded=world > 1)      # warm-up: module load, NCCL channels
        best = 1e30
        for _ in range(a.runs):
            if world > 1:
                dist.barrier()
            l0 = h.launches
            t0 = time.perf_counter()
            res = capi.gmwb_solve(h, W, A, [0., 10.], Z, a.refine, steps, sharded=world > 1)
            secs = torch.tensor([time.perf_counter() - t0], dtype=torch.float64)
            if world > 1:
                dist.all_reduce(secs, op=dist.ReduceOp.MAX)
            best = min(best, float(secs.item()))
            launches = h.launches - l0
        if rank == 0:
            iw, ia = np.searchsorted(res["W"], 100.), np.searchsorted(res["A"], 100.)
            line = {"workload": "GMWB HJBQVI %d x %d nodes, %d steps" % (len(res["W"]), len(res["A"]), res["n_steps"]),
                    "driver": "library (qpde_gmwb2d_solve_sharded)" if world > 1 else "library (qpde_gmwb2d_solve)",
                    "n_gpus": world, "seconds": best, "mean_inner": res["mean_inner"], "status": res["status"],
                    "value_100_100": float(res["V"][ia, iw]), "launches_rank0": launches}
            if a.check:
                with capi.Handle(local) as h1:
                    whole = capi.gmwb_solve(h1, W, A, [0., 10.], Z, a.refine, steps)
                line["bit_identical_to_one_gpu"] = bool(np.array_equal(whole["V"], res["V"])
                                                        and whole["mean_inner"] == res["mean_inner"])
            print(json.dumps(line), flush=True)
            if a.out:
                np.savez(a.out, V=res["V"], mean_inner=res["mean_inner"])
        if world > 1:
            dist.barrier()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
