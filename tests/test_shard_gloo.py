"""CPU, world_size 2, gloo: the N > 1 host-side logic — contract slices are a
disjoint ordered cover, each rank sweeps only its slice, and the gathered
per-contract values equal a single-rank run.  (The per-rank "sweep" here is the
oracle on tiny contracts: this is a test, the GPU path is exercised by
bench.py --gpus N and the gpu-marked tests.)"""
import os
import socket

import numpy as np
import pytest
import torch.multiprocessing as mp

from quantpde_b200 import shard


def test_contract_slices_cover_in_order():
    for total in (0, 1, 7, 64, 65536, 65537):
        for world in (1, 2, 3, 8):
            cuts = [shard.contract_slice(total, r, world) for r in range(world)]
            assert cuts[0][0] == 0 and cuts[-1][1] == total
            assert all(cuts[r][1] == cuts[r + 1][0] for r in range(world - 1))
            sizes = [e - b for b, e in cuts]
            assert max(sizes) - min(sizes) <= 1
    w = np.array([1.0] * 10 + [9.0] * 10)
    cuts = [shard.contract_slice(20, r, 2, w) for r in range(2)]
    assert cuts[0][1] == cuts[1][0] and cuts[0][0] == 0 and cuts[1][1] == 20
    loads = [w[b:e].sum() for b, e in cuts]
    assert abs(loads[0] - loads[1]) <= 2 * 9.0      # balanced by weight (within one contract), not by count
    assert cuts[0][1] > 10
    with pytest.raises(ValueError):
        shard.contract_slice(4, 2, 2)


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, total, out_dir):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import pyoracle as po
    rng = np.random.default_rng(1234)
    K = rng.uniform(50, 150, total); vol = rng.uniform(.1, .6, total)
    b, e = shard.contract_slice(total, rank, world)
    local = [po.interp(*(lambda o: (o["S"], o["V"]))(po.american_put(K[c], .04, vol[c], 1., 24, 1)), K[c])
             for c in range(b, e)]
    full = shard.gather_values(local, total, rank, world)
    dist.barrier()
    np.save(os.path.join(out_dir, "rank%d.npy" % rank), full)
    dist.destroy_process_group()


def test_two_rank_gather_equals_single_rank(tmp_path):
    total, world = 7, 2            # ragged: 4 + 3 contracts
    mp.spawn(_worker, args=(world, _free_port(), total, str(tmp_path)), nprocs=world, join=True)
    got = [np.load(tmp_path / ("rank%d.npy" % r)) for r in range(world)]
    assert np.array_equal(got[0], got[1])
    from oracle import pyoracle as po
    rng = np.random.default_rng(1234)
    K = rng.uniform(50, 150, total); vol = rng.uniform(.1, .6, total)
    want = np.array([po.interp(*(lambda o: (o["S"], o["V"]))(po.american_put(K[c], .04, vol[c], 1., 24, 1)), K[c])
                     for c in range(total)])
    assert np.array_equal(got[0], want)
