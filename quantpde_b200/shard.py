"""Contract sharding across ranks.

1-D contract batches are independent units: rank k sweeps a contiguous slice of
the contracts on its own GPU and there is NO collective on the data path.  The
only communication is an optional gather of per-contract scalars (prices,
iteration counts) at the end, through torch.distributed (NCCL on GPUs, gloo in
the CPU tests).
"""
import numpy as np


def contract_slice(total, rank, world, weights=None):
    """[begin, end) of the contracts rank `rank` of `world` owns.

    Without weights the slices differ in size by at most one contract.  With
    `weights` (expected cost per contract, e.g. the number of time steps times the
    expected inner iterations) the cut points balance the summed weight instead,
    still contiguous and in order."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError("rank %d of world %d" % (rank, world))
    if weights is None:
        base, extra = divmod(total, world)
        begin = rank * base + min(rank, extra)
        return begin, begin + base + (1 if rank < extra else 0)
    w = np.asarray(weights, dtype=np.float64)
    if len(w) != total or np.any(w < 0):
        raise ValueError("weights must be non-negative, one per contract")
    cum = np.concatenate([[0.0], np.cumsum(w)])
    targets = cum[-1] * np.arange(world + 1) / world
    cuts = np.searchsorted(cum, targets, side="left")
    cuts[0], cuts[-1] = 0, total
    cuts = np.maximum.accumulate(cuts)
    return int(cuts[rank]), int(cuts[rank + 1])


def gather_values(local, total, rank, world, weights=None, group=None):
    """All ranks receive the [total] vector assembled from every rank's slice
    (float64).  Uses all_gather on equally padded buffers, so it works with the
    nccl and gloo backends alike; world == 1 needs no process group."""
    local = np.ascontiguousarray(local, dtype=np.float64)
    b, e = contract_slice(total, rank, world, weights)
    if len(local) != e - b:
        raise ValueError("rank %d holds %d values for a slice of %d" % (rank, len(local), e - b))
    if world == 1:
        return local.copy()
    import torch
    import torch.distributed as dist
    sizes = [contract_slice(total, k, world, weights) for k in range(world)]
    longest = max(hi - lo for lo, hi in sizes)
    device = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")
    mine = torch.zeros(longest, dtype=torch.float64, device=device)
    mine[:len(local)] = torch.from_numpy(local).to(device)
    parts = [torch.zeros(longest, dtype=torch.float64, device=device) for _ in range(world)]
    dist.all_gather(parts, mine, group=group)
    out = np.empty(total)
    for k, (lo, hi) in enumerate(sizes):
        out[lo:hi] = parts[k][:hi - lo].cpu().numpy()
    return out
