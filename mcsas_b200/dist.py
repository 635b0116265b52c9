# -*- coding: utf-8 -*-
"""Multi-GPU plumbing: chains are independent, so repetitions / datasets
are partitioned over GPUs with NO data-path collective; one all_gather (NCCL over
NVLink on GPUs, gloo on CPU for tests) collects the final contribution sets, scalings,
backgrounds, chi-squared and iteration counts for the histogram statistics."""
import numpy as np


def partition(n_items, world, rank):
    """Contiguous block partition of n_items work items (dataset x repetition chains):
    returns (first, count) of `rank`; sizes differ by at most one."""
    base, rem = divmod(n_items, world)
    first = rank * base + min(rank, rem)
    return first, base + (1 if rank < rem else 0)


def _pack(res):
    """contribs [N,P,R] + per-repetition scalars -> one float64 row block [R, N*P + 5]."""
    c = np.asarray(res["contribs"])
    N, P, R = c.shape
    rows = np.empty((R, N * P + 5))
    rows[:, :N * P] = c.reshape(N * P, R).T
    rows[:, N * P + 0] = res["scalingAll"]
    rows[:, N * P + 1] = res["backgroundAll"]
    rows[:, N * P + 2] = res["conval"]
    rows[:, N * P + 3] = res["numIterAll"]
    rows[:, N * P + 4] = res["numMoves"]
    return rows, (N, P)


def _unpack(rows, N, P):
    R = rows.shape[0]
    return dict(contribs=rows[:, :N * P].T.reshape(N, P, R),
                scalingAll=rows[:, N * P], backgroundAll=rows[:, N * P + 1],
                conval=rows[:, N * P + 2], numIterAll=rows[:, N * P + 3].astype(np.int64),
                numMoves=rows[:, N * P + 4].astype(np.int64))


def gather_results(res, device=0):
    """all_gather of every rank's result block (equal repetition counts per rank)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return res
    rows, (N, P) = _pack(res)
    backend = dist.get_backend()
    dev = torch.device("cuda", device) if backend == "nccl" else torch.device("cpu")
    mine = torch.from_numpy(rows).to(dev)
    out = [torch.empty_like(mine) for _ in range(dist.get_world_size())]
    dist.all_gather(out, mine)
    allrows = torch.cat(out, 0).cpu().numpy()
    return _unpack(allrows, N, P)
