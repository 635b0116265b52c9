# -*- coding: utf-8 -*-
"""Multi-GPU plumbing: chains are independent, so repetitions / datasets
are partitioned over GPUs with NO data-path collective; one all-gather (NCCL over
NVLink on GPUs, gloo on CPU for tests) collects the final contribution sets, scalings,
backgrounds, chi-squared and iteration counts for the histogram statistics.

The gather is off the critical path of a fit: `Gatherer.start()` packs a rank's result block
into a pre-allocated pinned buffer, copies it to the device and issues
`all_gather_into_tensor` asynchronously; `wait()` is called once, later (bench.py overlaps it
with the next fit)."""
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


class Gatherer(object):
    """Asynchronous all-gather of every rank's result block.  Ranks may hold different numbers of
    repetitions (block partition) or none at all (a fit that returned no result): a small header
    all-gather (R, N, P per rank) sizes the padded payload, and a rank without a result makes EVERY
    rank raise -- no rank is left waiting in a collective."""

    def __init__(self, device=0):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.active = dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
        self.world = dist.get_world_size() if self.active else 1
        self.cuda = self.active and dist.get_backend() == "nccl"
        self.dev = torch.device("cuda", device) if self.cuda else torch.device("cpu")
        self._cap = 0
        self._work = None
        self._pending = None
        self.result = None

    def _buffers(self, n):
        torch = self.torch
        if n > self._cap:
            pin = self.cuda
            self._h_in = torch.empty(n, dtype=torch.float64, pin_memory=pin)
            self._h_out = torch.empty(n * self.world, dtype=torch.float64, pin_memory=pin)
            self._d_in = torch.empty(n, dtype=torch.float64, device=self.dev)
            self._d_out = torch.empty(n * self.world, dtype=torch.float64, device=self.dev)
            self._cap = n

    def start(self, res):
        torch, dist = self.torch, self.dist
        if not self.active:
            self.result = res
            return
        if self._work is not None:
            self.wait()
        if res is not None:
            rows, (N, P) = _pack(res)
            hdr = [rows.shape[0], N, P]
        else:
            rows, hdr = None, [-1, 0, 0]
        h = torch.tensor(hdr, dtype=torch.int64, device=self.dev)
        hs = torch.empty(3 * self.world, dtype=torch.int64, device=self.dev)
        self._all_gather(hs, h, False)
        hs = hs.cpu().numpy().reshape(self.world, 3)
        if (hs[:, 0] < 0).any():
            raise RuntimeError("gather_results: rank(s) %s hold no result (fit not converged and showIncomplete "
                               "off); nothing gathered" % np.nonzero(hs[:, 0] < 0)[0].tolist())
        if len(set(map(tuple, hs[:, 1:].tolist()))) != 1:
            raise RuntimeError("gather_results: ranks disagree on (numContribs, nActive): %s" % hs[:, 1:].tolist())
        N, P = int(hs[0, 1]), int(hs[0, 2])
        rmax, width = int(hs[:, 0].max()), N * P + 5
        n = rmax * width
        self._buffers(n)
        src = self._h_in[:n].view(rmax, width)
        src.zero_()
        src[:rows.shape[0]] = torch.from_numpy(rows)
        self._d_in[:n].copy_(self._h_in[:n], non_blocking=True)
        self._work = self._all_gather(self._d_out[:n * self.world], self._d_in[:n], True)
        self._pending = (hs[:, 0].copy(), N, P, rmax, width)

    def _all_gather(self, out, inp, async_op):
        try:
            return self.dist.all_gather_into_tensor(out, inp, async_op=async_op)
        except (RuntimeError, NotImplementedError):          # a backend without the flat form
            parts = list(out.view(self.world, -1).unbind(0))
            return self.dist.all_gather(parts, inp, async_op=async_op)

    def wait(self):
        if not self.active or self._work is None:
            return self.result
        self._work.wait()
        counts, N, P, rmax, width = self._pending
        n = rmax * width * self.world
        self._h_out[:n].copy_(self._d_out[:n], non_blocking=True)
        if self.cuda:
            self.torch.cuda.current_stream().synchronize()
        blocks = self._h_out[:n].view(self.world, rmax, width).numpy()
        rows = np.concatenate([blocks[r, :int(counts[r])] for r in range(self.world)], 0)
        self.result = _unpack(rows.copy(), N, P)
        self._work, self._pending = None, None
        return self.result


def gather_results(res, device=0):
    """Blocking form: every rank's result block, concatenated in rank order, on every rank."""
    g = Gatherer(device)
    g.start(res)
    return g.wait()
