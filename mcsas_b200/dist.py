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
    """Views into `rows` (a block this call owns): contribs as [N, P, R] without the transposing copy
    (3.9 MB per fit and rank at 8 x 100 repetitions -- it was half of the host time of a gather)."""
    R = rows.shape[0]
    return dict(contribs=rows[:, :N * P].reshape(R, N, P).transpose(1, 2, 0),
                scalingAll=rows[:, N * P], backgroundAll=rows[:, N * P + 1],
                conval=rows[:, N * P + 2], numIterAll=rows[:, N * P + 3].astype(np.int64),
                numMoves=rows[:, N * P + 4].astype(np.int64))


class Gatherer(object):
    """Asynchronous all-gather of every rank's result block.  Ranks may hold different numbers of
    repetitions (block partition) or none at all (a fit that returned no result).

    With `max_rows` (the largest repetition count any rank can hold; the usual case: it is known
    from the partition) a gather is ONE fixed-size asynchronous collective whose first row is a
    header (rows, numContribs, nActive), so start() never waits for another rank -- the ranks' fits
    differ in length by several ms and a blocking header exchange would make every fit wait for
    the slowest rank.  Without it, a small blocking header all-gather sizes the payload first.
    A rank without a result makes EVERY rank raise (in wait()) -- none is left in a collective."""

    def __init__(self, device=0, max_rows=None, n_contribs=None, n_active=None):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.active = dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
        self.world = dist.get_world_size() if self.active else 1
        self.cuda = self.active and dist.get_backend() == "nccl"
        self.dev = torch.device("cuda", device) if self.cuda else torch.device("cpu")
        if max_rows is not None and not (n_contribs and n_active):
            raise ValueError("Gatherer(max_rows=...) needs n_contribs and n_active: the fixed-size form must not "
                             "depend on whether this rank has a result")
        self.max_rows = max_rows
        self.shape = (int(n_contribs), int(n_active)) if n_contribs else None
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
        torch = self.torch
        if not self.active:
            self.result = res
            return
        if self._work is not None:
            self.wait()
        rows, shape = (None, None) if res is None else _pack(res)
        if self.max_rows is not None:
            if shape is not None and tuple(shape) != self.shape:
                raise ValueError("gather: result of shape %s, Gatherer built for %s" % (shape, self.shape))
            N, P = self.shape
            rmax, width = int(self.max_rows), N * P + 5
            if rows is not None and rows.shape[0] > rmax:
                raise ValueError("gather: %d repetitions exceed max_rows %d" % (rows.shape[0], rmax))
            counts = None                                   # read from the gathered headers
        else:
            hdr = [-1, 0, 0] if rows is None else [rows.shape[0], shape[0], shape[1]]
            h = torch.tensor(hdr, dtype=torch.int64, device=self.dev)
            hs = torch.empty(3 * self.world, dtype=torch.int64, device=self.dev)
            self._all_gather(hs, h, False)
            hs = hs.cpu().numpy().reshape(self.world, 3)
            self._check_headers(hs)
            N, P = int(hs[0, 1]), int(hs[0, 2])
            rmax, width = int(hs[:, 0].max()), N * P + 5
            counts = hs[:, 0].copy()
        n = (rmax + 1) * width                              # row 0: header
        self._buffers(n)
        src = self._h_in[:n].view(rmax + 1, width)
        src.zero_()
        src[0, 0] = -1. if rows is None else float(rows.shape[0])
        src[0, 1], src[0, 2] = float(N), float(P)
        if rows is not None:
            src[1:1 + rows.shape[0]] = torch.from_numpy(rows)
        self._d_in[:n].copy_(self._h_in[:n], non_blocking=True)
        self._work = self._all_gather(self._d_out[:n * self.world], self._d_in[:n], True)
        self._pending = (counts, N, P, rmax, width)

    @staticmethod
    def _check_headers(hs):
        if (hs[:, 0] < 0).any():
            raise RuntimeError("gather_results: rank(s) %s hold no result (fit not converged and showIncomplete "
                               "off); nothing gathered" % np.nonzero(hs[:, 0] < 0)[0].tolist())
        if len(set(map(tuple, hs[:, 1:].tolist()))) != 1:
            raise RuntimeError("gather_results: ranks disagree on (numContribs, nActive): %s" % hs[:, 1:].tolist())

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
        n = (rmax + 1) * width * self.world
        self._h_out[:n].copy_(self._d_out[:n], non_blocking=True)
        if self.cuda:
            self.torch.cuda.current_stream().synchronize()
        blocks = self._h_out[:n].view(self.world, rmax + 1, width).numpy()
        self._work, self._pending = None, None
        hs = np.rint(blocks[:, 0, :3]).astype(np.int64)
        self._check_headers(hs)                             # the same verdict on every rank
        # ONE pass over the gathered block: the ranks' rows without their header rows, out of the pinned
        # buffer (which the next gather overwrites) into an array the result owns
        rows = np.concatenate([blocks[r, 1:1 + int(hs[r, 0])] for r in range(self.world)], 0)
        self.result = _unpack(rows, N, P)
        return self.result


def gather_results(res, device=0):
    """Blocking form: every rank's result block, concatenated in rank order, on every rank."""
    g = Gatherer(device)
    g.start(res)
    return g.wait()
