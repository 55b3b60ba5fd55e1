"""Gradient ingest: the step BEFORE the aggregation path.

Reference behaviour mirrored (file:line in /root/reference):
    flatten_params(params)               user.py:17-18   np.concatenate([p.data.cpu().numpy().flatten() for p in params])
    row_into_parameters(row, parameters) user.py:21-28   inverse: consecutive slices of the flat row, in order
    usr.grads = flatten_params(grads)    user.py:92      one flat fp32 vector per client and round
    users_grads[idx, :] = usr.grads      server.py:81-83 N row copies into the N x D matrix

The flatten contract is therefore: parameters in `net.parameters()` order, each flattened in C order, concatenated;
column k of the matrix is the same scalar parameter for every client.  `ParamLayout` records that order once and
writes a client's gradients STRAIGHT into its row of the (column-sharded) device matrix - no concatenated temporary,
no host round trip when the gradients already live on the device.  `ShardIngest` copies host rows (what the
reference's CPU clients produce) through pinned staging buffers on a dedicated copy stream, so the H2D transfer of
round t+1 overlaps the aggregation kernels of round t; only the columns of this rank's shard are transferred.
"""
from __future__ import annotations

from typing import Iterable, List, Sequence, Tuple

import numpy as np
import torch


class ParamLayout:
    """Offsets of every parameter tensor inside the flat D-vector (user.py:17-28 order)."""

    def __init__(self, params: Iterable[torch.Tensor]):
        self.shapes: List[Tuple[int, ...]] = []
        self.offsets: List[int] = []
        off = 0
        for p in params:
            self.shapes.append(tuple(p.shape))
            self.offsets.append(off)
            off += int(np.prod(p.shape)) if len(p.shape) else 1
        self.dim = off

    def sizes(self):
        return [int(np.prod(s)) if len(s) else 1 for s in self.shapes]

    # -- user.py:17-18 (used for weights and, at user.py:92, for gradients)
    def flatten(self, tensors: Sequence[torch.Tensor], out: torch.Tensor | None = None, c0: int = 0, c1: int | None = None):
        """Concatenation of the flattened tensors, restricted to columns [c0, c1), written into `out` (any device)."""
        c1 = self.dim if c1 is None else c1
        if out is None:
            out = torch.empty(c1 - c0, dtype=torch.float32, device=tensors[0].device if len(tensors) else "cpu")
        assert out.numel() == c1 - c0
        for t, off, size in zip(tensors, self.offsets, self.sizes()):
            lo, hi = max(off, c0), min(off + size, c1)
            if lo < hi:
                out[lo - c0:hi - c0].copy_(t.detach().reshape(-1)[lo - off:hi - off], non_blocking=True)
        return out

    # -- user.py:21-28
    def row_into_parameters(self, row, parameters: Sequence[torch.Tensor]):
        row_t = row if isinstance(row, torch.Tensor) else torch.from_numpy(np.asarray(row))
        for p, off, size, shape in zip(parameters, self.offsets, self.sizes(), self.shapes):
            p.data[...] = row_t[off:off + size].reshape(shape).to(p.device, p.dtype)


class ShardIngest:
    """Fills rows of a device-resident [N, d_local] shard (columns [c0, c1) of the N x D matrix) from client gradients.

    Host rows go through two pinned staging buffers and a copy stream; `wait()` makes the current stream wait for
    the last copy, so everything enqueued before `collect` (the previous round's kernels) overlaps the transfer."""

    def __init__(self, matrix: torch.Tensor, c0: int = 0, c1: int | None = None, chunk_rows: int = 16):
        self.matrix = matrix
        self.n, self.d_local = matrix.shape
        self.c0 = c0
        self.c1 = c0 + self.d_local if c1 is None else c1
        assert self.c1 - self.c0 == self.d_local
        self.cuda = matrix.is_cuda
        self.chunk_rows = max(1, min(chunk_rows, self.n))
        if self.cuda:
            self.stream = torch.cuda.Stream(device=matrix.device)
            self.stage = [torch.empty((self.chunk_rows, self.d_local), dtype=torch.float32).pin_memory() for _ in range(2)]
            self.stage_free = [torch.cuda.Event() for _ in range(2)]
            self.done = torch.cuda.Event()
        self.bytes_h2d = 0

    def collect(self, users):
        """server.py:81-83 for this shard: row idx <- usr.grads[c0:c1] (NumPy / CPU tensors staged, CUDA tensors copied
        device to device), asynchronously on the copy stream."""
        if not self.cuda:
            for idx, usr in enumerate(users):
                g = usr.grads if isinstance(usr.grads, torch.Tensor) else torch.from_numpy(np.asarray(usr.grads))
                self.matrix[idx].copy_(g[self.c0:self.c1])
            return
        users = list(users)
        with torch.cuda.stream(self.stream):
            slot = 0
            i = 0
            while i < len(users):
                j = min(len(users), i + self.chunk_rows)
                host_rows = [k for k in range(i, j) if not (isinstance(users[k].grads, torch.Tensor) and users[k].grads.is_cuda)]
                if host_rows:
                    st = self.stage[slot]
                    self.stage_free[slot].synchronize()            # the copy that last used this buffer has finished
                    for k in host_rows:
                        g = users[k].grads
                        g = g.numpy() if isinstance(g, torch.Tensor) else np.asarray(g)
                        st[k - i].numpy()[:] = g[self.c0:self.c1]   # the only host pass: straight into pinned memory
                    lo, hi = host_rows[0], host_rows[-1] + 1
                    if len(host_rows) == hi - lo:                   # contiguous run of host rows: one 2-D copy
                        self.matrix[lo:hi].copy_(st[lo - i:hi - i], non_blocking=True)
                    else:
                        for k in host_rows:
                            self.matrix[k].copy_(st[k - i], non_blocking=True)
                    self.bytes_h2d += len(host_rows) * self.d_local * 4
                    self.stage_free[slot].record(self.stream)
                    slot ^= 1
                for k in range(i, j):
                    g = users[k].grads
                    if isinstance(g, torch.Tensor) and g.is_cuda:
                        self.matrix[k].copy_(g[self.c0:self.c1], non_blocking=True)
                i = j
            self.done.record(self.stream)

    def wait(self):
        if self.cuda:
            torch.cuda.current_stream(self.matrix.device).wait_event(self.done)
