"""GPU-resident mirror of the aggregation half of the reference's `Server` (server.py:15-37, 81-90):
the stacked client-gradient matrix, `collect_gradients` and `defend` (robust aggregation + the
momentum update).  Training, evaluation and checkpointing stay with the caller."""
from __future__ import annotations

import torch

from . import defences
from . import _device as dev
from .ingest import ShardIngest


class AggregationServer:
    def __init__(self, users_count, dim, mal_prop, learning_rate, momentum=0.9, device="cuda", dtype=torch.float32,
                 initial_weights=None):
        self.mal_prop = mal_prop
        self.learning_rate = learning_rate
        self.momentum = momentum
        self.users_count = users_count
        ld = (dim + 31) // 32 * 32                       # padded pitch so TMA / 16-byte loads always apply
        self._storage = torch.empty((users_count, ld), dtype=dtype, device=device)
        self.users_grads = self._storage[:, :dim]        # server.py:35
        self.velocity = torch.zeros(dim, dtype=torch.float32, device=device)   # server.py:36
        self.current_weights = (torch.zeros(dim, dtype=torch.float32, device=device) if initial_weights is None
                                else torch.as_tensor(initial_weights, dtype=torch.float32).to(device).clone())
        self._ingest = ShardIngest(self.users_grads) if dtype == torch.float32 else None

    def collect_gradients(self, users):
        """server.py:81-83: users_grads[idx, :] = usr.grads.  Host rows are staged through pinned memory on a copy
        stream (ingest.ShardIngest), so the transfer overlaps whatever the compute stream is still running; device
        rows are copied device to device.  `defend` waits for the copies."""
        if self._ingest is not None:
            self._ingest.collect(users)
            return
        for idx, usr in enumerate(users):
            g = usr.grads
            if not isinstance(g, torch.Tensor):
                g = torch.from_numpy(g)
            self.users_grads[idx].copy_(g, non_blocking=True)

    def defend(self, defence_method, cur_epoch=0):
        """server.py:86-90 (keeps the reference's use of the base learning rate)."""
        n = self.users_count
        if self._ingest is not None:
            self._ingest.wait()
        current_grads = defences.defend[defence_method](self.users_grads, n, int(n * self.mal_prop))
        if current_grads.dtype != torch.float32:
            current_grads = current_grads.float()
        dev.momentum_step(self.current_weights, self.velocity, current_grads.contiguous(), self.momentum,
                          self.learning_rate)
        return current_grads
