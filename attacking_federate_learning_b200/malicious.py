"""Drop-in for the reference's `malicious.py` (Attack / DriftAttack, the "A Little Is Enough"
perturbation), backed by the fused mu/sigma/perturb kernel `afl_alie`.

Reference surface mirrored: Attack.__init__/attack (malicious.py:4-27), DriftAttack._attack_grads
(malicious.py:30-36).  `users` are duck-typed objects with `.grads`, `.original_params`,
`.learning_rate` exactly as the reference expects; `.grads` may be NumPy float32 vectors or
torch.cuda vectors.  As in the reference, after `attack()` every malicious user holds THE SAME array
object, which is also `self.grads_mean` (mutated in place); `self.grads_stdev` keeps sigma.
"""
from __future__ import annotations

import numpy as np


class Attack(object):
    def __init__(self, num_std):
        self.num_std = num_std
        self.grads_mean = None
        self.grads_stdev = None

    def attack(self, users):
        if len(users) == 0:
            return
        import torch
        from . import _device as dev

        first = users[0].grads
        on_gpu = isinstance(first, torch.Tensor) and first.is_cuda
        if on_gpu:
            rows = torch.stack([u.grads for u in users])
        else:
            rows = torch.from_numpy(np.ascontiguousarray(np.stack([np.asarray(u.grads, np.float32) for u in users]))).cuda()
        # num_std == 0: statistics only, gradients untouched (malicious.py:21-22)
        crafted, mu, sigma = self._device_attack(rows, dev)
        if on_gpu:
            self.grads_mean, self.grads_stdev = mu, sigma
        else:
            self.grads_mean, self.grads_stdev = mu.cpu().numpy(), sigma.cpu().numpy()
            crafted = self.grads_mean
        if self.num_std == 0:
            return
        mal_grads = crafted
        for usr in users:
            usr.grads = mal_grads

    def _device_attack(self, rows, dev):
        raise NotImplementedError

    def attack_rows(self, users_grads, corrupted_count):
        """GPU-resident form: the malicious users are rows 0..f-1 of the stacked matrix (main.py:28);
        computes the crafted vector and writes it into those rows in place (what server.py:82-83 does
        next with the aliased `usr.grads`).  Returns the crafted vector."""
        from . import _device as dev
        if corrupted_count <= 0:
            return None
        mal = users_grads[:corrupted_count]
        if self.num_std == 0:
            _, mu, sigma = dev.alie(mal, 0.0, None, alias_mean=False)
            self.grads_mean, self.grads_stdev = mu, sigma
            return None
        import torch
        bcast = users_grads if users_grads.dtype == torch.float32 else None
        crafted, mu, sigma = dev.alie(mal, self.num_std, bcast, alias_mean=True)
        if bcast is None:
            users_grads[:corrupted_count] = crafted.to(users_grads.dtype)
        self.grads_mean, self.grads_stdev = mu, sigma
        return crafted


class DriftAttack(Attack):
    def __init__(self, num_std):
        super(DriftAttack, self).__init__(num_std)

    def _device_attack(self, rows, dev):
        if self.num_std == 0:
            crafted, mu, sigma = dev.alie(rows, 0.0, None, alias_mean=False)
            return None, mu, sigma
        return dev.alie(rows, self.num_std, None, alias_mean=True)

    def _attack_grads(self, grads_mean, grads_stdev, original_params, learning_rate):
        # malicious.py:34-36 — kept for callers that use the template-method hook directly
        grads_mean[:] -= self.num_std * grads_stdev[:]
        return grads_mean
