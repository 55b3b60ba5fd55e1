"""Generates tests/golden/golden_backdoor_v1.npz from the UNMODIFIED reference (run in the build
container, where /root/reference is mounted):

    python tests/golden/make_golden_backdoor.py

backdoor.BackdoorAttack._attack_grads (backdoor.py:52-65) is called on the real class; only
`train_malicious_network` (backdoor.py:108, model training on MNIST/CIFAR, not available offline and not
on the aggregation path) is replaced by a fixed affine map of the parameters so that the vector
arithmetic and the clip are exercised deterministically.  `Attack.attack` (malicious.py:10-27) runs
unmodified on top of it for the *_attack cases.
"""
import os
import sys

import numpy as np

sys.path.insert(0, "/root/reference")
import backdoor   # noqa: E402
import malicious  # noqa: E402


def fake_training(p):
    return (p * np.float32(0.9) + np.float32(0.01)).astype(np.float32)


class Harness(backdoor.BackdoorAttack):
    def __init__(self, num_std):
        malicious.Attack.__init__(self, num_std)   # skip the dataset download of backdoor.py:14-44

    def train_malicious_network(self, initial_params_flat):
        return fake_training(initial_params_flat)


class User:
    def __init__(self, g, w, lr):
        self.grads, self.original_params, self.learning_rate = g, w, lr


def main():
    rng = np.random.default_rng(20260922)
    out = {}
    for name, d, z, lr in [("d8", 8, 1.5, 0.1), ("d1000", 1000, 1.0, 0.1), ("d4099_tight", 4099, 0.05, 0.05),
                           ("d257_wide", 257, 30.0, 0.2)]:
        mu = rng.standard_normal(d).astype(np.float32)
        sd = np.abs(rng.standard_normal(d)).astype(np.float32)
        if name == "d1000":
            sd[::7] = 0.0                                  # degenerate band: clip to mu exactly
        w = rng.standard_normal(d).astype(np.float32)
        out[f"hook_{name}/mean"], out[f"hook_{name}/stdev"], out[f"hook_{name}/params"] = mu, sd, w
        out[f"hook_{name}/z"], out[f"hook_{name}/lr"] = np.float64(z), np.float64(lr)
        out[f"hook_{name}/want"] = Harness(z)._attack_grads(mu.copy(), sd.copy(), w.copy(), lr)
    for name, f, d, z, lr in [("f5_d300", 5, 300, 1.5, 0.1), ("f24_d2051", 24, 2051, 0.5, 0.1)]:
        rows = (rng.standard_normal((f, d)) * np.exp(0.25 * rng.standard_normal((f, 1)))).astype(np.float32)
        w = rng.standard_normal(d).astype(np.float32)
        users = [User(rows[i].copy(), w, lr) for i in range(f)]
        att = Harness(z)
        att.attack(users)
        out[f"attack_{name}/rows"], out[f"attack_{name}/params"] = rows, w
        out[f"attack_{name}/z"], out[f"attack_{name}/lr"] = np.float64(z), np.float64(lr)
        out[f"attack_{name}/mean"], out[f"attack_{name}/stdev"] = att.grads_mean, att.grads_stdev
        out[f"attack_{name}/grads0"] = users[0].grads
        out[f"attack_{name}/aliased"] = np.bool_(all(u.grads is users[0].grads for u in users))
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden_backdoor_v1.npz"), **out)
    print("wrote", len(out), "arrays")


if __name__ == "__main__":
    main()
