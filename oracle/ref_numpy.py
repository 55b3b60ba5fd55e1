"""CPU oracle for the Byzantine-robust aggregation hot path.  TEST INFRASTRUCTURE ONLY.

This module is a NumPy restatement of the algorithms in the upstream reference
(`/root/reference/defences.py`, `/root/reference/malicious.py`).  It exists so that the CUDA
path can be checked on machines where the reference tree is not present (the GPU box).  Only
`tests/`, `__graft_entry__.smoke()` and `bench.py`'s CPU-baseline / `--impl reference` legs may
import it.  The product package (`attacking_federate_learning_b200`) never imports it and has no
CPU fallback.

Parity status: PINNED.  `tests/golden/make_golden.py` imports the unmodified reference in the
build container and stores its outputs; `tests/test_oracle_golden.py` checks every function below
bit-for-bit against those fixtures (and, when `/root/reference` is present, against the live
reference on fresh random inputs).

Data structures differ from the reference on purpose (dense tables + explicit visit order instead of
a dict-of-dicts), the arithmetic does not: each numbered step cites the reference line it follows.

Two flavours are provided for every rule:
  * `*_f32`  : same operations, same order, same precision as the reference (fp32 NumPy).  This is
               the parity oracle and the "port" CPU baseline (same per-pair / per-column Python
               loops as the reference, so it costs what the reference costs).
  * `*_f64`  : same semantics evaluated in float64 ("arbiter").  Used to decide near-ties where the
               fp32 reference itself is within rounding noise of a different answer.
"""
from __future__ import annotations

import numpy as np

__all__ = [
    "visit_order", "pairwise_distances_f32", "pairwise_distances_f64", "krum_scores",
    "krum_select", "krum", "bulyan_select", "bulyan", "trimmed_mean", "trimmed_mean_f64",
    "no_defense", "alie_stats", "alie_attack", "momentum_step",
]


# ----------------------------------------------------------------------------------------------
# Krum
# ----------------------------------------------------------------------------------------------
def visit_order(n: int) -> list[int]:
    """Order in which the reference scores users.

    defences.py:17-20 fills a defaultdict with `distances[i][j] = distances[j][i] = ...` for
    i in range(n), j in range(i).  The right-most target `distances[i]` is subscripted first, so the
    first pair (i=1, j=0) creates key 1 and then key 0; afterwards keys appear as i grows.
    Result: [1, 0, 2, 3, ...]; empty for n < 2 (no pair is ever visited).
    """
    if n < 2:
        return []
    return [1, 0] + list(range(2, n))


def pairwise_distances_f32(G: np.ndarray) -> np.ndarray:
    """Dense symmetric table of ||g_i - g_j||_2, zero diagonal.   (defences.py:16-21)

    Per pair, exactly as the reference: fp32 element-wise difference, then `np.linalg.norm`, which
    for a real 1-D vector is sqrt(dot(x, x)) — fp32 BLAS dot, fp32 sqrt.
    """
    G = np.asarray(G)
    n = G.shape[0]
    table = np.zeros((n, n), dtype=np.float32)
    for i in range(n):
        gi = G[i]
        for j in range(i):
            delta = gi - G[j]                                   # defences.py:20  (fp32 subtract)
            table[i, j] = table[j, i] = np.sqrt(np.dot(delta, delta))   # == np.linalg.norm(delta)
    return table


def pairwise_distances_f64(G: np.ndarray, block: int = 1 << 16) -> np.ndarray:
    """Arbiter: exact-in-float64 pairwise distances of the *fp32-rounded differences*.

    Uses the same fl32(g_i - g_j) operand as the reference (defences.py:20) but accumulates the
    squares in float64, so the only error left is one final rounding.
    """
    G = np.asarray(G)
    n, dim = G.shape
    acc = np.zeros((n, n), dtype=np.float64)
    for c0 in range(0, dim, block):
        blk = G[:, c0:c0 + block]
        for i in range(n):
            delta = (blk[i][None, :] - blk[:i]).astype(np.float64)   # fp32 subtract, then widen
            acc[i, :i] += np.einsum("jk,jk->j", delta, delta)
    acc = acc + acc.T
    return np.sqrt(acc)


def krum_scores(table: np.ndarray, alive: list[int], keep: int, dtype=np.float32) -> dict[int, object]:
    """Score of every alive user = sum of its `keep` smallest distances to the other alive users.

    defences.py:33-34: `errors = sorted(distances[user].values()); sum(errors[:keep])`.
    `sum` starts from int 0 and adds np.float32 scalars left to right -> sequential fp32 sum in
    ascending order.  A slice longer than the list just takes everything.
    """
    scores = {}
    for u in alive:
        vals = sorted(dtype(table[u, v]) for v in alive if v != u)
        total = 0
        for x in vals[:keep]:                 # plain Python slice semantics, as in the reference
            total = total + x
        scores[u] = total
    return scores


def krum_select(table: np.ndarray, alive: list[int], users_count: int, corrupted_count: int,
                dtype=np.float32, with_margin: bool = False):
    """Index chosen by the reference's scan (defences.py:26-37).

    keep = users_count - corrupted_count; strict `<` against a running minimum that starts at 1e20
    with index -1; users are visited in `alive` order, so exact ties go to the earliest visited.
    NaN scores never win (`nan < x` is False); if nobody wins the result is -1.
    """
    keep = users_count - corrupted_count
    # Python slice semantics for a negative stop (never reached by the reference's own callers,
    # which assert n >= 2f+1 / n >= 4f+3) are reproduced by krum_scores.
    scores = krum_scores(table, alive, keep, dtype)
    best, best_idx = 1e20, -1
    second = np.inf
    for u in alive:
        s = scores[u]
        if s < best:
            second = best
            best, best_idx = s, u
        elif s < second and s == s:
            second = s
    if with_margin:
        margin = float("inf") if not np.isfinite(second) or best == 0 else (float(second) - float(best)) / abs(float(best))
        return best_idx, margin
    return best_idx


def krum(G: np.ndarray, users_count: int, corrupted_count: int, table=None, return_index=False,
         dtype=np.float32):
    """defences.py:23-42.  Returns a *view* of the winning row (or its index)."""
    if not return_index:
        assert users_count >= 2 * corrupted_count + 1, (
            "users_count>=2*corrupted_count + 3", users_count, corrupted_count)
    if table is None:
        table = pairwise_distances_f32(G) if dtype == np.float32 else pairwise_distances_f64(G)
    idx = krum_select(table, visit_order(len(G)), users_count, corrupted_count, dtype)
    return idx if return_index else G[idx]


# ----------------------------------------------------------------------------------------------
# Bulyan
# ----------------------------------------------------------------------------------------------
def bulyan_select(table: np.ndarray, n: int, f: int, dtype=np.float32, with_margins: bool = False):
    """Selection sequence of Bulyan's first stage (defences.py:57-68).

    theta = n - 2f rounds; round r calls krum with users_count = n - r on the *same* distance table
    with the already-selected users popped (dict `pop` keeps the relative order of the survivors).
    """
    alive = visit_order(table.shape[0])
    chosen, margins = [], []
    while len(chosen) < n - 2 * f:
        idx, margin = krum_select(table, alive, n - len(chosen), f, dtype, with_margin=True)
        chosen.append(idx)
        margins.append(margin)
        if idx < 0:            # reference would index row -1 and then KeyError on pop(-1)
            raise KeyError(idx)
        alive.remove(idx)
    return (chosen, margins) if with_margins else chosen


def bulyan(G: np.ndarray, users_count: int, corrupted_count: int, table=None, dtype=np.float32):
    """defences.py:55-70."""
    assert users_count >= 4 * corrupted_count + 3
    if table is None:
        table = pairwise_distances_f32(G) if dtype == np.float32 else pairwise_distances_f64(G)
    chosen = bulyan_select(table, users_count, corrupted_count, dtype)
    picked = np.array([G[i] for i in chosen])              # rows in SELECTION order (defences.py:70)
    return trimmed_mean(picked, len(chosen), 2 * corrupted_count)


# ----------------------------------------------------------------------------------------------
# Trimmed mean around the median
# ----------------------------------------------------------------------------------------------
def trimmed_mean(G: np.ndarray, users_count, corrupted_count, col_block: int = 4096) -> np.ndarray:
    """defences.py:44-52, column-blocked but arithmetically identical per column.

    Per coordinate: med = np.median(col) (fp32; even N -> mean of the two middle order statistics);
    deviations fl32(col - med); keep the k = N - f - 1 deviations of smallest magnitude, ties in
    client order (Python's `sorted(..., key=abs)` is stable); result = np.mean(kept) + med.
    `np.mean` of a 1-D fp32 sequence is NumPy's pairwise fp32 sum divided by k; reducing a
    C-contiguous [cols, k] block along its last axis runs the same pairwise routine per row.
    """
    G = np.asarray(G)
    n, dim = G.shape
    k = int(n - corrupted_count) - 1                                         # defences.py:45
    out = np.empty((dim,), G.dtype)
    for c0 in range(0, dim, col_block):
        blk = G[:, c0:c0 + col_block]
        med = np.median(blk, axis=0)                                          # defences.py:49
        dev = blk - med                                                       # defences.py:50
        order = np.argsort(np.abs(dev), axis=0, kind="stable")
        kept = np.take_along_axis(dev, order[:k], axis=0)
        kept_rows = np.ascontiguousarray(kept.T)                              # [cols, k]
        out[c0:c0 + col_block] = kept_rows.mean(axis=1) + med                 # defences.py:51
    return out


def trimmed_mean_f64(G: np.ndarray, users_count, corrupted_count, col_block: int = 4096) -> np.ndarray:
    """Arbiter: same kept set as the fp32 rule (fp32 median, fp32 deviations, stable order), but the
    mean of the kept deviations is accumulated in float64."""
    G = np.asarray(G)
    n, dim = G.shape
    k = int(n - corrupted_count) - 1
    out = np.empty((dim,), np.float64)
    for c0 in range(0, dim, col_block):
        blk = G[:, c0:c0 + col_block]
        med = np.median(blk, axis=0)
        dev = blk - med
        order = np.argsort(np.abs(dev), axis=0, kind="stable")
        kept = np.take_along_axis(dev, order[:k], axis=0).astype(np.float64)
        out[c0:c0 + col_block] = kept.mean(axis=0) + med.astype(np.float64)
    return out


# ----------------------------------------------------------------------------------------------
# Plain mean, ALIE attack, server momentum step
# ----------------------------------------------------------------------------------------------
def no_defense(G: np.ndarray, users_count=None, corrupted_count=None) -> np.ndarray:
    """defences.py:13-14."""
    return np.mean(G, axis=0)


def alie_stats(rows) -> tuple[np.ndarray, np.ndarray]:
    """malicious.py:18-19: mean and population standard deviation over the malicious users' rows."""
    stack = np.asarray(rows)
    mu = np.mean(stack, axis=0)
    sigma = np.var(stack, axis=0) ** 0.5
    return mu, sigma


def alie_attack(rows, num_std: float):
    """malicious.py:10-27,34-36.  Returns (crafted, mu_after, sigma); `crafted is mu_after`
    (the reference mutates grads_mean in place and hands the same array to every malicious user).
    With num_std == 0 or no rows the gradients are left untouched (returns crafted=None)."""
    if len(rows) == 0:
        return None, None, None
    mu, sigma = alie_stats(rows)
    if num_std == 0:
        return None, mu, sigma
    mu[:] -= num_std * sigma[:]                                               # malicious.py:35
    return mu, mu, sigma


def backdoor_attack_grads(grads_mean, grads_stdev, original_params, learning_rate: float, num_std: float,
                          train_malicious_network):
    """backdoor.py:52-65 (BackdoorAttack._attack_grads); `train_malicious_network` stands for
    backdoor.py:108 (model training, outside the aggregation path)."""
    initial_params_flat = original_params - learning_rate * grads_mean                    # backdoor.py:54
    mal_net_params = train_malicious_network(initial_params_flat)                          # backdoor.py:56
    new_params = mal_net_params + learning_rate * grads_mean                              # backdoor.py:59
    new_grads = (initial_params_flat - new_params) / learning_rate                        # backdoor.py:60
    return np.clip(new_grads, grads_mean - num_std * grads_stdev,                         # backdoor.py:62-63
                   grads_mean + num_std * grads_stdev)


def momentum_step(weights, velocity, grads, momentum: float, learning_rate: float):
    """server.py:89-90:  v = momentum*v - lr*g ;  w += v   (returns new (w, v), fp32)."""
    velocity = momentum * velocity - learning_rate * grads
    weights = weights + velocity
    return weights, velocity
