"""Flatten / ingest contract: parameter order and shard slicing, on CPU tensors."""
import functools

import numpy as np
import torch

from attacking_federate_learning_b200.ingest import ParamLayout, ShardIngest


def ref_flatten(params):                                   # user.py:17-18, restated
    return np.concatenate([p.data.cpu().numpy().flatten() for p in params])


def ref_row_into_parameters(row, parameters):              # user.py:21-28, restated
    offset = 0
    for param in parameters:
        size = functools.reduce(lambda x, y: x * y, param.shape)
        param.data[:] = torch.from_numpy(row[offset:offset + size].reshape(param.shape))
        offset += size


def make_net():
    torch.manual_seed(0)
    return torch.nn.Sequential(torch.nn.Linear(7, 5), torch.nn.ReLU(), torch.nn.Linear(5, 3), torch.nn.Conv2d(2, 3, 3))


def test_layout_matches_reference_order():
    net = make_net()
    params = list(net.parameters())
    lay = ParamLayout(params)
    flat = ref_flatten(params)
    assert lay.dim == flat.size
    assert np.array_equal(lay.flatten(params).numpy(), flat)
    for c0, c1 in ((0, lay.dim), (3, 41), (35, 36), (40, lay.dim)):      # shard slices cut through parameter tensors
        assert np.array_equal(lay.flatten(params, c0=c0, c1=c1).numpy(), flat[c0:c1])


def test_row_into_parameters_round_trip():
    a, b = make_net(), make_net()
    row = np.random.default_rng(0).standard_normal(ParamLayout(a.parameters()).dim).astype(np.float32)
    ParamLayout(a.parameters()).row_into_parameters(row, list(a.parameters()))
    ref_row_into_parameters(row, list(b.parameters()))
    for p, q in zip(a.parameters(), b.parameters()):
        assert torch.equal(p, q)
    assert np.array_equal(ref_flatten(a.parameters()), row)


def test_shard_ingest_cpu_rows():
    class U:
        def __init__(self, g): self.grads = g
    rng = np.random.default_rng(1)
    n, d = 7, 101
    G = rng.standard_normal((n, d)).astype(np.float32)
    for c0, c1 in ((0, d), (32, 64), (96, d)):
        shard = torch.empty((n, c1 - c0))
        ing = ShardIngest(shard, c0, c1)
        ing.collect([U(G[i]) if i % 2 else U(torch.from_numpy(G[i])) for i in range(n)])
        ing.wait()
        assert np.array_equal(shard.numpy(), G[:, c0:c1])
