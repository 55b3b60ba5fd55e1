"""CPU tests of the host-side mirror: reference-compatible names/signatures, the DistanceTable
mapping protocol, assertion behaviour, shard arithmetic."""
import inspect

import numpy as np
import pytest

from attacking_federate_learning_b200 import defences as D, malicious as M
from attacking_federate_learning_b200.sharded import shard_bounds


def test_surface_matches_reference_names():
    assert D.DefenseTypes.NoDefense == 'NoDefense' and D.DefenseTypes.Krum == 'Krum'
    assert D.DefenseTypes.TrimmedMean == 'TrimmedMean' and D.DefenseTypes.Bulyan == 'Bulyan'
    assert set(D.defend) == {'NoDefense', 'Krum', 'TrimmedMean', 'Bulyan'}
    assert D.defend['Krum'] is D.krum and D.defend['Bulyan'] is D.bulyan
    assert list(inspect.signature(D.krum).parameters) == ['users_grads', 'users_count', 'corrupted_count', 'distances', 'return_index', 'debug']
    assert list(inspect.signature(D.trimmed_mean).parameters) == ['users_grads', 'users_count', 'corrupted_count']
    assert list(inspect.signature(D.no_defense).parameters) == ['users_grads', 'users_count', 'corrupted_count']
    assert list(inspect.signature(D.bulyan).parameters)[:3] == ['users_grads', 'users_count', 'corrupted_count']
    assert list(inspect.signature(M.DriftAttack._attack_grads).parameters) == ['self', 'grads_mean', 'grads_stdev', 'original_params', 'learning_rate']
    a = M.DriftAttack(1.5)
    assert a.num_std == 1.5 and a.grads_mean is None and a.grads_stdev is None
    assert a.attack([]) is None


def test_reference_asserts_fire_before_any_gpu_work():
    G = np.zeros((10, 4), np.float32)
    with pytest.raises(AssertionError):
        D.krum(G, 10, 5)                      # users_count >= 2f+1     (defences.py:24-25)
    with pytest.raises(AssertionError):
        D.bulyan(G, 10, 2)                    # users_count >= 4f+3     (defences.py:56)


def test_distance_table_mapping_protocol():
    dense = np.arange(25, dtype=np.float32).reshape(5, 5)
    dense = dense + dense.T
    t = D.DistanceTable(dense)
    assert t.keys() == [1, 0, 2, 3, 4] and len(t) == 5
    assert list(t[2].keys()) == [0, 1, 3, 4]
    assert list(t[2].values()) == [dense[2, 0], dense[2, 1], dense[2, 3], dense[2, 4]]
    t.pop(0)
    assert t.keys() == [1, 2, 3, 4] and 0 not in t[2]
    with pytest.raises(KeyError):
        t.pop(0)


def test_attack_hooks_have_no_cpu_fallback():
    """The hooks compute on the device even for NumPy arguments: without a GPU they must fail loudly,
    not fall back to host arithmetic (the arithmetic itself is checked by the -m gpu tests)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present; covered by tests/test_gpu_parity.py::test_alie_band_semantics")
    mu = np.array([1.0, 2.0], np.float32); sd = np.array([0.5, 0.25], np.float32)
    with pytest.raises(Exception):
        M.DriftAttack(1.5)._attack_grads(mu, sd, None, None)
    with pytest.raises(Exception):
        M.BackdoorAttack(1.5, lambda p: p)._attack_grads(mu, sd, mu.copy(), 0.1)
    with pytest.raises(Exception):
        M.backdoor_clip(mu, mu, sd, 1.0)
    assert np.array_equal(mu, np.array([1.0, 2.0], np.float32))          # untouched


def test_attack_success_metrics():
    from attacking_federate_learning_b200 import metrics
    assert metrics.krum_attack_success(0, 24) and metrics.krum_attack_success(23, 24)
    assert not metrics.krum_attack_success(24, 24) and not metrics.krum_attack_success(-1, 24)
    assert metrics.bulyan_attack_success([0, 1, 50, 51], 2) == 0.5
    assert metrics.bulyan_attack_success([], 2) == 0.0


@pytest.mark.parametrize("dim,world", [(25_000_000, 8), (11_200_000, 4), (79_510, 2), (100, 8), (31, 2)])
def test_shard_bounds_partition_columns(dim, world):
    edges = [shard_bounds(dim, world, r) for r in range(world)]
    assert edges[0][0] == 0 and edges[-1][1] == dim
    for (a0, a1), (b0, b1) in zip(edges, edges[1:]):
        assert a1 == b0 and a0 <= a1
    assert all(lo % 32 == 0 for lo, _ in edges)
