"""world_size-2 gloo tests (CPU) of the D-sharded path: column partition, the single all-reduce of
partial squared-distance tables, replicated selection, sharded outputs.  The per-shard arithmetic is
a NumPy stand-in built from the oracle, so this exercises exactly the host logic of
attacking_federate_learning_b200/sharded.py."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import ref_numpy as orc


class NumpyKernels:
    """Same method names as attacking_federate_learning_b200._device, CPU tensors in and out."""

    def sqdist_partial(self, G, flags=0):
        return torch.from_numpy(orc.pairwise_distances_f64(G.numpy()) ** 2)

    def sqdist_to_dist(self, d2):
        return torch.sqrt(d2.clamp_min(0)).float()

    def krum_select(self, dist_t, n, f):
        t = dist_t.numpy()
        return torch.tensor([orc.krum_select(t, orc.visit_order(t.shape[0]), n, f)], dtype=torch.int32)

    def bulyan_select(self, dist_t, n, f):
        return torch.tensor(orc.bulyan_select(dist_t.numpy(), n, f), dtype=torch.int32)

    def trimmed_mean(self, G, f, row_index=None):
        A = G.numpy() if row_index is None else G.numpy()[row_index.numpy()]
        return torch.from_numpy(orc.trimmed_mean(A, len(A), f))

    def mean(self, G):
        return torch.from_numpy(orc.no_defense(G.numpy()))

    def alie(self, rows, z, bcast=None, alias_mean=True):
        crafted, mu, sigma = orc.alie_attack([r.copy() for r in rows.numpy()], z)
        if bcast is not None:
            bcast[:rows.shape[0]] = torch.from_numpy(crafted)
        return torch.from_numpy(crafted), torch.from_numpy(mu), torch.from_numpy(sigma)


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, n, d, f, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from attacking_federate_learning_b200.sharded import ShardedAggregator, shard_bounds
    rng = np.random.default_rng(99)
    G = (0.1 * rng.standard_normal(d) + np.exp(0.25 * rng.standard_normal((n, 1))) * rng.standard_normal((n, d))).astype(np.float32)
    c0, c1 = shard_bounds(d, world, rank)
    shard = torch.from_numpy(np.ascontiguousarray(G[:, c0:c1]))
    agg = ShardedAggregator(kernels=NumpyKernels())
    out = {}
    out["crafted"] = agg.alie(shard, f, 1.5).numpy()                      # also rewrites rows 0..f-1 of the shard
    out["krum_idx"] = agg.krum(shard, n, f, return_index=True)
    out["krum_row"] = agg.krum(shard, n, f).numpy()
    b, sel = agg.bulyan(shard, n, f, return_selection=True)
    out["bulyan"] = agg.gather_output(b, d).numpy(); out["sel"] = sel.tolist()
    out["tm"] = agg.gather_output(agg.trimmed_mean(shard, n, f), d).numpy()
    out["mean"] = agg.gather_output(agg.no_defense(shard), d).numpy()
    out["bounds"] = (c0, c1)
    q.put((rank, out))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n,d,f", [(11, 200, 2), (23, 333, 5)])
def test_sharded_matches_unsharded(n, d, f):
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, d, f, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0

    rng = np.random.default_rng(99)
    G = (0.1 * rng.standard_normal(d) + np.exp(0.25 * rng.standard_normal((n, 1))) * rng.standard_normal((n, d))).astype(np.float32)
    crafted, _, _ = orc.alie_attack([G[i].copy() for i in range(f)], 1.5)
    G[:f] = crafted
    table = orc.pairwise_distances_f64(G)
    idx = orc.krum_select(table, orc.visit_order(n), n, f, dtype=np.float64)
    sel = orc.bulyan_select(table, n, f, dtype=np.float64)
    for r in range(world):
        c0, c1 = got[r]["bounds"]
        np.testing.assert_array_equal(got[r]["crafted"], crafted[c0:c1])
        assert got[r]["krum_idx"] == idx                                   # replicated, identical on every rank
        np.testing.assert_array_equal(got[r]["krum_row"], G[idx, c0:c1])   # each rank returns its slice
        assert got[r]["sel"] == sel
        np.testing.assert_array_equal(got[r]["bulyan"], orc.trimmed_mean(G[sel], len(sel), 2 * f))
        np.testing.assert_array_equal(got[r]["tm"], orc.trimmed_mean(G, n, f))
        np.testing.assert_array_equal(got[r]["mean"], orc.no_defense(G))
