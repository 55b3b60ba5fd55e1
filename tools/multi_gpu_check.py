"""Run under torchrun on >= 2 GPUs: sharded Krum / Bulyan / trimmed mean / ALIE against the
single-GPU result on the same (gathered) matrix.  Prints one JSON line on rank 0."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from attacking_federate_learning_b200 import defences as D, _device as dev
from attacking_federate_learning_b200.sharded import ShardedAggregator, shard_bounds

rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); lr = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
n, d, f = 103, 1_000_000, 25
g = torch.Generator(device="cuda").manual_seed(7)          # same seed on every rank -> same full matrix
full = (0.1 * torch.randn(d, generator=g, device="cuda") +
        torch.exp(0.25 * torch.randn(n, 1, generator=g, device="cuda")) * torch.randn(n, d, generator=g, device="cuda"))
c0, c1 = shard_bounds(d, world, rank)
shard = full[:, c0:c1].contiguous()
agg = ShardedAggregator()
res = {}
crafted = agg.alie(shard, f, 1.5)
ref_c, _, _ = dev.alie(full[:f], 1.5, full)                  # also writes rows 0..f-1 of `full`
res["alie"] = bool(torch.equal(crafted, ref_c[c0:c1]))
idx = agg.krum(shard, n, f, return_index=True)
res["krum"] = [idx, D.krum(full, n, f, return_index=True)]
out, sel = agg.bulyan(shard, n, f, return_selection=True)
ref_out, ref_sel = D.bulyan(full, n, f, return_selection=True)
res["bulyan_sel"] = bool(torch.equal(sel, ref_sel))
res["bulyan_out"] = bool(torch.equal(out, ref_out[c0:c1]))
res["tm"] = bool(torch.equal(agg.trimmed_mean(shard, n, f), D.trimmed_mean(full, n, f)[c0:c1]))
gathered = agg.gather_output(out, d)
res["gather"] = bool(torch.equal(gathered, ref_out))
res["exchange"] = agg.exchange_name()
# C4-like client count through the tile-pair kernel and the peer-memory exchange (reduced D)
n2, d2, f2 = 500, 262144, 100
g2 = torch.Generator(device="cuda").manual_seed(11)
full2 = torch.exp(0.25 * torch.randn(n2, 1, generator=g2, device="cuda")) * torch.randn(n2, d2, generator=g2, device="cuda")
a0, a1 = shard_bounds(d2, world, rank)
sh2 = full2[:, a0:a1].contiguous()
o2, s2 = agg.bulyan(sh2, n2, f2, return_selection=True)
r2, rs2 = D.bulyan(full2, n2, f2, return_selection=True)
res["bulyan500_sel"] = bool(torch.equal(s2, rs2))
res["bulyan500_out"] = bool(torch.allclose(o2, r2[a0:a1], rtol=1e-5, atol=1e-6))
res["krum500"] = [agg.krum(sh2, n2, f2, return_index=True), D.krum(full2, n2, f2, return_index=True)]
flags = torch.tensor([int(res["alie"]), int(res["krum"][0] == res["krum"][1]), int(res["bulyan_sel"]),
                      int(res["bulyan_out"]), int(res["tm"]), int(res["gather"]), int(res["bulyan500_sel"]),
                      int(res["bulyan500_out"]), int(res["krum500"][0] == res["krum500"][1])], device="cuda")
dist.all_reduce(flags, op=dist.ReduceOp.MIN)
if rank == 0:
    print("MULTIGPU " + json.dumps({"world": world, "all_ranks_ok": flags.tolist(), "rank0": res}))
dist.barrier(); dist.destroy_process_group()
