"""Tiny driver used under ncu: runs one kernel family a few times.  usage: run_kernel.py tm|tm_bf16|gram|gram500 [iters]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from attacking_federate_learning_b200 import _device as dev
what = sys.argv[1]; iters = int(sys.argv[2]) if len(sys.argv) > 2 else 3
g = torch.Generator(device="cuda").manual_seed(1)
if what.startswith("tm"):
    G = torch.randn(1000, 262144, generator=g, device="cuda")
    if what == "tm_bf16":
        G = G.bfloat16()
    for _ in range(iters):
        dev.trimmed_mean(G, 240)
elif what == "gram_bf16":
    G = torch.randn(100, 11_200_000, generator=g, device="cuda")
    for _ in range(iters):
        dev.sqdist_partial(G, 32)
elif what == "gram":
    G = torch.randn(100, 11_200_000, generator=g, device="cuda")
    for _ in range(iters):
        dev.sqdist_partial(G)
elif what == "pair1000":
    G = torch.randn(1000, 1 << 19, generator=g, device="cuda")
    for _ in range(iters):
        dev.sqdist_partial(G)
elif what == "pair500":
    G = torch.randn(500, 2_500_000 // 32 * 32, generator=g, device="cuda")
    for _ in range(iters):
        dev.sqdist_partial(G)
elif what == "gram500":
    G = torch.randn(500, 1 << 20, generator=g, device="cuda")
    for _ in range(iters):
        dev.sqdist_partial(G)
torch.cuda.synchronize()
print("done", what)
