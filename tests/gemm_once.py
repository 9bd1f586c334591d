"""Profiling helper (not a test): a few launches of the two dense contractions at the eurlex shape."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mpvae_b200.probit import contract_nt, contract_tn

dev = "cuda:0"
M, N, K = 10240, 3993, 3993
a = torch.randn(M, K, device=dev)
b = (torch.rand(N, K, device=dev) - 0.5) * 0.06
g = torch.randn(M, N, device=dev) * 1e-4
for _ in range(int(os.environ.get("REPS", "3"))):
    c = contract_nt(a, b, engine=2)
    d = contract_tn(g, a, engine=2)
torch.cuda.synchronize()
print("ok", float(c[0, 0]), float(d[0, 0]))
