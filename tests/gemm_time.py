"""Timing probe (not a test): GEMM-only time of the dense contractions at the eurlex shape (pre-passes excluded by
timing the library's kernels through CUDA events around repeated contract_nt calls minus a split-only baseline)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mpvae_b200.probit import contract_nt
dev = "cuda:0"
M, N, K = 10240, 3993, 3993
a = torch.randn(M, K, device=dev); b = (torch.rand(N, K, device=dev) - 0.5) * 0.06
for _ in range(3): contract_nt(a, b, engine=2)
torch.cuda.synchronize()
ts = []
for _ in range(7):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); contract_nt(a, b, engine=2); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
print(os.environ.get("MPVAE_TC_DEBUG"), os.environ.get("MPVAE_TC_KC"), os.environ.get("MPVAE_TC_CTA"), "ms", sorted(ts)[3])
