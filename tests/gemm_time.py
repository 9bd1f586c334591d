"""Timing probe (not a test): GEMM-kernel-only time of the dense contractions at the eurlex shape.  The operand planes
are prepared once (engine ENGINE_PREP), the timed calls reuse them (engine ENGINE_RUN: 3 = three passes, 5 = two)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mpvae_b200.probit import contract_nt, contract_workspace
dev = "cuda:0"
M, N, K = 10240, 3993, 3993
prep, run = int(os.environ.get("ENGINE_PREP", "4")), int(os.environ.get("ENGINE_RUN", "5"))
a = torch.randn(M, K, device=dev).half().float(); b = (torch.rand(N, K, device=dev) - 0.5) * 0.06
ws = contract_workspace(M, N, K, dev, 2)
contract_nt(a, b, engine=prep, ws=ws)
for _ in range(3): contract_nt(a, b, engine=run, ws=ws)
torch.cuda.synchronize()
ts = []
for _ in range(7):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); contract_nt(a, b, engine=run, ws=ws); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
print("engine", run, "dbg", os.environ.get("MPVAE_TC_DEBUG"), "kc", os.environ.get("MPVAE_TC_KC"), "cta", os.environ.get("MPVAE_TC_CTA"),
      "ms", sorted(ts)[3])
