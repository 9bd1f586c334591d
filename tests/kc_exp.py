"""Experiment (not a test): accuracy / speed of the tcgen05 engine vs the TMEM chunk length MPVAE_TC_KC."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mpvae_b200.probit import contract_nt, contract_tn

kc = os.environ.get("MPVAE_TC_KC")
dev = "cuda:0"
for (M, N, K) in [(10240, 983, 983), (10240, 3993, 3993)]:
    g = torch.Generator(device="cpu").manual_seed(1)
    a = torch.randn(M, K, generator=g).to(dev)
    b = ((torch.rand(N, K, generator=g) - 0.5) * 0.06).to(dev)
    want = a[:512].double() @ b.double().T
    res = {}
    for eng in (2, 1):
        for _ in range(2):
            out = contract_nt(a, b, engine=eng)
        torch.cuda.synchronize()
        ts = []
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); out = contract_nt(a, b, engine=eng); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        d = (out[:512].double() - want)
        res[eng] = dict(ms=sorted(ts)[2], max=(d.abs().max() / want.abs().max()).item(),
                        rms=(d.pow(2).mean().sqrt() / want.pow(2).mean().sqrt()).item(),
                        bias=(d * want.sign()).mean().item() / want.abs().mean().item())
    print(json.dumps(dict(kc=kc, M=M, N=N, K=K, res=res)))
