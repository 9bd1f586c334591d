"""Experiment (not a test): accuracy / speed of the tcgen05 engine vs operand kind (MPVAE_TC_KIND = tf32 | f16)
and TMEM chunk length (MPVAE_TC_KC k-blocks).  Prints one JSON line per shape."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mpvae_b200.probit import contract_nt, contract_tn

kc, kind = os.environ.get("MPVAE_TC_KC"), os.environ.get("MPVAE_TC_KIND")
dev = "cuda:0"
ENGINES = tuple(int(e) for e in os.environ.get("KC_EXP_ENGINES", "2,4,1").split(","))


def stats(out, want):
    d = out.double() - want
    return dict(max=(d.abs().max() / want.abs().max()).item(), rms=(d.pow(2).mean().sqrt() / want.pow(2).mean().sqrt()).item(),
                bias=(d * want.sign()).mean().item() / want.abs().mean().item())


def timed(fn):
    for _ in range(2):
        out = fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); out = fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return out, sorted(ts)[2]


for (M, N, K) in [(10240, 983, 983), (10240, 3993, 3993)]:
    g = torch.Generator(device="cpu").manual_seed(1)
    a = torch.randn(M, K, generator=g).half().float().to(dev)     # noise on the fp16 grid, like the library's own
    b = ((torch.rand(N, K, generator=g) - 0.5) * 0.06).to(dev)
    want = a[:512].double() @ b.double().T
    res = {}
    for eng in ENGINES:
        out, ms = timed(lambda: contract_nt(a, b, engine=eng))
        res["nt%d" % eng] = dict(ms=ms, **stats(out[:512], want))
    # tn: C[N1, N2] = A[M, N1]^T B[M, N2] with gradient-like magnitudes on A
    ga = (torch.randn(M, N, generator=g) * 1e-4 * torch.rand(M, 1, generator=g)).to(dev)
    nb = torch.randn(M, K, generator=g).half().float().to(dev)
    want_t = ga[:, :256].double().T @ nb.double()
    for eng in ENGINES:
        out, ms = timed(lambda: contract_tn(ga, nb, engine=eng))
        res["tn%d" % eng] = dict(ms=ms, **stats(out[:256], want_t))
    print(json.dumps(dict(kind=kind, kc=kc, M=M, N=N, K=K, res=res)))
