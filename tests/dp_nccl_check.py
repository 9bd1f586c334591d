"""Multi-GPU check (run under torchrun, one rank per GPU; not collected by pytest; MPVAE_PEER_TIMEOUT_S bounds every wait):
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tests/dp_nccl_check.py
Two ranks on half batches must walk the same trajectory as one rank on the full batch (Philox noise is keyed by
the global row, gradients are averaged over NVLink), for both the eager and the CUDA-graph step."""
import os
import sys
from types import SimpleNamespace

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mpvae_b200 import synth
from mpvae_b200.mpvae import VAE
from mpvae_b200.train import DataParallelStep, GraphedTrainStep


def build(dev, L, Z, F, fused=False):
    args = SimpleNamespace(feature_dim=F, label_dim=L, latent_dim=50, z_dim=Z, keep_prob=0.0, scale_coeff=1.0,
                           residue_sigma="", n_train_sample=10, n_test_sample=10, mode="train", nll_coeff=0.5,
                           c_coeff=10.0, noise_seed=5)
    np.random.seed(4); torch.manual_seed(0)
    vae = VAE(args).to(dev)
    with torch.no_grad():
        for head in (vae.fe_logvar, vae.fx_logvar):
            head.weight.zero_(); head.bias.fill_(-30.0)
    if fused:
        from mpvae_b200.optim import FusedAdam
        opt = FusedAdam(vae.parameters(), lr=1e-3, weight_decay=1e-5)
    else:
        opt = torch.optim.SGD(vae.parameters(), lr=0.05)
    return vae, opt, args


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    ok = True
    # (peer_g_r, fused, peer_all): NCCL bucket + SGD; g_R summed over NVLink peer memory inside the backward; the same
    # with the fused clip + Adam (the 1 / world of the gradient mean folded into its gradient multiplier); peer_all = the
    # whole bucket (and the loss terms) over peer memory, no NCCL call in the step
    for (L, Z, F, B, peer, fused, peer_all) in ((38, 38, 100, 128, False, False, False), (983, 983, 64, 256, False, False, False),
                                                (983, 983, 64, 256, True, False, False), (983, 983, 64, 256, True, True, False),
                                                (983, 983, 64, 256, False, True, False), (38, 38, 100, 128, False, False, True),
                                                (983, 983, 64, 256, True, True, True), (983, 983, 64, 255, False, True, True)):
        rng = np.random.RandomState(1)
        x = torch.from_numpy(synth.features(B, F, rng)).to(dev)
        y = torch.from_numpy(synth.labels(B, L, 0.1, rng)).to(dev)
        vae, opt, args = build(dev, L, Z, F, fused)
        step = DataParallelStep(vae, opt, None, args, clip_norm=100.0, peer_g_r=peer, peer_all=peer_all)
        assert (step.pbucket is not None) == peer_all
        outs = [step.step(y, x) for _ in range(3)]
        # reference trajectory: the same three steps on the full batch by a single (group-less) stepper
        vae1, opt1, args1 = build(dev, L, Z, F, fused)
        solo = DataParallelStep(vae1, opt1, None, args1, clip_norm=100.0, distributed=False)
        outs1 = [solo.step(y, x) for _ in range(3)]
        worst, who = 0.0, ""
        for (n, p), (_, q) in zip(vae.named_parameters(), vae1.named_parameters()):
            # relative to the parameter's scale, with an absolute floor (the zeroed logvar heads only hold rounding noise)
            d = (p.double() - q.double()).abs().max().item() / max(q.double().abs().max().item(), 1e-3)
            if d > worst:
                worst, who = d, n
        dl = max(abs(float(a.total_loss) - float(b.total_loss)) / abs(float(b.total_loss)) for a, b in zip(outs, outs1))
        if rank == 0:
            print(f"L={L} Z={Z} B={B} peer_g_r={peer} fused_adam={fused} peer_all={peer_all}: worst param rel diff {worst:.2e} ({who}), "
                  f"worst loss rel diff {dl:.2e}", flush=True)
        for ring in (step.ring, step.pbucket):
            if ring is not None:
                ring.check()
                ring.close()
        ok &= worst < 2e-4 and dl < 1e-5
    # the CUDA-graph step under N ranks (NCCL-free: peer_all): same trajectory as the eager N-rank step
    for (L, Z, F, B, peer) in ((38, 38, 100, 128, False), (983, 983, 64, 256, False), (983, 983, 64, 256, True)):
        rng = np.random.RandomState(2)
        x = torch.from_numpy(synth.features(B, F, rng)).to(dev)
        y = torch.from_numpy(synth.labels(B, L, 0.1, rng)).to(dev)
        losses = []
        for graphed in (False, True):
            vae, _, args = build(dev, L, Z, F, True)
            from mpvae_b200.optim import FusedAdam
            opt = FusedAdam(vae.parameters(), lr=torch.tensor(1e-3, device=dev), weight_decay=1e-5)
            step = DataParallelStep(vae, opt, None, args, clip_norm=100.0, peer_g_r=peer, peer_all=True)
            fn = GraphedTrainStep(step).step if graphed else step.step
            losses.append([float(fn(y, x).total_loss) for _ in range(5)])
            torch.cuda.synchronize()
            for ring in (step.ring, step.pbucket):
                if ring is not None:
                    ring.check()
                    ring.close()
        dl = max(abs(a - b) / abs(b) for a, b in zip(losses[1], losses[0]))
        if rank == 0:
            print(f"graphed vs eager, {world} ranks, L={L} peer_g_r={peer} peer_all=True: worst loss rel diff {dl:.2e} "
                  f"(losses {losses[1][0]:.4f} -> {losses[1][-1]:.4f})", flush=True)
        ok &= dl < 2e-4
    t = torch.tensor([1.0 if ok else 0.0], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("DP_NCCL_CHECK", "PASS" if t.item() == 1.0 else "FAIL", flush=True)
    dist.destroy_process_group()
    sys.exit(0 if t.item() == 1.0 else 1)


if __name__ == "__main__":
    main()
