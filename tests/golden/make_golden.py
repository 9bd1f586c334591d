"""Generate golden vectors by running the UNMODIFIED reference (`/root/reference/mpvae.py`).

Run in the build container only (the GPU box has no /root/reference):
    python tests/golden/make_golden.py
Writes tests/golden/<case>.npz.  Each file holds the inputs (or the seeds to regenerate the large
ones with `mpvae_b200.synth`), the noise the reference drew at mpvae.py:162 (captured by seeding the
CPU generator right before the call), the 8 outputs of compute_loss (mpvae.py:210) and the autograd
gradients w.r.t. fe_out, fx_out, mu/logvar x4 and r_sqrt_sigma.
"""
import os
import sys
from types import SimpleNamespace

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")

import mpvae as ref_mpvae            # noqa: E402  (the reference, unmodified)
from mpvae_b200 import synth         # noqa: E402

GRAD_KEYS = ["fe_out", "fe_mu", "fe_logvar", "fx_out", "fx_mu", "fx_logvar", "r_sqrt_sigma"]
OUT_KEYS = ["total_loss", "nll_loss", "nll_loss_x", "c_loss", "c_loss_x", "kl_loss",
            "indiv_prob", "indiv_prob_label"]

# name -> (L, Z, B, S, mode, sigma, label_rate, nll_coeff, c_coeff, data_seed, noise_seed, extras)
CASES = {
    "mirflickr_b16":   dict(L=38, Z=38, B=16, S=10, mode="train", sigma=1.0, rate=0.1),
    "yeast_b16":       dict(L=14, Z=14, B=16, S=10, mode="train", sigma=0.5, rate=0.1),
    "nuswide_test_b8": dict(L=81, Z=81, B=8, S=100, mode="test", sigma=1.0, rate=0.1),
    "delicious_b2":    dict(L=983, Z=983, B=2, S=10, mode="train", sigma=1.0, rate=20.0 / 983,
                            big_r=True),
    "eurlex_z10_b2":   dict(L=3993, Z=10, B=2, S=10, mode="train", sigma=1.0, rate=20.0 / 3993),
    "mirflickr_tails": dict(L=38, Z=38, B=16, S=10, mode="train", sigma=3.0, rate=0.1),
    "ragged_s1":       dict(L=23, Z=7, B=5, S=1, mode="train", sigma=1.0, rate=0.2,
                            nll_coeff=0.1, c_coeff=200.0, D=17),
    "fair_upstream":   dict(L=38, Z=10, B=8, S=10, mode="train", sigma=1.0, rate=0.1, upstream=True),
    "degenerate_rows": dict(L=12, Z=12, B=6, S=4, mode="train", sigma=1.0, rate=0.3, degenerate=True),
    "lowrank_z10":     dict(L=81, Z=10, B=32, S=10, mode="train", sigma=1.0, rate=0.1),
}


def run_reference(inp, noise_seed, cfg, upstream=None):
    L, Z, S = cfg["L"], cfg["Z"], cfg["S"]
    args = SimpleNamespace(label_dim=L, z_dim=Z, n_train_sample=S, n_test_sample=S, mode=cfg["mode"],
                           nll_coeff=cfg.get("nll_coeff", 0.5), c_coeff=cfg.get("c_coeff", 10.0))
    t = {k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in inp.items() if k != "noise"}
    train = cfg["mode"] == "train"
    leaves = {k: t[k].clone().requires_grad_(train) for k in GRAD_KEYS}
    # the draw the reference is about to make (mpvae.py:162): same seed, same generator, same call
    torch.manual_seed(noise_seed)
    noise = torch.normal(0, 1, size=(S, t["y"].shape[0], Z))
    torch.manual_seed(noise_seed)
    ctx = torch.enable_grad() if train else torch.no_grad()
    with ctx:
        outs = ref_mpvae.compute_loss(t["y"], leaves["fe_out"], leaves["fe_mu"], leaves["fe_logvar"],
                                      leaves["fx_out"], leaves["fx_mu"], leaves["fx_logvar"],
                                      leaves["r_sqrt_sigma"], args)
    assert len(outs) == 8
    res = {"out_" + k: o.detach().numpy() for k, o in zip(OUT_KEYS, outs)}
    if train:
        objective = outs[0]
        if upstream is not None:
            objective = objective + (outs[6] * upstream["indiv_prob"]).sum() \
                + (outs[7] * upstream["indiv_prob_label"]).sum()
        grads = torch.autograd.grad(objective, [leaves[k] for k in GRAD_KEYS])
        for k, g in zip(GRAD_KEYS, grads):
            res["grad_" + k] = g.numpy()
    return noise.numpy(), res


def main():
    os.makedirs(HERE, exist_ok=True)
    for idx, (name, cfg) in enumerate(CASES.items()):
        data_seed, noise_seed = 100 + idx, 900 + idx
        inp = synth.loss_inputs(cfg["L"], cfg["Z"], cfg["B"], cfg["S"], seed=data_seed,
                                sigma=cfg["sigma"], label_rate=cfg["rate"],
                                latent_dim=cfg.get("D", 50), with_noise=False)
        if cfg.get("degenerate"):
            inp["y"][1, :] = 0.0     # no positive label
            inp["y"][4, :] = 1.0     # no negative label
        upstream = None
        if cfg.get("upstream"):
            rng = np.random.RandomState(5000 + idx)
            upstream = {"indiv_prob": torch.from_numpy(rng.standard_normal((cfg["B"], cfg["L"])).astype(np.float32)),
                        "indiv_prob_label": torch.from_numpy(rng.standard_normal((cfg["B"], cfg["L"])).astype(np.float32))}
        noise, res = run_reference(inp, noise_seed, cfg, upstream)
        payload = dict(res)
        payload["noise"] = noise
        payload["meta"] = np.array([cfg["L"], cfg["Z"], cfg["B"], cfg["S"], cfg.get("D", 50), data_seed, noise_seed],
                                   dtype=np.int64)
        payload["coeffs"] = np.array([cfg.get("nll_coeff", 0.5), cfg.get("c_coeff", 10.0), cfg["sigma"], cfg["rate"]],
                                     dtype=np.float64)
        payload["mode"] = np.array(cfg["mode"])
        payload["torch_version"] = np.array(torch.__version__)
        payload["degenerate"] = np.array(bool(cfg.get("degenerate")))
        if upstream is not None:
            payload["up_indiv_prob"] = upstream["indiv_prob"].numpy()
            payload["up_indiv_prob_label"] = upstream["indiv_prob_label"].numpy()
        big_r = cfg.get("big_r", False)
        for k, v in inp.items():
            if k == "r_sqrt_sigma" and big_r:
                continue             # regenerated from data_seed by synth.loss_inputs (7.7 MB fp64)
            payload["in_" + k] = v
        if big_r and "grad_r_sqrt_sigma" in payload:
            # keep g_R as a checkable digest: a strided subset + row / column sums + two projections
            g = payload.pop("grad_r_sqrt_sigma")
            rng = np.random.RandomState(77)
            u = rng.standard_normal(g.shape[1])
            v = rng.standard_normal(g.shape[0])
            payload["grad_r_digest_sub"] = g[::29, ::31].copy()
            payload["grad_r_digest_rowsum"] = g.sum(1)
            payload["grad_r_digest_colsum"] = g.sum(0)
            payload["grad_r_digest_gu"] = g @ u
            payload["grad_r_digest_vg"] = v @ g
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, **payload)
        print(f"{name:18s} total={float(res['out_total_loss']):.8f}  -> {os.path.getsize(path)/1024:.1f} KiB")


if __name__ == "__main__":
    main()
