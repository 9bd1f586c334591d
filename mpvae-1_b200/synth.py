"""Seeded synthetic stand-ins for the datasets the reference does not ship.

Shape-only inputs for the probit-ELBO path: what `VAE.forward` (mpvae.py:97-100) would hand to
`compute_loss` (mpvae.py:145).  numpy `RandomState` (frozen MT19937 stream) so fixtures regenerate
bit-identically anywhere.  Names follow BASELINE.json's configs.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np


@dataclass(frozen=True)
class Shape:
    name: str
    feature_dim: int   # F
    label_dim: int     # L
    z_dim: int         # Z (script convention z_dim = label_dim, run_train_mirflickr.sh:1)
    batch: int         # B
    n_sample: int      # S
    mode: str          # 'train' | 'test'
    label_rate: float  # Bernoulli rate of positive labels
    latent_dim: int = 50


# BASELINE.json configs[0..4]
SHAPES = {
    "mirflickr": Shape("mirflickr", 1000, 38, 38, 128, 10, "train", 0.1),
    "yeast": Shape("yeast", 103, 14, 14, 128, 10, "train", 0.1),
    "nuswide": Shape("nuswide", 128, 81, 81, 128, 100, "test", 0.1),
    "delicious": Shape("delicious", 500, 983, 983, 128, 10, "train", 20.0 / 983),
    "eurlex": Shape("eurlex", 5000, 3993, 3993, 1024, 10, "train", 20.0 / 3993),
}


def xavier_r(label_dim: int, z_dim: int, rng: np.random.RandomState) -> np.ndarray:
    """r_sqrt_sigma init of mpvae.py:47-48: float64 uniform(+-sqrt(6/(L+Z))), shape (L, Z)."""
    bound = np.sqrt(6.0 / (label_dim + z_dim))
    return rng.uniform(-bound, bound, (label_dim, z_dim))


def labels(batch: int, label_dim: int, rate: float, rng: np.random.RandomState) -> np.ndarray:
    """{0,1} float32 label rows; every row has >=1 positive and >=1 negative label (else the
    reference backward is NaN, mpvae.py:118-121)."""
    y = (rng.uniform(size=(batch, label_dim)) < rate).astype(np.float32)
    if label_dim >= 2:
        y[:, 0] = 1.0
        y[:, 1] = 0.0
    return y


def loss_inputs(label_dim: int, z_dim: int, batch: int, n_sample: int, *, seed: int = 7,
                sigma: float = 1.0, label_rate: float = 0.1, latent_dim: int = 50,
                with_noise: bool = True, mulv_std: float = 0.5) -> dict:
    """Direct inputs of the loss (kernel-only benches of SURVEY.md section 8d): logits ~ N(0, sigma^2),
    mu / logvar ~ N(0, 0.25), R = Xavier fp64, noise ~ N(0, 1) fp32 of shape (S, B, Z)."""
    rng = np.random.RandomState(seed)
    out = {
        "y": labels(batch, label_dim, label_rate, rng),
        "fe_out": (rng.standard_normal((batch, label_dim)) * sigma).astype(np.float32),
        "fx_out": (rng.standard_normal((batch, label_dim)) * sigma).astype(np.float32),
        "fe_mu": (rng.standard_normal((batch, latent_dim)) * mulv_std).astype(np.float32),
        "fe_logvar": (rng.standard_normal((batch, latent_dim)) * mulv_std).astype(np.float32),
        "fx_mu": (rng.standard_normal((batch, latent_dim)) * mulv_std).astype(np.float32),
        "fx_logvar": (rng.standard_normal((batch, latent_dim)) * mulv_std).astype(np.float32),
        "r_sqrt_sigma": xavier_r(label_dim, z_dim, rng),
    }
    if with_noise:
        out["noise"] = rng.standard_normal((n_sample, batch, z_dim)).astype(np.float32)
    return out


def features(batch: int, feature_dim: int, rng: np.random.RandomState) -> np.ndarray:
    return rng.standard_normal((batch, feature_dim)).astype(np.float32)


def make_args(label_dim: int, z_dim: int, n_train_sample: int = 10, n_test_sample: int = 100, mode: str = "train",
              nll_coeff: float = 0.5, c_coeff: float = 10.0, **extra):
    """The fields `compute_loss` reads from `args` (mpvae.py:153,158,162,207-208), as the namespace main.py builds."""
    from types import SimpleNamespace
    return SimpleNamespace(label_dim=label_dim, z_dim=z_dim, n_train_sample=n_train_sample,
                           n_test_sample=n_test_sample, mode=mode, nll_coeff=nll_coeff, c_coeff=c_coeff, **extra)
