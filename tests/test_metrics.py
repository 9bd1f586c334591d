"""Per-step metrics: the numpy oracle against golden vectors of the reference's evals.py (CPU), and
the CUDA kernels against the oracle (GPU)."""
import os

import numpy as np
import pytest
import torch

from oracle import evals_oracle as ev

KEYS = ["ACC", "HA", "ebF1", "miF1", "maF1", "p_at_1", "p_at_3", "p_at_5"]
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "metrics_cases.npz")
CASES = ["yeast", "mirflickr", "nuswide", "delicious", "tiny"]


@pytest.mark.parametrize("name", CASES)
def test_oracle_matches_reference_evals(name):
    z = np.load(GOLD)
    got = ev.batch_metrics(z[name + "_p"], z[name + "_y"], 0.5)
    want = z[name + "_m"]
    for k, w in zip(KEYS, want):
        assert float(got[k]) == w, (k, float(got[k]), w)       # same numpy ops in the same order: identical


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_device_metrics_match_reference(name):
    from mpvae_b200.metrics import batch_metrics
    z = np.load(GOLD)
    p, y = torch.from_numpy(z[name + "_p"]).cuda(), torch.from_numpy(z[name + "_y"]).cuda()
    got = batch_metrics(p, y, 0.5)
    want = dict(zip(KEYS, z[name + "_m"]))
    for k in ("ACC", "HA", "p_at_1", "p_at_3", "p_at_5"):                     # integer counts / exact ratios
        assert abs(got[k].item() - want[k]) <= 1e-12, (k, got[k].item(), want[k])
    for k in ("ebF1", "miF1", "maF1"):                                        # the reference rounds these through fp32
        assert abs(got[k].item() - want[k]) <= 2e-6 * max(want[k], 1e-3), (k, got[k].item(), want[k])


@pytest.mark.gpu
def test_device_metrics_full_size():
    """eurlex-sized batch (1024 x 3993) against the numpy oracle.  (Tied scores are not exercised here: the
    reference ranks with numpy's default introsort, which leaves the order of ties unspecified for long rows; the
    kernel resolves ties towards the higher index, i.e. what a stable argsort + reverse gives.)"""
    from mpvae_b200.metrics import batch_metrics
    rng = np.random.RandomState(9)
    B, L = 1024, 3993
    y = (rng.uniform(size=(B, L)) < 0.005).astype(np.float32)
    y[:, 0], y[:, 1] = 1.0, 0.0
    p = rng.uniform(size=(B, L)).astype(np.float32)
    got = batch_metrics(torch.from_numpy(p).cuda(), torch.from_numpy(y).cuda(), 0.5)
    want = ev.batch_metrics(p, y, 0.5)
    for k in KEYS:
        tol = 1e-12 if k in ("ACC", "HA") else 3e-6
        assert abs(got[k].item() - float(want[k])) <= tol * max(abs(float(want[k])), 1.0), (k, got[k].item(), float(want[k]))
