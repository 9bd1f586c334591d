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


@pytest.mark.gpu
@pytest.mark.parametrize("B,L", [(7, 3), (9, 5), (33, 37), (64, 130), (16, 517), (128, 3993)])
def test_device_precision_at_k_with_ties(B, L):
    """Heavily tied scores: the kernel's ranking is 'stable argsort, reversed' (ties towards the higher index), one
    pass per row with per-thread top-5 lists merged by warps -- every merge level sees ties here."""
    from mpvae_b200.metrics import batch_metrics
    rng = np.random.RandomState(B * 1000 + L)
    p = (np.floor(rng.uniform(size=(B, L)) * 6) / 8).astype(np.float32)
    p[0, :] = 0.25                                                       # a row of one single value
    y = (rng.uniform(size=(B, L)) < 0.3).astype(np.float32)
    got = batch_metrics(torch.from_numpy(p).cuda(), torch.from_numpy(y).cuda(), 0.5)
    order = np.argsort(p, axis=1, kind="stable")[:, ::-1]
    for k in (1, 3, 5):
        kk = min(k, L)
        hits = np.take_along_axis(y, order[:, :kk], axis=1).astype(np.float64).sum(axis=1)
        want = float(np.mean(hits / k))
        assert abs(got[f"p_at_{k}"].item() - want) <= 1e-12, (k, got[f"p_at_{k}"].item(), want)
    want = ev.batch_metrics(p, y, 0.5)
    for k in ("ACC", "HA"):
        assert abs(got[k].item() - float(want[k])) <= 1e-12


# ----------------------------------------------------------------------------- threshold sweep
CURVES = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "metrics_curves.npz")
CURVE_CASES = ["yeast", "nuswide", "ties", "delicious"]


def _close(a, b, tol):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    if not np.array_equal(np.isnan(a), np.isnan(b)):
        return False
    m = ~np.isnan(a)
    return bool(np.all(np.abs(a[m] - b[m]) <= tol * np.maximum(np.abs(b[m]), 1.0)))


@pytest.mark.parametrize("name", CURVE_CASES)
def test_curve_oracle_matches_reference_evals(name):
    """compute_metrics(..., all_metrics=True) of the reference (scikit-learn curves) against the numpy restatement:
    per-label AUC / AUPR / FDR arrays, their mean / median / variance (NaN where sklearn yields NaN), and the
    thresholded metrics at every threshold of the fixture."""
    z = np.load(CURVES)
    p, y = z[name + "_p"], z[name + "_y"]
    keys = [str(k) for k in z["scalar_keys"]]
    auc, aupr, fdr = ev.label_curve_metrics(p, y)
    assert _close(auc, z[name + "_allAUC"], 1e-12)
    assert _close(aupr, z[name + "_allAUPR"], 1e-12)
    assert _close(fdr, z[name + "_allFDR"], 1e-12)
    for t in z["thresholds"]:
        got = ev.full_metrics(p, y, float(t))
        want = dict(zip(keys, z[f"{name}_t{t}_scal"]))
        for k in keys:
            assert _close(float(got[k]), want[k], 1e-12), (t, k, float(got[k]), want[k])


@pytest.mark.gpu
@pytest.mark.parametrize("name", CURVE_CASES)
def test_device_sweep_matches_reference(name):
    """mpvae_b200.metrics.sweep_metrics (one device sort + label_curves_kernel, batch metrics per threshold) against the
    reference's compute_metrics(..., all_metrics=True) at every threshold of the fixture."""
    from mpvae_b200.metrics import sweep_metrics
    z = np.load(CURVES)
    p, y = torch.from_numpy(z[name + "_p"]).cuda(), torch.from_numpy(z[name + "_y"]).cuda()
    keys = [str(k) for k in z["scalar_keys"]]
    got = sweep_metrics(p, y, [float(t) for t in z["thresholds"]])
    for k in ("allAUC", "allAUPR", "allFDR"):
        assert _close(got[0][k].cpu().numpy(), z[f"{name}_{k}"], 1e-12), k
    for t, m in zip(z["thresholds"], got):
        want = dict(zip(keys, z[f"{name}_t{t}_scal"]))
        for k in keys:
            if name == "ties" and k.startswith("p_at_"):
                continue      # top-k among TIED scores: the reference's numpy introsort leaves their order unspecified
            tol = 3e-6 if k in ("ebF1", "miF1", "maF1") else 1e-12        # the reference rounds these three through fp32
            assert _close(m[k].item(), want[k], tol), (float(t), k, m[k].item(), want[k])


@pytest.mark.gpu
def test_device_curves_full_size():
    """eurlex-sized evaluation set (3 809 x 3 993, sparse labels, some without positives) against the numpy oracle."""
    from mpvae_b200.metrics import label_curves
    rng = np.random.RandomState(12)
    N, L = 3809, 3993
    y = (rng.uniform(size=(N, L)) < 0.002).astype(np.float32)
    p = (rng.uniform(size=(N, L)) * (0.3 + 0.7 * y)).astype(np.float32)
    p = (np.round(p * 4096) / 4096).astype(np.float32)                    # ties
    auc, aupr, fdr = label_curves(torch.from_numpy(p).cuda(), torch.from_numpy(y).cuda())
    cols = rng.choice(L, 200, replace=False)                              # the oracle is a Python loop over labels
    w_auc, w_aupr, w_fdr = ev.label_curve_metrics(p[:, cols], y[:, cols])
    assert _close(auc.cpu().numpy()[cols], w_auc, 1e-12)
    assert _close(aupr.cpu().numpy()[cols], w_aupr, 1e-12)
    assert _close(fdr.cpu().numpy()[cols], w_fdr, 1e-12)
