"""ORACLE -- TEST INFRASTRUCTURE ONLY.  numpy restatement of the per-step metrics the reference computes on the
host from `indiv_prob`: `evals.compute_metrics(pred, target, threshold, all_metrics=False)`
(`/root/reference/evals.py:178-239`), i.e. ACC / HA / ebF1 / miF1 / maF1 / p@1,3,5.  The AUC / AUPR / FDR
entries are 0 in that mode (evals.py:186-189) and are not restated.  Pinned against the unmodified reference's
outputs in tests/golden/metrics_*.npz (tests/golden/make_golden_metrics.py).
"""
from __future__ import annotations

import numpy as np


def threshold_predictions(scores: np.ndarray, threshold: float) -> np.ndarray:
    """evals.py:201-202: p < t -> 0, p >= t -> 1 (in the scores' dtype)."""
    out = scores.copy()
    out[scores < threshold] = 0
    out[scores >= threshold] = 1
    return out


def label_counts(targets: np.ndarray, preds: np.ndarray):
    """evals.py:61-67 with axis=0: per-label tp / fp / fn as float32."""
    tp = np.sum(targets * preds, axis=0).astype("float32")
    fp = np.sum(np.logical_not(targets) * preds, axis=0).astype("float32")
    fn = np.sum(targets * np.logical_not(preds), axis=0).astype("float32")
    return tp, fp, fn


def micro_f1(tp, fp, fn):
    """evals.py:97-99."""
    return 2 * np.sum(tp) / float(2 * np.sum(tp) + np.sum(fp) + np.sum(fn))


def macro_f1(tp, fp, fn):
    """evals.py:101-109: mean over labels of 2tp / (2tp + fp + fn + 1e-6), non-finite entries dropped."""
    with np.errstate(divide="ignore", invalid="ignore"):
        c = np.true_divide(2 * tp, 2 * tp + fp + fn + 1e-6)
    return np.mean(c[np.isfinite(c)])


def example_f1(targets, preds):
    """evals.py:70-88 with axis=1: per-row 2tp / (|t| + |p|), rows with an empty denominator dropped, then the mean."""
    tp = np.sum(targets * preds, axis=1).astype("float32")
    den = np.sum(targets, axis=1).astype("float32") + np.sum(preds, axis=1).astype("float32")
    keep = den != 0
    return np.mean((2 * tp)[keep] / den[keep])


def precision_at_k(targets, scores, k):
    """evals.py:13-44: fraction of positive labels among the k highest scores (argsort ascending, reversed)."""
    uniq = np.unique(targets)
    if len(uniq) > 2:
        raise ValueError("Only supported for two relevance levels.")
    pos = uniq[1]
    order = np.argsort(scores, axis=1)[:, ::-1]
    top = np.array([row[idx] for row, idx in zip(targets, order[:, :k])])
    return np.average(np.sum(top == pos, axis=1).astype(float) / k)


def batch_metrics(scores: np.ndarray, targets: np.ndarray, threshold: float = 0.5) -> dict:
    """evals.py:178-239 with all_metrics=False."""
    out = {f"p_at_{k}": precision_at_k(targets, scores, k) for k in (1, 3, 5)}
    preds = threshold_predictions(scores, threshold)
    out["ACC"] = np.mean(np.all(targets == preds, axis=1))                       # evals.py:47-51
    out["HA"] = 1 - np.mean(np.mean(np.logical_xor(targets, preds), axis=1))     # evals.py:54-58, :209-211
    out["ebF1"] = example_f1(targets, preds)
    tp, fp, fn = label_counts(targets, preds)
    out["miF1"] = micro_f1(tp, fp, fn)
    out["maF1"] = macro_f1(tp, fp, fn)
    out["tp"], out["fp"], out["fn"] = tp, fp, fn
    return out
