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


# ----------------------------------------------------------------------------------------------------------------
# the threshold-independent per-label curves of `compute_metrics(..., all_metrics=True)`
# (evals.py:129-175).  The reference delegates them to scikit-learn (NOT vendored; this container has 1.9.0):
#   compute_auc   -> metrics.roc_auc_score(y, s)                      = trapezoid area under the ROC points
#   compute_aupr  -> metrics.precision_recall_curve + metrics.auc     = trapezoid area under the PR points
#   compute_fdr   -> recall at the first PR point (increasing threshold) whose 1 - precision <= cutoff
# Restated from the published algorithm (sklearn/metrics/_ranking.py 1.9.0: one curve point per DISTINCT score, in
# decreasing score order, tps / fps cumulative; PR curve closed with the point (recall 0, precision 1); no truncation at
# full recall; a label without positives gets recall 1 everywhere; a label with a single class gets AUC = NaN) and
# pinned against the reference run with that sklearn (tests/golden/metrics_curves.npz).
def label_curve_metrics(scores: np.ndarray, targets: np.ndarray, fdr_cutoff: float = 0.5):
    """Per-label (AUC, AUPR, FDR-recall) arrays of length L, fp64."""
    n, L = scores.shape
    auc, aupr, fdr = np.empty(L), np.empty(L), np.empty(L)
    for l in range(L):
        order = np.argsort(-scores[:, l].astype(np.float64), kind="stable")
        s, y = scores[order, l], targets[order, l] != 0
        last_of_group = np.r_[s[1:] != s[:-1], True]            # one point per distinct score
        tps = np.cumsum(y)[last_of_group].astype(np.float64)
        fps = np.cumsum(~y)[last_of_group].astype(np.float64)
        n_pos, n_neg = tps[-1], fps[-1]
        # ROC: (0,0) then (fps/n_neg, tps/n_pos)
        if n_pos == 0 or n_neg == 0:
            auc[l] = np.nan
        else:
            fpr, tpr = np.r_[0.0, fps / n_neg], np.r_[0.0, tps / n_pos]
            auc[l] = np.sum(np.diff(fpr) * (tpr[1:] + tpr[:-1]) * 0.5)
        # PR in decreasing-threshold order, preceded by the closing point (recall 0, precision 1)
        prec = np.r_[1.0, tps / (tps + fps)]
        rec = np.r_[0.0, (tps / n_pos) if n_pos > 0 else np.ones_like(tps)]
        aupr[l] = np.sum(np.diff(rec) * (prec[1:] + prec[:-1]) * 0.5)
        # sklearn's arrays run the other way (increasing threshold): the FIRST point with fdr <= cutoff there is the LAST
        # one here; the closing point (fdr 0) always qualifies
        ok = np.nonzero(1.0 - prec <= fdr_cutoff)[0]
        fdr[l] = rec[ok[-1]]
    return auc, aupr, fdr


def _summary(a):
    return float(np.mean(a)), float(np.median(a)), float(np.var(a))


def full_metrics(scores: np.ndarray, targets: np.ndarray, threshold: float) -> dict:
    """evals.py:178-239 with all_metrics=True."""
    out = batch_metrics(scores, targets, threshold)
    auc, aupr, fdr = label_curve_metrics(scores, targets)
    for name, arr in (("AUC", auc), ("AUPR", aupr), ("FDR", fdr)):
        out["mean" + name], out["median" + name], out["var" + name] = _summary(arr)
        out["all" + name] = arr
    return out
