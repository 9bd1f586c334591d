"""Golden vectors for the threshold sweep from the UNMODIFIED reference `evals.py`
(compute_metrics(..., all_metrics=True), which calls scikit-learn -- version recorded in the file).
Run in the build container only:  python tests/golden/make_golden_curves.py  -> tests/golden/metrics_curves.npz"""
import os
import sys
import warnings

import numpy as np
import sklearn

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, "/root/reference")
import evals  # noqa: E402  (the reference, unmodified)

THRESHOLDS = [0.05, 0.3, 0.5, 0.75]        # a subset of train.py:22
SCAL = ["ACC", "HA", "ebF1", "miF1", "maF1", "p_at_1", "p_at_3", "p_at_5", "meanAUC", "medianAUC", "varAUC",
        "meanAUPR", "medianAUPR", "varAUPR", "meanFDR", "medianFDR", "varFDR"]
CASES = {"yeast": (300, 14, 0.3, False), "nuswide": (257, 81, 0.1, False), "ties": (200, 20, 0.2, True),
         "delicious": (150, 983, 0.02, False)}

payload = {"sklearn_version": np.array(sklearn.__version__), "thresholds": np.array(THRESHOLDS)}
for i, (name, (N, L, rate, ties)) in enumerate(CASES.items()):
    rng = np.random.RandomState(500 + i)
    y = (rng.uniform(size=(N, L)) < rate).astype(np.float32)
    y[0, :], y[1, :] = 1.0, 0.0                # every label has both classes ...
    if name in ("ties", "delicious"):
        y[:, 3] = 0.0                          # ... except a label without positives
        y[:, 4] = 1.0                          # and one without negatives
    logits = rng.standard_normal((N, L)).astype(np.float32) + (y * 2 - 1) * 0.8
    p = (1.0 / (1.0 + np.exp(-logits))).astype(np.float32)
    if ties:
        p = (np.round(p * 8) / 8).astype(np.float32)      # heavy ties: nine distinct scores
    payload[f"{name}_p"], payload[f"{name}_y"] = p, y
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for t in THRESHOLDS:
            m = evals.compute_metrics(p, y, t, all_metrics=True)
            payload[f"{name}_t{t}_scal"] = np.array([float(m[k]) for k in SCAL], dtype=np.float64)
        for k in ("allAUC", "allAUPR", "allFDR"):
            payload[f"{name}_{k}"] = np.asarray(m[k], dtype=np.float64)
    print(name, {k: float(m[k]) for k in SCAL[8:]}, len(m["allAUC"]), len(m["allAUPR"]), len(m["allFDR"]))
payload["scalar_keys"] = np.array(SCAL)
np.savez_compressed(os.path.join(HERE, "metrics_curves.npz"), **payload)
