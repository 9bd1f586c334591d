"""Golden vectors for the per-step metrics from the UNMODIFIED reference `evals.py`.
Run in the build container only:  python tests/golden/make_golden_metrics.py  -> tests/golden/metrics_cases.npz"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, "/root/reference")
import evals  # noqa: E402  (the reference, unmodified)

KEYS = ["ACC", "HA", "ebF1", "miF1", "maF1", "p_at_1", "p_at_3", "p_at_5"]
CASES = {"yeast": (128, 14, 0.3), "mirflickr": (128, 38, 0.1), "nuswide": (77, 81, 0.1), "delicious": (64, 983, 0.02),
         "tiny": (5, 6, 0.5)}

payload = {}
for i, (name, (B, L, rate)) in enumerate(CASES.items()):
    rng = np.random.RandomState(300 + i)
    y = (rng.uniform(size=(B, L)) < rate).astype(np.float32)
    y[:, 0], y[:, 1] = 1.0, 0.0
    logits = rng.standard_normal((B, L)).astype(np.float32) + (y * 2 - 1) * 0.8
    p = (1.0 / (1.0 + np.exp(-logits))).astype(np.float32)
    if name == "tiny":
        p[0] = 0.2            # a row predicting nothing (and y row forced to have positives): ebF1 denominator > 0
        y[1, :] = 0.0
        p[1] = 0.1            # empty target AND empty prediction: dropped from ebF1 (evals.py:79-82)
    m = evals.compute_metrics(p, y, 0.5, all_metrics=False)
    payload[f"{name}_p"], payload[f"{name}_y"] = p, y
    payload[f"{name}_m"] = np.array([float(m[k]) for k in KEYS], dtype=np.float64)
    print(name, {k: float(m[k]) for k in KEYS})
np.savez_compressed(os.path.join(HERE, "metrics_cases.npz"), **payload)
