"""Where the end-to-end step's extra time goes (not collected by pytest): device times of the pieces bench.py's `e2e`
leg adds to the loss step at the eurlex shape -- the byte -> float cast of the labels, the per-step metrics
(mpvae_b200.metrics), the host -> device copy of one batch.  python tests/e2e_breakdown.py [B L]"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mpvae_b200.metrics import batch_metrics_tensor


def timeit(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def main():
    B, L = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (1024, 3993)
    dev = torch.device("cuda", 0)
    rng = np.random.RandomState(0)
    p = torch.from_numpy(rng.uniform(size=(B, L)).astype(np.float32)).to(dev)
    y8h = torch.from_numpy((rng.uniform(size=(B, L)) < 0.005).astype(np.uint8)).pin_memory()
    ph = p.cpu().pin_memory()
    y8 = y8h.to(dev)
    y = y8.float()
    out = {"B": B, "L": L}
    out["label_cast_ms"] = timeit(lambda: y8.float())
    out["batch_metrics_ms"] = timeit(lambda: batch_metrics_tensor(p, y, 0.5))
    out["h2d_two_logit_matrices_and_labels_ms"] = timeit(lambda: (ph.to(dev, non_blocking=True), ph.to(dev, non_blocking=True),
                                                                   y8h.to(dev, non_blocking=True)))
    big = torch.empty(64 << 20, dtype=torch.float32).pin_memory()          # 256 MB in one copy: the link's rate
    bigd = torch.empty_like(big, device=dev)
    out["h2d_256MB_GBps"] = 268.435456 / timeit(lambda: bigd.copy_(big, non_blocking=True), 5)
    streams = [torch.cuda.Stream(dev) for _ in range(3)]
    ph2 = ph.clone().pin_memory()

    def three_streams():
        cur = torch.cuda.current_stream(dev)
        for st, src in zip(streams, (ph, ph2, y8h)):
            st.wait_stream(cur)
            with torch.cuda.stream(st):
                src.to(dev, non_blocking=True)
        for st in streams:
            cur.wait_stream(st)
    out["h2d_same_bytes_on_three_streams_ms"] = timeit(three_streams)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
