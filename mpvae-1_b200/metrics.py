"""Per-step training metrics on the device.

The reference evaluates `evals.compute_metrics(indiv_prob.cpu().data.numpy(), input_label.cpu().data.numpy(), 0.5,
all_metrics=False)` after EVERY optimizer step (train.py:131, fairsoft_train.py:149): a device->host sync, two
copies and a numpy pass that dwarf the GPU step at the small configurations.  `batch_metrics` computes the same
eight numbers (reference evals.py:178-239) with integer tp/fp/fn counts on the GPU and returns device scalars, so
nothing synchronises until the caller actually prints them."""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib

NAMES = ("ACC", "HA", "ebF1", "miF1", "maF1", "p_at_1", "p_at_3", "p_at_5")


def batch_metrics(indiv_prob: torch.Tensor, input_label: torch.Tensor, threshold: float = 0.5) -> dict:
    """Same keys as the reference's metrics_dict (the AUC/AUPR/FDR entries, 0 in this mode, are omitted).
    Values are 0-dim float64 CUDA tensors."""
    out = batch_metrics_tensor(indiv_prob, input_label, threshold)
    return dict(zip(NAMES, out.unbind(0)))


def batch_metrics_tensor(indiv_prob: torch.Tensor, input_label: torch.Tensor, threshold: float = 0.5) -> torch.Tensor:
    """The eight metrics in the order of `NAMES` as ONE (8,) float64 CUDA tensor (one device-to-host copy fetches them)."""
    if not (indiv_prob.is_cuda and input_label.is_cuda):
        raise RuntimeError("mpvae_b200.batch_metrics runs on CUDA tensors only (use the reference's evals on the host)")
    p = indiv_prob.detach().float().contiguous()
    y = input_label.detach().float().contiguous()
    if p.shape != y.shape or p.dim() != 2:
        raise ValueError(f"shapes disagree: {tuple(p.shape)} vs {tuple(y.shape)}")
    B, L = p.shape
    lib = _lib.lib()
    out = torch.empty(8, dtype=torch.float64, device=p.device)
    if B == 0:
        return out.fill_(float("nan"))
    with torch.cuda.device(p.device):
        nbytes = int(lib.mpvae_batch_metrics_workspace(B, L))
        ws = torch.empty(nbytes, dtype=torch.uint8, device=p.device)
        stream = C.c_void_p(torch.cuda.current_stream(p.device).cuda_stream)
        _lib.check(lib.mpvae_batch_metrics(C.c_void_p(p.data_ptr()), C.c_void_p(y.data_ptr()), B, L, float(threshold),
                                           C.c_void_p(out.data_ptr()), C.c_void_p(ws.data_ptr()), nbytes, stream),
                   "mpvae_batch_metrics")
    return out


# ----------------------------------------------------------------------------------------------------------------
# the test-time threshold sweep (test.py:90-101, train.py:277-289): for each of 27 thresholds the
# reference calls compute_metrics(pred, label, t, all_metrics=True), which recomputes the threshold-INDEPENDENT
# per-label scikit-learn curves (AUC, AUPR, FDR-recall: 3 x L sklearn calls) every time.  Here the curves are computed
# once on the device (one sort + one kernel), the thresholded counts once per threshold (three small kernels each).
CURVE_NAMES = ("AUC", "AUPR", "FDR")


def label_curves(indiv_prob: torch.Tensor, input_label: torch.Tensor, fdr_cutoff: float = 0.5):
    """(allAUC, allAUPR, allFDR): three (L,) float64 CUDA tensors (evals.py:129-175 with scikit-learn 1.9 semantics)."""
    if not (indiv_prob.is_cuda and input_label.is_cuda):
        raise RuntimeError("mpvae_b200.label_curves runs on CUDA tensors only")
    p = indiv_prob.detach().float().contiguous()
    y = input_label.detach().float().contiguous()
    if p.shape != y.shape or p.dim() != 2 or p.shape[0] == 0:
        raise ValueError(f"bad shapes: {tuple(p.shape)} vs {tuple(y.shape)}")
    N, L = p.shape
    lib = _lib.lib()
    with torch.cuda.device(p.device):
        scores, order = torch.sort(p, dim=0, descending=True)        # plumbing: the library kernel walks sorted columns
        targets = torch.gather(y, 0, order)
        out = torch.empty((3, L), dtype=torch.float64, device=p.device)
        stream = C.c_void_p(torch.cuda.current_stream(p.device).cuda_stream)
        _lib.check(lib.mpvae_label_curves(C.c_void_p(scores.data_ptr()), C.c_void_p(targets.data_ptr()), N, L,
                                          float(fdr_cutoff), C.c_void_p(out.data_ptr()), stream), "mpvae_label_curves")
    return out[0], out[1], out[2]


def _np_median(t: torch.Tensor) -> torch.Tensor:
    """numpy.median semantics on a 1-D device tensor: mean of the two middle values, NaN if any NaN."""
    s, _ = torch.sort(t)
    n = s.numel()
    mid = (s[(n - 1) // 2] + s[n // 2]) * 0.5
    return torch.where(torch.isnan(t).any(), torch.full_like(mid, float("nan")), mid)


def sweep_metrics(indiv_prob: torch.Tensor, input_label: torch.Tensor, thresholds, fdr_cutoff: float = 0.5) -> list:
    """One metrics dict per threshold with the keys of the reference's compute_metrics(..., all_metrics=True)
    (evals.py:218-239); values are device tensors (0-dim float64, the `all*` entries (L,) float64)."""
    auc, aupr, fdr = label_curves(indiv_prob, input_label, fdr_cutoff)
    shared = {}
    for name, arr in zip(CURVE_NAMES, (auc, aupr, fdr)):
        shared["mean" + name] = arr.mean()
        shared["median" + name] = _np_median(arr)
        shared["var" + name] = arr.var(unbiased=False)
        shared["all" + name] = arr
    out = []
    for t in thresholds:
        m = batch_metrics(indiv_prob, input_label, float(t))
        m.update(shared)
        out.append(m)
    return out
