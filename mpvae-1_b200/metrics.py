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
        return {k: out.new_full((), float("nan")) for k in NAMES}
    with torch.cuda.device(p.device):
        nbytes = int(lib.mpvae_batch_metrics_workspace(B, L))
        ws = torch.empty(nbytes, dtype=torch.uint8, device=p.device)
        stream = C.c_void_p(torch.cuda.current_stream(p.device).cuda_stream)
        _lib.check(lib.mpvae_batch_metrics(C.c_void_p(p.data_ptr()), C.c_void_p(y.data_ptr()), B, L, float(threshold),
                                           C.c_void_p(out.data_ptr()), C.c_void_p(ws.data_ptr()), nbytes, stream),
                   "mpvae_batch_metrics")
    return dict(zip(NAMES, out.unbind(0)))
