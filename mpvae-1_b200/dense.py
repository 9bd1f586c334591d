"""The large nn.Linear layers of VAE.forward (mpvae.py:51-84) on the library's tcgen05 engine.

SURVEY.md 8f-N3: at the eurlex shape the MLP around the probit layer spends 2 ms per training step in cuBLAS SIMT
SGEMM (fp32 without TF32, as the reference runs it): the first layers (F+L -> 512, F -> 256, F+D -> 256) and the two
512 -> L heads, forward, input gradient and weight gradient.  The split-precision tensor-core product that computes
noise.R^T (csrc/contract_tc.cu: every fp32 operand as two fp16 pieces, three MMA passes, fp32 accumulation) has the
accuracy of an fp32 SGEMM (rms error 4e-7 at K = 3993 vs 1.1e-6 for SGEMM) at ten times its speed, so the same
kernels serve these layers:

    y  = x . W^T + b       mpvae_tc_gemm_nt           (M = batch, N = out, K = in)
    gx = gy . W            mpvae_tc_gemm_nt on W^T    (skipped when x needs no gradient: the data layers)
    gW = gy^T . x          mpvae_tc_gemm_tn           (reduction over the batch)

Operands are split into planes once and reused (mpvae_tc_split): x for y and gW, gy for gx and gW.  Few output tiles
and a long K (batch 1024 x 256 outputs x 5000 inputs is four 256x256 tiles) would leave most SMs idle, so these
calls allow the K-sliced partial wave.  Small layers stay on cuBLAS.
MPVAE_DENSE=torch switches the whole thing off.
"""
from __future__ import annotations

import ctypes as C
import os

import torch

from . import _lib

MIN_WEIGHT_ELEMS = 1 << 20     # in_features * out_features from which the tensor engine pays for its operand split
MIN_BATCH = 128


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


class Planes:
    """An fp32 matrix [rows][cols] split once into the engine's operand planes; reusable by several products."""

    __slots__ = ("buf", "absmax", "rows", "cols")

    def __init__(self, t: torch.Tensor):
        t = t.contiguous()
        assert t.is_cuda and t.dtype == torch.float32 and t.dim() == 2
        lib = _lib.lib()
        self.rows, self.cols = t.shape
        self.buf = torch.empty(int(lib.mpvae_tc_planes_bytes(self.rows, self.cols)), dtype=torch.uint8, device=t.device)
        self.absmax = torch.empty(1, dtype=torch.int32, device=t.device)
        with torch.cuda.device(t.device):
            stream = C.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)
            _lib.check(lib.mpvae_tc_split(_ptr(t), self.rows, self.cols, _ptr(self.buf), _ptr(self.absmax), stream), "mpvae_tc_split")


def _tail(device):
    return torch.empty(int(_lib.lib().mpvae_tc_tail_scratch_bytes()), dtype=torch.uint8, device=device)


def gemm_nt(a: Planes, b: Planes) -> torch.Tensor:
    """C[M,N] = A[M,K] . B[N,K]^T, K-sliced partial waves allowed."""
    assert a.cols == b.cols
    lib = _lib.lib()
    dev = a.buf.device
    out = torch.empty((a.rows, b.rows), dtype=torch.float32, device=dev)
    tail = _tail(dev)
    with torch.cuda.device(dev):
        stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        _lib.check(lib.mpvae_tc_gemm_nt(_ptr(a.buf), _ptr(b.buf), _ptr(out), a.rows, b.rows, a.cols, b.rows, _ptr(a.absmax),
                                        _ptr(b.absmax), 1, _ptr(tail), tail.numel(), stream), "mpvae_tc_gemm_nt")
    return out


def gemm_tn(a: Planes, b: Planes) -> torch.Tensor:
    """C[N1,N2] = A[M,N1]^T . B[M,N2]."""
    assert a.rows == b.rows
    lib = _lib.lib()
    dev = a.buf.device
    out = torch.empty((a.cols, b.cols), dtype=torch.float32, device=dev)
    tail = _tail(dev)
    with torch.cuda.device(dev):
        stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        _lib.check(lib.mpvae_tc_gemm_tn(_ptr(a.buf), _ptr(b.buf), _ptr(out), a.rows, a.cols, b.cols, _ptr(a.absmax),
                                        _ptr(b.absmax), _ptr(tail), tail.numel(), stream), "mpvae_tc_gemm_tn")
    return out


class TensorLinear(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias):
        xp = Planes(x)
        y = gemm_nt(xp, Planes(weight))
        if bias is not None:
            y += bias
        ctx.save_for_backward(weight)
        ctx.xp = xp if weight.requires_grad else None      # the input planes serve the weight gradient again
        ctx.has_bias = bias is not None
        return y

    @staticmethod
    def backward(ctx, gy):
        (weight,) = ctx.saved_tensors
        gx = gw = gb = None
        gp = Planes(gy)                                     # one split, two products
        if ctx.needs_input_grad[0]:
            gx = gemm_nt(gp, Planes(weight.t()))
        if ctx.needs_input_grad[1]:
            gw = gemm_tn(gp, ctx.xp)
        if ctx.has_bias and ctx.needs_input_grad[2]:
            # column sums as a (1 x B) . (B x out) product: cuBLAS gemv streams gy once at memory speed, where torch's
            # generic dim-0 reduction takes 3-4x as long on a (1024 x 3993) gradient
            gb = torch.mv(gy.t(), gy.new_ones(gy.shape[0]))
        ctx.xp = None
        return gx, gw, gb


def uses_tensor_engine(layer: torch.nn.Linear, x: torch.Tensor) -> bool:
    return (x.is_cuda and x.dim() == 2 and x.dtype == torch.float32 and layer.weight.dtype == torch.float32
            and x.shape[0] >= MIN_BATCH and min(layer.in_features, layer.out_features) >= 128
            and layer.in_features * layer.out_features >= MIN_WEIGHT_ELEMS
            and os.environ.get("MPVAE_DENSE", "tensor") != "torch")


def linear(layer: torch.nn.Linear, x: torch.Tensor) -> torch.Tensor:
    """`layer(x)`, on the tcgen05 engine when the layer is large enough to be a dense-GEMM problem."""
    if uses_tensor_engine(layer, x):
        return TensorLinear.apply(x, layer.weight, layer.bias)
    return layer(x)
