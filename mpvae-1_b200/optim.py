"""clip_grad_norm_ + Adam of the reference's training step (train.py:93,126-128: `optim.Adam(vae.parameters(),
lr, weight_decay=1e-5)`, `clip_grad_norm_(vae.parameters(), 100)`, `optimizer.step()`) as three launches of
libmpvae_b200 (csrc/optim.cu) over flat buffers.

torch runs that tail as ~40 foreach / elementwise kernels per step; the fp64 `r_sqrt_sigma` Parameter alone costs
eight fp64 passes over 16 M elements at the eurlex shape, plus an fp32->fp64 cast of its gradient and an fp64->fp32
cast of the updated parameter for the next forward.  Here every parameter of one dtype lives in ONE flat buffer
(`p.data` become views, so state_dict / checkpoints look exactly as before), the moments likewise, and

    mpvae_grad_norm   : ||g||_2 of the flat gradient bucket -> norm, clip coefficient, bias corrections, lr
    mpvae_adam_step   : fp32 parameters           (p, m, v updated in one pass)
    mpvae_adam_step   : fp64 r_sqrt_sigma          (reads the fp32 g_R, writes the fp32 shadow the next forward uses)

The arithmetic is torch.optim.Adam's (amsgrad=False, maximize=False, L2 weight decay), in each parameter's dtype.
The step count, clip coefficient and learning rate live on the device, so the step is CUDA-graph capturable.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional

import torch

from . import _lib


def _ptr(t: Optional[torch.Tensor]):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


class _Segment:
    """All parameters of one dtype, flattened."""

    def __init__(self, params, dtype, device):
        self.params = params
        self.dtype = dtype
        n = sum(p.numel() for p in params)
        self.flat_p = torch.empty(n, dtype=dtype, device=device)
        self.flat_m = torch.zeros(n, dtype=dtype, device=device)
        self.flat_v = torch.zeros(n, dtype=dtype, device=device)
        self.offsets = []
        off = 0
        for p in params:
            self.offsets.append(off)
            self.flat_p[off:off + p.numel()].copy_(p.detach().reshape(-1))
            p.data = self.flat_p[off:off + p.numel()].view(p.shape)
            off += p.numel()
        self.numel = n
        self.grad_staging = None      # only when the caller's gradients are not already one flat fp32 run

    def views(self, flat):
        return [flat[o:o + p.numel()].view(p.shape) for o, p in zip(self.offsets, self.params)]


class FusedAdam(torch.optim.Optimizer):
    """Drop-in for `torch.optim.Adam(params, lr, betas, eps, weight_decay)` on CUDA parameters (one param group).

    `step()` additionally takes what the data-parallel step knows, so that nothing has to be gathered or cast:
      max_norm     clip_grad_norm_ threshold (None: no clipping; the norm is still computed)
      grad_scale   multiplier applied to every gradient first (1 / world_size when the bucket holds the all-reduced sum)
      flat_grad    ONE fp32 tensor holding every gradient (any order) -- the norm is taken over it
      f32_grads    {fp64 parameter: fp32 gradient tensor}   (the bucket's g_R for r_sqrt_sigma)
      f32_shadows  {fp64 parameter: fp32 tensor}            receives the updated parameter as fp32
    After the call `grad_norm` is a 0-dim device tensor (what clip_grad_norm_ returns)."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0):
        if not 0.0 <= eps or not 0.0 <= betas[0] < 1.0 or not 0.0 <= betas[1] < 1.0:
            raise ValueError("invalid Adam hyper-parameters")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        if len(self.param_groups) != 1:
            raise NotImplementedError("FusedAdam supports the reference's single parameter group")
        self._segments = None
        self._scal = None
        self.shadowed = set()         # fp64 parameters whose fp32 shadow the last step() wrote

    # ------------------------------------------------------------------------------------------------ layout
    def _build(self):
        group = self.param_groups[0]
        params = [p for p in group["params"] if p.requires_grad]
        if not params or not all(p.is_cuda for p in params):
            raise RuntimeError("FusedAdam needs CUDA parameters (no CPU fallback); use torch.optim.Adam on the host")
        dev = params[0].device
        self._device = dev
        self._segments = []
        for dt in (torch.float32, torch.float64):
            ps = [p for p in params if p.dtype == dt]
            if ps:
                self._segments.append(_Segment(ps, dt, dev))
        other = [p for p in params if p.dtype not in (torch.float32, torch.float64)]
        if other:
            raise TypeError(f"FusedAdam: unsupported parameter dtype {other[0].dtype}")
        lib = _lib.lib()
        self._scal = torch.zeros(8, dtype=torch.float64, device=dev)       # step, norm, multiplier, bc1, sqrt(bc2), lr
        self._ws = torch.empty(int(lib.mpvae_grad_norm_workspace()), dtype=torch.uint8, device=dev)
        # torch.optim.Adam's per-parameter state, as views (so state_dict() keeps the reference's layout)
        step_view = self._scal[0]
        for seg in self._segments:
            for p, m, v in zip(seg.params, seg.views(seg.flat_m), seg.views(seg.flat_v)):
                self.state[p] = {"step": step_view, "exp_avg": m, "exp_avg_sq": v}

    def load_state_dict(self, state_dict):
        if self._segments is None:
            self._build()
        keep = {p: dict(s) for p, s in self.state.items()}
        super().load_state_dict(state_dict)
        # copy the loaded moments into the flat buffers and point the state back at the views
        step = None
        for p, mine in keep.items():
            loaded = self.state.get(p, {})
            if "exp_avg" in loaded:
                mine["exp_avg"].copy_(loaded["exp_avg"])
                mine["exp_avg_sq"].copy_(loaded["exp_avg_sq"])
                step = loaded.get("step", step)
            self.state[p] = mine
        if step is not None:
            self._scal[0] = float(step)

    @property
    def grad_norm(self) -> torch.Tensor:
        return self._scal[1].float()

    # ------------------------------------------------------------------------------------------------ the step
    def _segment_grad(self, seg, f32_grads):
        """The segment's gradients as one fp32 run aligned with its flat parameter buffer."""
        grads = []
        for p in seg.params:
            g = f32_grads.get(p) if f32_grads else None
            if g is None:
                g = p.grad
            if g is None:
                raise RuntimeError("FusedAdam: a parameter has no gradient (torch.optim.Adam would skip it; "
                                   "freeze it with requires_grad_(False) before building the optimizer)")
            grads.append(g)
        if all(g.dtype == torch.float32 and g.is_contiguous() for g in grads):
            base = grads[0].data_ptr()
            if all(g.data_ptr() == base + 4 * o for g, o in zip(grads, seg.offsets)):
                g0 = grads[0]                                              # already flat and in order: span the run
                return g0.as_strided((seg.numel,), (1,), g0.storage_offset()), base
        if seg.grad_staging is None:
            seg.grad_staging = torch.empty(seg.numel, dtype=torch.float32, device=self._device)
        for g, v in zip(grads, seg.views(seg.grad_staging)):
            v.copy_(g)
        return seg.grad_staging, seg.grad_staging.data_ptr()

    @torch.no_grad()
    def step(self, closure=None, *, max_norm: Optional[float] = None, grad_scale: float = 1.0,
             flat_grad: Optional[torch.Tensor] = None, f32_grads: Optional[Dict] = None,
             f32_shadows: Optional[Dict] = None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        if self._segments is None:
            self._build()
        lib = _lib.lib()
        group = self.param_groups[0]
        beta1, beta2 = group["betas"]
        lr = group["lr"]
        lr_dev = lr if (isinstance(lr, torch.Tensor) and lr.is_cuda and lr.dtype == torch.float32) else None
        lr_host = 0.0 if lr_dev is not None else float(lr)
        seg_grads = [self._segment_grad(seg, f32_grads) for seg in self._segments]
        if flat_grad is None:
            if len(seg_grads) == 1:
                flat_grad = seg_grads[0][0]
                n_norm = self._segments[0].numel
            else:
                flat_grad = torch.cat([g for g, _ in seg_grads])            # slow path: the caller has no flat bucket
                n_norm = flat_grad.numel()
        else:
            assert flat_grad.dtype == torch.float32 and flat_grad.is_contiguous()
            n_norm = flat_grad.numel()
        with torch.cuda.device(self._device):
            stream = C.c_void_p(torch.cuda.current_stream(self._device).cuda_stream)
            _lib.check(lib.mpvae_grad_norm(C.c_void_p(flat_grad.data_ptr()), n_norm,
                                           float(max_norm) if max_norm is not None else 0.0, float(grad_scale),
                                           _ptr(lr_dev), lr_host, float(beta1), float(beta2), _ptr(self._scal),
                                           _ptr(self._ws), self._ws.numel(), stream), "mpvae_grad_norm")
            self.shadowed = set()
            for seg, (_, gptr) in zip(self._segments, seg_grads):
                shadow = None
                if f32_shadows and len(seg.params) == 1:
                    shadow = f32_shadows.get(seg.params[0])
                    if shadow is not None:
                        assert shadow.dtype == torch.float32 and shadow.is_contiguous() and shadow.numel() == seg.numel
                        self.shadowed.add(seg.params[0])
                _lib.check(lib.mpvae_adam_step(_ptr(seg.flat_p), 1 if seg.dtype == torch.float64 else 0, C.c_void_p(gptr),
                                               _ptr(seg.flat_m), _ptr(seg.flat_v), _ptr(shadow), seg.numel,
                                               _ptr(self._scal), float(beta1), float(beta2), float(group["eps"]),
                                               float(group["weight_decay"]), stream), "mpvae_adam_step")
        return loss
