"""mpvae_b200 -- B200-native (sm_100a) implementation of MPVAE's probit-ELBO hot path.

Drop-in for `/root/reference/mpvae.py::compute_loss` / `VAE.forward` (same signatures, same 8-tuple).
The compute path is hand-written CUDA behind the C-ABI in `include/mpvae_b200.h`; there is NO CPU
fallback: calling into it without the built library or without a CUDA device raises.
"""
from . import synth  # noqa: F401  (numpy only)

__all__ = ["synth"]


def __getattr__(name):
    # Lazy: importing the package (e.g. for `synth`) must not require torch or the CUDA library.
    if name in ("compute_loss", "VAE", "ProbitELBO", "probit_elbo"):
        from . import mpvae as _m
        return getattr(_m, name)
    raise AttributeError(name)
