"""ORACLE -- TEST INFRASTRUCTURE ONLY.  Not part of the product path.

CPU restatement of the multivariate-probit ELBO of lliutianc/MPVAE-1
(`/root/reference/mpvae.py:103-135` ranking loss, `:145-210` compute_loss).
Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline /
`--impl reference` legs may import this file; the shipped package
(`mpvae-1_b200/`) never does and has no CPU fallback.

Pinning status: the reference ships NO tests / golden vectors for this path
(SURVEY.md section 4), so this restatement is pinned against outputs of the
reference itself, produced in the build container by importing the unmodified
`/root/reference/mpvae.py` (script: `tests/golden/make_golden.py`, fixtures:
`tests/golden/*.npz`; checked by `tests/test_oracle_golden.py`).

The arithmetic lives in third-party torch (unpinned by the reference; the
fixtures record torch 2.11.0+cu128).  Everything here is stated with torch
tensor ops in the reference's op order so that, on a given device, rounding
follows the same ATen kernels the reference would hit:

  * `probit_elbo(...)`            faithful fp32 restatement, O(S*B*L^2) pairwise
                                  ranking loss exactly as mpvae.py:103-123
  * `probit_elbo(..., ranking="factorised")`
                                  the exact algebraic factorisation
                                  sum_{i in pos, j in neg} exp(-5(E_i-E_j))
                                    = (sum_pos exp(-5E_i)) * (sum_neg exp(5E_j))
                                  used for label sets where (S,B,L,L) does not fit
  * `probit_elbo(..., accum=torch.float64)`
                                  same fp32 cell arithmetic (erf/log/exp in fp32)
                                  but every reduction accumulated in fp64: the
                                  "exact-sum" truth used to bound summation-order
                                  noise at large L
  * `probit_elbo(..., contract=torch.float64)`
                                  the product noise.R^T accumulated in fp64 and rounded once (the exact
                                  contraction; bounds the fp32 SGEMM's own rounding at Z >= 983)
  * `probit_elbo(..., dtype=torch.float64)`
                                  everything in fp64 (error budgeting)

The only deliberate difference from the reference is that the Gaussian noise
is an ARGUMENT (the reference draws it at mpvae.py:162 from the CPU default
generator); `reference_noise(seed, S, B, Z)` reproduces that exact draw.
"""
from __future__ import annotations

import math
from types import SimpleNamespace
from typing import NamedTuple, Optional

import torch


class ElboTerms(NamedTuple):
    """Same order as the 8-tuple returned at mpvae.py:210."""

    total_loss: torch.Tensor
    nll_loss: torch.Tensor
    nll_loss_x: torch.Tensor
    c_loss: torch.Tensor
    c_loss_x: torch.Tensor
    kl_loss: torch.Tensor
    indiv_prob: torch.Tensor
    indiv_prob_label: torch.Tensor


def reference_noise(seed: int, n_sample: int, n_batch: int, z_dim: int) -> torch.Tensor:
    """The tensor mpvae.py:162 would draw if `torch.manual_seed(seed)` ran just before
    compute_loss: `torch.normal(0, 1, size=(S, B, Z))` on the CPU default generator."""
    gen_state = torch.get_rng_state()
    try:
        torch.manual_seed(seed)
        return torch.normal(0, 1, size=(n_sample, n_batch, z_dim))
    finally:
        torch.set_rng_state(gen_state)


def gaussian_kl(fe_mu, fe_logvar, fx_mu, fx_logvar, accum=None):
    """mpvae.py:147-148 -- analytic KL between label- and feature-encoder posteriors."""
    per_dim = (fx_logvar - fe_logvar) - 1 + torch.exp(fe_logvar - fx_logvar) \
        + torch.square(fx_mu - fe_mu) / (torch.exp(fx_logvar) + 1e-6)
    if accum is not None:
        per_dim = per_dim.to(accum)
    return torch.mean(0.5 * torch.sum(per_dim, dim=1))


def standard_normal_cdf(x):
    """torch.distributions.Normal(0, 1).cdf as written in torch/distributions/normal.py:104-110:
    0.5 * (1 + erf((value - loc) * scale.reciprocal() / sqrt(2))) with loc=0, scale=1.
    (x - 0) and (* 1) are exact, so only the division by the python scalar sqrt(2) rounds:
    a true fp32 division on CPU, a multiply by the fp32 reciprocal on CUDA.)"""
    return 0.5 * (1 + torch.erf(x / math.sqrt(2)))


def clamped_probit(x, eps1):
    """mpvae.py:177,180 -- E = cdf * (1 - eps1) + eps1 * 0.5 (mul then add, eps1 a 1-elem tensor)."""
    return standard_normal_cdf(x) * (1 - eps1) + eps1 * 0.5


def ranking_loss_pairwise(E, y, accum=None):
    """mpvae.py:103-123 (+ pairwise_and :126-129, pairwise_sub :132-135), op for op.

    E: (S, B, L) probabilities, y: (B, L) labels.  Materialises (S, B, L, L)."""
    y = y.float() if E.dtype == torch.float32 else y.to(E.dtype)
    is_pos = torch.eq(y, torch.ones_like(y))
    is_neg = torch.eq(y, torch.zeros_like(y))
    truth = torch.logical_and(is_pos.unsqueeze(2), is_neg.unsqueeze(1)).to(E.dtype)   # (B, L, L)
    diff = E.unsqueeze(3) - E.unsqueeze(2)                                               # (S, B, L, L)
    masked = torch.exp(-5 * diff) * truth
    if accum is not None:
        masked = masked.to(accum)
    sums = torch.sum(masked, dim=[2, 3])                                                 # (S, B)
    n_pos = torch.sum(is_pos.to(sums.dtype), dim=1)
    n_neg = torch.sum(is_neg.to(sums.dtype), dim=1)
    per_row = torch.div(sums, 5 * (n_pos * n_neg))
    bad = torch.logical_or(torch.isinf(per_row), torch.isnan(per_row))
    per_row = torch.where(bad, torch.zeros_like(per_row), per_row)
    return torch.mean(per_row)


def ranking_loss_factorised(E, y, accum=None):
    """Exact O(S*B*L) factorisation of `ranking_loss_pairwise` (SURVEY.md section 7 hard part 2)."""
    y = y.float() if E.dtype == torch.float32 else y.to(E.dtype)
    is_pos = torch.eq(y, torch.ones_like(y)).to(E.dtype)
    is_neg = torch.eq(y, torch.zeros_like(y)).to(E.dtype)
    e_pos = torch.exp(-5 * E) * is_pos
    e_neg = torch.exp(5 * E) * is_neg
    if accum is not None:
        e_pos, e_neg = e_pos.to(accum), e_neg.to(accum)
    sums = torch.sum(e_pos, dim=2) * torch.sum(e_neg, dim=2)                             # (S, B)
    n_pos = torch.sum(is_pos.to(sums.dtype), dim=1)
    n_neg = torch.sum(is_neg.to(sums.dtype), dim=1)
    per_row = torch.div(sums, 5 * (n_pos * n_neg))
    bad = torch.logical_or(torch.isinf(per_row), torch.isnan(per_row))
    per_row = torch.where(bad, torch.zeros_like(per_row), per_row)
    return torch.mean(per_row)


def bernoulli_log_mean_exp(E, y, accum=None):
    """mpvae.py:182-190 -- per-sample Bernoulli log-likelihood and the log-mean-exp over samples."""
    cell = -(torch.log(E) * y + torch.log(1 - E) * (1 - y))
    if accum is not None:
        cell = cell.to(accum)
    logprob = -torch.sum(cell, dim=2)                                                    # (S, B)
    peak = torch.max(logprob, dim=0)[0]
    mean_exp = torch.mean(torch.exp(logprob - peak), dim=0)
    return torch.mean(-torch.log(mean_exp) - peak)


def probit_elbo(input_label, fe_out, fe_mu, fe_logvar, fx_out, fx_mu, fx_logvar, r_sqrt_sigma,
                noise, nll_coeff, c_coeff, *, ranking: str = "pairwise",
                dtype: torch.dtype = torch.float32,
                accum: Optional[torch.dtype] = None,
                contract: Optional[torch.dtype] = None) -> ElboTerms:
    """Restatement of compute_loss (mpvae.py:145-210) with the noise of :162 as an argument.

    The dead code at mpvae.py:150-153 (sigma / covariance, never used) is not reproduced.
    `noise` is (S, B, Z); S plays the role of args.n_train_sample / n_test_sample (:158)."""
    dev = input_label.device
    y = input_label.to(dtype)
    fe_out, fx_out = fe_out.to(dtype), fx_out.to(dtype)
    kl = gaussian_kl(fe_mu.to(dtype), fe_logvar.to(dtype), fx_mu.to(dtype), fx_logvar.to(dtype), accum)

    eps1 = torch.tensor([1e-6]).float().to(dev).to(dtype)           # :156
    noise = noise.to(dev).to(dtype)
    basis = r_sqrt_sigma.T.float().to(dev).to(dtype) if dtype == torch.float32 \
        else r_sqrt_sigma.T.to(dev).to(dtype)                        # :165
    if contract is not None:
        # exact-contraction truth: noise.R^T accumulated in `contract` (fp64), rounded once to the working dtype
        nr = torch.tensordot(noise.to(contract), basis.to(contract), dims=1).to(dtype)
        sample_r, sample_r_x = nr + fe_out, nr + fx_out
    else:
        sample_r = torch.tensordot(noise, basis, dims=1) + fe_out        # :168
        sample_r_x = torch.tensordot(noise, basis, dims=1) + fx_out      # :170
    E = clamped_probit(sample_r, eps1)                               # :177
    E_x = clamped_probit(sample_r_x, eps1)                           # :180

    rank = ranking_loss_pairwise if ranking == "pairwise" else ranking_loss_factorised
    nll = bernoulli_log_mean_exp(E, y, accum)
    c = rank(E, y, accum)
    nll_x = bernoulli_log_mean_exp(E_x, y, accum)
    c_x = rank(E_x, y, accum)

    if accum is not None:
        indiv_prob = torch.mean(E_x.to(accum), dim=0)                # :203
        indiv_prob_label = torch.mean(E.to(accum), dim=0)            # :204
    else:
        indiv_prob = torch.mean(E_x, dim=0)
        indiv_prob_label = torch.mean(E, dim=0)

    total = (nll + nll_x) * nll_coeff + (c + c_x) * c_coeff + kl * 1.1   # :207-208
    return ElboTerms(total, nll, nll_x, c, c_x, kl, indiv_prob, indiv_prob_label)


def compute_loss(input_label, fe_out, fe_mu, fe_logvar, fx_out, fx_mu, fx_logvar, r_sqrt_sigma, args,
                 noise=None, **kw) -> ElboTerms:
    """Reference-signature wrapper (mpvae.py:145).  If `noise` is None it is drawn exactly as at
    mpvae.py:162 (CPU default generator), so seeding before the call reproduces the reference."""
    n_sample = args.n_train_sample if args.mode == "train" else args.n_test_sample   # :158
    if noise is None:
        noise = torch.normal(0, 1, size=(n_sample, fe_out.shape[0], args.z_dim))
    return probit_elbo(input_label, fe_out, fe_mu, fe_logvar, fx_out, fx_mu, fx_logvar, r_sqrt_sigma,
                       noise, args.nll_coeff, args.c_coeff, **kw)


def probit_elbo_with_grads(inputs: dict, noise, nll_coeff, c_coeff, *, upstream=None, **kw):
    """Run `probit_elbo` under autograd and return (terms, grads).

    `inputs` holds y, fe_out, fe_mu, fe_logvar, fx_out, fx_mu, fx_logvar, r_sqrt_sigma.
    `upstream` optionally maps output names (total_loss, nll_loss, ..., indiv_prob,
    indiv_prob_label) to cotangents; default is d total_loss = 1 (train.py:125).
    Row-chunking for big L is the caller's job (all terms are batch means)."""
    names = ["fe_out", "fe_mu", "fe_logvar", "fx_out", "fx_mu", "fx_logvar", "r_sqrt_sigma"]
    leaves = {k: inputs[k].detach().clone().requires_grad_(True) for k in names}
    terms = probit_elbo(inputs["y"], leaves["fe_out"], leaves["fe_mu"], leaves["fe_logvar"],
                        leaves["fx_out"], leaves["fx_mu"], leaves["fx_logvar"], leaves["r_sqrt_sigma"],
                        noise, nll_coeff, c_coeff, **kw)
    if upstream is None:
        objective = terms.total_loss
    else:
        objective = sum((getattr(terms, k) * v.to(getattr(terms, k).dtype)).sum() for k, v in upstream.items())
    grads = torch.autograd.grad(objective, [leaves[k] for k in names], allow_unused=True)
    return terms, {k: g for k, g in zip(names, grads)}


def chunked_probit_elbo_with_grads(inputs: dict, noise, nll_coeff, c_coeff, rows_per_chunk: int, **kw):
    """Large-L helper: run the oracle on row micro-batches and recombine.
    Every loss term is a mean over rows (mpvae.py:147,122,190), so term = sum_c (B_c/B) term_c and
    gradients add with the same weights.  d total_loss = 1 only."""
    y = inputs["y"]
    n_rows = y.shape[0]
    acc_terms = None
    acc_grads = None
    preds, preds_label = [], []
    row_keys = ["y", "fe_out", "fe_mu", "fe_logvar", "fx_out", "fx_mu", "fx_logvar"]
    for lo in range(0, n_rows, rows_per_chunk):
        hi = min(n_rows, lo + rows_per_chunk)
        part = {k: inputs[k][lo:hi] for k in row_keys}
        part["r_sqrt_sigma"] = inputs["r_sqrt_sigma"]
        terms, grads = probit_elbo_with_grads(part, noise[:, lo:hi], nll_coeff, c_coeff, **kw)
        wgt = (hi - lo) / n_rows
        scal = [t.detach().double() * wgt for t in terms[:6]]
        acc_terms = scal if acc_terms is None else [a + s for a, s in zip(acc_terms, scal)]
        preds.append(terms.indiv_prob.detach())
        preds_label.append(terms.indiv_prob_label.detach())
        if acc_grads is None:
            acc_grads = {k: [] for k in grads}
            acc_grads["r_sqrt_sigma"] = torch.zeros_like(grads["r_sqrt_sigma"], dtype=torch.float64)
        for k, g in grads.items():
            if k == "r_sqrt_sigma":
                acc_grads[k] += g.double() * wgt
            else:
                acc_grads[k].append(g * wgt)
    for k in list(acc_grads):
        if k != "r_sqrt_sigma":
            acc_grads[k] = torch.cat(acc_grads[k], dim=0)
    out = ElboTerms(*acc_terms, torch.cat(preds, 0), torch.cat(preds_label, 0))
    return out, acc_grads


def make_args(label_dim, z_dim, n_train_sample=10, n_test_sample=100, mode="train",
              nll_coeff=0.5, c_coeff=10.0, **extra):
    """The fields compute_loss reads from `args` (mpvae.py:153,158,162,207-208)."""
    return SimpleNamespace(label_dim=label_dim, z_dim=z_dim, n_train_sample=n_train_sample,
                           n_test_sample=n_test_sample, mode=mode, nll_coeff=nll_coeff,
                           c_coeff=c_coeff, **extra)
