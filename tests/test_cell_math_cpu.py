"""CPU: the kernels' per-cell arithmetic compiled for the host (csrc/probit_math.cuh is __host__ __device__) against the
formulas of mpvae.py:171-190 / 103-123 evaluated by torch in fp64, and the closed-form backward
against autograd.  No GPU, no CUDA toolkit: plain g++."""
import os
import shutil
import subprocess

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "host", "cell_math_host.cpp")


@pytest.fixture(scope="module")
def cell_binary(tmp_path_factory):
    gxx = shutil.which("g++")
    if gxx is None:
        pytest.skip("g++ not available")
    out = str(tmp_path_factory.mktemp("cell") / "cell_math_host")
    # -ffp-contract=off: the faithful arithmetic rounds after every multiply and add, like the __fmul_rn/__fadd_rn device build
    subprocess.check_call([gxx, "-O2", "-std=c++17", "-ffp-contract=off", "-o", out, SRC])
    return out


def run_cells(binary, recs):
    text = f"{len(recs)}\n" + "\n".join(" ".join(repr(float(v)) for v in r) for r in recs) + "\n"
    res = subprocess.run([binary], input=text, capture_output=True, text=True, check=True)
    return np.array([[float(v) for v in ln.split()] for ln in res.stdout.strip().splitlines()], dtype=np.float64)


def truth(recs):
    """fp64 evaluation of the reference's formulas for one cell and autograd of the cell's share of the objective."""
    r = torch.tensor(recs, dtype=torch.float64)
    x = r[:, 0].clone().requires_grad_(True)
    y, cn, cp, cq, gp = r[:, 1], r[:, 2], r[:, 3], r[:, 4], r[:, 5]
    eps = torch.tensor(1e-6, dtype=torch.float32).double()              # fp32(1e-6), mpvae.py:156
    cdf = 0.5 * (1.0 + torch.erf(x / np.sqrt(2.0)))
    E = cdf * (1.0 - eps) + eps * 0.5                                   # mpvae.py:177
    ll = y * torch.log(E) + (1.0 - y) * torch.log(1.0 - E)             # mpvae.py:184
    epos = torch.where(y == 1.0, torch.exp(-5.0 * E), torch.zeros_like(E))
    eneg = torch.where(y == 0.0, torch.exp(5.0 * E), torch.zeros_like(E))
    # d objective / dE = cn * dll/dE + cp * e^{-5E} [y = 1] + cq * e^{5E} [y = 0] + gp   (cell_backward's contract)
    obj = cn * ll + cp * (-0.2) * epos + cq * 0.2 * eneg + gp * E
    (g,) = torch.autograd.grad(obj.sum(), x)
    return E.detach().numpy(), ll.detach().numpy(), epos.detach().numpy(), eneg.detach().numpy(), g.numpy()


def make_records(n, lo, hi, seed):
    rng = np.random.RandomState(seed)
    x = rng.uniform(lo, hi, n)
    y = (rng.uniform(size=n) < 0.4).astype(np.float64)
    return np.stack([x, y, -rng.uniform(0.001, 0.02, n), -rng.uniform(0.0, 0.01, n), rng.uniform(0.0, 0.01, n),
                     rng.standard_normal(n) * 1e-3], axis=1).astype(np.float32).astype(np.float64)


def rel(a, b):
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-30))


def test_cell_forward_and_backward_against_fp64(cell_binary):
    recs = make_records(4000, -3.0, 3.0, seed=1)               # inside the range where fp32 cdf has no cancellation
    got = run_cells(cell_binary, recs)
    E, ll, epos, eneg, g = truth(recs)
    for mode, o in (("faithful", 0), ("stable", 5)):
        assert rel(got[:, o + 0], E) <= 2e-7, mode
        # log E carries the absolute rounding of E when E is close to 1 (the bound is absolute below |ll| = 1); the
        # faithful mode additionally inherits the reference's cancellation in 0.5 (1 + erf): 3e-8 absolute on a cdf of
        # 1.3e-3 at |x| = 3
        tol_ll = 2e-5 if mode == "faithful" else 5e-7
        assert np.max(np.abs(got[:, o + 1] - ll) / np.maximum(np.abs(ll), 1.0)) <= tol_ll, mode
        assert rel(got[:, o + 2], epos) <= 1e-6 and rel(got[:, o + 3], eneg) <= 1e-6, mode
        tol_g = 5e-5 if mode == "faithful" else 2e-6
        assert np.max(np.abs(got[:, o + 4] - g)) <= tol_g * np.max(np.abs(g)), mode


def test_stable_cdf_keeps_relative_accuracy_in_the_tails(cell_binary):
    """Deep tails: the reference's fp32 form quantises cdf to multiples of 2^-25 (faithful mode follows it); the erfc form
    keeps log E and the gradient accurate."""
    recs = np.concatenate([make_records(2000, -5.2, -3.5, seed=2), make_records(2000, 3.5, 5.2, seed=3)])
    got = run_cells(cell_binary, recs)
    E, ll, _, _, g = truth(recs)
    # the cells whose log-likelihood is LARGE: a positive label deep in the lower tail, a negative one in the upper tail
    hard = ((recs[:, 0] < 0) & (recs[:, 1] == 1.0)) | ((recs[:, 0] > 0) & (recs[:, 1] == 0.0))
    assert hard.sum() > 500
    err_f = np.abs(got[hard, 1] - ll[hard]) / np.abs(ll[hard])
    err_s = np.abs(got[hard, 6] - ll[hard]) / np.abs(ll[hard])
    assert err_s.max() <= 1e-6
    assert err_f.max() >= 20 * err_s.max()                      # the reference's own arithmetic is visibly off here
    # the easy cells (log of a number next to 1): absolute accuracy of an fp32 value near 1
    assert np.max(np.abs(got[~hard, 6] - ll[~hard])) <= 2e-7
    gerr_s = np.abs(got[hard, 9] - g[hard]) / np.maximum(np.abs(g[hard]), 1e-12)
    gerr_f = np.abs(got[hard, 4] - g[hard]) / np.maximum(np.abs(g[hard]), 1e-12)
    assert np.median(gerr_s) <= 1e-6 and gerr_s.max() <= 1e-4
    assert np.median(gerr_f) >= 10 * np.median(gerr_s)


def test_soft_labels_follow_the_reference_formula(cell_binary):
    recs = make_records(500, -2.0, 2.0, seed=4)
    recs[:, 1] = np.random.RandomState(5).uniform(0.1, 0.9, 500).astype(np.float32)
    got = run_cells(cell_binary, recs)
    E, ll, epos, eneg, _ = truth(recs)
    assert rel(got[:, 1], ll) <= 2e-6
    assert np.all(got[:, 2] == 0.0) and np.all(got[:, 3] == 0.0)      # a soft label is in neither ranking set


def test_gxs_bound_constants_are_upper_bounds():
    """gxs_bound_kernel (probit_rows.cu) scales the fp16 planes of gxs with the a-priori bound
        |dL/dx| <= 5 |cn| + 17 max(|cp|, |cq|) + 0.4 |gp|
    i.e. phi/E <= 5, phi/(1-E) <= 5, e^{5E} phi <= 17, phi <= 0.4 over all x, for the clamped E of mpvae.py:177.
    Checked on a dense grid in fp64 (the maxima are 4.27, 4.27, 16.4 and 0.399)."""
    x = torch.linspace(-12.0, 12.0, 2_400_001, dtype=torch.float64)
    eps = torch.tensor(1e-6, dtype=torch.float32).double()
    cdf = 0.5 * torch.erfc(-x / np.sqrt(2.0))
    E = cdf * (1.0 - eps) + eps * 0.5
    om = (1.0 - cdf) * (1.0 - eps) + eps * 0.5
    phi = torch.exp(-0.5 * x * x) / np.sqrt(2.0 * np.pi)
    m1, m2 = float((phi / E).max()), float((phi / om).max())
    m3, m4 = float((torch.exp(5.0 * E) * phi).max()), float(phi.max())
    assert 4.0 < m1 < 4.5 and 4.0 < m2 < 4.5 and 16.0 < m3 < 16.8 and m4 < 0.4
    assert max(m1, m2) <= 5.0 / 1.15 and m3 <= 17.0 / 1.03       # the margins the kernel's constants leave
