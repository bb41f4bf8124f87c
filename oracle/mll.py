"""Oracle: exact-GP marginal log likelihood assembly (test infrastructure only).

Follows gpytorch/distributions/multivariate_normal.py:221-252 (log_prob =
-0.5 (inv_quad + logdet + N log 2 pi)), gpytorch/likelihoods/gaussian_likelihood.py:117-121
(K_hat = K + sigma^2 I) and gpytorch/mlls/exact_marginal_log_likelihood.py:54-89 (divide by N).
The default reference path materialises K once (lazy_evaluated_kernel_tensor.py:343-373) and
runs dense K @ V inside CG; ``mll_bbmm`` does the same on CPU.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import torch

from . import kernels, linalg


@dataclass
class MLLResult:
    mll: float
    log_prob: float
    inv_quad: float
    logdet: float
    iters: int = 0
    info: object = None
    precond: object = None
    solves: torch.Tensor | None = None
    t_mat: torch.Tensor | None = None


def make_probe_noise(n, k, tp, seed, dtype=torch.float32):
    """Deterministic base samples shared by the oracle and the CUDA path:
    eps1 [k, tp] and eps2 [n, tp] ~ N(0,1) (z = L eps1 + sigma eps2), rademacher [n, tp]."""
    g = torch.Generator().manual_seed(seed)
    eps1 = torch.randn(max(k, 1), tp, generator=g, dtype=torch.float64).to(dtype)
    eps2 = torch.randn(n, tp, generator=g, dtype=torch.float64).to(dtype)
    rad = (torch.randint(0, 2, (n, tp), generator=g).to(dtype) * 2 - 1)
    return eps1, eps2, rad


def _noise_diag(noise, n, dtype):
    if torch.is_tensor(noise) and noise.numel() > 1:
        return torch.diag(noise.to(dtype).reshape(-1))
    return float(noise) * torch.eye(n, dtype=dtype)


def mll_cholesky(kind, x, y, mean, lengthscale, outputscale, noise):
    """Dense ground truth (the reference's own N <= max_cholesky_size branch)."""
    n = x.size(-2)
    K = kernels.kernel_matrix(kind, x, x, lengthscale, outputscale, True)
    Khat = K + _noise_diag(noise, n, x.dtype)   # scalar sigma^2, or a per-row vector (FixedNoiseGaussianLikelihood)
    Lc = torch.linalg.cholesky(Khat)
    r = (y - mean).unsqueeze(-1)
    sol = torch.cholesky_solve(r, Lc)
    inv_quad = float((r * sol).sum())
    logdet = float(2 * Lc.diagonal().log().sum())
    lp = -0.5 * (inv_quad + logdet + n * math.log(2 * math.pi))
    return MLLResult(mll=lp / n, log_prob=lp, inv_quad=inv_quad, logdet=logdet, solves=sol)


def mll_bbmm(
    kind,
    x,
    y,
    mean,
    lengthscale,
    outputscale,
    noise,
    probe_noise,
    precond_size=linalg.MAX_PRECONDITIONER_SIZE,
    min_precond_size=linalg.MIN_PRECONDITIONING_SIZE,
    tolerance=linalg.CG_TOLERANCE,
    max_iter=linalg.MAX_CG_ITERATIONS,
    max_tridiag_iter=linalg.MAX_LANCZOS_QUADRATURE_ITERATIONS,
    precond_tol=linalg.PRECONDITIONER_TOLERANCE,
    K=None,
):
    """mBCG/SLQ evaluation of the exact MLL, the reference's N > max_cholesky_size branch.

    probe_noise = (eps1 [k,tp], eps2 [n,tp], rademacher [n,tp]) from make_probe_noise.
    """
    n = x.size(-2)
    if K is None:
        K = kernels.kernel_matrix(kind, x, x, lengthscale, outputscale, True)

    per_row = torch.is_tensor(noise) and noise.numel() > 1
    nz = noise.to(x.dtype).reshape(-1) if per_row else noise

    def matmul(v):
        if per_row:
            return K @ v + (nz.unsqueeze(-1) * v if v.dim() > 1 else nz * v)
        return K @ v + nz * v

    eps1, eps2, rad = probe_noise
    precond = None
    if precond_size > 0 and n >= min_precond_size:
        diag = torch.full((n,), float(outputscale), dtype=x.dtype)
        L, piv = linalg.pivoted_cholesky(diag, lambda i: K[i], precond_size, precond_tol)
        if not torch.isnan(L).any():
            precond = linalg.build_preconditioner(L, noise, piv)
    if precond is not None:
        probes = precond.probes(eps1[: precond.L.size(1)], eps2)
    else:
        probes = rad
    r = y - mean
    inv_quad, logdet, info, solves, t_mat = linalg.inv_quad_logdet(
        matmul, n, r, probes, precond, tolerance, max_iter, max_tridiag_iter, return_info=True
    )
    lp = -0.5 * (inv_quad + logdet + n * math.log(2 * math.pi))
    return MLLResult(
        mll=lp / n, log_prob=lp, inv_quad=inv_quad, logdet=logdet, iters=info.iters, info=info,
        precond=precond, solves=solves, t_mat=t_mat,
    )


def synthetic_problem(n, d, seed=0, dtype=torch.float32):
    """BASELINE.md section 2 inputs: X~U[0,1]^{n x d}, y = sin(3 sum_d x) + 0.1 eps."""
    g = torch.Generator().manual_seed(seed)
    x = torch.rand(n, d, generator=g, dtype=torch.float64)
    y = torch.sin(3 * x.sum(-1)) + 0.1 * torch.randn(n, generator=g, dtype=torch.float64)
    return x.to(dtype), y.to(dtype)
