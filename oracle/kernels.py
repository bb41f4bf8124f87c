"""Oracle: covariance functions (test infrastructure only; see oracle/__init__.py).

Every function states the reference lines it follows (paths relative to
/root/reference/gpytorch).
"""
from __future__ import annotations

import math

import torch


def sq_dist(x1: torch.Tensor, x2: torch.Tensor, x1_eq_x2: bool = False) -> torch.Tensor:
    """Pairwise squared distance.  Follows kernels/kernel.py:26-49.

    mean-centre by x1's mean, augmented GEMM [-2 x1, |x1|^2, 1] . [x2, 1, |x2|^2]^T,
    zero the diagonal when x1 is x2, clamp at 0.
    """
    adjustment = x1.mean(-2, keepdim=True)
    x1 = x1 - adjustment
    x1_norm = x1.pow(2).sum(dim=-1, keepdim=True)
    x1_pad = torch.ones_like(x1_norm)
    if x1_eq_x2:
        x2, x2_norm, x2_pad = x1, x1_norm, x1_pad
    else:
        x2 = x2 - adjustment
        x2_norm = x2.pow(2).sum(dim=-1, keepdim=True)
        x2_pad = torch.ones_like(x2_norm)
    x1_ = torch.cat([-2.0 * x1, x1_norm, x1_pad], dim=-1)
    x2_ = torch.cat([x2, x2_pad, x2_norm], dim=-1)
    res = x1_.matmul(x2_.transpose(-2, -1))
    if x1_eq_x2:
        res.diagonal(dim1=-2, dim2=-1).fill_(0)
    return res.clamp_min_(0)


def dist(x1: torch.Tensor, x2: torch.Tensor, x1_eq_x2: bool = False) -> torch.Tensor:
    """Pairwise distance.  Follows kernels/kernel.py:52-60."""
    if not x1_eq_x2:
        return torch.cdist(x1, x2).clamp_min(1e-15)
    res = sq_dist(x1, x2, x1_eq_x2=True)
    return res.clamp_min_(1e-30).sqrt_()


def _same(x1, x2):
    return x1 is x2 or (x1.shape == x2.shape and torch.equal(x1, x2))


def rbf(x1, x2, lengthscale, x1_eq_x2=None):
    """k = exp(-0.5 |(x1-x2)/l|^2).  Follows kernels/rbf_kernel.py:68-85 and
    functions/rbf_covariance.py:14-19.  ``lengthscale`` is a scalar or a [d] (ARD) tensor."""
    if x1_eq_x2 is None:
        x1_eq_x2 = _same(x1, x2)
    ls = torch.as_tensor(lengthscale, dtype=x1.dtype)
    return sq_dist(x1 / ls, x2 / ls, x1_eq_x2).div_(-2.0).exp_()


def matern(x1, x2, lengthscale, nu, x1_eq_x2=None):
    """Matern nu in {0.5,1.5,2.5}.  Follows kernels/matern_kernel.py:85-110 and
    functions/matern_covariance.py:17-47 (centre by x1.mean, scale, dist, poly*exp)."""
    if x1_eq_x2 is None:
        x1_eq_x2 = _same(x1, x2)
    ls = torch.as_tensor(lengthscale, dtype=x1.dtype)
    mean = x1.mean(dim=-2, keepdim=True)
    x1_ = (x1 - mean) / ls
    x2_ = x1_ if x1_eq_x2 else (x2 - mean) / ls
    r = dist(x1_, x2_, x1_eq_x2)
    e = torch.exp(-math.sqrt(2 * nu) * r)
    if nu == 0.5:
        c = 1.0
    elif nu == 1.5:
        c = (math.sqrt(3) * r).add(1)
    elif nu == 2.5:
        c = (math.sqrt(5) * r).add(1).add(5.0 / 3.0 * r**2)
    else:
        raise RuntimeError("nu expected to be 0.5, 1.5, or 2.5")  # matern_kernel.py:80-81
    return c * e


def kernel_matrix(kind, x1, x2, lengthscale, outputscale=1.0, x1_eq_x2=None):
    """ScaleKernel(base)(x1,x2) dense.  kernels/scale_kernel.py:108-118 (K <- s*K)."""
    if kind == "rbf":
        k = rbf(x1, x2, lengthscale, x1_eq_x2)
    elif kind in ("matern12", "matern32", "matern52"):
        nu = {"matern12": 0.5, "matern32": 1.5, "matern52": 2.5}[kind]
        k = matern(x1, x2, lengthscale, nu, x1_eq_x2)
    else:
        raise ValueError(kind)
    return k * outputscale


def kernel_diag(kind, x, outputscale=1.0):
    """diag K(x,x) for a stationary kernel = outputscale (lazy_evaluated_kernel_tensor.py:107-133)."""
    return torch.full(x.shape[:-1], float(outputscale), dtype=x.dtype)


def dk_dlengthscale(kind, x1, x2, lengthscale, x1_eq_x2=None):
    """dK/d(lengthscale) for scalar lengthscale: functions/rbf_covariance.py:20-22 (sq*k/l) and
    functions/matern_covariance.py:27-45."""
    if x1_eq_x2 is None:
        x1_eq_x2 = _same(x1, x2)
    ls = float(lengthscale)
    if kind == "rbf":
        sq = sq_dist(x1 / ls, x2 / ls, x1_eq_x2)
        return sq * torch.exp(-0.5 * sq) / ls
    nu = {"matern12": 0.5, "matern32": 1.5, "matern52": 2.5}[kind]
    mean = x1.mean(dim=-2, keepdim=True)
    x1_ = (x1 - mean) / ls
    x2_ = x1_ if x1_eq_x2 else (x2 - mean) / ls
    rho = dist(x1_, x2_, x1_eq_x2) * math.sqrt(2 * nu)
    e = torch.exp(-rho)
    if nu == 0.5:
        return rho / ls * e
    if nu == 1.5:
        return rho**2 / ls * e
    return (rho + 1) * (rho**2 / 3) * e / ls
