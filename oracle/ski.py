"""CPU restatement of the SKI / KISS-GP operator  K_ski = W (T_0 x T_1 x ... x T_{d-1}) W^T  (SURVEY.md section 8f row 3).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Groundwork for the next row of the scope table: the CUDA path does
not exist yet; this file and tests/test_oracle_ski.py pin the restatement first.

* `cubic_interp_weights`, `interpolate`   -- gpytorch/utils/interpolation.py:15-167 (Keys cubic convolution, 4 points per
  dimension, one-hot snapping in the first / last grid cell).  PINNED: tests/golden/ski_golden.npz holds the outputs of
  the reference's own code (tests/golden/make_golden_ski.py) and the known-answer case of
  test/utils/test_interpolation.py:26-120; indices are compared bit-exactly.
* `create_grid`, `choose_grid_size`      -- gpytorch/utils/grid.py:95-110, 142-180.  PINNED the same way.
* `left_interp`, `left_t_interp`, `toeplitz_matmul`, `kron_toeplitz_matmul` -- linear_operator (absent here, see
  oracle/linalg.py): published algorithms (sparse W products; symmetric Toeplitz product through a circulant embedding
  and FFT; Kronecker product applied one mode at a time).  parity unpinned; cross-checked against dense matrices.
* `grid_toeplitz_columns`, `ski_matmul`  -- gpytorch/kernels/grid_kernel.py:107-177 and
  kernels/grid_interpolation_kernel.py:132-213: per-dimension first columns k(u_i[0], u_i[:]) of the 1-D base kernel
  (`last_dim_is_batch=True`), K_grid = T_0 x ... x T_{d-1} with dimension 0 the slowest index -- the ordering that
  `Interpolation.interpolate` uses for its flat indices (interpolation.py:157-163).
"""
from __future__ import annotations

import math
from functools import reduce
from operator import mul

import torch

from . import kernels as ok

INTERP_OFFSETS = (-2, -1, 0, 1)   # interp_points=range(-2, 2)


def choose_grid_size(num_data: int, num_dim: int, ratio: float = 1.0) -> int:
    """utils/grid.py:95-110 (kronecker_structure=True)."""
    return int(ratio * math.pow(num_data, 1.0 / num_dim))


def create_grid(grid_sizes, grid_bounds, extend=True, dtype=torch.float32):
    """utils/grid.py:142-180: one linspace per dimension, extended by one cell on both sides."""
    axes = []
    for size, (lo, hi) in zip(grid_sizes, grid_bounds):
        step = float(hi - lo) / (size - 2)
        a, b = (lo - step, hi + step) if extend else (lo, hi)
        axes.append(torch.linspace(a, b, size, dtype=dtype))
    return axes


def cubic_interp_weights(s: torch.Tensor) -> torch.Tensor:
    """Keys (1981) cubic convolution kernel with a = -1/2, evaluated in the reference's Horner order
    (utils/interpolation.py:33-43): |s| < 1 -> 1.5|s|^3 - 2.5|s|^2 + 1 ; otherwise -0.5|s|^3 + 2.5|s|^2 - 4|s| + 2."""
    u = s.abs()
    near = ((1.5 * u - 2.5) * u) * u + 1
    far = ((-0.5 * u + 2.5) * u - 4) * u + 2
    inside = 1 - u.floor().clamp(0, 1)
    return near * inside + far * (1 - inside)


def interpolate(grid_axes, x: torch.Tensor, eps: float = 1e-10):
    """Sparse interpolation matrix W [n, prod(G_i)] as (indices [n, 4^d] int64, values [n, 4^d]).
    utils/interpolation.py:45-167.  Column p of the result belongs to the base-4 digit string (c_0 ... c_{d-1}) of p with
    dimension 0 the most significant digit; the flat grid index is row-major with dimension 0 slowest."""
    n, d = x.shape
    assert d == len(grid_axes)
    sizes = [int(g.numel()) for g in grid_axes]
    lo = torch.stack([g.min() for g in grid_axes]).to(x)
    hi = torch.stack([g.max() for g in grid_axes]).to(x)
    if bool(((x.min(0)[0] - lo) < -1e-7).any()) or bool(((x.max(0)[0] - hi) > 1e-7).any()):
        raise RuntimeError("Received data that was out of bounds for the specified grid.")
    npt = len(INTERP_OFFSETS)
    offs = torch.tensor(INTERP_OFFSETS, dtype=grid_axes[0].dtype)
    per_dim = []
    for i, g in enumerate(grid_axes):
        delta = (g[1] - g[0]).clamp_min(eps)
        t = (x[:, i] - g[0]) / delta
        cell = torch.floor(t)
        frac = t - cell
        first = cell - offs.max()                                       # left-most of the 4 points, in index space
        w = cubic_interp_weights(frac.unsqueeze(-1) + offs.flip(0).unsqueeze(0))   # distances +1+f, f, f-1, f-2
        left = first < 0
        if bool(left.any()):                                            # first cell: snap to the nearest of the first 4 nodes
            near = (g[:npt].unsqueeze(0) - x[left, i].unsqueeze(1)).abs().argmin(1)
            w[left] = torch.nn.functional.one_hot(near, npt).to(w)
            first = torch.where(left, torch.zeros_like(first), first)
        right = first > sizes[i] - npt
        if bool(right.any()):                                           # last cell: nearest of the last 4 nodes
            near = (g[-npt:].unsqueeze(0) - x[right, i].unsqueeze(1)).abs().argmin(1)
            w[right] = torch.nn.functional.one_hot(near, npt).to(w)
            first = torch.where(right, torch.full_like(first, sizes[i] - npt), first)
        idx = first.long().unsqueeze(-1) + torch.arange(npt)
        per_dim.append((idx, w))
    indices = torch.zeros(n, 1, dtype=torch.long)
    values = torch.ones(n, 1, dtype=grid_axes[0].dtype)
    for i, (idx, w) in enumerate(per_dim):                              # digit i is more significant than digit i+1
        stride = reduce(mul, sizes[i + 1:], 1)
        indices = (indices.unsqueeze(-1) + (idx * stride).unsqueeze(1)).reshape(n, -1)
        values = (values.unsqueeze(-1) * w.unsqueeze(1)).reshape(n, -1)
    return indices, values


def left_interp(indices, values, rhs):
    """W @ rhs for rhs [M, t]  (linear_operator.utils.interpolation.left_interp)."""
    return (values.unsqueeze(-1) * rhs[indices]).sum(-2)


def left_t_interp(indices, values, rhs, size):
    """W^T @ rhs for rhs [n, t] -> [size, t]  (left_t_interp): scatter-add."""
    out = torch.zeros(size, rhs.size(-1), dtype=rhs.dtype)
    contrib = (values.unsqueeze(-1) * rhs.unsqueeze(1)).reshape(-1, rhs.size(-1))
    out.index_add_(0, indices.reshape(-1), contrib)
    return out


def toeplitz_matmul(col, v):
    """Symmetric Toeplitz T (first column `col` [G]) times v [G, t] through the circulant embedding of size 2G-2 and FFT."""
    g = col.numel()
    if g == 1:
        return col * v
    c = torch.cat([col, col[1:-1].flip(0)])                             # first column of the circulant, length 2G-2
    vp = torch.cat([v, torch.zeros(g - 2, v.size(1), dtype=v.dtype)], 0)
    out = torch.fft.irfft(torch.fft.rfft(c).unsqueeze(-1) * torch.fft.rfft(vp, dim=0), n=2 * g - 2, dim=0)
    return out[:g].to(v.dtype)


def kron_toeplitz_matmul(cols, v):
    """(T_0 x T_1 x ... x T_{d-1}) v with dimension 0 the slowest index of v's rows; v [prod G_i, t]."""
    sizes = [int(c.numel()) for c in cols]
    t = v.size(1)
    cur = v.reshape(*sizes, t)
    for i, col in enumerate(cols):                                      # apply T_i along mode i
        cur = cur.movedim(i, 0)
        shp = cur.shape
        cur = toeplitz_matmul(col, cur.reshape(sizes[i], -1)).reshape(shp).movedim(0, i)
    return cur.reshape(-1, t)


def grid_toeplitz_columns(kind, grid_axes, lengthscale):
    """First columns k_1d(u_i[0], u_i[:]) of the per-dimension Toeplitz factors (grid_kernel.py:138-157: the base kernel
    is evaluated with last_dim_is_batch=True, i.e. on every dimension separately)."""
    ls = torch.as_tensor(lengthscale, dtype=grid_axes[0].dtype).reshape(-1)
    cols = []
    for i, g in enumerate(grid_axes):
        li = ls[0] if ls.numel() == 1 else ls[i]
        cols.append(ok.kernel_matrix(kind, g[:1].unsqueeze(-1), g.unsqueeze(-1), li, 1.0, False).reshape(-1))
    return cols


def ski_matmul(kind, x, grid_axes, lengthscale, outputscale, v):
    """K_ski v = s W (T_0 x ... x T_{d-1}) W^T v   (grid_interpolation_kernel.py:132-213 + scale_kernel.py:108-118)."""
    idx, val = interpolate(grid_axes, x)
    m = reduce(mul, [int(g.numel()) for g in grid_axes], 1)
    cols = grid_toeplitz_columns(kind, grid_axes, lengthscale)
    return outputscale * left_interp(idx, val, kron_toeplitz_matmul(cols, left_t_interp(idx, val, v, m)))
