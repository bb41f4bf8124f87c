"""CPU tests of the SKI / KISS-GP oracle against vectors produced by the reference's own code
(tests/golden/ski_golden.npz, generator tests/golden/make_golden_ski.py) and against dense linear algebra."""
import os

import numpy as np
import pytest
import torch

from oracle import kernels as ok
from oracle import ski

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "ski_golden.npz"))


@pytest.mark.parametrize("tag,d", [("d1", 1), ("d2", 2), ("d3", 3), ("d3c5", 3)])
@pytest.mark.parametrize("dt", ["f32", "f64"])
def test_interpolation_matches_reference_outputs(tag, d, dt):
    key = f"{tag}_{dt}"
    x = torch.from_numpy(GOLD[f"{key}_x"])
    grid = [torch.from_numpy(GOLD[f"{key}_grid{i}"]) for i in range(d)]
    idx, val = ski.interpolate(grid, x)
    assert idx.dtype == torch.int64 and torch.equal(idx, torch.from_numpy(GOLD[f"{key}_idx"]))      # index work: bit exact
    ref = torch.from_numpy(GOLD[f"{key}_val"])
    assert torch.equal(val, ref) or (val - ref).abs().max().item() <= (1e-6 if dt == "f32" else 1e-14)
    assert val.sum(-1).sub(1).abs().max().item() < (1e-5 if dt == "f32" else 1e-12)                 # partition of unity


def test_interpolation_known_answer_case_of_the_reference_tests():
    # test/utils/test_interpolation.py:26-120: 4 points on an 11^3 grid; first / last index rows quoted from the test
    x = torch.from_numpy(GOLD["ka_x"])
    grid = [torch.linspace(0.0, 1.0, 11) for _ in range(3)]
    idx, val = ski.interpolate(grid, x)
    assert idx[0, :13].tolist() == [146, 147, 148, 149, 157, 158, 159, 160, 168, 169, 170, 171, 179]
    assert idx[3, -12:].tolist() == [1259, 1260, 1261, 1262, 1270, 1271, 1272, 1273, 1281, 1282, 1283, 1284]
    assert torch.equal(idx, torch.from_numpy(GOLD["ka_idx"]))
    assert (val - torch.from_numpy(GOLD["ka_val"])).abs().max().item() < 1e-6
    assert val[0, :4].tolist() == pytest.approx([-0.0002, 0.0022, 0.0022, -0.0002], abs=6e-5)


def test_cubic_interpolation_reproduces_a_quadratic():
    # test/utils/test_interpolation.py:13-24
    x = torch.linspace(0.01, 1, 100).unsqueeze(1)
    grid = [torch.linspace(-0.05, 1.05, 50)]
    idx, val = ski.interpolate(grid, x)
    f = ski.left_interp(idx, val, grid[0].pow(2).unsqueeze(1)).squeeze()
    assert (f - x.squeeze().pow(2)).abs().max().item() < 1e-4


def test_out_of_bounds_is_rejected_like_the_reference():
    with pytest.raises(RuntimeError, match="out of bounds"):
        ski.interpolate([torch.linspace(0, 1, 10)], torch.tensor([[1.2]]))


def test_grid_helpers_match_reference():
    assert ski.choose_grid_size(1000, 3) == int(GOLD["choose_grid_size_1000_3"])
    assert ski.choose_grid_size(10 ** 6, 3) == int(GOLD["choose_grid_size_1e6_3"])       # BASELINE C5: 100^3 grid
    for dt, name in ((torch.float32, "f32"), (torch.float64, "f64")):
        for i, g in enumerate(ski.create_grid([8, 9, 10], [(0.0, 1.0)] * 3, dtype=dt)):
            assert torch.equal(g, torch.from_numpy(GOLD[f"d3_{name}_grid{i}"]))


def test_sparse_products_and_toeplitz_kronecker_against_dense():
    torch.manual_seed(0)
    n, sizes = 40, [7, 6, 9]
    x = torch.rand(n, 3, dtype=torch.float64)
    grid = ski.create_grid(sizes, [(0.0, 1.0)] * 3, dtype=torch.float64)
    idx, val = ski.interpolate(grid, x)
    m = sizes[0] * sizes[1] * sizes[2]
    W = torch.zeros(n, m, dtype=torch.float64)
    W.scatter_add_(1, idx, val)
    v = torch.randn(m, 3, dtype=torch.float64)
    r = torch.randn(n, 3, dtype=torch.float64)
    assert (ski.left_interp(idx, val, v) - W @ v).abs().max().item() < 1e-12
    assert (ski.left_t_interp(idx, val, r, m) - W.T @ r).abs().max().item() < 1e-12
    cols = ski.grid_toeplitz_columns("rbf", grid, [0.3, 0.5, 0.4])
    Ts = [c[(torch.arange(c.numel()).unsqueeze(0) - torch.arange(c.numel()).unsqueeze(1)).abs()] for c in cols]
    assert (ski.toeplitz_matmul(cols[2], v[:9]) - Ts[2] @ v[:9]).abs().max().item() < 1e-12
    Kg = torch.kron(torch.kron(Ts[0], Ts[1]), Ts[2])
    assert (ski.kron_toeplitz_matmul(cols, v) - Kg @ v).abs().max().item() < 1e-11
    # the grid covariance is the product RBF kernel evaluated at the grid nodes (dimension 0 slowest)
    nodes = torch.cartesian_prod(*grid)
    Kdirect = ok.kernel_matrix("rbf", nodes, nodes, torch.tensor([0.3, 0.5, 0.4], dtype=torch.float64), 1.0, True)
    assert (Kg - Kdirect).abs().max().item() < 1e-10
    out = ski.ski_matmul("rbf", x, grid, [0.3, 0.5, 0.4], 1.7, r)
    assert (out - 1.7 * (W @ (Kg @ (W.T @ r)))).abs().max().item() < 1e-10


def test_ski_approximates_the_exact_kernel():
    # the point of SKI: W K_uu W^T ~= K_xx for a smooth kernel and a fine grid (grid_interpolation_kernel.py docstring)
    torch.manual_seed(1)
    x = torch.rand(200, 2, dtype=torch.float64)
    grid = ski.create_grid([40, 40], [(0.0, 1.0)] * 2, dtype=torch.float64)
    v = torch.randn(200, 2, dtype=torch.float64)
    exact = ok.kernel_matrix("rbf", x, x, 0.3, 1.0, True) @ v
    approx = ski.ski_matmul("rbf", x, grid, 0.3, 1.0, v)
    assert ((approx - exact).norm() / exact.norm()).item() < 2e-3


@pytest.mark.parametrize("kind", ["rbf", "matern52"])
def test_ski_hyperparameter_gradients_autograd_vs_finite_differences(kind):
    """The GPU parity test of gp_bilinear_grad on the SKI backend (tests/test_gpu_ski.py) takes its reference from autograd through
    ski_matmul; pin that reference itself: central finite differences of sum(L * (K_ski R)) in the lengthscale and the outputscale."""
    g = torch.Generator().manual_seed(3)
    n, d, sizes = 300, 2, [14, 11]
    x = torch.rand(n, d, generator=g, dtype=torch.float64)
    axes = ski.create_grid(sizes, [(0.0, 1.0)] * d, dtype=torch.float64)
    left = torch.randn(n, 3, generator=g, dtype=torch.float64)
    right = torch.randn(n, 3, generator=g, dtype=torch.float64)

    def f(ls, osc):
        return (left * ski.ski_matmul(kind, x, axes, ls, osc, right)).sum()

    ls = torch.tensor([0.4], dtype=torch.float64, requires_grad=True)
    osc = torch.tensor(1.3, dtype=torch.float64, requires_grad=True)
    f(ls, osc).backward()
    h = 1e-6
    fd_ls = (f(torch.tensor([0.4 + h], dtype=torch.float64), 1.3) - f(torch.tensor([0.4 - h], dtype=torch.float64), 1.3)) / (2 * h)
    fd_os = (f(torch.tensor([0.4], dtype=torch.float64), 1.3 + h) - f(torch.tensor([0.4], dtype=torch.float64), 1.3 - h)) / (2 * h)
    assert ls.grad.item() == pytest.approx(fd_ls.item(), rel=1e-6)
    assert osc.grad.item() == pytest.approx(fd_os.item(), rel=1e-7)


def test_additive_kernel_oracle_mll_matches_dense_cholesky():
    """Reference for tests/test_gpu_sum.py: the oracle's mBCG evaluation on the dense sum K_1 + K_2 (what AdditiveKernel builds,
    kernels/kernel.py:612-621) agrees with dense Cholesky; the preconditioner's pivoted Cholesky runs on rows of the sum."""
    from oracle import mll as om

    n = 600
    x, y = om.synthetic_problem(n, 4, 1, torch.float64)
    K = ok.kernel_matrix("rbf", x[:, :2], x[:, :2], 0.5, 0.9, True) + ok.kernel_matrix("matern52", x, x, 1.3, 0.6, True)
    pn = tuple(a.double() for a in om.make_probe_noise(n, 20, 10, 2))
    res = om.mll_bbmm("rbf", x, y, 0.0, 1.0, 1.5, 0.2, pn, precond_size=20, min_precond_size=100, tolerance=1e-6, K=K)
    Lc = torch.linalg.cholesky(K + 0.2 * torch.eye(n, dtype=torch.float64))
    iq = (y @ torch.cholesky_solve(y.unsqueeze(-1), Lc)).item()
    ld = 2 * Lc.diagonal().log().sum().item()
    assert res.inv_quad == pytest.approx(iq, rel=1e-5)
    assert res.logdet == pytest.approx(ld, rel=0.1)            # 10 probes: stochastic
    assert res.precond is not None and res.precond.L.shape == (n, 20)
