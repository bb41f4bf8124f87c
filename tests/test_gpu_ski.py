"""GPU parity of the SKI / KISS-GP backend (csrc/ski.cu; BASELINE configs[4]) against the pinned CPU
oracle (oracle/ski.py: interpolation checked against outputs of the reference's own code, tests/golden/ski_golden.npz).
Hyper-parameter gradients (gp_bilinear_grad on the SKI backend: < W^T L, dK_uu/dtheta W^T R > on the grid) against fp64 autograd
through the oracle's interpolated product.
Tolerances: products rel-l2 <= 2e-5 (fp32 interpolation weights and mode products, atomics in arbitrary order; the oracle runs in
fp64), MLL by the Krylov rule of tests/test_gpu_configs.py."""
import math
import warnings

import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import linalg as ol, mll as om, ski  # noqa: E402


def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-300)).item()


def _grid(sizes, bounds):
    axes = ski.create_grid(sizes, bounds, dtype=torch.float32)
    return axes, [float(a[0]) for a in axes], [float(a[1] - a[0]) for a in axes]


@pytest.mark.parametrize("d,sizes", [(1, [24]), (2, [20, 16]), (3, [20, 16, 12]), (3, [9, 33, 7])])
@pytest.mark.parametrize("kind", ["rbf", "matern52"])
def test_ski_matmul_matches_oracle(cuda_dev, d, sizes, kind):
    from gpytorch_b200.engine import Plan

    g = torch.Generator().manual_seed(d * 10 + len(kind))
    n, t = 3000, 11
    bounds = [(0.0, 1.0)] * d
    axes, lo, step = _grid(sizes, bounds)
    x = torch.rand(n, d, generator=g)
    # points inside the first / last grid cell (one-hot snapping, interpolation.py:84-131) and exactly on nodes
    x[:20] = torch.tensor(lo) + torch.rand(20, d, generator=g) * torch.tensor(step) * 0.999     # first cell: [lo, lo + spacing)
    x[20:40] = torch.tensor([float(a[-1]) for a in axes]) - torch.rand(20, d, generator=g) * torch.tensor(step) * 0.999
    x[40] = torch.tensor([float(a[3]) for a in axes])
    v = torch.randn(n, t, generator=g)
    ls = 0.3
    ref = ski.ski_matmul(kind, x.double(), [a.double() for a in axes], ls, 1.7, v.double())
    p = Plan(x.to(cuda_dev)).set_ski(sizes, lo, step).set_hypers(kind, ls, 1.7, 0.2)
    assert p.info()["backend"] == "ski"
    out = p.kmv(v.to(cuda_dev))
    assert rel(out, ref) < 2e-5
    assert rel(p.kmv(v.to(cuda_dev), add_noise=True), ref + 0.2 * v.double()) < 2e-5
    # symmetry of the interpolated operator
    u = torch.randn(n, t, generator=g)
    a = (u.double() * out.double().cpu()).sum()
    b = (p.kmv(u.to(cuda_dev)).double().cpu() * v.double()).sum()
    assert abs(a - b) <= 1e-4 * abs(a)
    p.close()


def test_ski_out_of_bounds_is_rejected_like_the_reference(cuda_dev):
    from gpytorch_b200.engine import Plan

    axes, lo, step = _grid([10, 10], [(0.0, 1.0)] * 2)
    x = torch.rand(100, 2)
    x[5, 1] = 1.5
    with pytest.raises(RuntimeError, match="out of bounds"):
        Plan(x.to(cuda_dev)).set_ski([10, 10], lo, step).set_hypers("rbf", 0.3, 1.0, 0.1)


def test_ski_mll_matches_oracle_and_api(cuda_dev):
    """MLL of the interpolated operator (no preconditioner: Rademacher probes) against the oracle's mBCG run on the dense K_ski,
    then the same through ScaleKernel(GridInterpolationKernel(RBFKernel())) + ExactMarginalLogLikelihood."""
    from gpytorch_b200.engine import Plan

    n, d, sizes, ls, osc, nz = 2500, 2, [30, 30], 0.25, 1.4, 0.3
    x, y = om.synthetic_problem(n, d, 5, torch.float32)
    axes, lo, step = _grid(sizes, [(0.0, 1.0)] * d)
    Kd = ski.ski_matmul("rbf", x.double(), [a.double() for a in axes], ls, osc, torch.eye(n, dtype=torch.float64))
    Kd = 0.5 * (Kd + Kd.t())
    pn = om.make_probe_noise(n, 15, 10, 3)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        o64 = om.mll_bbmm("rbf", x.double(), y.double(), 0.0, ls, osc, nz, tuple(a.double() for a in pn), precond_size=0, K=Kd)
        o32 = om.mll_bbmm("rbf", x, y, 0.0, ls, osc, nz, pn, precond_size=0, K=Kd.float())
    p = Plan(x.to(cuda_dev)).set_ski(sizes, lo, step).set_hypers("rbf", ls, osc, nz)
    res, sol = p.mll(y.to(cuda_dev), pn[0].to(cuda_dev), pn[1].to(cuda_dev), pn[2].to(cuda_dev), 10, 15, 2000, want_solve=True)
    assert res.precond_rank == 0 and res.cg_iters == o64.iters
    assert abs(res.inv_quad - o64.inv_quad) <= max(2e-4 * abs(o64.inv_quad), 3 * abs(o32.inv_quad - o64.inv_quad))
    assert abs(res.logdet - o64.logdet) <= max(2e-4 * abs(o64.logdet), 3 * abs(o32.logdet - o64.logdet))
    p.close()
    # public API
    import gpytorch_b200 as gp
    from gpytorch_b200 import settings

    lik = gp.likelihoods.GaussianLikelihood().to(cuda_dev)
    lik.noise = nz

    class M(gp.models.ExactGP):
        def __init__(self):
            super().__init__(x.to(cuda_dev), y.to(cuda_dev), lik)
            self.mean_module = gp.means.ZeroMean()
            self.covar_module = gp.kernels.ScaleKernel(gp.kernels.GridInterpolationKernel(gp.kernels.RBFKernel(), grid_size=30, num_dims=2,
                                                                                       grid_bounds=[(0.0, 1.0)] * 2))

        def forward(self, xx):
            return gp.distributions.MultivariateNormal(self.mean_module(xx), self.covar_module(xx))

    model = M().to(cuda_dev)
    model.covar_module.base_kernel.base_kernel.lengthscale = ls
    model.covar_module.outputscale = osc
    model.train(); lik.train()
    with torch.no_grad(), settings.probe_seed(3), settings.cg_tolerance(1e-3), settings.num_trace_samples(15), settings.max_preconditioner_size(0):
        out = gp.mlls.ExactMarginalLogLikelihood(lik, model)(model(x.to(cuda_dev)), y.to(cuda_dev))
    Lc = torch.linalg.cholesky(Kd + nz * torch.eye(n, dtype=torch.float64))
    alpha = torch.cholesky_solve(y.double().unsqueeze(-1), Lc)[:, 0]
    exact = -0.5 * (float(y.double() @ alpha) + float(2 * Lc.diagonal().log().sum()) + n * math.log(2 * math.pi)) / n
    assert abs(out.item() - exact) < 0.03 * abs(exact) + 2e-3


@pytest.mark.parametrize("kind,d,sizes,ard", [("rbf", 2, [24, 18], False), ("matern52", 3, [14, 12, 10], False), ("matern32", 2, [20, 20], True),
                                               ("matern12", 1, [40], False)])
def test_ski_bilinear_derivative_matches_oracle_autograd(cuda_dev, kind, d, sizes, ard):
    from gpytorch_b200.engine import Plan

    g = torch.Generator().manual_seed(11 + d)
    n, s_cols = 2000, 19                                  # 19 columns: two sweeps of 16
    axes, lo, step = _grid(sizes, [(0.0, 1.0)] * d)
    x = torch.rand(n, d, generator=g)
    left = torch.randn(n, s_cols, generator=g)
    right = torch.randn(n, s_cols, generator=g)
    ls0 = [0.3 + 0.1 * i for i in range(d)] if ard else [0.35]
    ls = torch.tensor(ls0, dtype=torch.float64, requires_grad=True)
    osc = torch.tensor(1.6, dtype=torch.float64, requires_grad=True)
    val = (left.double() * ski.ski_matmul(kind, x.double(), [a.double() for a in axes], ls, osc, right.double())).sum()
    val.backward()
    p = Plan(x.to(cuda_dev)).set_ski(sizes, lo, step).set_hypers(kind, ls0, 1.6, 0.1)
    gl, go = p.bilinear_grad(left.to(cuda_dev), right.to(cuda_dev))
    scale = max(abs(float(v)) for v in ls.grad) + 1e-12
    assert go == pytest.approx(osc.grad.item(), rel=2e-4, abs=2e-4 * abs(val.item()))
    for a, b in zip(gl, ls.grad.tolist()):
        assert a == pytest.approx(b, rel=5e-4, abs=5e-4 * scale)
    p.close()


def test_ski_training_step_through_the_api(cuda_dev):
    """One optimiser-style step on ScaleKernel(GridInterpolationKernel(RBFKernel())): loss and gradients of lengthscale / outputscale /
    noise against dense fp64 autograd on the oracle's K_ski (stochastic trace estimate: ~15 %)."""
    import gpytorch_b200 as gp
    from gpytorch_b200 import settings

    n, d, sizes, ls0, os0, nz0 = 2000, 2, [26, 26], 0.3, 1.2, 0.25
    x, y = om.synthetic_problem(n, d, 6, torch.float32)
    axes, lo, step = _grid(sizes, [(0.0, 1.0)] * d)
    lik = gp.likelihoods.GaussianLikelihood().to(cuda_dev)
    lik.noise = nz0

    class M(gp.models.ExactGP):
        def __init__(self):
            super().__init__(x.to(cuda_dev), y.to(cuda_dev), lik)
            self.mean_module = gp.means.ZeroMean()
            self.covar_module = gp.kernels.ScaleKernel(gp.kernels.GridInterpolationKernel(gp.kernels.RBFKernel(), grid_size=26, num_dims=2,
                                                                                       grid_bounds=[(0.0, 1.0)] * 2))

        def forward(self, xx):
            return gp.distributions.MultivariateNormal(self.mean_module(xx), self.covar_module(xx))

    model = M().to(cuda_dev)
    model.covar_module.base_kernel.base_kernel.lengthscale = ls0
    model.covar_module.outputscale = os0
    model.train(); lik.train()
    with settings.probe_seed(5), settings.cg_tolerance(1e-3), settings.num_trace_samples(15):
        loss = -gp.mlls.ExactMarginalLogLikelihood(lik, model)(model(x.to(cuda_dev)), y.to(cuda_dev))
        loss.backward()
    ls = torch.tensor(ls0, dtype=torch.float64, requires_grad=True)
    osc = torch.tensor(os0, dtype=torch.float64, requires_grad=True)
    nz = torch.tensor(nz0, dtype=torch.float64, requires_grad=True)
    Kd = ski.ski_matmul("rbf", x.double(), [a.double() for a in axes], ls, osc, torch.eye(n, dtype=torch.float64))
    Kd = 0.5 * (Kd + Kd.t()) + nz * torch.eye(n, dtype=torch.float64)
    Lc = torch.linalg.cholesky(Kd)
    r = y.double().unsqueeze(-1)
    ref = 0.5 * ((r * torch.cholesky_solve(r, Lc)).sum() + 2 * Lc.diagonal().log().sum() + n * math.log(2 * math.pi)) / n
    ref.backward()
    assert loss.item() == pytest.approx(ref.item(), rel=3e-2, abs=2e-3)
    k = model.covar_module

    def raw_grad(p):
        return p.grad.item() / torch.sigmoid(p).item()

    assert raw_grad(k.base_kernel.base_kernel.raw_lengthscale) == pytest.approx(ls.grad.item(), rel=0.2, abs=3e-3)
    assert raw_grad(k.raw_outputscale) == pytest.approx(osc.grad.item(), rel=0.2, abs=3e-3)
    assert raw_grad(lik.raw_noise) == pytest.approx(nz.grad.item(), rel=0.2, abs=3e-3)


def test_ski_prediction_through_the_api(cuda_dev):
    """Posterior mean / covariance of a KISS-GP model: the joint train + test operator is sliced (W[r] K_uu W[c]^T) -- against the dense
    posterior built from the oracle's interpolated covariance of the joint point set."""
    import gpytorch_b200 as gp
    from gpytorch_b200 import settings

    n, m, d, sizes, ls0, os0, nz0 = 1500, 30, 2, [24, 24], 0.3, 1.1, 0.2
    x, y = om.synthetic_problem(n, d, 8, torch.float32)
    xt = torch.rand(m, d, generator=torch.Generator().manual_seed(2))
    axes, lo, step = _grid(sizes, [(0.0, 1.0)] * d)
    lik = gp.likelihoods.GaussianLikelihood().to(cuda_dev)
    lik.noise = nz0

    class M(gp.models.ExactGP):
        def __init__(self):
            super().__init__(x.to(cuda_dev), y.to(cuda_dev), lik)
            self.mean_module = gp.means.ZeroMean()
            self.covar_module = gp.kernels.ScaleKernel(gp.kernels.GridInterpolationKernel(gp.kernels.RBFKernel(), grid_size=24, num_dims=2,
                                                                                       grid_bounds=[(0.0, 1.0)] * 2))

        def forward(self, xx):
            return gp.distributions.MultivariateNormal(self.mean_module(xx), self.covar_module(xx))

    model = M().to(cuda_dev)
    model.covar_module.base_kernel.base_kernel.lengthscale = ls0
    model.covar_module.outputscale = os0
    model.eval(); lik.eval()
    with torch.no_grad(), settings.eval_cg_tolerance(1e-5):
        pred = model(xt.to(cuda_dev))
    xj = torch.cat([x, xt]).double()
    Kj = ski.ski_matmul("rbf", xj, [a.double() for a in axes], ls0, os0, torch.eye(n + m, dtype=torch.float64))
    Kj = 0.5 * (Kj + Kj.t())
    Lc = torch.linalg.cholesky(Kj[:n, :n] + nz0 * torch.eye(n, dtype=torch.float64))
    mean_ref = Kj[n:, :n] @ torch.cholesky_solve(y.double().unsqueeze(-1), Lc).squeeze(-1)
    cov_ref = Kj[n:, n:] - Kj[n:, :n] @ torch.cholesky_solve(Kj[:n, n:], Lc)
    assert rel(pred.mean, mean_ref) < 2e-3
    assert rel(pred.covariance_matrix, cov_ref) < 2e-2
