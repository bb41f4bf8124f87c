"""Kernel sums on the device (run with -m gpu): AdditiveKernel (reference kernels/kernel.py:541-545, :592-621) as ONE engine
operator (gp_plan_set_sum, csrc/sum.cu) -- products, pivoted Cholesky of the sum, the preconditioned MLL -- against the oracle
on the dense sum K_1 + K_2, and the public API (k1 + k2, ScaleKernel over a sum, active_dims per term, gradients of every term's
hyper-parameters vs dense fp64 autograd, prediction)."""
import math
import warnings

import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import kernels as ok, linalg as ol, mll as om  # noqa: E402


def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-300)).item()


@pytest.fixture(scope="module")
def Plan(cuda_dev):
    from gpytorch_b200.engine import Plan as P

    return P


@pytest.mark.parametrize("backends", [("tcgen05", "tcgen05"), ("tcgen05", "simt")])
def test_sum_plan_products_pivots_and_mll_match_oracle(Plan, cuda_dev, backends):
    n, d, rank = 3000, 6, 40
    x, y = om.synthetic_problem(n, d, 2, torch.float32)
    xd = x.to(cuda_dev)
    xa = x[:, :3].contiguous()                  # term A sees three of the six dimensions (active_dims)
    K1 = ok.kernel_matrix("rbf", xa.double(), xa.double(), 0.6, 1.2, True)
    K2 = ok.kernel_matrix("matern52", x.double(), x.double(), 1.5, 0.7, True)
    K = K1 + K2
    pa = Plan(xa.to(cuda_dev), backend=backends[0]).set_hypers("rbf", 0.6, 1.2, 0.0)
    pb = Plan(xd, backend=backends[1]).set_hypers("matern52", 1.5, 0.7, 0.0)
    ps = Plan(xd).set_sum([pa, pb]).set_hypers("rbf", [1.0], 1.0, 0.1)
    assert ps.info()["backend"] == "sum"
    g = torch.Generator().manual_seed(5)
    v = torch.randn(n, 7, generator=g)
    assert rel(ps.kmv(v.to(cuda_dev), add_noise=True), K @ v.double() + 0.1 * v.double()) < 5e-6
    assert rel(ps.kmv(v.to(cuda_dev)), K @ v.double()) < 5e-6
    # pivoted Cholesky evaluates rows of the SUM: pivots bit-identical to the oracle's greedy choice on the dense sum
    lt, piv, st = ps.pivoted_cholesky(rank, 1e-3)
    L, piv_o = ol.pivoted_cholesky(torch.full((n,), 1.9, dtype=torch.float64), lambda i: K[i], rank)
    assert st == 0 and torch.equal(piv.cpu(), piv_o)
    assert rel(lt.t(), L) < 1e-3
    # the preconditioned MLL with identical probe base samples
    pn = om.make_probe_noise(n, rank, 10, 7)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        o64 = om.mll_bbmm("rbf", x.double(), y.double(), 0.0, 1.0, 1.9, 0.1, tuple(a.double() for a in pn), precond_size=rank, K=K)
        o32 = om.mll_bbmm("rbf", x, y, 0.0, 1.0, 1.9, 0.1, pn, precond_size=rank, K=K.float())
    res, sol = ps.mll(y.to(cuda_dev), pn[0].to(cuda_dev), pn[1].to(cuda_dev), pn[2].to(cuda_dev), 10, rank, 2000, want_solve=True)
    assert res.cg_iters == o64.iters and res.precond_rank == rank
    assert abs(res.inv_quad - o64.inv_quad) <= max(1e-4 * abs(o64.inv_quad), 3 * abs(o32.inv_quad - o64.inv_quad))
    assert abs(res.logdet - o64.logdet) <= max(1e-4 * abs(o64.logdet), 3 * abs(o32.logdet - o64.logdet))
    Khat = K + 0.1 * torch.eye(n, dtype=torch.float64)
    Lc = torch.linalg.cholesky(Khat)
    dense_lp = -0.5 * ((y.double() @ torch.cholesky_solve(y.double().unsqueeze(-1), Lc)).item() + 2 * Lc.diagonal().log().sum().item()
                       + n * math.log(2 * math.pi)) / n
    assert abs(res.mll - dense_lp) < 0.02 * abs(dense_lp) + 1e-3
    # re-packing a term (new outputscale / lengthscale) is picked up by the next call on the sum
    pb.set_hypers("matern52", 1.1, 2.0, 0.0)
    K2b = ok.kernel_matrix("matern52", x.double(), x.double(), 1.1, 2.0, True)
    assert rel(ps.kmv(v.to(cuda_dev)), (K1 + K2b) @ v.double()) < 5e-6
    # Lanczos on the sum (LOVE): T = Q^T K_hat Q
    q, t = ps.lanczos(torch.randn(n, generator=g).to(cuda_dev), 12)
    qd = q.double().cpu()
    Kb = K1 + K2b + 0.1 * torch.eye(n, dtype=torch.float64)
    assert rel(t, qd.t() @ Kb @ qd) < 1e-3
    ps.close(); pa.close(); pb.close()


def test_sum_plan_rejects_mismatched_terms(Plan, cuda_dev):
    x = torch.rand(500, 3, device=cuda_dev)
    pa = Plan(x).set_hypers("rbf", 0.6, 1.0, 0.0)
    pb = Plan(x[:400].contiguous()).set_hypers("rbf", 0.6, 1.0, 0.0)
    ps = Plan(x)
    ps.set_sum([pa, pb])
    with pytest.raises(RuntimeError, match="differs from the sum"):
        ps.set_hypers("rbf", [1.0], 1.0, 0.1)
    with pytest.raises(RuntimeError, match="1 to 4 terms"):
        Plan(x).set_sum([])
    for p in (ps, pa, pb):
        p.close()


def test_api_additive_kernel_mll_gradients_and_prediction(cuda_dev):
    import gpytorch_b200 as gp
    from gpytorch_b200 import settings
    from gpytorch_b200.operators import SumKernelLinearOperator

    n, d = 2500, 4
    x, y = om.synthetic_problem(n, d, 1, torch.float32)
    xd, yd = x.to(cuda_dev), y.to(cuda_dev)
    lik = gp.likelihoods.GaussianLikelihood()

    class M(gp.models.ExactGP):
        def __init__(self):
            super().__init__(xd, yd, lik)
            self.mean_module = gp.means.ZeroMean()
            self.covar_module = (gp.kernels.ScaleKernel(gp.kernels.RBFKernel(active_dims=[0, 1]))
                                 + gp.kernels.ScaleKernel(gp.kernels.MaternKernel(nu=2.5)))

        def forward(self, xx):
            return gp.distributions.MultivariateNormal(self.mean_module(xx), self.covar_module(xx))

    model = M().to(cuda_dev)
    lik = lik.to(cuda_dev)
    ka, kb = model.covar_module.kernels
    assert isinstance(model.covar_module, gp.kernels.AdditiveKernel)
    ka.base_kernel.lengthscale = 0.5; ka.outputscale = 0.9
    kb.base_kernel.lengthscale = 1.3; kb.outputscale = 0.6
    lik.noise = 0.2
    op = model.covar_module(xd)
    assert isinstance(op, SumKernelLinearOperator) and len(op.ops) == 2
    # dense pieces of the public operator
    K1 = ok.kernel_matrix("rbf", x[:, :2].double(), x[:, :2].double(), 0.5, 0.9, True)
    K2 = ok.kernel_matrix("matern52", x.double(), x.double(), 1.3, 0.6, True)
    assert rel(op.diagonal(), (K1 + K2).diagonal()) < 1e-6
    assert rel(op[:5].to_dense(), (K1 + K2)[:5]) < 1e-5
    v = torch.randn(n, 3, device=cuda_dev)
    assert rel(op.matmul(v), (K1 + K2) @ v.double().cpu()) < 5e-6
    mll = gp.mlls.ExactMarginalLogLikelihood(lik, model)
    model.train(); lik.train()
    with settings.max_preconditioner_size(50), settings.cg_tolerance(1e-3), settings.num_trace_samples(15), settings.probe_seed(3):
        loss = -mll(model(xd), yd)
        loss.backward()
    # dense fp64 autograd reference of the same objective
    ls1 = torch.tensor(0.5, dtype=torch.float64, requires_grad=True); os1 = torch.tensor(0.9, dtype=torch.float64, requires_grad=True)
    ls2 = torch.tensor(1.3, dtype=torch.float64, requires_grad=True); os2 = torch.tensor(0.6, dtype=torch.float64, requires_grad=True)
    nz = torch.tensor(0.2, dtype=torch.float64, requires_grad=True)
    Kd = (ok.kernel_matrix("rbf", x[:, :2].double(), x[:, :2].double(), ls1, os1, True)
          + ok.kernel_matrix("matern52", x.double(), x.double(), ls2, os2, True) + nz * torch.eye(n, dtype=torch.float64))
    Lc = torch.linalg.cholesky(Kd)
    r = y.double().unsqueeze(-1)
    ref = 0.5 * ((r * torch.cholesky_solve(r, Lc)).sum() + 2 * Lc.diagonal().log().sum() + n * math.log(2 * math.pi)) / n
    ref.backward()
    assert loss.item() == pytest.approx(ref.item(), rel=2e-2)

    def raw_grad(mod, name):   # chain rule through softplus: d raw = d value * sigmoid(raw)
        p = getattr(mod, name)
        return p.grad.item() / torch.sigmoid(p).item()

    # stochastic trace estimate with 15 probes: gradients agree to ~15 %
    assert raw_grad(ka.base_kernel, "raw_lengthscale") == pytest.approx(ls1.grad.item(), rel=0.2, abs=3e-3)
    assert raw_grad(ka, "raw_outputscale") == pytest.approx(os1.grad.item(), rel=0.2, abs=3e-3)
    assert raw_grad(kb.base_kernel, "raw_lengthscale") == pytest.approx(ls2.grad.item(), rel=0.2, abs=3e-3)
    assert raw_grad(kb, "raw_outputscale") == pytest.approx(os2.grad.item(), rel=0.2, abs=3e-3)
    assert raw_grad(lik, "raw_noise") == pytest.approx(nz.grad.item(), rel=0.2, abs=3e-3)
    # prediction through the joint covariance of the sum: mean vs the dense posterior mean
    model.eval(); lik.eval()
    xt = torch.rand(40, d, generator=torch.Generator().manual_seed(9))
    with torch.no_grad(), settings.eval_cg_tolerance(1e-4), settings.skip_posterior_variances(True):
        pred = model(xt.to(cuda_dev))
    Kst = (ok.kernel_matrix("rbf", xt[:, :2].double(), x[:, :2].double(), 0.5, 0.9, False)
           + ok.kernel_matrix("matern52", xt.double(), x.double(), 1.3, 0.6, False))
    mean_ref = Kst @ torch.cholesky_solve(r, Lc.detach()).squeeze(-1)
    assert rel(pred.mean, mean_ref) < 2e-3


def test_api_scale_kernel_over_additive_kernel(cuda_dev):
    import gpytorch_b200 as gp

    n = 1200
    x = torch.rand(n, 3, generator=torch.Generator().manual_seed(4))
    k = gp.kernels.ScaleKernel(gp.kernels.RBFKernel() + gp.kernels.MaternKernel(nu=1.5)).to(cuda_dev)
    k.outputscale = 1.7
    k.base_kernel.kernels[0].lengthscale = 0.4
    k.base_kernel.kernels[1].lengthscale = 0.8
    op = k(x.to(cuda_dev))
    Kd = 1.7 * (ok.kernel_matrix("rbf", x.double(), x.double(), 0.4, 1.0, True) + ok.kernel_matrix("matern32", x.double(), x.double(), 0.8, 1.0, True))
    v = torch.randn(n, 2, device=cuda_dev)
    assert rel(op.matmul(v), Kd @ v.double().cpu()) < 5e-6
    # the gradient of the shared outputscale flows through both terms
    out = op.matmul(v).sum()
    out.backward()
    vs = v.double().cpu()
    ref = ((Kd / 1.7) @ vs).sum().item() * torch.sigmoid(k.raw_outputscale).item()
    assert k.raw_outputscale.grad.item() == pytest.approx(ref, rel=1e-3)
