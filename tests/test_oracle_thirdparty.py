"""Third-party anchors for the oracle (CPU).  oracle/linalg.py restates linear_operator, which is absent ("parity unpinned",
oracle/__init__.py); these checks tie the oracle to implementations that are NOT this repository's restatement:
scikit-learn's GaussianProcessRegressor (exact log marginal likelihood by dense Cholesky, RBF and Matern kernels) and
scipy.sparse.linalg.cg (conjugate gradients)."""
import math
import warnings

import numpy as np
import pytest
import torch

from oracle import kernels as ok, linalg as ol, mll as om


def _sk(kind, ls, osc):
    from sklearn.gaussian_process.kernels import RBF, ConstantKernel, Matern

    base = RBF(ls, "fixed") if kind == "rbf" else Matern(ls, "fixed", nu={"matern12": 0.5, "matern32": 1.5, "matern52": 2.5}[kind])
    return ConstantKernel(osc, "fixed") * base


@pytest.mark.parametrize("kind", ["rbf", "matern12", "matern32", "matern52"])
def test_kernel_matrix_and_exact_mll_match_sklearn(kind):
    from sklearn.gaussian_process import GaussianProcessRegressor

    n, d, ls, osc, nz = 900, 4, 0.8, 1.6, 0.07
    x, y = om.synthetic_problem(n, d, 1, torch.float64)
    sk = _sk(kind, ls, osc)
    K = ok.kernel_matrix(kind, x, x, ls, osc, True)
    assert np.abs(K.numpy() - sk(x.numpy())).max() < 1e-10
    gpr = GaussianProcessRegressor(kernel=sk, alpha=nz, optimizer=None).fit(x.numpy(), y.numpy())
    r = om.mll_cholesky(kind, x, y, 0.0, ls, osc, nz)
    assert r.log_prob == pytest.approx(gpr.log_marginal_likelihood_value_, rel=1e-9)


def test_mbcg_solution_matches_scipy_cg_and_bbmm_mll_matches_sklearn():
    from scipy.sparse.linalg import cg as scipy_cg
    from sklearn.gaussian_process import GaussianProcessRegressor

    n, d, kind, ls, osc, nz = 2500, 5, "rbf", 0.9, 1.2, 0.1
    x, y = om.synthetic_problem(n, d, 2, torch.float64)
    K = ok.kernel_matrix(kind, x, x, ls, osc, True)
    Khat = K + nz * torch.eye(n, dtype=torch.float64)
    # linear_cg (tight tolerance, pivoted-Cholesky preconditioner) == scipy's CG == dense solve
    L, piv = ol.pivoted_cholesky(torch.full((n,), osc, dtype=torch.float64), lambda i: K[i], 40)
    pre = ol.build_preconditioner(L, nz, piv)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        sol = ol.linear_cg(lambda v: Khat @ v, y.unsqueeze(-1), tolerance=1e-10, max_iter=2000, preconditioner=pre.apply)
    ref, info = scipy_cg(Khat.numpy(), y.numpy(), rtol=1e-12, maxiter=10000)
    assert info == 0
    dense = torch.linalg.solve(Khat, y).numpy()
    assert np.linalg.norm(sol[:, 0].numpy() - ref) / np.linalg.norm(ref) < 1e-5      # two Krylov solvers, kappa ~ 1e4
    assert np.linalg.norm(sol[:, 0].numpy() - dense) / np.linalg.norm(dense) < 1e-5
    # the stochastic estimate (default knobs) against sklearn's exact value
    gpr = GaussianProcessRegressor(kernel=_sk(kind, ls, osc), alpha=nz, optimizer=None).fit(x.numpy(), y.numpy())
    pn = tuple(a.double() for a in om.make_probe_noise(n, 40, 10, 3))
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        r = om.mll_bbmm(kind, x, y, 0.0, ls, osc, nz, pn, precond_size=40, tolerance=1e-6, max_tridiag_iter=40)
    exact_ld = float(torch.linalg.slogdet(Khat)[1])
    assert r.inv_quad == pytest.approx(float(y @ torch.from_numpy(dense)), rel=1e-6)
    assert abs(r.logdet - exact_ld) < 0.02 * abs(exact_ld)
    assert abs(r.log_prob - gpr.log_marginal_likelihood_value_) < 0.02 * abs(gpr.log_marginal_likelihood_value_)
