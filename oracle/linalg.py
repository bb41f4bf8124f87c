"""Oracle: the linear-algebra half of the BBMM path (test infrastructure only).

Restates the published algorithms of ``linear_operator>=0.6.1`` (pinned at
/root/reference/setup.py:44; source NOT under /root/reference, so there are no
golden vectors: **parity unpinned**, see oracle/__init__.py).  Reference call
sites that anchor the signatures and semantics:

* ``linear_cg``            gpytorch/variational/ciq_variational_strategy.py:56-64,
                           patched at test/lazy/test_lazy_evaluated_kernel_tensor.py:82-83
* ``inv_quad_logdet``      gpytorch/distributions/multivariate_normal.py:248-249,
                           gpytorch/__init__.py:118-145
* ``pivoted_cholesky``     gpytorch/__init__.py:146-173
* ``lanczos_tridiag``      via root_inv_decomposition, models/exact_prediction_strategies.py:268-272
* settings defaults        gpytorch/settings.py:6-31 (re-exports), SURVEY.md Appendix A.1

All routines are single-problem (no leading batch dims) and dtype generic.
"""
from __future__ import annotations

import math
import warnings
from dataclasses import dataclass, field

import torch

# defaults (linear_operator.settings; names attested at gpytorch/settings.py:6-31)
CG_TOLERANCE = 1.0
EVAL_CG_TOLERANCE = 0.01  # gpytorch/settings.py:173-180
MAX_CG_ITERATIONS = 1000
MAX_LANCZOS_QUADRATURE_ITERATIONS = 20
MAX_PRECONDITIONER_SIZE = 15
MIN_PRECONDITIONING_SIZE = 2000
NUM_TRACE_SAMPLES = 10
PRECONDITIONER_TOLERANCE = 1e-3
MAX_CHOLESKY_SIZE = 800


@dataclass
class CGInfo:
    iters: int = 0
    tolerance_reached: bool = False
    residual_norms: torch.Tensor | None = None
    alphas: list = field(default_factory=list)
    betas: list = field(default_factory=list)


def linear_cg(
    matmul_closure,
    rhs,
    n_tridiag=0,
    tolerance=None,
    eps=1e-10,
    stop_updating_after=1e-10,
    max_iter=None,
    max_tridiag_iter=None,
    initial_guess=None,
    preconditioner=None,
    return_info=False,
):
    """Modified batched preconditioned CG (mBCG), linear_operator.utils.linear_cg.

    rhs [N, t].  Returns solves [N, t] (and t_mat [n_tridiag, J, J] when n_tridiag>0).
    """
    is_vector = rhs.dim() == 1
    if is_vector:
        rhs = rhs.unsqueeze(-1)
    if max_iter is None:
        max_iter = MAX_CG_ITERATIONS
    if max_tridiag_iter is None:
        max_tridiag_iter = MAX_LANCZOS_QUADRATURE_ITERATIONS
    if initial_guess is None:
        initial_guess = torch.zeros_like(rhs)
    if tolerance is None:
        tolerance = CG_TOLERANCE
    if max_tridiag_iter > max_iter:
        raise RuntimeError("Getting a tridiagonalization larger than the number of CG iterations run is not possible!")
    num_rows = rhs.size(-2)
    n_iter = max_iter  # terminate_cg_by_size defaults off
    n_tridiag_iter = min(max_tridiag_iter, num_rows)
    t = rhs.size(-1)

    rhs_norm = rhs.norm(2, dim=-2, keepdim=True)
    rhs_is_zero = rhs_norm.lt(eps)
    rhs_norm = rhs_norm.masked_fill(rhs_is_zero, 1)
    rhs = rhs / rhs_norm

    residual = rhs - matmul_closure(initial_guess)
    result = initial_guess.clone()
    if not torch.equal(residual, residual):
        raise RuntimeError("NaNs encountered when trying to perform matrix-vector multiplication")

    residual_norm = residual.norm(2, dim=-2, keepdim=True)
    has_converged = residual_norm < stop_updating_after
    info = CGInfo()

    if has_converged.all() and not n_tridiag:
        n_iter = 0
    else:
        precond_residual = preconditioner(residual) if preconditioner is not None else residual.clone()
        curr_conjugate_vec = precond_residual.clone()
        residual_inner_prod = (precond_residual * residual).sum(-2, keepdim=True)

    if n_tridiag:
        t_mat = torch.zeros(n_tridiag_iter, n_tridiag_iter, n_tridiag, dtype=rhs.dtype)
        prev_alpha_reciprocal = torch.empty(n_tridiag, dtype=rhs.dtype)
        prev_beta = torch.empty(n_tridiag, dtype=rhs.dtype)

    update_tridiag = True
    last_tridiag_iter = 0
    tolerance_reached = False
    k = -1
    for k in range(n_iter):
        mvms = matmul_closure(curr_conjugate_vec)
        alpha = (curr_conjugate_vec * mvms).sum(-2, keepdim=True)
        is_zero = alpha < eps
        alpha = alpha.masked_fill(is_zero, 1)
        alpha = residual_inner_prod / alpha
        alpha = alpha.masked_fill(is_zero, 0)
        alpha = alpha.masked_fill(has_converged, 0)

        residual = residual - alpha * mvms
        precond_residual = preconditioner(residual) if preconditioner is not None else residual.clone()

        result = result + alpha * curr_conjugate_vec
        beta = residual_inner_prod.clone()
        residual_inner_prod = (residual * precond_residual).sum(-2, keepdim=True)
        is_zero = beta < eps
        beta = beta.masked_fill(is_zero, 1)
        beta = residual_inner_prod / beta
        beta = beta.masked_fill(is_zero, 0)
        curr_conjugate_vec = curr_conjugate_vec * beta + precond_residual

        residual_norm = residual.norm(2, dim=-2, keepdim=True)
        residual_norm = residual_norm.masked_fill(rhs_is_zero, 0)
        has_converged = residual_norm < stop_updating_after
        info.alphas.append(alpha.reshape(-1).clone())
        info.betas.append(beta.reshape(-1).clone())

        if (
            k >= min(10, max_iter - 1)
            and bool(residual_norm.mean() < tolerance)
            and not (n_tridiag and k < min(n_tridiag_iter, max_iter - 1))
        ):
            tolerance_reached = True
            break

        if n_tridiag and k < n_tridiag_iter and update_tridiag:
            alpha_tridiag = alpha.reshape(-1)[:n_tridiag].clone()
            beta_tridiag = beta.reshape(-1)[:n_tridiag].clone()
            alpha_is_zero = alpha_tridiag == 0
            alpha_reciprocal = 1.0 / alpha_tridiag.masked_fill(alpha_is_zero, 1)
            if k == 0:
                t_mat[k, k] = alpha_reciprocal
            else:
                t_mat[k, k] = alpha_reciprocal + prev_beta * prev_alpha_reciprocal
                off = prev_beta.sqrt() * prev_alpha_reciprocal
                t_mat[k, k - 1] = off
                t_mat[k - 1, k] = off
                if t_mat[k - 1, k].max() < 1e-6:
                    update_tridiag = False
            last_tridiag_iter = k
            prev_alpha_reciprocal = alpha_reciprocal.clone()
            prev_beta = beta_tridiag.clone()

    result = result * rhs_norm
    info.iters = k + 1 if n_iter > 0 else 0
    info.tolerance_reached = tolerance_reached
    info.residual_norms = residual_norm.reshape(-1).clone()
    if not tolerance_reached and n_iter > 0:
        warnings.warn(
            f"CG terminated in {k + 1} iterations with average residual norm {residual_norm.mean().item()}"
            f" which is larger than the tolerance of {tolerance}",
            RuntimeWarning,
        )
    if is_vector:
        result = result.squeeze(-1)
    out = [result]
    if n_tridiag:
        t_mat = t_mat[: last_tridiag_iter + 1, : last_tridiag_iter + 1]
        out.append(t_mat.permute(2, 0, 1).contiguous())
    if return_info:
        out.append(info)
    return out[0] if len(out) == 1 else tuple(out)


def pivoted_cholesky(diag, get_row, rank, error_tol=PRECONDITIONER_TOLERANCE):
    """Greedy pivoted partial Cholesky, linear_operator.functions._pivoted_cholesky.

    diag: [N] diagonal of K (a copy is modified).  get_row(i) -> K[i, :] [N].
    Returns (L [N, m], pivots [m] int64) with K ~= L L^T.
    """
    matrix_diag = diag.clone()
    n = matrix_diag.numel()
    max_iter = min(rank, n)
    L = torch.zeros(max_iter, n, dtype=diag.dtype)
    orig_error = matrix_diag.max()
    errors = matrix_diag.abs().sum() / orig_error
    permutation = torch.arange(n, dtype=torch.long)
    pivots = []
    m = 0
    while m == 0 or (m < max_iter and errors > error_tol):
        permuted_diags = matrix_diag[permutation[m:]]
        max_val, max_idx = torch.max(permuted_diags, -1)
        max_idx = int(max_idx) + m
        old_pi_m = int(permutation[m])
        permutation[m] = permutation[max_idx]
        permutation[max_idx] = old_pi_m
        pi_m = int(permutation[m])
        pivots.append(pi_m)
        L_m = L[m]
        L_m[pi_m] = max_val.sqrt()
        row = get_row(pi_m)
        if m + 1 < n:
            pi_i = permutation[m + 1 :]
            L_m_new = row[pi_i].clone()
            if m > 0:
                L_prev = L[:m][:, pi_i]
                update = L[:m, pi_m].unsqueeze(-1)
                L_m_new -= (update * L_prev).sum(0)
            L_m_new /= L_m[pi_m]
            L_m[pi_i] = L_m_new
            matrix_diag[pi_i] = matrix_diag[pi_i] - L_m_new**2
            errors = matrix_diag[pi_i].abs().sum() / orig_error
        m += 1
    return L[:m].t().contiguous(), torch.tensor(pivots, dtype=torch.long)


@dataclass
class Preconditioner:
    L: torch.Tensor  # [N, k] pivoted Cholesky factor
    Q: torch.Tensor  # [N, k] top block of qr([L; sigma I])   (per-row noise: of qr([D^-1/2 L; I]), rows scaled by D^-1/2)
    noise: object    # float (constant diagonal) or tensor [N] (per-row noise, FixedNoiseGaussianLikelihood)
    logdet: float
    pivots: torch.Tensor | None = None

    def _d(self, like):
        d = self.noise
        return d.to(like).reshape(-1, 1) if torch.is_tensor(d) else d

    def apply(self, v):
        """P^{-1} v = (v - Q Q^T v) / sigma^2 (AddedDiagLinearOperator._preconditioner closure); with a per-row diagonal
        P^{-1} v = v / d - Q~ Q~^T v, Q~ = D^-1/2 Q  (the non-constant-diagonal closure)."""
        if torch.is_tensor(self.noise):
            return v / self._d(v) - self.Q @ (self.Q.t() @ v)
        return (v - self.Q @ (self.Q.t() @ v)) / self.noise

    def probes(self, eps1, eps2):
        """z ~ N(0, P), P = L L^T + D, as z = L eps1 + D^1/2 eps2."""
        if torch.is_tensor(self.noise):
            return self.L @ eps1 + self._d(eps2).sqrt() * eps2
        return self.L @ eps1 + math.sqrt(self.noise) * eps2


def build_preconditioner(L, noise, pivots=None):
    """QR of [L; sigma I], log det P = 2 sum log|R_ii| + (N-k) log sigma^2
    (AddedDiagLinearOperator._init_cache_for_constant_diag).  A tensor `noise` [N] takes the non-constant-diagonal branch
    (_init_cache_for_non_constant_diag): QR of [D^-1/2 L; I], Q~ = D^-1/2 Q[:N], log det P = 2 sum log|R_ii| + sum log d_i."""
    n, k = L.shape
    eye = torch.eye(k, dtype=L.dtype)
    if torch.is_tensor(noise) and noise.numel() > 1:
        d = noise.to(L.dtype).reshape(-1, 1)
        Q, R = torch.linalg.qr(torch.cat((L / d.sqrt(), eye), dim=-2))
        Q = Q[:n] / d.sqrt()
        logdet = float(R.diagonal().abs().log().sum() * 2 + d.log().sum())
        return Preconditioner(L=L, Q=Q, noise=noise.to(L.dtype).reshape(-1), logdet=logdet, pivots=pivots)
    noise = float(noise)
    Q, R = torch.linalg.qr(torch.cat((L, math.sqrt(noise) * eye), dim=-2))
    Q = Q[:n]
    logdet = float(R.diagonal().abs().log().sum() * 2 + (n - k) * math.log(noise))
    return Preconditioner(L=L, Q=Q, noise=float(noise), logdet=logdet, pivots=pivots)


def tridiag_to_diag(t_mat):
    """linear_operator.utils.lanczos.lanczos_tridiag_to_diag: eigh, negative eigenvalues -> 1 with zeroed vectors."""
    evals, evecs = torch.linalg.eigh(t_mat)
    mask = evals >= 0
    evecs = evecs * mask.to(evecs.dtype).unsqueeze(-2)
    evals = evals.masked_fill(~mask, 1)
    return evals, evecs


def slq_logdet(t_mat, n):
    """StochasticLQ.to_dense with f=log: (N/t_p) sum_i sum_j (V_i[0,j])^2 log lambda_ij."""
    evals, evecs = tridiag_to_diag(t_mat)
    tp = t_mat.size(0)
    first = evecs[..., 0, :]
    return float(n / float(tp) * (first.pow(2) * evals.log()).sum())


def inv_quad_logdet(
    matmul_closure,
    n,
    inv_quad_rhs,
    probes,
    preconditioner: Preconditioner | None = None,
    tolerance=None,
    max_iter=None,
    max_tridiag_iter=None,
    return_info=False,
):
    """LinearOperator.inv_quad_logdet / InvQuadLogdet.forward, CG branch.

    probes [N, t_p] are an input (un-normalised; Rademacher without a preconditioner,
    N(0,P) with one).  Returns (inv_quad, logdet[, info, solves, t_mat]).
    """
    if inv_quad_rhs.dim() == 1:
        inv_quad_rhs = inv_quad_rhs.unsqueeze(-1)
    tp = probes.size(-1)
    probe_norms = probes.norm(2, dim=-2, keepdim=True)
    z = probes / probe_norms
    rhs = torch.cat([z, inv_quad_rhs], -1)
    solves, t_mat, info = linear_cg(
        matmul_closure,
        rhs,
        n_tridiag=tp,
        tolerance=tolerance,
        max_iter=max_iter,
        max_tridiag_iter=max_tridiag_iter,
        preconditioner=(preconditioner.apply if preconditioner is not None else None),
        return_info=True,
    )
    if torch.isnan(t_mat).any():
        logdet = float("nan")
    else:
        logdet = slq_logdet(t_mat, n)
        if preconditioner is not None:
            logdet += preconditioner.logdet
    inv_quad = float((solves[:, tp:] * inv_quad_rhs).sum())
    if return_info:
        return inv_quad, logdet, info, solves, t_mat
    return inv_quad, logdet


def lanczos_tridiag(matmul_closure, max_iter, init_vecs, tol=1e-5):
    """Lanczos with full re-orthogonalisation, linear_operator.utils.lanczos.lanczos_tridiag.

    init_vecs [N, b].  Returns (Q [b, N, J], T [b, J, J]).
    """
    n, b = init_vecs.shape
    num_iter = min(max_iter, n)
    dtype = init_vecs.dtype
    q_mat = torch.zeros(num_iter, n, b, dtype=dtype)
    t_mat = torch.zeros(num_iter, num_iter, b, dtype=dtype)
    q_0 = init_vecs / init_vecs.norm(2, dim=-2, keepdim=True)
    q_mat[0] = q_0
    r_vec = matmul_closure(q_0)
    alpha_0 = (q_0 * r_vec).sum(-2)
    r_vec = r_vec - alpha_0.unsqueeze(-2) * q_0
    beta_0 = r_vec.norm(2, dim=-2)
    t_mat[0, 0] = alpha_0
    k = 0
    if num_iter > 1:
        t_mat[0, 1] = beta_0
        t_mat[1, 0] = beta_0
        q_mat[1] = r_vec / beta_0.unsqueeze(-2)
    for k in range(1, num_iter):
        q_prev = q_mat[k - 1]
        q_curr = q_mat[k]
        beta_prev = t_mat[k, k - 1].unsqueeze(-2)
        r_vec = matmul_closure(q_curr) - q_prev * beta_prev
        alpha_curr = (q_curr * r_vec).sum(-2, keepdim=True)
        t_mat[k, k] = alpha_curr.squeeze(-2)
        if (k + 1) < num_iter:
            r_vec = r_vec - alpha_curr * q_curr
            correction = (r_vec.unsqueeze(0) * q_mat[: k + 1]).sum(-2, keepdim=True)
            correction = (q_mat[: k + 1] * correction).sum(0)
            r_vec = r_vec - correction
            r_norm = r_vec.norm(2, dim=-2, keepdim=True)
            r_vec = r_vec / r_norm
            beta_curr = r_norm.squeeze(-2)
            t_mat[k, k + 1] = beta_curr
            t_mat[k + 1, k] = beta_curr
            inner = (q_mat[: k + 1] * r_vec.unsqueeze(0)).sum(-2)
            could_reorth = False
            for _ in range(10):
                if not torch.sum(inner > tol):
                    could_reorth = True
                    break
                correction = (r_vec.unsqueeze(0) * q_mat[: k + 1]).sum(-2, keepdim=True)
                correction = (q_mat[: k + 1] * correction).sum(0)
                r_vec = r_vec - correction
                r_norm = r_vec.norm(2, dim=-2, keepdim=True)
                r_vec = r_vec / r_norm
                inner = (q_mat[: k + 1] * r_vec.unsqueeze(0)).sum(-2)
            q_mat[k + 1] = r_vec
            if torch.sum(beta_curr.abs() > 1e-6) == 0 or not could_reorth:
                break
    num_iter = k + 1
    q_out = q_mat[:num_iter].permute(2, 1, 0).contiguous()
    t_out = t_mat[:num_iter, :num_iter].permute(2, 0, 1).contiguous()
    return q_out, t_out


def root_inv_decomposition(matmul_closure, max_iter, init_vec, tol=1e-5):
    """Lanczos root of A^{-1}: R [N, J] with R R^T ~= A^{-1} (linear_operator RootDecomposition, inverse branch, reached
    from gpytorch/models/exact_prediction_strategies.py:268-272 -- the LOVE covariance cache).  parity unpinned.

    Q, T = lanczos_tridiag(A, max_iter, init); T = V diag(lam) V^T with negative eigenvalues masked as in
    lanczos_tridiag_to_diag; R = Q V diag(lam^-1/2).
    """
    q, t = lanczos_tridiag(matmul_closure, max_iter, init_vec.reshape(-1, 1), tol)
    evals, evecs = tridiag_to_diag(t)
    return q[0] @ (evecs[0] / evals[0].sqrt())


def love_predictive_covar(k_ss, k_sx, inv_root):
    """K** - (K*x R)(K*x R)^T  (exact_prediction_strategies.py:464-478)."""
    root = k_sx @ inv_root
    return k_ss - root @ root.transpose(-1, -2)
