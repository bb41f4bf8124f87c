"""MultivariateNormal with a lazy covariance (gpytorch/distributions/multivariate_normal.py:44-64, :221-252)."""
import math

import torch



class MultivariateNormal:
    def __init__(self, mean, covariance_matrix):
        self.loc = mean
        self._covar = covariance_matrix
        if mean.shape[-1] != covariance_matrix.shape[-1]:
            raise RuntimeError("mean and covariance sizes do not match")

    @property
    def mean(self):
        return self.loc

    @property
    def lazy_covariance_matrix(self):
        return self._covar

    @property
    def covariance_matrix(self):
        c = self._covar
        return c if torch.is_tensor(c) else c.to_dense()

    @property
    def variance(self):
        c = self._covar
        return c.diagonal(dim1=-2, dim2=-1) if torch.is_tensor(c) else c.diagonal()

    @property
    def event_shape(self):
        return self.loc.shape[-1:]

    def log_prob(self, value):
        """-0.5 (r^T K^-1 r + log det K + N log 2 pi), r = value - mean (multivariate_normal.py:221-252)."""
        mean, covar = self.loc, self._covar
        diff = value - mean
        if torch.is_tensor(covar):  # dense covariance: plain Cholesky
            chol = torch.linalg.cholesky(covar)
            sol = torch.cholesky_solve(diff.unsqueeze(-1), chol)
            inv_quad = (diff.unsqueeze(-1) * sol).sum((-2, -1))
            logdet = 2 * chol.diagonal(dim1=-2, dim2=-1).log().sum(-1)
        else:
            covar = covar.evaluate_kernel()
            inv_quad, logdet = covar.inv_quad_logdet(inv_quad_rhs=diff.unsqueeze(-1), logdet=True)
        return -0.5 * sum([inv_quad, logdet, diff.size(-1) * math.log(2 * math.pi)])
