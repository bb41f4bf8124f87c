"""GaussianLikelihood (gpytorch/likelihoods/gaussian_likelihood.py:117-121; noise_models.py:29-30,57-92)."""
import torch

from .constraints import GreaterThan
from .distributions import MultivariateNormal
from .module import Module
from .operators import ConstantDiagLinearOperator


class GaussianLikelihood(Module):
    def __init__(self, noise_prior=None, noise_constraint=None, batch_shape=torch.Size(), **kwargs):
        super().__init__()
        self.register_parameter("raw_noise", torch.nn.Parameter(torch.zeros(1)))
        self.register_constraint("raw_noise", noise_constraint or GreaterThan(1e-4))  # noise_models.py:29-30

    @property
    def noise(self):
        return self.raw_noise_constraint.transform(self.raw_noise)

    @noise.setter
    def noise(self, value):
        self._set_noise(value)

    def _set_noise(self, value):
        self._set_constrained("raw_noise", value)

    def marginal(self, function_dist: MultivariateNormal, *params, **kwargs):
        """p(y) = N(mean, K + sigma^2 I): `covar + noise_covar` (gaussian_likelihood.py:117-121)."""
        mean, covar = function_dist.mean, function_dist.lazy_covariance_matrix
        noise_covar = ConstantDiagLinearOperator(self.noise, mean.shape[-1])
        if torch.is_tensor(covar):
            full = covar + noise_covar.to_dense()
        else:
            full = covar + noise_covar
        return function_dist.__class__(mean, full)

    def __call__(self, input, *params, **kwargs):
        if isinstance(input, MultivariateNormal):
            return self.marginal(input, *params, **kwargs)
        raise RuntimeError("Likelihoods expects a MultivariateNormal input to make marginal predictions")
