"""ExactMarginalLogLikelihood (gpytorch/mlls/exact_marginal_log_likelihood.py:54-89)."""
from .distributions import MultivariateNormal
from .likelihoods import _GaussianLikelihoodBase
from .module import Module


class ExactMarginalLogLikelihood(Module):
    def __init__(self, likelihood, model):
        if not isinstance(likelihood, _GaussianLikelihoodBase):
            raise RuntimeError("Likelihood must be Gaussian for exact inference")
        super().__init__()
        self.likelihood = likelihood
        self.model = model

    def forward(self, function_dist, target, *params, **kwargs):
        if not isinstance(function_dist, MultivariateNormal):
            raise RuntimeError("ExactMarginalLogLikelihood can only operate on Gaussian random variables")
        output = self.likelihood(function_dist, *params, **kwargs)
        res = output.log_prob(target)
        num_data = function_dist.event_shape.numel()
        return res.div_(num_data) if not res.requires_grad else res / num_data

    def __call__(self, *a, **k):
        return self.forward(*a, **k)
