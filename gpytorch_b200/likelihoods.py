"""GaussianLikelihood / FixedNoiseGaussianLikelihood (gpytorch/likelihoods/gaussian_likelihood.py:117-121, :245-363;
noise_models.py:26-92, :150-190).  Module / parameter names follow the reference (`noise_covar.raw_noise`,
`second_noise_covar.raw_noise`) so that state dicts are interchangeable."""
import warnings

import torch

from .constraints import GreaterThan
from .distributions import MultivariateNormal
from .module import Module
from .operators import ConstantDiagLinearOperator, DiagLinearOperator


class HomoskedasticNoise(Module):
    """noise_models.py:26-92: one learned sigma^2 (per batch element), returned as a constant-diagonal operator."""

    def __init__(self, noise_prior=None, noise_constraint=None, batch_shape=torch.Size(), num_tasks=1):
        super().__init__()
        self.register_parameter("raw_noise", torch.nn.Parameter(torch.zeros(*batch_shape, num_tasks)))
        self.register_constraint("raw_noise", noise_constraint or GreaterThan(1e-4))  # noise_models.py:29-30

    @property
    def noise(self):
        return self.raw_noise_constraint.transform(self.raw_noise)

    @noise.setter
    def noise(self, value):
        self._set_noise(value)

    def _set_noise(self, value):
        self._set_constrained("raw_noise", value)

    def forward(self, *params, shape=None, **kwargs):
        if shape is None:
            p = params[0] if torch.is_tensor(params[0]) else params[0][0]
            shape = p.shape if p.dim() == 1 else p.shape[:-1]
        return ConstantDiagLinearOperator(self.noise, shape[-1])

    __call__ = forward


class FixedGaussianNoise(Module):
    """noise_models.py:150-190: known per-observation noise variances."""

    def __init__(self, noise):
        super().__init__()
        self.noise = noise

    def forward(self, *params, shape=None, noise=None, **kwargs):
        if shape is None:
            p = params[0] if torch.is_tensor(params[0]) else params[0][0]
            shape = p.shape if p.dim() == 1 else p.shape[:-1]
        if noise is not None:
            return DiagLinearOperator(noise)
        if shape[-1] == self.noise.shape[-1]:
            return DiagLinearOperator(self.noise)
        return None   # ZeroLinearOperator in the reference: sizes do not match and no noise was passed

    __call__ = forward

    def _apply(self, fn):
        self.noise = fn(self.noise)
        return super()._apply(fn)


class _GaussianLikelihoodBase(Module):
    def _shaped_noise_covar(self, base_shape, *params, **kwargs):
        return self.noise_covar(*params, shape=base_shape, **kwargs)

    def marginal(self, function_dist: MultivariateNormal, *params, **kwargs):
        """p(y) = N(mean, K + noise_covar): `covar + noise_covar` (gaussian_likelihood.py:117-121)."""
        mean, covar = function_dist.mean, function_dist.lazy_covariance_matrix
        noise_covar = self._shaped_noise_covar(mean.shape, *params, **kwargs)
        if noise_covar is None:
            return function_dist
        if torch.is_tensor(covar):
            full = covar + noise_covar.to_dense()
        else:
            full = covar + noise_covar
        return function_dist.__class__(mean, full)

    def __call__(self, input, *params, **kwargs):
        if isinstance(input, MultivariateNormal):
            return self.marginal(input, *params, **kwargs)
        raise RuntimeError("Likelihoods expects a MultivariateNormal input to make marginal predictions")


class GaussianLikelihood(_GaussianLikelihoodBase):
    def __init__(self, noise_prior=None, noise_constraint=None, batch_shape=torch.Size(), **kwargs):
        super().__init__()
        self.noise_covar = HomoskedasticNoise(noise_prior=noise_prior, noise_constraint=noise_constraint,
                                              batch_shape=torch.Size(batch_shape))
        self._register_load_state_dict_pre_hook(self._rename_flat_raw_noise)

    @staticmethod
    def _rename_flat_raw_noise(state_dict, prefix, *args):
        # state dicts written by round 1 of this package kept raw_noise at the top level
        if prefix + "raw_noise" in state_dict:
            state_dict[prefix + "noise_covar.raw_noise"] = state_dict.pop(prefix + "raw_noise")

    @property
    def noise(self):
        return self.noise_covar.noise

    @noise.setter
    def noise(self, value):
        self.noise_covar._set_noise(value)

    def _set_noise(self, value):
        self.noise_covar._set_noise(value)

    @property
    def raw_noise(self):
        return self.noise_covar.raw_noise

    @raw_noise.setter
    def raw_noise(self, value):
        self.noise_covar.initialize(raw_noise=value)


class FixedNoiseGaussianLikelihood(_GaussianLikelihoodBase):
    """Known heteroscedastic observation noise (+ optionally a learned homoskedastic term): gaussian_likelihood.py:245-363."""

    def __init__(self, noise, learn_additional_noise=False, batch_shape=torch.Size(), **kwargs):
        super().__init__()
        self.noise_covar = FixedGaussianNoise(noise=noise)
        self.second_noise_covar = None
        if learn_additional_noise:
            self.second_noise_covar = HomoskedasticNoise(noise_prior=kwargs.get("noise_prior"),
                                                         noise_constraint=kwargs.get("noise_constraint"),
                                                         batch_shape=torch.Size(batch_shape))

    @property
    def noise(self):
        return self.noise_covar.noise + self.second_noise

    @noise.setter
    def noise(self, value):
        self.noise_covar.noise = value

    @property
    def second_noise(self):
        return 0.0 if self.second_noise_covar is None else self.second_noise_covar.noise

    @second_noise.setter
    def second_noise(self, value):
        if self.second_noise_covar is None:
            raise RuntimeError("Attempting to set secondary learned noise for FixedNoiseGaussianLikelihood, "
                               "but learn_additional_noise must have been False!")
        self.second_noise_covar._set_noise(value)

    def _shaped_noise_covar(self, base_shape, *params, **kwargs):
        res = self.noise_covar(*params, shape=base_shape, **kwargs)
        if self.second_noise_covar is not None:
            second = self.second_noise_covar(*params, shape=base_shape, **kwargs)
            res = second if res is None else res + second
        elif res is None:
            warnings.warn("You have passed data through a FixedNoiseGaussianLikelihood that did not match the size "
                          "of the fixed noise, *and* you did not specify noise. This is treated as a no-op.")
        return res
