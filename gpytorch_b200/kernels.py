"""gpytorch.kernels.{Kernel, RBFKernel, MaternKernel, ScaleKernel} for the accelerated path.

Same constructor kwargs and call contract as the reference (kernels/kernel.py:163-171, :454-534;
rbf_kernel.py:68-85; matern_kernel.py:79-110; scale_kernel.py:64-118), but `forward` returns an
engine-backed KernelLinearOperator (the KeOps plug-in pattern, kernels/keops/rbf_kernel.py:44-55)
instead of a dense tensor: K is never materialised.
"""
from __future__ import annotations

import torch

from .constraints import Positive
from .module import Module
from .operators import KernelLinearOperator


class Kernel(Module):
    has_lengthscale = False
    kind = None

    def __init__(self, ard_num_dims=None, batch_shape=None, active_dims=None, lengthscale_prior=None,
                 lengthscale_constraint=None, eps=1e-6, **kwargs):
        super().__init__()
        if batch_shape is not None and len(batch_shape) > 0:
            raise NotImplementedError("batched hyper-parameters: launch one operator per batch element (SURVEY.md section 8, C4)")
        self.ard_num_dims = ard_num_dims
        self.active_dims = None if active_dims is None else torch.as_tensor(active_dims, dtype=torch.long)
        self.eps = eps
        if self.has_lengthscale:
            n = 1 if ard_num_dims is None else ard_num_dims
            self.register_parameter("raw_lengthscale", torch.nn.Parameter(torch.zeros(1, n)))
            self.register_constraint("raw_lengthscale", lengthscale_constraint or Positive())

    @property
    def lengthscale(self):
        return self.raw_lengthscale_constraint.transform(self.raw_lengthscale) if self.has_lengthscale else None

    @lengthscale.setter
    def lengthscale(self, value):
        self._set_lengthscale(value)

    def _set_lengthscale(self, value):
        if not self.has_lengthscale:
            raise RuntimeError("Kernel has no lengthscale.")
        self._set_constrained("raw_lengthscale", value)

    def forward(self, x1, x2, diag=False, **params):
        raise NotImplementedError

    def __call__(self, x1, x2=None, diag=False, **params):
        # kernels/kernel.py:454-534: active dims, 1-D -> 2-D, x2=None -> x1, size check
        if self.active_dims is not None:
            idx = self.active_dims.to(x1.device)
            x1 = x1.index_select(-1, idx)
            if x2 is not None:
                x2 = x2.index_select(-1, idx)
        if x1.dim() == 1:
            x1 = x1.unsqueeze(1)
        if x2 is not None:
            if x2.dim() == 1:
                x2 = x2.unsqueeze(1)
            if x1.size(-1) != x2.size(-1):
                raise RuntimeError("x1_ and x2_ must have the same number of dimensions!")
        if x1.dim() != 2:
            raise NotImplementedError("batched inputs are evaluated one operator per batch element")
        if self.ard_num_dims is not None and self.ard_num_dims != x1.size(-1):
            raise RuntimeError(f"Expected the input to have {self.ard_num_dims} dimensionality "
                               f"(based on the ard_num_dims argument). Got {x1.size(-1)}.")
        same = x2 is None
        res = self.forward(x1.contiguous(), x1.contiguous() if same else x2.contiguous(), diag=diag, _same=same, **params)
        return res


class _StationaryKernel(Kernel):
    has_lengthscale = True

    def forward(self, x1, x2, diag=False, _same=False, **params):
        ls = self.lengthscale.reshape(-1)
        ls = ls[0] if ls.numel() == 1 else ls
        op = KernelLinearOperator(x1, None if _same else x2, self.kind, ls)
        if diag:
            return op.diagonal()
        return op


class RBFKernel(_StationaryKernel):
    """k = exp(-0.5 |x1 - x2|^2 / l^2)  (kernels/rbf_kernel.py)."""
    kind = "rbf"


class MaternKernel(_StationaryKernel):
    """Matern nu in {0.5, 1.5, 2.5}  (kernels/matern_kernel.py:79-110)."""

    def __init__(self, nu=2.5, **kwargs):
        if nu not in {0.5, 1.5, 2.5}:
            raise RuntimeError("nu expected to be 0.5, 1.5, or 2.5")
        super().__init__(**kwargs)
        self.nu = nu

    @property
    def kind(self):
        return {0.5: "matern12", 1.5: "matern32", 2.5: "matern52"}[self.nu]


class ScaleKernel(Kernel):
    """K <- outputscale * base(K)  (kernels/scale_kernel.py:64-118); the scale is folded into the fused kernel."""

    def __init__(self, base_kernel, outputscale_prior=None, outputscale_constraint=None, **kwargs):
        if base_kernel.active_dims is not None:
            kwargs["active_dims"] = base_kernel.active_dims
        super().__init__(**kwargs)
        self.base_kernel = base_kernel
        self.register_parameter("raw_outputscale", torch.nn.Parameter(torch.zeros(())))
        self.register_constraint("raw_outputscale", outputscale_constraint or Positive())

    @property
    def outputscale(self):
        return self.raw_outputscale_constraint.transform(self.raw_outputscale)

    @outputscale.setter
    def outputscale(self, value):
        self._set_outputscale(value)

    def _set_outputscale(self, value):
        self._set_constrained("raw_outputscale", value)

    def forward(self, x1, x2, diag=False, _same=False, **params):
        base = self.base_kernel.forward(x1, x2, diag=False, _same=_same, **params)
        op = KernelLinearOperator(base.x1, base.x2, base.kind, base.lengthscale, self.outputscale)
        return op.diagonal() if diag else op

    def __call__(self, x1, x2=None, diag=False, **params):
        self.active_dims = None  # selection happens once, here (base active_dims were lifted in __init__)
        if self.base_kernel.active_dims is not None:
            idx = self.base_kernel.active_dims.to(x1.device)
            x1 = x1.index_select(-1, idx)
            x2 = None if x2 is None else x2.index_select(-1, idx)
        return Kernel.__call__(self, x1, x2, diag=diag, **params)
