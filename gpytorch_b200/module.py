"""Minimal Module with constrained parameters (mirror of gpytorch/module.py:238-349 for the hot path)."""
import torch

from .constraints import Interval


class Module(torch.nn.Module):
    def register_constraint(self, param_name: str, constraint: Interval):
        self.add_module(param_name + "_constraint", constraint)

    def constraint_for(self, param_name: str):
        return getattr(self, param_name + "_constraint", None)

    def initialize(self, **kwargs):
        """Set parameters by (constrained) value, e.g. kernel.initialize(lengthscale=2.) (module.py:88-141)."""
        for name, val in kwargs.items():
            if hasattr(self, "raw_" + name):
                setter = getattr(self, "_set_" + name)
                setter(val)
            elif isinstance(getattr(self, name, None), torch.nn.Parameter):
                with torch.no_grad():
                    getattr(self, name).copy_(torch.as_tensor(val).expand_as(getattr(self, name)))
            else:
                raise AttributeError(f"Unknown parameter {name} for {self.__class__.__name__}")
        return self

    def _set_constrained(self, raw_name, value):
        raw = getattr(self, raw_name)
        c = self.constraint_for(raw_name)
        value = torch.as_tensor(value, dtype=raw.dtype, device=raw.device)
        with torch.no_grad():
            raw.copy_((c.inverse_transform(value) if c is not None else value).expand_as(raw))
