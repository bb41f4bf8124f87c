"""Parameter constraints (mirror of gpytorch/constraints/constraints.py:17-194): softplus transforms; the bounds are
registered buffers named `lower_bound` / `upper_bound` (:44-45) so that state dicts are interchangeable."""
import math

import torch
from torch.nn.functional import softplus


def inv_softplus(x):
    return x + torch.log(-torch.expm1(-x))


class Interval(torch.nn.Module):
    def __init__(self, lower_bound=-math.inf, upper_bound=math.inf):
        super().__init__()
        dtype = torch.get_default_dtype()
        lower_bound = torch.as_tensor(lower_bound).to(dtype)
        upper_bound = torch.as_tensor(upper_bound).to(dtype)
        if torch.any(torch.ge(lower_bound, upper_bound)):
            raise ValueError("Got parameter bounds with empty intervals.")   # constraints.py:31-32
        self.register_buffer("lower_bound", lower_bound)
        self.register_buffer("upper_bound", upper_bound)

    def _load_from_state_dict(self, state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys, error_msgs):
        # constraints.py:62-77: the bound buffers may be absent from older state dicts -> never strict here
        return super()._load_from_state_dict(state_dict, prefix, local_metadata, False, missing_keys, unexpected_keys, error_msgs)

    def check(self, tensor) -> bool:
        return bool(torch.all(tensor <= self.upper_bound) and torch.all(tensor >= self.lower_bound))

    def transform(self, raw):
        return raw

    def inverse_transform(self, value):
        return value


class GreaterThan(Interval):
    def __init__(self, lower_bound):
        super().__init__(lower_bound, math.inf)

    def transform(self, raw):
        return softplus(raw) + self.lower_bound

    def inverse_transform(self, value):
        return inv_softplus(value - self.lower_bound)


class Positive(GreaterThan):
    def __init__(self):
        super().__init__(0.0)
