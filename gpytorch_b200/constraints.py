"""Parameter constraints (mirror of gpytorch/constraints/constraints.py:156-194): softplus transforms."""
import math

import torch
from torch.nn.functional import softplus


def inv_softplus(x):
    return x + torch.log(-torch.expm1(-x))


class Interval(torch.nn.Module):
    def __init__(self, lower_bound=-math.inf, upper_bound=math.inf):
        super().__init__()
        self.lower_bound = float(lower_bound)
        self.upper_bound = float(upper_bound)

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
