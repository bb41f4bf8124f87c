"""Solver knobs with the reference's names and defaults (gpytorch/settings.py:6-31 re-exports the
linear_operator settings; own knobs :173-180, :261-269).  Class-level context managers, as in the
reference (settings.py:84-144): `with settings.cg_tolerance(1e-4): ...`, `settings.cg_tolerance.value()`.
"""
from __future__ import annotations


class _value_context:
    _global_value = None

    @classmethod
    def value(cls):
        return cls._global_value

    @classmethod
    def _set_value(cls, value):
        cls._global_value = value

    def __init__(self, value):
        self._orig_value = self.__class__.value()
        self._instance_value = value

    def __enter__(self):
        self.__class__._set_value(self._instance_value)

    def __exit__(self, *args):
        self.__class__._set_value(self._orig_value)
        return False


class _feature_flag:
    _default = False
    _state = None

    @classmethod
    def on(cls):
        return cls._default if cls._state is None else cls._state

    @classmethod
    def off(cls):
        return not cls.on()

    @classmethod
    def _set_state(cls, state):
        cls._state = state

    def __init__(self, state=True):
        self.prev = self.__class__._state
        self.state = state

    def __enter__(self):
        self.__class__._set_state(self.state)

    def __exit__(self, *args):
        self.__class__._set_state(self.prev)
        return False


class cg_tolerance(_value_context):
    _global_value = 1.0


class eval_cg_tolerance(_value_context):
    _global_value = 0.01  # gpytorch/settings.py:173-180


class max_cg_iterations(_value_context):
    _global_value = 1000


class max_cholesky_size(_value_context):
    _global_value = 800


class max_lanczos_quadrature_iterations(_value_context):
    _global_value = 20


class max_preconditioner_size(_value_context):
    _global_value = 15


class min_preconditioning_size(_value_context):
    _global_value = 2000


class num_trace_samples(_value_context):
    _global_value = 10


class preconditioner_tolerance(_value_context):
    _global_value = 1e-3


class max_root_decomposition_size(_value_context):
    _global_value = 100


class skip_logdet_forward(_feature_flag):
    _default = False


class fast_pred_var(_feature_flag):
    """LOVE predictive variances: K_hat^{-1} ~= R R^T from `max_root_decomposition_size` Lanczos steps
    (settings.py:183-222; models/exact_prediction_strategies.py:268-272, 464-478)."""
    _default = False


class skip_posterior_variances(_feature_flag):
    """Return a zero predictive covariance (models/exact_prediction_strategies.py:432-433)."""
    _default = False


class _use_eval_tolerance(_feature_flag):
    _default = False


# ---- engine-specific knobs (no reference counterpart) ----
class backend(_value_context):
    """'auto' | 'tcgen05' | 'simt': which fused K.V kernel the engine runs."""
    _global_value = "auto"


class probe_seed(_value_context):
    """Seed of the base samples for the SLQ probes (None = draw from torch's global CUDA generator)."""
    _global_value = None


# ---- propagation into worker threads (operators.BatchLinearOperator) ----
# The knobs are process-wide class attributes (as in the reference, settings.py:84-144), so worker threads already see the values
# set by the calling thread's `with` blocks; snapshot() / restore() exist so that the hand-over is explicit and testable.
def snapshot():
    import sys
    mod = sys.modules[__name__]
    out = {}
    for name, obj in vars(mod).items():
        if isinstance(obj, type) and issubclass(obj, _value_context) and obj is not _value_context:
            out[name] = ("v", obj.value())
        elif isinstance(obj, type) and issubclass(obj, _feature_flag) and obj is not _feature_flag:
            out[name] = ("f", obj._state)
    return out


class restore:
    def __init__(self, snap):
        self.snap = snap

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False
