"""Applied-field factories needed by the hot path (reference sources/constant.py:8-32)."""
from __future__ import annotations

import numpy as np


class Parameter:
    """A callable ``f(x, y[, z])`` with bound keyword arguments (minimal stand-in for
    superscreen/parameter.py; arithmetic composition is out of scope)."""

    def __init__(self, func, **kwargs):
        self.func = func
        self.kwargs = kwargs

    def __call__(self, x, y, z=None):
        x, y = np.atleast_1d(x, y)
        if z is None:
            return np.asarray(self.func(x, y, **self.kwargs))
        z = np.atleast_1d(z)
        if z.ndim > 1:
            z = np.squeeze(z, axis=tuple(range(1, z.ndim)))
        return np.asarray(self.func(x, y, z, **self.kwargs))

    def __repr__(self):
        return f"Parameter<{getattr(self.func, '__name__', 'f')}({self.kwargs})>"


class Constant(Parameter):
    def __init__(self, value, dimensions: int = 2):
        if dimensions == 2:
            super().__init__(lambda x, y, value=0: value * np.ones_like(x, dtype=float), value=value)
        else:
            super().__init__(lambda x, y, z, value=0: value * np.ones_like(x, dtype=float), value=value)


def constant(x, y, z, value=0):
    return value * np.ones_like(x, dtype=float)


def ConstantField(value: float = 0) -> Parameter:
    """A Parameter that returns ``value`` at all ``x, y, z``."""
    return Parameter(constant, value=float(value))
