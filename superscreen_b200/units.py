"""A very small dimensional-analysis unit system.

The reference leans on ``pint`` (not installable here) for a handful of conversions on the
solve path: ``field_conversion_factor`` / ``convert_field`` (solver/utils.py:350-437), the
vortex flux ``Phi_0 / mu_0`` (solver/solve.py:441), current strings such as ``"1 mA"``
(solver/utils.py:327-347), fluxoid / inductance units (solution.py:535-559,
device/device.py:637-639).  This module provides exactly that: unit expressions are parsed
with ``ast`` into (scale, dimension-vector) pairs over [length, mass, time, current].
"""
from __future__ import annotations

import ast
import functools
import math
import re
from typing import Tuple, Union

import numpy as np

MU_0 = 1.25663706212e-06  # N / A^2, the value in pint's default registry (CODATA 2018)
PHI_0 = 2.067833848461929e-15  # Wb  (h / 2e)

_DIMS = ("length", "mass", "time", "current")
_L, _M, _T, _I = (np.array(v, dtype=float) for v in ((1, 0, 0, 0), (0, 1, 0, 0), (0, 0, 1, 0), (0, 0, 0, 1)))
_ZERO = np.zeros(4)

_PREFIX = {"": 1.0, "G": 1e9, "M": 1e6, "k": 1e3, "c": 1e-2, "m": 1e-3, "u": 1e-6, "µ": 1e-6,
           "μ": 1e-6, "n": 1e-9, "p": 1e-12, "f": 1e-15, "a": 1e-18}

_B = _M - 2 * _T - _I                # tesla: kg s^-2 A^-1
_WB = _B + 2 * _L                    # weber
_H = _WB - _I                        # henry

_BASE = {
    "m": (1.0, _L), "meter": (1.0, _L), "metre": (1.0, _L),
    "g": (1e-3, _M), "s": (1.0, _T), "second": (1.0, _T),
    "A": (1.0, _I), "amp": (1.0, _I), "ampere": (1.0, _I),
    "T": (1.0, _B), "tesla": (1.0, _B), "G": (1e-4, _B), "gauss": (1e-4, _B),
    "Wb": (1.0, _WB), "weber": (1.0, _WB), "H": (1.0, _H), "henry": (1.0, _H),
    "Oe": (1e3 / (4 * math.pi), _I - _L), "oersted": (1e3 / (4 * math.pi), _I - _L),
    "N": (1.0, _L + _M - 2 * _T), "J": (1.0, 2 * _L + _M - 2 * _T),
    "V": (1.0, 2 * _L + _M - 3 * _T - _I), "ohm": (1.0, 2 * _L + _M - 3 * _T - 2 * _I),
}
_CONST = {
    "mu_0": (MU_0, _L + _M - 2 * _T - 2 * _I), "mu0": (MU_0, _L + _M - 2 * _T - 2 * _I),
    "Phi_0": (PHI_0, _WB), "Phi0": (PHI_0, _WB),
    "mu_B": (9.2740100783e-24, _I + 2 * _L), "bohr_magneton": (9.2740100783e-24, _I + 2 * _L),  # A m^2 (CODATA 2018)
    "dimensionless": (1.0, _ZERO),
}
_LONG_PREFIX = {"micro": 1e-6, "milli": 1e-3, "nano": 1e-9, "pico": 1e-12, "kilo": 1e3}


class DimensionalityError(ValueError):
    pass


def _lookup(name: str) -> Tuple[float, np.ndarray]:
    if name in _CONST:
        return _CONST[name]
    if name in _BASE:
        return _BASE[name]
    for lp, f in _LONG_PREFIX.items():
        if name.startswith(lp) and name[len(lp):] in _BASE:
            s, d = _BASE[name[len(lp):]]
            return f * s, d
    for k in (1, 2):
        pre, base = name[:k], name[k:]
        if pre in _PREFIX and pre and base in _BASE:
            s, d = _BASE[base]
            return _PREFIX[pre] * s, d
    raise ValueError(f"Unknown unit {name!r}")


def _eval(node) -> Tuple[float, np.ndarray]:
    if isinstance(node, ast.Expression):
        return _eval(node.body)
    if isinstance(node, ast.Constant) and isinstance(node.value, (int, float)):
        return float(node.value), _ZERO
    if isinstance(node, ast.Name):
        return _lookup(node.id)
    if isinstance(node, ast.UnaryOp) and isinstance(node.op, (ast.USub, ast.UAdd)):
        s, d = _eval(node.operand)
        return (-s if isinstance(node.op, ast.USub) else s), d
    if isinstance(node, ast.BinOp):
        ls, ld = _eval(node.left)
        if isinstance(node.op, ast.Pow):
            rs, rd = _eval(node.right)
            if np.any(rd):
                raise ValueError("exponent must be dimensionless")
            return ls**rs, ld * rs
        rs, rd = _eval(node.right)
        if isinstance(node.op, ast.Mult):
            return ls * rs, ld + rd
        if isinstance(node.op, ast.Div):
            return ls / rs, ld - rd
    raise ValueError(f"Cannot parse unit expression node {ast.dump(node)}")


def parse(expr: str) -> Tuple[float, np.ndarray]:
    """``"1 mA"``, ``"uA / um"``, ``"mT * um ** 2"``, ``"Phi_0 / A"`` -> (SI scale, dims)."""
    if isinstance(expr, Unit):
        return expr.scale, expr.dims
    return _parse_text(str(expr))


@functools.lru_cache(maxsize=512)
def _parse_text(text: str) -> Tuple[float, np.ndarray]:
    text = text.strip().replace("^", "**")
    # implicit multiplication between a leading number and a unit: "1 mA" -> "1 * mA"
    text = re.sub(r"^([-+]?[0-9.]+(?:[eE][-+]?[0-9]+)?)\s+(?=[A-Za-zµμ])", r"\1 * ", text)
    text = text.replace("µ", "u").replace("μ", "u")
    return _eval(ast.parse(text, mode="eval"))


class Unit:
    def __init__(self, expr):
        self.expr = str(expr)
        self.scale, self.dims = parse(expr)

    def __repr__(self):
        return f"Unit({self.expr!r})"


def conversion_factor(old: str, new: str) -> float:
    """Multiply a magnitude in ``old`` units by this to express it in ``new`` units."""
    if isinstance(old, str) and isinstance(new, str):
        return _conversion_factor_text(old, new)
    return _conversion_factor(old, new)


@functools.lru_cache(maxsize=1024)
def _conversion_factor_text(old: str, new: str) -> float:
    return _conversion_factor(old, new)


def _conversion_factor(old, new) -> float:
    so, do = parse(old)
    sn, dn = parse(new)
    if not np.allclose(do, dn):
        raise DimensionalityError(f"Cannot convert {old!r} to {new!r}")
    return so / sn


def same_dimension(a: str, b: str) -> bool:
    return bool(np.allclose(parse(a)[1], parse(b)[1]))


class Quantity:
    """Minimal stand-in for ``pint.Quantity``: magnitude + unit string."""

    __array_priority__ = 1000

    def __init__(self, magnitude, units: str):
        self.magnitude = magnitude
        self.units = str(units)

    @property
    def m(self):
        return self.magnitude

    def to(self, units: str) -> "Quantity":
        return Quantity(self.magnitude * conversion_factor(self.units, units), units)

    def __iter__(self):
        for v in np.atleast_1d(self.magnitude):
            yield Quantity(v, self.units)

    def __add__(self, other):
        if isinstance(other, Quantity):
            return Quantity(self.magnitude + other.to(self.units).magnitude, self.units)
        if np.isscalar(other) and other == 0:
            return self
        return NotImplemented

    __radd__ = __add__

    def __mul__(self, other):
        if isinstance(other, Quantity):
            return Quantity(self.magnitude * other.magnitude, f"({self.units}) * ({other.units})")
        return Quantity(self.magnitude * other, self.units)

    __rmul__ = __mul__

    def __truediv__(self, other):
        if isinstance(other, Quantity):
            return Quantity(self.magnitude / other.magnitude, f"({self.units}) / ({other.units})")
        return Quantity(self.magnitude / other, self.units)

    def __repr__(self):
        return f"<Quantity({self.magnitude!r}, {self.units!r})>"

    def __float__(self):
        return float(self.magnitude)

    def __array__(self, dtype=None, copy=None):
        return np.asarray(self.magnitude, dtype=dtype)


def to_quantity(value: Union[str, float, Quantity], default_units: str) -> Quantity:
    """``"1 mA"`` / 3.0 / Quantity -> Quantity (floats are taken to be in ``default_units``)."""
    if isinstance(value, Quantity):
        return value
    if isinstance(value, str):
        m = re.match(r"^\s*([-+]?[0-9.]+(?:[eE][-+]?[0-9]+)?)\s*(.*)$", value)
        if m and m.group(2).strip():
            return Quantity(float(m.group(1)), m.group(2).strip())
        if m:
            return Quantity(float(m.group(1)), default_units)
        return Quantity(1.0, value)
    return Quantity(value, default_units)
